// attention.cu — tcgen05 self-attention for the transformer expert (d_model 1024, 16 heads x 64, seq 512;
// reference: nn.MultiheadAttention inside /root/reference/experiments/throughput/layers.py:22-51).
//
// FORWARD (this file): ping-pong flash attention.  One CTA owns a PAIR of 128-query tiles of one (batch, head); K / V stream
// through a 2-stage TMA pipeline in 128-key blocks shared by both tiles; warps 0-3 / 4-7 are the softmax groups of the two
// tiles (thread = query row: the 128 scores of a block are read from TMEM once and stay in registers), warp 8 issues
// tcgen05.mma, warp 9 drives TMA; while one group turns S_j into P_j the tensor core works for the other; online softmax
// with the running output O in TMEM, rescaled in place (tcgen05.ld -> mul -> tcgen05.st) only when the row maximum grew by
// more than 2^8 (lazy rescale).  S and P never touch HBM.  The kernel also emits the row log-sum-exp (base 2) that the
// BACKWARD kernel (attention_bwd.cu) needs to recompute P without a second softmax pass.
//
// Input : qkv [T = batch*512, 3*D] bf16 (output of the fused in_proj GEMM: [q | k | v] per token, heads contiguous)
// Output: out [T, D] bf16 (heads concatenated, ready for out_proj); lse2 [T, heads] fp32 (optional)
#include "sm100.cuh"
#include <stdlib.h>

namespace lah {
namespace attn {

constexpr int S_LEN = 512;      // keys per sequence
constexpr int HEAD_DIM = 64;
constexpr int Q_TILE = 128;

namespace v2 {

constexpr int KB = 128;                       // keys per block
constexpr int NUM_KB = S_LEN / KB;            // 4
constexpr int NUM_THREADS2 = 320;
constexpr int TILE_BYTES = 128 * HEAD_DIM * 2;            // 16 KB: a 128 x 64 bf16 tile (Q tile, K block, V block)
constexpr int P_BYTES = Q_TILE * KB * 2;                  // 32 KB per query tile
constexpr int OFF_Q2 = 0;                                 // Q_A, Q_B
constexpr int OFF_K2 = OFF_Q2 + 2 * TILE_BYTES;           // 2 stages
constexpr int OFF_V2 = OFF_K2 + 2 * TILE_BYTES;
constexpr int OFF_P2 = OFF_V2 + 2 * TILE_BYTES;           // P_A, P_B
constexpr int OFF_BAR2 = OFF_P2 + 2 * P_BYTES;
constexpr int NUM_BARS = 1 + 8 + 8;
constexpr int SMEM_TOTAL2 = OFF_BAR2 + NUM_BARS * 8 + 16 + 1024;
constexpr int COL_S = 0, COL_O = 256;

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// (320 threads x 200 registers did not launch (cudaErrorLaunchOutOfResources); the launch-bounds build uses 168)
__global__ void __launch_bounds__(NUM_THREADS2, 1)
attention_fwd_v2_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ out, float* __restrict__ lse2,
                        int d_model, int num_heads, float scale_log2e) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR2);
    uint64_t* bar_q = bars;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // [2 tiles]  MMA -> softmax
    uint64_t* s_free = bars + 11;   // [2 tiles]  softmax (128 threads) -> MMA
    uint64_t* p_full = bars + 13;   // [2 tiles]  softmax (128 threads) -> MMA
    uint64_t* pv_done = bars + 15;  // [2 tiles]  MMA -> softmax
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qpair = blockIdx.x & 1;
    const int head = (blockIdx.x >> 1) % num_heads;
    const int batch = (blockIdx.x >> 1) / num_heads;
    const int seq_row0 = batch * S_LEN;

    if (warp == 9 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 128);
            mbar_init(&p_full[i], 128);
            mbar_init(&pv_done[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_ptr, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 9) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            mbar_arrive_expect_tx(bar_q, 2 * TILE_BYTES);
            tma_load_2d(smem + OFF_Q2, &tm_qkv, bar_q, head * HEAD_DIM, seq_row0 + (2 * qpair) * Q_TILE);
            tma_load_2d(smem + OFF_Q2 + TILE_BYTES, &tm_qkv, bar_q, head * HEAD_DIM, seq_row0 + (2 * qpair + 1) * Q_TILE);
            for (int j = 0; j < NUM_KB; ++j) {
                const int st = j & 1;
                const uint32_t par = ((j >> 1) & 1) ^ 1;
                mbar_wait(&k_empty[st], par);
                mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
                tma_load_2d(smem + OFF_K2 + st * TILE_BYTES, &tm_qkv, &k_full[st], d_model + head * HEAD_DIM, seq_row0 + j * KB);
                mbar_wait(&v_empty[st], par);
                mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
                tma_load_2d(smem + OFF_V2 + st * TILE_BYTES, &tm_qkv, &v_full[st], 2 * d_model + head * HEAD_DIM, seq_row0 + j * KB);
            }
        }
    } else if (warp == 8) {
        if (lane == 0) {
            // ------------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc_s = make_idesc_bf16_f32(Q_TILE, KB, 0u, 0u);
            constexpr uint32_t idesc_o = make_idesc_bf16_f32(Q_TILE, HEAD_DIM, 0u, 1u);
            const uint32_t sq = smem_u32(smem + OFF_Q2), sk = smem_u32(smem + OFF_K2), sv = smem_u32(smem + OFF_V2),
                           sp = smem_u32(smem + OFF_P2);
            auto issue_s = [&](int t, int j) {
                const uint32_t a = sq + t * TILE_BYTES, b = sk + (j & 1) * TILE_BYTES;
#pragma unroll
                for (int ks = 0; ks < HEAD_DIM / 16; ++ks)
                    umma_bf16_ss(tmem_base + COL_S + t * KB, make_smem_desc_sw128(a + ks * 32, 0, 1024),
                                 make_smem_desc_sw128(b + ks * 32, 0, 1024), idesc_s, ks > 0 ? 1u : 0u);
                umma_commit(&s_full[t]);
            };
            mbar_wait(bar_q, 0);
            mbar_wait(&k_full[0], 0);
            tcgen05_fence_after();
            issue_s(0, 0);
            issue_s(1, 0);
            umma_commit(&k_empty[0]);
            for (int j = 0; j < NUM_KB; ++j) {
                const int st = j & 1;
                for (int t = 0; t < 2; ++t) {
                    if (j + 1 < NUM_KB) {   // S of the NEXT block first: the softmax warps hold S_t(j) in registers already
                        if (t == 0) mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
                        mbar_wait(&s_free[t], j & 1);
                        tcgen05_fence_after();
                        issue_s(t, j + 1);
                        if (t == 1) umma_commit(&k_empty[(j + 1) & 1]);
                    }
                    mbar_wait(&p_full[t], j & 1);
                    if (t == 0) mbar_wait(&v_full[st], (j >> 1) & 1);
                    tcgen05_fence_after();
                    const uint32_t a = sp + t * P_BYTES, b = sv + st * TILE_BYTES;
#pragma unroll
                    for (int kb = 0; kb < KB / 64; ++kb) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_bf16_ss(tmem_base + COL_O + t * HEAD_DIM,
                                         make_smem_desc_sw128(a + kb * (Q_TILE * 128) + ks * 32, 0, 1024),
                                         make_smem_desc_sw128(b + (kb * 64 + ks * 16) * 128, 0, 1024), idesc_o,
                                         (j | kb | ks) ? 1u : 0u);
                    }
                    umma_commit(&pv_done[t]);
                    if (t == 1) umma_commit(&v_empty[st]);
                }
            }
        }
    } else {
        // ---------------------------------------------------------------------- softmax + epilogue: thread = query row
        const int t = warp >> 2;                 // query tile of the pair
        const int row = (warp & 3) * 32 + lane;  // row inside the tile == TMEM lane
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t ts = lane_base + COL_S + t * KB, to = lane_base + COL_O + t * HEAD_DIM;
        uint8_t* pbase = smem + OFF_P2 + t * P_BYTES;
        // m = the reference maximum the probabilities are currently expressed against (raw score units).  LAZY rescale:
        // m only moves when the row maximum grew by more than 8 / scale_log2e, i.e. probabilities stay below 2^8 — exact
        // after the final division by l, and the TMEM round trip of O (ld -> mul -> st) is skipped for most blocks.
        float m = -INFINITY, l = 0.f;
        const float lazy_margin = 8.f / scale_log2e;
#pragma unroll 1
        for (int j = 0; j < NUM_KB; ++j) {
            mbar_wait(&s_full[t], j & 1);
            tcgen05_fence_after();
            uint32_t r[KB / 32][32];   // the whole 128-score row of this block: one pass over TMEM
#pragma unroll
            for (int c = 0; c < KB / 32; ++c) tmem_ld_32x32(ts + c * 32, r[c]);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(&s_free[t]);   // S_t is in registers: the issuer may already compute the next block's S_t
            // 8 independent max chains (a single 128-deep fmaxf chain is 128 x the ALU latency on the critical path)
            float mxs8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) mxs8[i] = -INFINITY;
#pragma unroll
            for (int c = 0; c < KB / 32; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) mxs8[i & 7] = fmaxf(mxs8[i & 7], __uint_as_float(r[c][i]));
            const float mx = fmaxf(fmaxf(fmaxf(mxs8[0], mxs8[1]), fmaxf(mxs8[2], mxs8[3])),
                                   fmaxf(fmaxf(mxs8[4], mxs8[5]), fmaxf(mxs8[6], mxs8[7])));
            float alpha = 1.f;
            if (mx > m + lazy_margin || j == 0) {
                alpha = exp2f((m - mx) * scale_log2e);   // j == 0: m = -inf -> 0 (O is not read then)
                m = mx;
            }
            const float ms = m * scale_log2e;
            if (j > 0) {   // P buffer free and O_t stable once PV_t(j-1) completed
                mbar_wait(&pv_done[t], (j - 1) & 1);
                tcgen05_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {   // rescale the running output in place (warp-collective)
#pragma unroll 1
                    for (int c = 0; c < HEAD_DIM / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32(to + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32(to + c * 32, o);
                    }
                    tmem_st_wait();
                }
            }
            float sum4[4] = {0.f, 0.f, 0.f, 0.f};   // independent partial sums: short dependency chains
#pragma unroll
            for (int c = 0; c < KB / 32; ++c) {
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p0 = exp2f(__uint_as_float(r[c][2 * i]) * scale_log2e - ms);
                    const float p1 = exp2f(__uint_as_float(r[c][2 * i + 1]) * scale_log2e - ms);
                    const uint32_t pk = pack_bf16x2(p0, p1);
                    const float2 back = unpack_bf16x2(pk);   // sum what the tensor core will actually see
                    sum4[i & 3] += back.x + back.y;
                    packed[i] = pk;
                }
                // keys [c*32, c*32+32) of the block live in P tile kb = c/2 (64 keys), 16B chunks (c%2)*4 .. +3 of the row
                uint8_t* tile_row = pbase + (c >> 1) * (Q_TILE * 128) + row * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = (c & 1) * 4 + q;
                    int4 v;
                    v.x = packed[4 * q + 0]; v.y = packed[4 * q + 1]; v.z = packed[4 * q + 2]; v.w = packed[4 * q + 3];
                    *reinterpret_cast<int4*>(tile_row + ((chunk ^ (row & 7)) << 4)) = v;
                }
            }
            l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
            tcgen05_fence_before();
            fence_proxy_async_smem();       // generic-proxy smem writes (P) -> visible to the tensor core
            mbar_arrive(&p_full[t]);
        }
        mbar_wait(&pv_done[t], (NUM_KB - 1) & 1);
        tcgen05_fence_after();
        const float inv = 1.f / l;
        const long long token = seq_row0 + (2 * qpair + t) * Q_TILE + row;
        if (lse2) lse2[token * num_heads + head] = m * scale_log2e + log2f(l);   // log2 sum_j exp2(s_j * scale_log2e)
        bf16* op = out + token * d_model + head * HEAD_DIM;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(to + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int4 v;
                v.x = pack_bf16x2(__uint_as_float(r[8 * i + 0]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
                v.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
                v.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
                v.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
                *reinterpret_cast<int4*>(op + c * 32 + 8 * i) = v;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace v2

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace attn
}  // namespace lah

using namespace lah;
using namespace lah::attn;

extern "C" {

// qkv: [tokens, 3*d_model] bf16, tokens = batch * 512; out: [tokens, d_model] bf16; lse2: [tokens, heads] fp32 or NULL
int lah_attention_fwd(const void* qkv, void* out, float* lse2, int batch, int num_heads, int d_model, cudaStream_t st) {
    if (d_model != num_heads * HEAD_DIM) return -2;
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr)
            return -100;
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)3 * d_model, (cuuint64_t)batch * S_LEN};
    cuuint64_t strides[1] = {(cuuint64_t)3 * d_model * 2};
    cuuint32_t box[2] = {HEAD_DIM, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -1000 - (int)r;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(v2::attention_fwd_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM_TOTAL2);
        if (e != cudaSuccess) return -(int)e;
        configured = true;
    }
    if (batch <= 0) return 0;
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)HEAD_DIM);
    v2::attention_fwd_v2_kernel<<<batch * num_heads * 2, v2::NUM_THREADS2, v2::SMEM_TOTAL2, st>>>(
        tm, (bf16*)out, lse2, d_model, num_heads, scale_log2e);
    return -(int)cudaGetLastError();
}

}  // extern "C"
