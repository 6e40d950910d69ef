// attention.cu — tcgen05 self-attention forward for the transformer expert (d_model 1024, 16 heads x 64, seq 512;
// reference: nn.MultiheadAttention inside /root/reference/experiments/throughput/layers.py:22-51).
//
// One CTA = one (batch, head, 128-query tile).  Everything between the QKV projection and the output projection stays on
// chip:  TMA (Q tile, all 512 keys, all 512 values of the head) -> smem;  S = Q K^T with tcgen05.mma into TMEM
// (128 lanes x 512 fp32 columns = the whole TMEM);  exact softmax straight out of TMEM (one query row per thread:
// tcgen05.ld, row max, exp2, row sum);  P is written back to shared memory as bf16 in the 128B-swizzled K-major layout the
// tensor core expects;  O = P V with tcgen05.mma (V consumed as an MN-major operand, no transpose), normalised by the
// row sums in the epilogue.  S and P never touch HBM (the unfused formulation moves 2 x 512 x 512 x 2 B per head).
//
// Input : qkv [T = batch*512, 3*D] bf16 (output of the fused in_proj GEMM: [q | k | v] per token, heads contiguous)
// Output: out [T, D] bf16 (heads concatenated, ready for out_proj)
#include "sm100.cuh"
#include <stdlib.h>

namespace lah {
namespace attn {

constexpr int S_LEN = 512;      // keys per sequence
constexpr int HEAD_DIM = 64;
constexpr int Q_TILE = 128;
constexpr int NUM_THREADS = 160;   // 4 softmax/epilogue warps + 1 control warp (TMA + MMA issue)

constexpr int Q_BYTES = Q_TILE * HEAD_DIM * 2;      // 16 KB
constexpr int KV_BYTES = S_LEN * HEAD_DIM * 2;      // 64 KB each
constexpr int P_HALF_BYTES = Q_TILE * 256 * 2;      // 64 KB: probabilities of 256 keys
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;              // reused for P(keys 0..255) once S has been computed
constexpr int OFF_V = OFF_K + KV_BYTES;
constexpr int OFF_P1 = OFF_V + KV_BYTES;            // P(keys 256..511)
constexpr int OFF_BAR = OFF_P1 + P_HALF_BYTES;
constexpr int SMEM_TOTAL = OFF_BAR + 8 * 8 + 16 + 1024;

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ out, int d_model, int num_heads,
                     float scale_log2e) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* bar_s = bar_load + 1;      // S = QK^T complete
    uint64_t* bar_p0 = bar_load + 2;     // P(0..255) written by all 128 softmax threads
    uint64_t* bar_p1 = bar_load + 3;
    uint64_t* bar_o = bar_load + 4;      // O complete
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_load + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x & 3;
    const int head = (blockIdx.x >> 2) % num_heads;
    const int batch = (blockIdx.x >> 2) / num_heads;
    const int seq_row0 = batch * S_LEN;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p0, 128);
        mbar_init(bar_p1, 128);
        mbar_init(bar_o, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 4) {
        if (lane == 0) {
            // ---------------------------------------------------------------- loads
            mbar_arrive_expect_tx(bar_load, Q_BYTES + 2 * KV_BYTES);
            tma_load_2d(smem + OFF_Q, &tm_qkv, bar_load, head * HEAD_DIM, seq_row0 + qt * Q_TILE);
#pragma unroll
            for (int i = 0; i < S_LEN / 128; ++i) {
                tma_load_2d(smem + OFF_K + i * 128 * 128, &tm_qkv, bar_load, d_model + head * HEAD_DIM, seq_row0 + i * 128);
                tma_load_2d(smem + OFF_V + i * 128 * 128, &tm_qkv, bar_load, 2 * d_model + head * HEAD_DIM, seq_row0 + i * 128);
            }
            mbar_wait(bar_load, 0);
            tcgen05_fence_after();
            // ---------------------------------------------------------------- S = Q K^T  (two N = 256 halves)
            const uint32_t sq = smem_u32(smem + OFF_Q), sk = smem_u32(smem + OFF_K);
            constexpr uint32_t idesc_s = make_idesc_bf16_f32(Q_TILE, 256, 0u, 0u);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int ks = 0; ks < HEAD_DIM / 16; ++ks) {
                    const uint64_t da = make_smem_desc_sw128(sq + ks * 32, 0, 1024);
                    const uint64_t db = make_smem_desc_sw128(sk + half * 256 * 128 + ks * 32, 0, 1024);
                    umma_bf16_ss(tmem_base + half * 256, da, db, idesc_s, ks > 0 ? 1u : 0u);
                }
            }
            umma_commit(bar_s);
            // ---------------------------------------------------------------- O = P V  (V is an MN-major operand)
            const uint32_t sv = smem_u32(smem + OFF_V);
            constexpr uint32_t idesc_o = make_idesc_bf16_f32(Q_TILE, HEAD_DIM, 0u, 1u);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                mbar_wait(half == 0 ? bar_p0 : bar_p1, 0);
                tcgen05_fence_after();
                const uint32_t sp = smem_u32(smem + (half == 0 ? OFF_K : OFF_P1));
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {        // 64 keys per P tile
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {    // 16 keys per MMA
                        const uint64_t da = make_smem_desc_sw128(sp + kb * (Q_TILE * 128) + ks * 32, 0, 1024);
                        const uint64_t db = make_smem_desc_sw128(sv + (half * 256 + kb * 64 + ks * 16) * 128, 0, 1024);
                        umma_bf16_ss(tmem_base, da, db, idesc_o, (half | kb | ks) ? 1u : 0u);
                    }
                }
            }
            umma_commit(bar_o);
        }
    } else {
        // ==================================================================== softmax + epilogue: thread = query row
        const int row = warp * 32 + lane;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        mbar_wait(bar_s, 0);
        tcgen05_fence_after();
        // pass A: row maximum over the 512 scores
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < S_LEN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        const float mx_scaled = mx * scale_log2e;
        // pass B: p = exp2(s*scale*log2e - max), row sum, bf16 P tiles in the swizzled K-major layout
        float sum = 0.f;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint8_t* pbase = smem + (half == 0 ? OFF_K : OFF_P1);
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {   // 32 keys per chunk, 8 chunks per half
                uint32_t r[32];
                tmem_ld_32x32(taddr + half * 256 + c * 32, r);
                tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float p0 = exp2f(__uint_as_float(r[2 * j]) * scale_log2e - mx_scaled);
                    const float p1 = exp2f(__uint_as_float(r[2 * j + 1]) * scale_log2e - mx_scaled);
                    // sum what the tensor core will actually see (bf16-rounded probabilities)
                    const uint32_t pk = pack_bf16x2(p0, p1);
                    const float2 back = unpack_bf16x2(pk);
                    sum += back.x + back.y;
                    packed[j] = pk;
                }
                // keys [c*32, c*32+32) of this half live in P tile kb = c/2, 16B-chunks (c%2)*4 .. +3 of row `row`
                uint8_t* tile_row = pbase + (c >> 1) * (Q_TILE * 128) + row * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = (c & 1) * 4 + q;
                    int4 v;
                    v.x = packed[4 * q + 0]; v.y = packed[4 * q + 1]; v.z = packed[4 * q + 2]; v.w = packed[4 * q + 3];
                    *reinterpret_cast<int4*>(tile_row + ((chunk ^ (row & 7)) << 4)) = v;
                }
            }
            fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tcgen05_fence_before();
            mbar_arrive(half == 0 ? bar_p0 : bar_p1);
        }
        // epilogue: O / sum -> bf16 -> out[token, head*64 ...]
        mbar_wait(bar_o, 0);
        tcgen05_fence_after();
        const float inv = 1.f / sum;
        bf16* op = out + static_cast<long long>(seq_row0 + qt * Q_TILE + row) * d_model + head * HEAD_DIM;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int4 v;
                v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
                v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
                v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
                v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
                *reinterpret_cast<int4*>(op + c * 32 + 8 * j) = v;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ================================================================================================================
// v2: ping-pong flash attention.  One CTA = one (batch, head, PAIR of 128-query tiles); 10 warps:
//   warps 0-3 softmax of tile A, warps 4-7 softmax of tile B (thread = query row), warp 8 = MMA issuer, warp 9 = TMA.
// K / V stream through a 2-stage TMA pipeline in blocks of 128 keys and are shared by both tiles (half the L2 traffic
// of v1); while the softmax warps of one tile turn S_j into P_j, the tensor core computes S / PV of the other tile, so
// MUFU (the real bottleneck: 512 exp2 per query row) and tcgen05 overlap.  Online softmax over the 4 key blocks:
//   m' = max(m, rowmax(S_j)),  alpha = 2^((m - m') * c),  l = l * alpha + sum_k 2^((s_k - m') * c),  O = O * alpha + P_j V_j
// with O living in TMEM (rescaled in place by the softmax threads: tcgen05.ld -> mul -> tcgen05.st).
// TMEM: S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384).
// ================================================================================================================
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
constexpr int OFF_XCHG3 = OFF_BAR2 + NUM_BARS * 8 + 16;               // v3: max / sum exchange between row halves
constexpr int SMEM_TOTAL2 = OFF_XCHG3 + 2 * 2 * 2 * Q_TILE * 4 + 1024;
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
attention_fwd_v2_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ out, int d_model, int num_heads,
                        float scale_log2e) {
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
        bf16* op = out + static_cast<long long>(seq_row0 + (2 * qpair + t) * Q_TILE + row) * d_model + head * HEAD_DIM;
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

// v3: same pipeline, but TWO threads per query row (each owns 64 of the 128 scores of a block): 16 softmax warps instead
// of 8, i.e. four warps per SM sub-partition to hide the tcgen05.ld / MUFU / barrier latencies that dominate v2 (ncu: XU
// pipe 34 %, tensor 17 %).  The two halves of a row agree on the block maximum through shared memory + a named barrier
// per tile group; each keeps its own partial row sum (same reference maximum), summed once in the epilogue.
constexpr int NUM_THREADS3 = 576;   // 16 softmax warps + MMA warp + TMA warp
__global__ void __launch_bounds__(NUM_THREADS3, 1)
attention_fwd_v3_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ out, int d_model, int num_heads,
                        float scale_log2e) {
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
    float* xchg = reinterpret_cast<float*>(smem + OFF_XCHG3);   // [2 buffers][2 tiles][2 halves][128 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qpair = blockIdx.x & 1;
    const int head = (blockIdx.x >> 1) % num_heads;
    const int batch = (blockIdx.x >> 1) / num_heads;
    const int seq_row0 = batch * S_LEN;

    if (warp == 17 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 256);
            mbar_init(&p_full[i], 256);
            mbar_init(&pv_done[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 16) tmem_alloc(tmem_ptr, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 17) {
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
    } else if (warp == 16) {
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
        // ---------------------------------------------------------------------- softmax + epilogue: 2 threads per query row
        const int t = warp >> 3;                 // query tile of the pair
        const int hc = (warp >> 2) & 1;          // which 64 of the 128 scores of a block (and which 32 output columns)
        const int row = (warp & 3) * 32 + lane;  // row inside the tile == TMEM lane
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t ts = lane_base + COL_S + t * KB + hc * 64, to = lane_base + COL_O + t * HEAD_DIM + hc * 32;
        uint8_t* pbase = smem + OFF_P2 + t * P_BYTES + hc * (Q_TILE * 128);   // P tile of keys [64 hc, 64 hc + 64)
        float m = -INFINITY, l = 0.f;
        const float lazy_margin = 8.f / scale_log2e;
#pragma unroll 1
        for (int j = 0; j < NUM_KB; ++j) {
            mbar_wait(&s_full[t], j & 1);
            tcgen05_fence_after();
            uint32_t r[2][32];
            tmem_ld_32x32(ts, r[0]);
            tmem_ld_32x32(ts + 32, r[1]);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(&s_free[t]);
            float mxs8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) mxs8[i] = -INFINITY;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) mxs8[i & 7] = fmaxf(mxs8[i & 7], __uint_as_float(r[c][i]));
            float mx = fmaxf(fmaxf(fmaxf(mxs8[0], mxs8[1]), fmaxf(mxs8[2], mxs8[3])),
                             fmaxf(fmaxf(mxs8[4], mxs8[5]), fmaxf(mxs8[6], mxs8[7])));
            // agree on the block maximum of the row with the thread that owns the other 64 scores
            float* xb = xchg + (((j & 1) * 2 + t) * 2) * Q_TILE;
            xb[hc * Q_TILE + row] = mx;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "r"(256) : "memory");
            mx = fmaxf(mx, xb[(hc ^ 1) * Q_TILE + row]);
            float alpha = 1.f;
            if (mx > m + lazy_margin || j == 0) {
                alpha = exp2f((m - mx) * scale_log2e);
                m = mx;
            }
            const float ms = m * scale_log2e;
            if (j > 0) {
                mbar_wait(&pv_done[t], (j - 1) & 1);
                tcgen05_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {   // my 32 columns of the running output
                    uint32_t o[32];
                    tmem_ld_32x32(to, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_32x32(to, o);
                    tmem_st_wait();
                }
            }
            float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p0 = exp2f(__uint_as_float(r[c][2 * i]) * scale_log2e - ms);
                    const float p1 = exp2f(__uint_as_float(r[c][2 * i + 1]) * scale_log2e - ms);
                    const uint32_t pk = pack_bf16x2(p0, p1);
                    const float2 back = unpack_bf16x2(pk);
                    sum4[i & 3] += back.x + back.y;
                    packed[i] = pk;
                }
                uint8_t* tile_row = pbase + row * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = c * 4 + q;
                    int4 v;
                    v.x = packed[4 * q + 0]; v.y = packed[4 * q + 1]; v.z = packed[4 * q + 2]; v.w = packed[4 * q + 3];
                    *reinterpret_cast<int4*>(tile_row + ((chunk ^ (row & 7)) << 4)) = v;
                }
            }
            l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
            tcgen05_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(&p_full[t]);
        }
        // total row sum = the two partial sums (same reference maximum m)
        float* xb = xchg + ((0 * 2 + t) * 2) * Q_TILE;   // buffer 0 was last used by block NUM_KB - 2: free again
        asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "r"(256) : "memory");
        xb[hc * Q_TILE + row] = l;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "r"(256) : "memory");
        l += xb[(hc ^ 1) * Q_TILE + row];
        mbar_wait(&pv_done[t], (NUM_KB - 1) & 1);
        tcgen05_fence_after();
        const float inv = 1.f / l;
        bf16* op = out + static_cast<long long>(seq_row0 + (2 * qpair + t) * Q_TILE + row) * d_model + head * HEAD_DIM + hc * 32;
        {
            uint32_t r[32];
            tmem_ld_32x32(to, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int4 v;
                v.x = pack_bf16x2(__uint_as_float(r[8 * i + 0]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
                v.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
                v.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
                v.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
                *reinterpret_cast<int4*>(op + 8 * i) = v;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 16) {
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

// qkv: [tokens, 3*d_model] bf16, tokens = batch * 512; out: [tokens, d_model] bf16
int lah_attention_fwd(const void* qkv, void* out, int batch, int num_heads, int d_model, cudaStream_t st) {
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
    static int use_v1 = 0, use_version = 2;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
        if (e != cudaSuccess) return -(int)e;
        e = cudaFuncSetAttribute(v2::attention_fwd_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM_TOTAL2);
        if (e != cudaSuccess) return -(int)e;
        e = cudaFuncSetAttribute(v2::attention_fwd_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM_TOTAL2);
        if (e != cudaSuccess) return -(int)e;
        const char* env = getenv("LAH_ATTN_V1");   // A/B switches: LAH_ATTN_V1=1 one tile per CTA; LAH_ATTN=3 two threads / row
        use_v1 = env && atoi(env) == 1;
        env = getenv("LAH_ATTN");
        if (env) use_version = atoi(env);
        configured = true;
    }
    if (batch <= 0) return 0;
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)HEAD_DIM);
    if (use_v1)
        attention_fwd_kernel<<<batch * num_heads * (S_LEN / Q_TILE), NUM_THREADS, SMEM_TOTAL, st>>>(tm, (bf16*)out, d_model,
                                                                                              num_heads, scale_log2e);
    else if (use_version == 3)
        v2::attention_fwd_v3_kernel<<<batch * num_heads * 2, v2::NUM_THREADS3, v2::SMEM_TOTAL2, st>>>(
            tm, (bf16*)out, d_model, num_heads, scale_log2e);
    else
        v2::attention_fwd_v2_kernel<<<batch * num_heads * 2, v2::NUM_THREADS2, v2::SMEM_TOTAL2, st>>>(
            tm, (bf16*)out, d_model, num_heads, scale_log2e);
    return -(int)cudaGetLastError();
}

}  // extern "C"
