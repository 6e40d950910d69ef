// attention_bwd.cu — tcgen05 self-attention BACKWARD for the transformer expert (d_model 1024, 16 heads x 64, seq 512).
// The reference's transformer expert cannot be trained at all (in-place transpose of a leaf); this kernel is
// what makes the sm_100a transformer expert trainable without falling back to eager PyTorch.
//
// One CTA = one (batch, head, 128-key block j).  K_j / V_j stay in shared memory; the four 128-query blocks i stream through
// a 2-stage TMA pipeline (Q_i and dO_i).  Per query block, all on tensor cores with accumulators in TMEM:
//
//     S   = Q_i K_j^T                 (128 x 128)        dP  = dO_i V_j^T              (128 x 128)
//     P   = exp2(S * scale*log2e - LSE_i)   [recomputed from the forward's row log-sum-exp: no second softmax pass]
//     dS  = P o (dP - Delta_i) * scale      [Delta = rowsum(dO o O)]
//     dV_j += P^T dO_i   (128 x 64)    dK_j += dS^T Q_i  (128 x 64)    dQ_i^(j) = dS K_j  (128 x 64, one partial per key block j)
//
// P and dS are written ONCE to shared memory (bf16, 128B-swizzled [query rows][64 keys] atoms) and consumed three ways by
// tcgen05.mma without any transpose: as an MN-major A operand (P^T, dS^T) and as a K-major A operand (dS); Q_i, dO_i, K_j are
// consumed both K-major (S, dP) and MN-major (dV, dK, dQ) straight from their TMA tiles.  Nothing of size S x S touches HBM.
//
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2-5 = one thread per query row (tcgen05.ld of the S and
// dP rows, exp2, dS, smem stores, dQ reduction, final dK / dV epilogue).
#include "sm100.cuh"

namespace lah {
namespace attnb {

constexpr int S_LEN = 512;
constexpr int HEAD_DIM = 64;
constexpr int BLK = 128;                       // query block == key block
constexpr int NUM_QB = S_LEN / BLK;            // 4
constexpr int NUM_THREADS = 192;
constexpr int TILE = BLK * HEAD_DIM * 2;       // 16 KB: 128 x 64 bf16
constexpr int PS_BYTES = BLK * BLK * 2;        // 32 KB: P / dS (two [128 rows][64 keys] atoms)
constexpr int ATOM = BLK * 128;                // bytes between the two 64-wide atoms of P / dS
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + TILE;
constexpr int OFF_Q = OFF_V + TILE;            // 2 stages
constexpr int OFF_DO = OFF_Q + 2 * TILE;       // 2 stages
constexpr int OFF_P = OFF_DO + 2 * TILE;
constexpr int OFF_DS = OFF_P + PS_BYTES;
constexpr int OFF_BAR = OFF_DS + PS_BYTES;
constexpr int NUM_BARS = 1 + 2 + 2 + 1 + 1 + 1;
constexpr int SMEM_TOTAL = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
constexpr int COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                     const float* __restrict__ lse2, const float* __restrict__ delta, bf16* __restrict__ dqkv,
                     bf16* __restrict__ dq_part, long long total_tokens, int d_model, int num_heads, float scale,
                     float scale_log2e) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* kv_full = bars;
    uint64_t* q_full = bars + 1;     // [2]
    uint64_t* q_empty = bars + 3;    // [2]
    uint64_t* s_full = bars + 5;     // MMA -> row threads: S and dP of block i are in TMEM
    uint64_t* p_full = bars + 6;     // row threads (128) -> MMA: P and dS of block i are in shared memory
    uint64_t* dq_full = bars + 7;    // MMA -> row threads: dQ_i complete (and P / dS / Q_i / dO_i no longer read)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x & 3;
    const int head = (blockIdx.x >> 2) % num_heads;
    const int batch = (blockIdx.x >> 2) / num_heads;
    const int seq0 = batch * S_LEN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_do);
        mbar_init(kv_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_full, 128);
        mbar_init(dq_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            mbar_arrive_expect_tx(kv_full, 2 * TILE);
            tma_load_2d(smem + OFF_K, &tm_qkv, kv_full, d_model + head * HEAD_DIM, seq0 + j * BLK);
            tma_load_2d(smem + OFF_V, &tm_qkv, kv_full, 2 * d_model + head * HEAD_DIM, seq0 + j * BLK);
            for (int i = 0; i < NUM_QB; ++i) {
                const int st = i & 1;
                mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&q_full[st], 2 * TILE);
                tma_load_2d(smem + OFF_Q + st * TILE, &tm_qkv, &q_full[st], head * HEAD_DIM, seq0 + i * BLK);
                tma_load_2d(smem + OFF_DO + st * TILE, &tm_do, &q_full[st], head * HEAD_DIM, seq0 + i * BLK);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc_s = make_idesc_bf16_f32(BLK, BLK, 0u, 0u);          // S, dP: both operands K-major
            constexpr uint32_t idesc_t = make_idesc_bf16_f32(BLK, HEAD_DIM, 1u, 1u);     // dV, dK: A^T (MN-major) x MN-major B
            constexpr uint32_t idesc_q = make_idesc_bf16_f32(BLK, HEAD_DIM, 0u, 1u);     // dQ: K-major A x MN-major B
            const uint32_t sk = smem_u32(smem + OFF_K), sv = smem_u32(smem + OFF_V), sp = smem_u32(smem + OFF_P),
                           sds = smem_u32(smem + OFF_DS);
            mbar_wait(kv_full, 0);
            for (int i = 0; i < NUM_QB; ++i) {
                const int st = i & 1;
                const uint32_t sq = smem_u32(smem + OFF_Q + st * TILE), sdo = smem_u32(smem + OFF_DO + st * TILE);
                mbar_wait(&q_full[st], (i >> 1) & 1);
                tcgen05_fence_after();
#pragma unroll
                for (int ks = 0; ks < HEAD_DIM / 16; ++ks)
                    umma_bf16_ss(tmem_base + COL_S, make_smem_desc_sw128(sq + ks * 32, 0, 1024),
                                 make_smem_desc_sw128(sk + ks * 32, 0, 1024), idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < HEAD_DIM / 16; ++ks)
                    umma_bf16_ss(tmem_base + COL_DP, make_smem_desc_sw128(sdo + ks * 32, 0, 1024),
                                 make_smem_desc_sw128(sv + ks * 32, 0, 1024), idesc_s, ks > 0 ? 1u : 0u);
                umma_commit(s_full);
                mbar_wait(p_full, i & 1);
                tcgen05_fence_after();
#pragma unroll
                for (int ks = 0; ks < BLK / 16; ++ks) {   // reduction over the 128 queries of the block, 16 rows per step
                    umma_bf16_ss(tmem_base + COL_DV, make_smem_desc_sw128(sp + ks * 2048, ATOM, 1024),
                                 make_smem_desc_sw128(sdo + ks * 2048, ATOM, 1024), idesc_t, (i | ks) ? 1u : 0u);
                }
#pragma unroll
                for (int ks = 0; ks < BLK / 16; ++ks) {
                    umma_bf16_ss(tmem_base + COL_DK, make_smem_desc_sw128(sds + ks * 2048, ATOM, 1024),
                                 make_smem_desc_sw128(sq + ks * 2048, ATOM, 1024), idesc_t, (i | ks) ? 1u : 0u);
                }
#pragma unroll
                for (int ks = 0; ks < BLK / 16; ++ks) {   // reduction over the 128 keys: atom ks / 4, 32 B per step inside the row
                    umma_bf16_ss(tmem_base + COL_DQ, make_smem_desc_sw128(sds + (ks >> 2) * ATOM + (ks & 3) * 32, 0, 1024),
                                 make_smem_desc_sw128(sk + ks * 2048, ATOM, 1024), idesc_q, ks ? 1u : 0u);
                }
                umma_commit(&q_empty[st]);
                umma_commit(dq_full);
            }
        }
    } else {
        // ---------------------------------------------------------------------- one thread per query row (TMEM lane)
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        uint8_t* p_row = smem + OFF_P + row * 128;
        uint8_t* ds_row = smem + OFF_DS + row * 128;
#pragma unroll 1
        for (int i = 0; i < NUM_QB; ++i) {
            const long long token = seq0 + i * BLK + row;
            const float lse = __ldg(lse2 + token * num_heads + head);
            const float dl = __ldg(delta + token * num_heads + head);
            mbar_wait(s_full, i & 1);
            tcgen05_fence_after();
#pragma unroll 1
            for (int c = 0; c < BLK / 32; ++c) {
                uint32_t rs[32], rd[32];
                tmem_ld_32x32(lane_base + COL_S + c * 32, rs);
                tmem_ld_32x32(lane_base + COL_DP + c * 32, rd);
                tmem_ld_wait();
                uint32_t pp[16], dd[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float p0 = exp2f(__uint_as_float(rs[2 * e]) * scale_log2e - lse);
                    const float p1 = exp2f(__uint_as_float(rs[2 * e + 1]) * scale_log2e - lse);
                    pp[e] = pack_bf16x2(p0, p1);
                    dd[e] = pack_bf16x2(p0 * (__uint_as_float(rd[2 * e]) - dl) * scale,
                                        p1 * (__uint_as_float(rd[2 * e + 1]) - dl) * scale);
                }
                // keys [c*32, c*32+32) live in atom c/2, 16 B chunks (c%2)*4 .. +3 of the 128 B row (swizzled with row & 7)
                const int atom_off = (c >> 1) * ATOM;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = (((c & 1) * 4 + q) ^ (row & 7)) << 4;
                    *reinterpret_cast<int4*>(p_row + atom_off + chunk) = make_int4(pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]);
                    *reinterpret_cast<int4*>(ds_row + atom_off + chunk) = make_int4(dd[4 * q], dd[4 * q + 1], dd[4 * q + 2], dd[4 * q + 3]);
                }
            }
            tcgen05_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(p_full);
            mbar_wait(dq_full, i & 1);
            tcgen05_fence_after();
            // partial dQ of THIS key block: plain 16 B stores of the thread's 128 contiguous bytes into slice j of dq_part
            // ([4, T, D] bf16; attn_dq_reduce_kernel sums the four slices in fp32) — no atomics, half the bytes of fp32 partials
            int4* dq = reinterpret_cast<int4*>(dq_part + (static_cast<long long>(j) * total_tokens + token) * d_model + head * HEAD_DIM);
#pragma unroll 1
            for (int c = 0; c < HEAD_DIM / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(lane_base + COL_DQ + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    dq[c * 4 + e] = make_int4(pack_bf16x2(__uint_as_float(r[8 * e]), __uint_as_float(r[8 * e + 1])),
                                              pack_bf16x2(__uint_as_float(r[8 * e + 2]), __uint_as_float(r[8 * e + 3])),
                                              pack_bf16x2(__uint_as_float(r[8 * e + 4]), __uint_as_float(r[8 * e + 5])),
                                              pack_bf16x2(__uint_as_float(r[8 * e + 6]), __uint_as_float(r[8 * e + 7])));
            }
            tcgen05_fence_before();
        }
        // dK_j / dV_j: complete after the last dq_full (the commit covers every MMA issued before it)
        const long long key = seq0 + j * BLK + row;
        bf16* dk = dqkv + key * (3ll * d_model) + d_model + head * HEAD_DIM;
        bf16* dv = dqkv + key * (3ll * d_model) + 2 * d_model + head * HEAD_DIM;
#pragma unroll 1
        for (int c = 0; c < HEAD_DIM / 32; ++c) {
            uint32_t rk[32], rv[32];
            tmem_ld_32x32(lane_base + COL_DK + c * 32, rk);
            tmem_ld_32x32(lane_base + COL_DV + c * 32, rv);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int4 a, b;
                a.x = pack_bf16x2(__uint_as_float(rk[8 * q + 0]), __uint_as_float(rk[8 * q + 1]));
                a.y = pack_bf16x2(__uint_as_float(rk[8 * q + 2]), __uint_as_float(rk[8 * q + 3]));
                a.z = pack_bf16x2(__uint_as_float(rk[8 * q + 4]), __uint_as_float(rk[8 * q + 5]));
                a.w = pack_bf16x2(__uint_as_float(rk[8 * q + 6]), __uint_as_float(rk[8 * q + 7]));
                b.x = pack_bf16x2(__uint_as_float(rv[8 * q + 0]), __uint_as_float(rv[8 * q + 1]));
                b.y = pack_bf16x2(__uint_as_float(rv[8 * q + 2]), __uint_as_float(rv[8 * q + 3]));
                b.z = pack_bf16x2(__uint_as_float(rv[8 * q + 4]), __uint_as_float(rv[8 * q + 5]));
                b.w = pack_bf16x2(__uint_as_float(rv[8 * q + 6]), __uint_as_float(rv[8 * q + 7]));
                *reinterpret_cast<int4*>(dk + c * 32 + 8 * q) = a;
                *reinterpret_cast<int4*>(dv + c * 32 + 8 * q) = b;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// prologue: delta[t, h] = sum_d dO[t, h, d] * O[t, h, d]  (one warp per (token, head): 64 bf16 = one 128 B row segment each)
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ out,
                                                         float* __restrict__ delta, long long pairs) {
    const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= pairs) return;
    const __nv_bfloat162 a = reinterpret_cast<const __nv_bfloat162*>(dout + w * HEAD_DIM)[lane];
    const __nv_bfloat162 b = reinterpret_cast<const __nv_bfloat162*>(out + w * HEAD_DIM)[lane];
    float acc = __bfloat162float(a.x) * __bfloat162float(b.x) + __bfloat162float(a.y) * __bfloat162float(b.y);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) delta[w] = acc;
}

// epilogue: dQ = sum of the per-key-block partials, written as bf16 into the Q third of dqkv
__global__ void __launch_bounds__(256) attn_dq_reduce_kernel(const bf16* __restrict__ dq_part, bf16* __restrict__ dqkv,
                                                             long long tokens, int d_model, int parts) {
    const long long i8 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // 8-element (16 B) index inside [T, D]
    const long long n8 = tokens * d_model / 8;
    if (i8 >= n8) return;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int p = 0; p < parts; ++p) {
        const int4 t = reinterpret_cast<const int4*>(dq_part)[i8 + p * n8];
        const uint32_t w[4] = {(uint32_t)t.x, (uint32_t)t.y, (uint32_t)t.z, (uint32_t)t.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16x2(w[e]);
            acc[2 * e] += f.x;
            acc[2 * e + 1] += f.y;
        }
    }
    const long long el = i8 * 8, tok = el / d_model, col = el - tok * d_model;
    *reinterpret_cast<int4*>(dqkv + tok * 3 * d_model + col) =
        make_int4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
}

}  // namespace attnb
}  // namespace lah

using namespace lah;
using namespace lah::attnb;

extern "C" {

// qkv [T, 3D] bf16 (forward input), out [T, D] bf16 (forward output), dout [T, D] bf16, lse2 [T, H] fp32 (forward output)
// -> dqkv [T, 3D] bf16.  Scratch: delta [T, H] fp32 (rowsum(dout o out), computed here), dq_part [4, T, D] bf16 (the four
// per-key-block partials of dQ, reduced into the Q third of dqkv here).  Three launches, no PyTorch ops around them.
int lah_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, float* delta, void* dqkv,
                      void* dq_part, int batch, int num_heads, int d_model, cudaStream_t st) {
    if (d_model != num_heads * HEAD_DIM) return -2;
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr)
            return -100;
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    CUtensorMap tm_qkv, tm_do;
    cuuint32_t box[2] = {HEAD_DIM, BLK};
    cuuint32_t estr[2] = {1, 1};
    {
        cuuint64_t dims[2] = {(cuuint64_t)3 * d_model, (cuuint64_t)batch * S_LEN};
        cuuint64_t strides[1] = {(cuuint64_t)3 * d_model * 2};
        CUresult r = fn(&tm_qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return -1000 - (int)r;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d_model, (cuuint64_t)batch * S_LEN};
        cuuint64_t strides[1] = {(cuuint64_t)d_model * 2};
        CUresult r = fn(&tm_do, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dout), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return -1000 - (int)r;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
        if (e != cudaSuccess) return -(int)e;
        configured = true;
    }
    if (batch <= 0) return 0;
    const float scale = 1.f / sqrtf((float)HEAD_DIM);
    const long long tokens = (long long)batch * S_LEN, pairs = tokens * num_heads;
    attn_delta_kernel<<<(unsigned)((pairs * 32 + 255) / 256), 256, 0, st>>>((const bf16*)dout, (const bf16*)out, delta, pairs);
    attention_bwd_kernel<<<batch * num_heads * (S_LEN / BLK), NUM_THREADS, SMEM_TOTAL, st>>>(
        tm_qkv, tm_do, lse2, delta, (bf16*)dqkv, (bf16*)dq_part, tokens, d_model, num_heads, scale,
        scale * 1.4426950408889634f);
    attn_dq_reduce_kernel<<<(unsigned)((tokens * d_model / 8 + 255) / 256), 256, 0, st>>>((const bf16*)dq_part, (bf16*)dqkv, tokens, d_model,
                                                                                         S_LEN / BLK);
    return -(int)cudaGetLastError();
}

}  // extern "C"
