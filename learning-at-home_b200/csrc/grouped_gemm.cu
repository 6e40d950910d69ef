// grouped_gemm.cu — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
// One kernel template covers every GEMM-shaped op of the DMoE expert path
// (reference hot loops: experiments/throughput/layers.py:8-19 forward, lib/runtime/expert_backend.py:73-93 backward):
//
//   MODE_MGROUP  C[rows, N] = A[rows, K] * B[g(rows)]^T (+bias[g]) (+residual)
//                rows are grouped by expert, every group padded to a multiple of 128 rows; the group of a
//                128-row tile comes from a device-side table written by the dispatch kernel (no host sync).
//                B is either K-major  ([G, N, K], forward:  x @ W^T)
//                      or MN-major    ([G, K, N], dgrad:    dy @ W  with W stored [K=out, N=in]).
//   MODE_KGROUP  C[g][M, N] = A_g^T * B_g   (wgrad: dW = dY^T X, reduction over the tokens of expert g)
//                A = dY [tokens, M], B = X [tokens, N], both "MN-major" operands; fp32 output per group.
//
// Structure (one CTA per SM, 192 threads):
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1 (one lane)  MMA issuer:   tcgen05.mma cta_group::1 kind::f16, 128 x BLOCK_N x 16, accumulators in TMEM
//                      (2 accumulator stages => epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2..5         epilogue:     tcgen05.ld 32x32b -> registers -> bias/residual -> bf16|fp32 -> global
#include "sm100.cuh"

namespace lah {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int MODE_MGROUP = 0;
constexpr int MODE_KGROUP = 1;

struct GemmParams {
    int N;               // columns of C
    int K;               // MGROUP: reduction length
    int M;               // KGROUP: rows of C per group (multiple of 128). MGROUP: number of valid rows of C.
    int num_groups;      // KGROUP: number of groups
    int num_m_tiles;     // MGROUP: number of 128-row tiles to visit (upper bound; unused tiles have group -1)
    const int* tile_group;  // MGROUP: [num_m_tiles] group of each m tile, -1 = skip; nullptr => group 0
    const int* group_off;   // KGROUP: [G+1] padded token offsets (multiples of 128)
    void* C;
    long long ldc;
    long long c_group_stride;  // KGROUP: elements between consecutive groups of C
    const float* bias;         // MGROUP: [G, N] or nullptr
    const bf16* residual;      // MGROUP: [rows, ldr] or nullptr
    long long ldr;
    // receive-side fusion: the TMA producer waits until every source rank's dispatch flag reached `wait_epoch`
    const int* wait_flags;     // [wait_count] local flag words written by the peers (st.release.sys), or nullptr
    int wait_count, wait_epoch;
    const int* epoch_base;   // device-side epoch base added to wait_epoch (nullptr: 0), see moe.cu Peers::step_ctr
    int* status;
};

template <int BLOCK_N, int STAGES>
struct SmemLayout {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;  // + alignment slack
};

template <int BLOCK_N, int STAGES, int MODE, bool A_MN, bool B_MN, bool OUT_F32>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    using L = SmemLayout<BLOCK_N, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
    if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // ------------------------------------------------------------------ tile enumeration
    const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
    int total_tiles, tiles_per_group = 0;
    if (MODE == MODE_MGROUP) {
        total_tiles = p.num_m_tiles * n_tiles;
    } else {
        tiles_per_group = (p.M / BLOCK_M) * n_tiles;
        total_tiles = p.num_groups * tiles_per_group;
    }

    // decode one tile; returns false when the tile must be skipped (identical decision in all roles)
    auto decode = [&](int tile, int& m_row, int& n_col, int& group, int& k_begin, int& num_kb) -> bool {
        if (MODE == MODE_MGROUP) {
            const int m_tile = tile / n_tiles;
            const int n_tile = tile - m_tile * n_tiles;
            group = p.tile_group ? __ldg(p.tile_group + m_tile) : 0;
            m_row = m_tile * BLOCK_M;
            n_col = n_tile * BLOCK_N;
            k_begin = 0;
            num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
            return group >= 0;
        } else {
            group = tile / tiles_per_group;
            const int r = tile - group * tiles_per_group;
            const int m_tile = r / n_tiles;
            const int n_tile = r - m_tile * n_tiles;
            m_row = m_tile * BLOCK_M;
            n_col = n_tile * BLOCK_N;
            k_begin = __ldg(p.group_off + group);
            num_kb = (__ldg(p.group_off + group + 1) - k_begin) / BLOCK_K;
            return num_kb > 0;
        }
    };

    if (warp == 0 && lane == 0) {
        // =============================================================== TMA producer
        if (p.wait_flags) {  // rows pushed by peer GPUs over NVLink must have landed before the first TMA load
            const unsigned long long t_wait = globaltimer_ns();
            for (int sidx = 0; sidx < p.wait_count; ++sidx)
                spin_flag_ft(p.wait_flags + sidx, p.wait_epoch + (p.epoch_base ? p.epoch_base[0] : 0), p.status, sidx,
                             p.epoch_base ? p.epoch_base[1] : 0);
            // exposed communication wait (ns) of this rank: status[2..3] is a 64-bit counter (EngineContext.wait_ns)
            if (blockIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.status + 2), globaltimer_ns() - t_wait);
            fence_proxy_async_global();
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int m_row, n_col, g, k_begin, num_kb;
            if (!decode(tile, m_row, n_col, g, k_begin, num_kb)) continue;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * L::STAGE_BYTES;
                uint8_t* sb = sa + L::A_BYTES;
                const int k = k_begin + kb * BLOCK_K;
                mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                // ---- A
                if (!A_MN) {
                    tma_load_2d(sa, &tmA, &full_bar[stage], k, m_row);
                } else {
#pragma unroll
                    for (int i = 0; i < BLOCK_M / 64; ++i)
                        tma_load_2d(sa + i * (BLOCK_K * 128), &tmA, &full_bar[stage], m_row + i * 64, k);
                }
                // ---- B
                if (MODE == MODE_MGROUP) {
                    if (!B_MN) {
                        tma_load_3d(sb, &tmB, &full_bar[stage], k, n_col, g);
                    } else {
#pragma unroll
                        for (int i = 0; i < BLOCK_N / 64; ++i)
                            tma_load_3d(sb + i * (BLOCK_K * 128), &tmB, &full_bar[stage], n_col + i * 64, k, g);
                    }
                } else {
                    if (!B_MN) {
                        tma_load_2d(sb, &tmB, &full_bar[stage], k, n_col);
                    } else {
#pragma unroll
                        for (int i = 0; i < BLOCK_N / 64; ++i)
                            tma_load_2d(sb + i * (BLOCK_K * 128), &tmB, &full_bar[stage], n_col + i * 64, k);
                    }
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // =============================================================== MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16_f32(BLOCK_M, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
        // K-major:  SBO = 8 rows * 128B = 1024, LBO unused; advance 32 B per UMMA_K inside the swizzle row
        // MN-major: SBO = 1024 (8 k-rows), LBO = BLOCK_K * 128 B (next 64-wide MN atom); advance 16 k-rows = 2048 B
        constexpr uint32_t A_LBO = A_MN ? BLOCK_K * 128 : 0, B_LBO = B_MN ? BLOCK_K * 128 : 0;
        constexpr uint32_t A_KSTEP = A_MN ? UMMA_K * 128 : UMMA_K * 2, B_KSTEP = B_MN ? UMMA_K * 128 : UMMA_K * 2;
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int m_row, n_col, g, k_begin, num_kb;
            if (!decode(tile, m_row, n_col, g, k_begin, num_kb)) continue;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + as * BLOCK_N;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint64_t da = make_smem_desc_sw128(sa + k * A_KSTEP, A_LBO, 1024);
                    const uint64_t db = make_smem_desc_sw128(sb + k * B_KSTEP, B_LBO, 1024);
                    umma_bf16_ss(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);  // smem slot is free once these MMAs have read it
                if (kb == num_kb - 1) umma_commit(&tmem_full[as]);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            ++iter;
        }
    } else if (warp >= 2) {
        // =============================================================== epilogue (4 warps, one row per thread)
        const int lane_group = warp & 3;  // TMEM lane quarter this warp may access
        int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int m_row, n_col, g, k_begin, num_kb;
            if (!decode(tile, m_row, n_col, g, k_begin, num_kb)) continue;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tmem_full[as], aphase);
            tcgen05_fence_after();
            const int row = m_row + lane_group * 32 + lane;
            const bool row_ok = (MODE == MODE_KGROUP) ? true : (row < p.M);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                tmem_ld_wait();
                const int col = n_col + c * 32;
                if (col >= p.N || !row_ok) continue;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (MODE == MODE_MGROUP) {
                    if (p.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + static_cast<long long>(g) * p.N + col);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(bp + j);
                            v[4 * j + 0] += b.x;
                            v[4 * j + 1] += b.y;
                            v[4 * j + 2] += b.z;
                            v[4 * j + 3] += b.w;
                        }
                    }
                    if (p.residual) {
                        const int4* rp = reinterpret_cast<const int4*>(p.residual + static_cast<long long>(row) * p.ldr + col);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int4 q = __ldg(rp + j);
                            const uint32_t w[4] = {(uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w};
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float2 f = unpack_bf16x2(w[t]);
                                v[8 * j + 2 * t] += f.x;
                                v[8 * j + 2 * t + 1] += f.y;
                            }
                        }
                    }
                }
                if (OUT_F32) {
                    float* cp = reinterpret_cast<float*>(p.C) +
                                (MODE == MODE_KGROUP ? static_cast<long long>(g) * p.c_group_stride : 0ll) +
                                static_cast<long long>(row) * p.ldc + col;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(cp + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
                    bf16* cp = reinterpret_cast<bf16*>(p.C) +
                               (MODE == MODE_KGROUP ? static_cast<long long>(g) * p.c_group_stride : 0ll) +
                               static_cast<long long>(row) * p.ldc + col;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int4 q;
                        q.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                        q.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                        q.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                        q.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                        *reinterpret_cast<int4*>(cp + 8 * j) = q;
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            ++iter;
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// bf16 tensor map, up to 3 dims (dim0 = contiguous), 128B swizzle, zero OOB fill.
static int make_tmap_bf16(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return -100;
    cuuint64_t gdims[3];
    cuuint64_t gstr[2];
    cuuint32_t gbox[3];
    cuuint32_t estr[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
    }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdims, gstr, gbox, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}

template <int BLOCK_N, int STAGES, int MODE, bool A_MN, bool B_MN, bool OUT_F32>
static int launch(const GemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int max_ctas, cudaStream_t st) {
    using L = SmemLayout<BLOCK_N, STAGES>;
    auto kern = gemm_kernel<BLOCK_N, STAGES, MODE, A_MN, B_MN, OUT_F32>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) return -static_cast<int>(e);
        configured = true;
    }
    const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
    long long total = (MODE == MODE_MGROUP) ? 1ll * p.num_m_tiles * n_tiles : 1ll * p.num_groups * (p.M / BLOCK_M) * n_tiles;
    if (total <= 0) return 0;
    int grid = num_sms();
    if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
    if (total < grid) grid = static_cast<int>(total);
    kern<<<grid, NUM_THREADS, L::TOTAL, st>>>(p, tmA, tmB);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

}  // namespace lah

using namespace lah;

// ------------------------------------------------------------------------------------------------
// C ABI (called from python via ctypes; see ops/native.py)
// ------------------------------------------------------------------------------------------------
extern "C" const int* lah_get_epoch_base();

extern "C" {

// C[rows, N] = A[rows, K] @ B[g]^T (+bias) (+residual)
//   b_mn == 0: B is [G, N, K] (K contiguous);  b_mn == 1: B is [G, K, N] (N contiguous)
//   a_rows: rows of the A buffer (TMA bound); m_valid: rows of C that may be written
//   tile_group: device int[num_m_tiles] or null; out_f32: 0 => bf16 C, 1 => fp32 C
int lah_gemm_mgroup(const void* A, long long lda, int a_rows, const void* B, int G, int N, int K, int b_mn, void* C,
                    long long ldc, int out_f32, int m_valid, int num_m_tiles, const int* tile_group,
                    const float* bias, const void* residual, long long ldr, int block_n, int max_ctas,
                    const int* wait_flags, int wait_count, int wait_epoch, int* status, cudaStream_t stream) {
    if ((K % 8) || (N % 32) || (lda % 8)) return -2;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)a_rows};
        uint64_t str[1] = {(uint64_t)lda * 2};
        uint32_t box[2] = {BLOCK_K, BLOCK_M};
        int r = make_tmap_bf16(&tmA, A, 2, dims, str, box);
        if (r) return r;
    }
    if (!b_mn) {
        uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)G};
        uint64_t str[2] = {(uint64_t)K * 2, (uint64_t)N * K * 2};
        uint32_t box[3] = {BLOCK_K, (uint32_t)block_n, 1};
        int r = make_tmap_bf16(&tmB, B, 3, dims, str, box);
        if (r) return r;
    } else {
        uint64_t dims[3] = {(uint64_t)N, (uint64_t)K, (uint64_t)G};
        uint64_t str[2] = {(uint64_t)N * 2, (uint64_t)N * K * 2};
        uint32_t box[3] = {64, BLOCK_K, 1};
        int r = make_tmap_bf16(&tmB, B, 3, dims, str, box);
        if (r) return r;
    }
    GemmParams p;
    p.N = N; p.K = K; p.M = m_valid; p.num_groups = G; p.num_m_tiles = num_m_tiles; p.tile_group = tile_group;
    p.group_off = nullptr; p.C = C; p.ldc = ldc; p.c_group_stride = 0; p.bias = bias;
    p.residual = reinterpret_cast<const bf16*>(residual); p.ldr = ldr;
    p.wait_flags = wait_flags; p.wait_count = wait_count; p.wait_epoch = wait_epoch; p.epoch_base = lah_get_epoch_base(); p.status = status;
#define LAH_LAUNCH_M(BN, ST)                                                                                   \
    if (!b_mn && !out_f32) return launch<BN, ST, MODE_MGROUP, false, false, false>(p, tmA, tmB, max_ctas, stream); \
    if (b_mn && !out_f32) return launch<BN, ST, MODE_MGROUP, false, true, false>(p, tmA, tmB, max_ctas, stream);   \
    if (!b_mn && out_f32) return launch<BN, ST, MODE_MGROUP, false, false, true>(p, tmA, tmB, max_ctas, stream);   \
    return launch<BN, ST, MODE_MGROUP, false, true, true>(p, tmA, tmB, max_ctas, stream);
    if (block_n == 256) { LAH_LAUNCH_M(256, 4) }
    if (block_n == 128) { LAH_LAUNCH_M(128, 6) }
    if (block_n == 64) { LAH_LAUNCH_M(64, 8) }
#undef LAH_LAUNCH_M
    return -3;
}

// C[g][M, N] (fp32) = A[off[g]:off[g+1], :M]^T @ B[off[g]:off[g+1], :N]   (A, B row-major token matrices)
int lah_gemm_kgroup(const void* A, long long lda, const void* B, long long ldb, int total_rows, int G, int M, int N,
                    const int* group_off, float* C, long long ldc, long long c_group_stride, int block_n,
                    int max_ctas, cudaStream_t stream) {
    if ((M % 128) || (N % 32) || (lda % 8) || (ldb % 8)) return -2;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)M, (uint64_t)total_rows};
        uint64_t str[1] = {(uint64_t)lda * 2};
        uint32_t box[2] = {64, BLOCK_K};
        int r = make_tmap_bf16(&tmA, A, 2, dims, str, box);
        if (r) return r;
    }
    {
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)total_rows};
        uint64_t str[1] = {(uint64_t)ldb * 2};
        uint32_t box[2] = {64, BLOCK_K};
        int r = make_tmap_bf16(&tmB, B, 2, dims, str, box);
        if (r) return r;
    }
    GemmParams p;
    p.N = N; p.K = 0; p.M = M; p.num_groups = G; p.num_m_tiles = 0; p.tile_group = nullptr; p.group_off = group_off;
    p.C = C; p.ldc = ldc; p.c_group_stride = c_group_stride; p.bias = nullptr; p.residual = nullptr; p.ldr = 0;
    p.wait_flags = nullptr; p.wait_count = 0; p.wait_epoch = 0; p.epoch_base = nullptr; p.status = nullptr;
    if (block_n == 256) return launch<256, 4, MODE_KGROUP, true, true, true>(p, tmA, tmB, max_ctas, stream);
    if (block_n == 128) return launch<128, 6, MODE_KGROUP, true, true, true>(p, tmA, tmB, max_ctas, stream);
    if (block_n == 64) return launch<64, 8, MODE_KGROUP, true, true, true>(p, tmA, tmB, max_ctas, stream);
    return -3;
}

}  // extern "C"
