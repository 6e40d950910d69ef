// grouped_gemm_2cta.cu — CTA-pair (cta_group::2) version of the grouped GEMM: 256 x 256 output tile per pair.
//
// Why: a 128 x 256 tile per CTA needs (128+256)*64*2 B of operands per 4.2 MFLOP k-block = 85 FLOP/B from L2, which
// saturates the L2 before the tensor cores (measured: csrc/grouped_gemm.cu peaks at ~0.85-1.3 PFLOP/s).  Pairing two
// SMs on one 256 x 256 tile (tcgen05.mma.cta_group::2, M = 256) lets each CTA load only its own 128 rows of A and HALF of
// B (128 of the 256 columns); the tensor cores of both SMs read B from both shared memories -> 128 FLOP/B.
//
// Protocol (per pair; CTA 0 = leader):
//   * both CTAs run a TMA producer thread; loads use the `.cta_group::2` form and credit their bytes to the LEADER's
//     full barrier (count 2: one arrive.expect_tx from the leader's producer, one remote arrive from the peer's);
//   * only the leader's MMA thread issues tcgen05.mma; `tcgen05.commit ... multicast::cluster` (mask 0b11) releases the
//     smem stage in BOTH CTAs and publishes the accumulator to BOTH epilogues;
//   * each CTA's epilogue drains its own 128 TMEM lanes (= its 128 rows of the tile) and arrives on the leader's
//     tmem_empty barrier (count 8 = 4 warps x 2 CTAs).
// Same modes / operand majors / epilogues as grouped_gemm.cu; M-grouped mode requires expert groups padded to 256 rows.
#include "sm100.cuh"

namespace lah {
namespace pair {

constexpr int TILE_M = 256;   // per pair
constexpr int CTA_M = 128;    // per CTA
constexpr int TILE_N = 256;   // per pair
constexpr int CTA_N = 128;    // B rows loaded per CTA
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 6;
constexpr int NUM_THREADS = 320;   // TMA warp, MMA warp, up to 8 epilogue warps
constexpr int MODE_MGROUP = 0;
constexpr int MODE_KGROUP = 1;

constexpr int A_BYTES = CTA_M * BLOCK_K * 2;   // 16 KB
constexpr int B_BYTES = CTA_N * BLOCK_K * 2;   // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
constexpr int EPI_BYTES = 8 * 32 * 80;   // max(8 warps x 32 rows x 80 B (bf16 out), 4 warps x 32 rows x 144 B (fp32 out))
constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
constexpr int SMEM_TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;

struct Params {
    int N, K, M, num_groups, num_m_tiles;   // num_m_tiles counts 256-row tiles in MGROUP mode
    const int* tile_group;                  // MGROUP: group of every 128-row tile (we read entry 2*t)
    const int* group_off;                   // KGROUP
    void* C;
    long long ldc, c_group_stride;
    const float* bias;
    const bf16* residual;
    long long ldr;
    const int* wait_flags;   // receive-side fusion (see grouped_gemm.cu)
    int wait_count, wait_epoch;
    const int* epoch_base;   // device-side epoch base added to wait_epoch (nullptr: 0), see moe.cu Peers::step_ctr
    int* status;
    int act;                 // epilogue activation after the bias: 0 none, 1 ReLU, 2 GELU (erf)
    int accumulate;          // KGROUP (fp32 out): C += result (gradient accumulation across steps, update_every_*)
};

template <int MODE, bool A_MN, bool B_MN, bool OUT_F32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_kernel(const Params p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    // epilogue geometry: 32-column chunks; a chunk row is 64 B (bf16) or 128 B (fp32) + 16 B pad in the staging slab
    constexpr int EPI_WARPS = OUT_F32 ? 4 : 8;
    constexpr int EPI_ROW_BYTES = OUT_F32 ? 144 : 80;
    constexpr int EPI_WARP_BYTES = 32 * EPI_ROW_BYTES;
    static_assert(EPI_WARPS * EPI_WARP_BYTES <= EPI_BYTES, "epilogue staging");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 2);    // leader's: one arrival per CTA of the pair
            mbar_init(&empty_bar[i], 1);   // released by the multicast commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 2 * EPI_WARPS);  // leader's: every epilogue warp of both CTAs
        }
        fence_mbar_init();
    }
    cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / 2-SM allocation
    if (warp == 1) tmem_alloc_2sm(tmem_ptr, 512);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int n_tiles = (p.N + TILE_N - 1) / TILE_N;
    int total_tiles, tiles_per_group = 0;
    if (MODE == MODE_MGROUP) {
        total_tiles = p.num_m_tiles * n_tiles;
    } else {
        tiles_per_group = (p.M / TILE_M) * n_tiles;
        total_tiles = p.num_groups * tiles_per_group;
    }
    const int pair_id = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;

    auto decode = [&](int tile, int& m_row, int& n_col, int& group, int& k_begin, int& num_kb) -> bool {
        if (MODE == MODE_MGROUP) {
            const int m_tile = tile / n_tiles;
            const int n_tile = tile - m_tile * n_tiles;
            group = p.tile_group ? __ldg(p.tile_group + 2 * m_tile) : 0;
            m_row = m_tile * TILE_M;
            n_col = n_tile * TILE_N;
            k_begin = 0;
            num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
            return group >= 0;
        } else {
            group = tile / tiles_per_group;
            const int r = tile - group * tiles_per_group;
            const int m_tile = r / n_tiles;
            const int n_tile = r - m_tile * n_tiles;
            m_row = m_tile * TILE_M;
            n_col = n_tile * TILE_N;
            k_begin = __ldg(p.group_off + group);
            num_kb = (__ldg(p.group_off + group + 1) - k_begin) / BLOCK_K;
            return num_kb > 0;
        }
    };

    if (warp == 0 && lane == 0) {
        // =============================================================== TMA producer (both CTAs)
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
        for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
            int m_row, n_col, g, k_begin, num_kb;
            if (!decode(tile, m_row, n_col, g, k_begin, num_kb)) continue;
            const int my_m = m_row + cta_rank * CTA_M;   // my 128 rows of A / C
            const int my_n = n_col + cta_rank * CTA_N;   // my half of B
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                uint8_t* sb = sa + A_BYTES;
                const int k = k_begin + kb * BLOCK_K;
                if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                if (!A_MN) {
                    tma_load_2d_2sm(sa, &tmA, &full_bar[stage], k, my_m);
                } else {
#pragma unroll
                    for (int i = 0; i < CTA_M / 64; ++i)
                        tma_load_2d_2sm(sa + i * (BLOCK_K * 128), &tmA, &full_bar[stage], my_m + i * 64, k);
                }
                if (MODE == MODE_MGROUP) {
                    if (!B_MN) {
                        tma_load_3d_2sm(sb, &tmB, &full_bar[stage], k, my_n, g);
                    } else {
#pragma unroll
                        for (int i = 0; i < CTA_N / 64; ++i)
                            tma_load_3d_2sm(sb + i * (BLOCK_K * 128), &tmB, &full_bar[stage], my_n + i * 64, k, g);
                    }
                } else {
                    if (!B_MN) {
                        tma_load_2d_2sm(sb, &tmB, &full_bar[stage], k, my_n);
                    } else {
#pragma unroll
                        for (int i = 0; i < CTA_N / 64; ++i)
                            tma_load_2d_2sm(sb + i * (BLOCK_K * 128), &tmB, &full_bar[stage], my_n + i * 64, k);
                    }
                }
                if (!leader) mbar_arrive_cluster(&full_bar[stage], 0);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && lane == 0 && leader) {
        // =============================================================== MMA issuer (leader CTA only)
        constexpr uint32_t idesc = make_idesc_bf16_f32(TILE_M, TILE_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
        constexpr uint32_t A_LBO = A_MN ? BLOCK_K * 128 : 0, B_LBO = B_MN ? BLOCK_K * 128 : 0;
        constexpr uint32_t A_KSTEP = A_MN ? UMMA_K * 128 : UMMA_K * 2, B_KSTEP = B_MN ? UMMA_K * 128 : UMMA_K * 2;
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
            int m_row, n_col, g, k_begin, num_kb;
            if (!decode(tile, m_row, n_col, g, k_begin, num_kb)) continue;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + as * TILE_N;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t sb = sa + A_BYTES;
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint64_t da = make_smem_desc_sw128(sa + k * A_KSTEP, A_LBO, 1024);
                    const uint64_t db = make_smem_desc_sw128(sb + k * B_KSTEP, B_LBO, 1024);
                    umma_bf16_ss_2sm(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit_2sm(&empty_bar[stage], 0b11);
                if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[as], 0b11);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            ++iter;
        }
    } else if (warp >= 2 && warp < 2 + EPI_WARPS) {
        // =============================================================== epilogue (both CTAs; own 128 rows)
        // Each thread owns one accumulator row (tcgen05.ld 32x32b).  Storing rows straight from registers makes every
        // warp-level 16 B store touch 32 different 128 B lines, so the warp transposes each 32-column chunk through a
        // padded smem slab and writes row-contiguous segments instead.  bf16 output uses EIGHT warps (two per TMEM lane
        // quadrant, interleaved over the chunks): with four, one warp per SM sub-partition walked a ~100-instruction
        // dependent chain per chunk and the drain of a short-K tile took longer than its MMAs (ncu: tensor pipe 38 %).
        const int lane_group = warp & 3;
        const int chunk_first = (warp - 2) >> 2;             // 0 or 1 (always 0 with four warps)
        constexpr int CHUNK_STEP = EPI_WARPS / 4;
        int iter = 0;
        for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
            int m_row, n_col, g, k_begin, num_kb;
            if (!decode(tile, m_row, n_col, g, k_begin, num_kb)) continue;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tmem_full[as], aphase);
            tcgen05_fence_after();
            const int row_base = m_row + cta_rank * CTA_M + lane_group * 32;   // first row of this warp's slab
            const int row = row_base + lane;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + as * TILE_N;
            uint8_t* slab = smem + EPI_OFFSET + (warp - 2) * EPI_WARP_BYTES;
#pragma unroll 1
            for (int c0 = chunk_first * 32; c0 < TILE_N; c0 += CHUNK_STEP * 32) {
                const int col = n_col + c0;
                if (col >= p.N) break;
                uint32_t r[32];
                tmem_ld_32x32(taddr + c0, r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (MODE == MODE_MGROUP) {
                    if (p.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + static_cast<long long>(g) * p.N + col);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(bp + j);
                            v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                        }
                    }
                    if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (p.act == 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752f));
                    }
                    if (p.residual && row < p.M) {
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
                uint8_t* my = slab + lane * EPI_ROW_BYTES;
                if (OUT_F32) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(my + 16 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int4 q;
                        q.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                        q.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                        q.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                        q.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                        *reinterpret_cast<int4*>(my + 16 * j) = q;
                    }
                }
                __syncwarp();
                constexpr int ELEM = OUT_F32 ? 4 : 2;
                constexpr int LANES_PER_ROW = 32 * ELEM / 16;        // 16 B pieces of a 32-column row segment: 8 / 4
                constexpr int ROWS_PER_INSTR = 32 / LANES_PER_ROW;   // 4 / 8
                const int piece = lane % LANES_PER_ROW;
                uint8_t* cbase = reinterpret_cast<uint8_t*>(p.C) +
                                 ((MODE == MODE_KGROUP ? static_cast<long long>(g) * p.c_group_stride : 0ll) + col) * ELEM +
                                 piece * 16;
                const long long row_bytes = p.ldc * ELEM;
#pragma unroll
                for (int i = 0; i < 32 / ROWS_PER_INSTR; ++i) {
                    const int rl = i * ROWS_PER_INSTR + lane / LANES_PER_ROW;
                    const int grow = row_base + rl;
                    int4 q = *reinterpret_cast<const int4*>(slab + rl * EPI_ROW_BYTES + piece * 16);
                    if (MODE == MODE_KGROUP && OUT_F32 && p.accumulate) {
                        const float4 o = *reinterpret_cast<const float4*>(cbase + static_cast<long long>(grow) * row_bytes);
                        float4 n = *reinterpret_cast<float4*>(&q);
                        n.x += o.x; n.y += o.y; n.z += o.z; n.w += o.w;
                        q = *reinterpret_cast<int4*>(&n);
                    }
                    if (MODE == MODE_KGROUP || grow < p.M)
                        *reinterpret_cast<int4*>(cbase + static_cast<long long>(grow) * row_bytes) = q;
                }
                __syncwarp();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty[as], 0);
            ++iter;
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();  // the leader's MMAs read the peer's smem / write its TMEM: nobody leaves early
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled encode_fn() {
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

static int tmap_bf16(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
    PFN_encodeTiled fn = encode_fn();
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

template <int MODE, bool A_MN, bool B_MN, bool OUT_F32>
static int launch2(const Params& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int max_ctas, cudaStream_t st) {
    auto kern = gemm2_kernel<MODE, A_MN, B_MN, OUT_F32>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
        if (e != cudaSuccess) return -static_cast<int>(e);
        configured = true;
    }
    const int n_tiles = (p.N + TILE_N - 1) / TILE_N;
    const long long total = (MODE == MODE_MGROUP) ? 1ll * p.num_m_tiles * n_tiles
                                                  : 1ll * p.num_groups * (p.M / TILE_M) * n_tiles;
    if (total <= 0) return 0;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    int pairs = sms / 2;
    if (max_ctas > 0 && max_ctas / 2 < pairs) pairs = max_ctas / 2 > 0 ? max_ctas / 2 : 1;
    if (total < pairs) pairs = static_cast<int>(total);
    kern<<<pairs * 2, NUM_THREADS, SMEM_TOTAL, st>>>(p, tmA, tmB);  // cluster dims are a kernel attribute
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

}  // namespace pair
}  // namespace lah

using namespace lah;
using namespace lah::pair;

extern "C" const int* lah_get_epoch_base();

extern "C" {

// same contract as lah_gemm_mgroup, but expert groups must be padded to 256 rows; tile_group still has one entry per
// 128 rows (entries 2t and 2t+1 agree)
int lah_gemm_mgroup2(const void* A, long long lda, int a_rows, const void* B, int G, int N, int K, int b_mn, void* C,
                     long long ldc, int out_f32, int m_valid, int num_m_tiles128, const int* tile_group,
                     const float* bias, const void* residual, long long ldr, int max_ctas, const int* wait_flags,
                     int wait_count, int wait_epoch, int* status, int act, cudaStream_t stream) {
    if ((K % 8) || (N % 32) || (lda % 8)) return -2;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)a_rows};
        uint64_t str[1] = {(uint64_t)lda * 2};
        uint32_t box[2] = {BLOCK_K, CTA_M};
        int r = tmap_bf16(&tmA, A, 2, dims, str, box);
        if (r) return r;
    }
    if (!b_mn) {
        uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)G};
        uint64_t str[2] = {(uint64_t)K * 2, (uint64_t)N * K * 2};
        uint32_t box[3] = {BLOCK_K, CTA_N, 1};
        int r = tmap_bf16(&tmB, B, 3, dims, str, box);
        if (r) return r;
    } else {
        uint64_t dims[3] = {(uint64_t)N, (uint64_t)K, (uint64_t)G};
        uint64_t str[2] = {(uint64_t)N * 2, (uint64_t)N * K * 2};
        uint32_t box[3] = {64, BLOCK_K, 1};
        int r = tmap_bf16(&tmB, B, 3, dims, str, box);
        if (r) return r;
    }
    Params p;
    p.N = N; p.K = K; p.M = m_valid; p.num_groups = G; p.num_m_tiles = (num_m_tiles128 + 1) / 2; p.tile_group = tile_group;
    p.group_off = nullptr; p.C = C; p.ldc = ldc; p.c_group_stride = 0; p.bias = bias;
    p.residual = reinterpret_cast<const bf16*>(residual); p.ldr = ldr;
    p.wait_flags = wait_flags; p.wait_count = wait_count; p.wait_epoch = wait_epoch; p.epoch_base = lah_get_epoch_base(); p.status = status; p.act = act; p.accumulate = 0;
    if (!b_mn && !out_f32) return launch2<MODE_MGROUP, false, false, false>(p, tmA, tmB, max_ctas, stream);
    if (b_mn && !out_f32) return launch2<MODE_MGROUP, false, true, false>(p, tmA, tmB, max_ctas, stream);
    if (!b_mn && out_f32) return launch2<MODE_MGROUP, false, false, true>(p, tmA, tmB, max_ctas, stream);
    return launch2<MODE_MGROUP, false, true, true>(p, tmA, tmB, max_ctas, stream);
}

int lah_gemm_kgroup2(const void* A, long long lda, const void* B, long long ldb, int total_rows, int G, int M, int N,
                     const int* group_off, float* C, long long ldc, long long c_group_stride, int max_ctas,
                     int accumulate, cudaStream_t stream) {
    if ((M % 256) || (N % 32) || (lda % 8) || (ldb % 8)) return -2;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)M, (uint64_t)total_rows};
        uint64_t str[1] = {(uint64_t)lda * 2};
        uint32_t box[2] = {64, BLOCK_K};
        int r = tmap_bf16(&tmA, A, 2, dims, str, box);
        if (r) return r;
    }
    {
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)total_rows};
        uint64_t str[1] = {(uint64_t)ldb * 2};
        uint32_t box[2] = {64, BLOCK_K};
        int r = tmap_bf16(&tmB, B, 2, dims, str, box);
        if (r) return r;
    }
    Params p;
    p.N = N; p.K = 0; p.M = M; p.num_groups = G; p.num_m_tiles = 0; p.tile_group = nullptr; p.group_off = group_off;
    p.C = C; p.ldc = ldc; p.c_group_stride = c_group_stride; p.bias = nullptr; p.residual = nullptr; p.ldr = 0;
    p.wait_flags = nullptr; p.wait_count = 0; p.wait_epoch = 0; p.epoch_base = nullptr; p.status = nullptr; p.act = 0; p.accumulate = accumulate;
    return launch2<MODE_KGROUP, true, true, true>(p, tmA, tmB, max_ctas, stream);
}

}  // extern "C"
