// grouped_gemm_fp8.cu — block-scaled FP8 (MXFP8: E4M3 data, one UE8M0 scale per 1x32 block along K) grouped GEMM on
// tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale, plus the quantisation kernels that produce its operands.
//
//   C[r, :] = (A_q[r, :] * 2^sfa) @ (W_q[g(r)] * 2^sfb)^T (+ bias) (+ act) (+ residual)       bf16 / fp32 out
//
// Same CTA-pair protocol as grouped_gemm_2cta.cu (two TMA producers, one MMA thread in the leader, multicast commits,
// two TMEM accumulator stages, smem-transposed coalesced epilogue).  Differences:
//   * 8-bit operands: BLOCK_K = 128 elements (one 128 B swizzle row), UMMA_K = 32 -> twice the math per smem byte;
//   * scale factors travel with every K block: TMA -> smem (512 B per 128 rows x 4 k-steps, already in the UTCCP
//     "32 x 128 bit" order, see sf_word_index) -> TMEM with tcgen05.cp.32x128b.warpx4 issued by the MMA thread right
//     before the four MMAs of the block (tcgen05.cp and tcgen05.mma execute in issue order, so ONE scale buffer in
//     TMEM is enough); the a_sf_id/b_sf_id fields of the instruction descriptor select the byte (k-step) of each word;
//   * TMEM budget: 2 accumulator stages + 12 scale columns must fit in 512 columns -> TILE_N = 192 (2*192 + 4 + 8).
//
// Used for the expert FFN forward GEMMs (BASELINE.json config "fp8 expert GEMM"); dgrad / wgrad stay in bf16.
#include "sm100.cuh"
#include <cuda_fp8.h>

namespace lah {
namespace f8 {

constexpr int TILE_M = 256;   // per pair
constexpr int CTA_M = 128;    // per CTA
constexpr int TILE_N = 192;   // per pair
constexpr int CTA_N = 96;     // B rows loaded per CTA
constexpr int BLOCK_K = 128;  // elements == bytes
constexpr int UMMA_K = 32;
constexpr int STAGES = 6;
constexpr int NUM_THREADS = 320;   // TMA warp, MMA warp, up to 8 epilogue warps

constexpr int A_BYTES = CTA_M * BLOCK_K;        // 16 KB
constexpr int B_BYTES = CTA_N * BLOCK_K;        // 12 KB
constexpr int SFA_BYTES = 512;                  // 128 rows x 4 k-steps
constexpr int SFB_BYTES = 1024;                 // 2 atoms of 128 rows (192 used)
constexpr int STAGE_TX = A_BYTES + B_BYTES + SFA_BYTES + SFB_BYTES;
constexpr int STAGE_BYTES = 30 * 1024;          // padded to a multiple of 1024 (swizzle atom alignment)
constexpr int SFA_OFF = A_BYTES + B_BYTES;
constexpr int SFB_OFF = SFA_OFF + SFA_BYTES;
constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
constexpr int EPI_BYTES = 8 * 32 * 80;          // max(8 warps x 32 rows x 80 B (bf16), 4 warps x 32 rows x 144 B (fp32))
constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
constexpr int SMEM_TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;
constexpr int TMEM_SFA_COL = 2 * TILE_N;        // 384
constexpr int TMEM_SFB_COL = TMEM_SFA_COL + 4;  // 388 .. 395
static_assert(STAGE_TX <= STAGE_BYTES, "stage layout");
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");

struct Params {
    int N, K, M, num_groups, num_m_tiles, n_tiles, num_kb;
    const int* tile_group;
    void* C;
    long long ldc;
    const float* bias;
    const bf16* residual;
    long long ldr;
    const int* wait_flags;
    int wait_count, wait_epoch;
    const int* epoch_base;   // device-side epoch base added to wait_epoch (nullptr: 0), see moe.cu Peers::step_ctr
    int* status;
    int act;
};

// Instruction descriptor of kind::mxf8f6f4.block_scale (cute/arch/mma_sm100_desc.hpp, InstrDescriptorBlockScaled):
//   [4,6) b_sf_id  [7,10) a_format (0 = E4M3)  [10,13) b_format  [15] a_major  [16] b_major  [17,23) N>>3
//   [23] scale_format (1 = UE8M0)  [24,29) M>>4  [29,31) a_sf_id  [31] k_size (0 = K32)
__host__ __device__ constexpr uint32_t make_idesc_mxf8(uint32_t M, uint32_t N, uint32_t sf_id) {
    return (sf_id << 4) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24) | (sf_id << 29);
}

// smem descriptor of a scale-factor chunk for tcgen05.cp: no swizzle, 8-row x 16 B core matrices 128 B apart
__device__ __forceinline__ uint64_t make_sf_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((128u >> 4) & 0x3FFFu) << 32;  // SBO = 128 B
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}
// smem (32 rows x 128 bit) -> TMEM (32 lanes x 4 columns, replicated into the 4 lane quadrants) in BOTH CTAs of the pair
__device__ __forceinline__ void utccp_32x128b_2sm(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_mxf8_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}

template <bool OUT_F32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_fp8_kernel(const Params p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmSFA, const __grid_constant__ CUtensorMap tmSFB) {
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
        tma_prefetch_desc(&tmSFA);
        tma_prefetch_desc(&tmSFB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 2);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 2 * EPI_WARPS);   // every epilogue warp of both CTAs
        }
        fence_mbar_init();
    }
    cluster_sync_all();
    if (warp == 1) tmem_alloc_2sm(tmem_ptr, 512);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int n_tiles = p.n_tiles;
    const int num_kb = p.num_kb;
    const int total_tiles = p.num_m_tiles * n_tiles;
    const int pair_id = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;

    auto decode = [&](int tile, int& m_tile, int& n_tile, int& group) -> bool {
        m_tile = tile / n_tiles;
        n_tile = tile - m_tile * n_tiles;
        group = p.tile_group ? __ldg(p.tile_group + 2 * m_tile) : 0;
        return group >= 0;
    };

    if (warp == 0 && lane == 0) {
        // =============================================================== TMA producer (both CTAs)
        if (p.wait_flags) {
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
            int m_tile, n_tile, g;
            if (!decode(tile, m_tile, n_tile, g)) continue;
            const int my_m = m_tile * TILE_M + cta_rank * CTA_M;
            const int my_n = n_tile * TILE_N + cta_rank * CTA_N;
            const int sfa_row = (my_m >> 7) * num_kb;                      // [row tile of 128][kb] chunks of 512 B
            const int sfb_row = ((g * n_tiles + n_tile) * num_kb) * 2;     // [g][n tile][kb][2 atoms]
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_TX);
                tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BLOCK_K, my_m);
                tma_load_3d_2sm(sa + A_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, my_n, g);
                tma_load_2d_2sm(sa + SFA_OFF, &tmSFA, &full_bar[stage], 0, sfa_row + kb);
                tma_load_2d_2sm(sa + SFB_OFF, &tmSFB, &full_bar[stage], 0, sfb_row + 2 * kb);
                if (!leader) mbar_arrive_cluster(&full_bar[stage], 0);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && lane == 0 && leader) {
        // =============================================================== MMA issuer (leader CTA only)
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        const uint32_t tsfa = tmem_base + TMEM_SFA_COL;
        const uint32_t tsfb = tmem_base + TMEM_SFB_COL;
        for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
            int m_tile, n_tile, g;
            if (!decode(tile, m_tile, n_tile, g)) continue;
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
                utccp_32x128b_2sm(tsfa, make_sf_desc(sa + SFA_OFF));
                utccp_32x128b_2sm(tsfb, make_sf_desc(sa + SFB_OFF));
                utccp_32x128b_2sm(tsfb + 4, make_sf_desc(sa + SFB_OFF + 512));
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint64_t da = make_smem_desc_sw128(sa + k * UMMA_K, 0, 1024);
                    const uint64_t db = make_smem_desc_sw128(sb + k * UMMA_K, 0, 1024);
                    umma_mxf8_ss_2sm(tmem_d, da, db, make_idesc_mxf8(TILE_M, TILE_N, k), tsfa, tsfb,
                                     (kb > 0 || k > 0) ? 1u : 0u);
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
        // bf16 output: EIGHT warps — two per TMEM lane quadrant, interleaved over the 32-column chunks.  With four warps
        // one warp per SM sub-partition walked a ~100-instruction dependent chain per chunk (tcgen05.ld -> bias ->
        // pack -> smem transpose -> stores) and the drain of a tile took longer than its MMAs (ncu: tensor pipe 46 %).
        const int lane_group = warp & 3;
        const int chunk_first = (warp - 2) >> 2;             // 0 or 1 (always 0 with four warps)
        constexpr int CHUNK_STEP = EPI_WARPS / 4;
        int iter = 0;
        for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
            int m_tile, n_tile, g;
            if (!decode(tile, m_tile, n_tile, g)) continue;
            const int n_col = n_tile * TILE_N;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tmem_full[as], aphase);
            tcgen05_fence_after();
            const int row_base = m_tile * TILE_M + cta_rank * CTA_M + lane_group * 32;
            const int row = row_base + lane;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + as * TILE_N;
            uint8_t* slab = smem + EPI_OFFSET + (warp - 2) * EPI_WARP_BYTES;
#pragma unroll 1
            for (int c0 = chunk_first * 32; c0 < TILE_N; c0 += CHUNK_STEP * 32) {
                const int col = n_col + c0;
                if (col >= p.N) break;  // tail tile (N % 192 != 0): columns past N are never stored
                uint32_t r[32];
                tmem_ld_32x32(taddr + c0, r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
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
                // transpose through a padded smem slab so that the global stores are row-contiguous
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
                uint8_t* cbase = reinterpret_cast<uint8_t*>(p.C) + (static_cast<long long>(col) * ELEM + piece * 16);
                const long long row_bytes = p.ldc * ELEM;
#pragma unroll
                for (int i = 0; i < 32 / ROWS_PER_INSTR; ++i) {
                    const int rl = i * ROWS_PER_INSTR + lane / LANES_PER_ROW;
                    const int grow = row_base + rl;
                    const int4 q = *reinterpret_cast<const int4*>(slab + rl * EPI_ROW_BYTES + piece * 16);
                    if (grow < p.M) *reinterpret_cast<int4*>(cbase + static_cast<long long>(grow) * row_bytes) = q;
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
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ quantisation
// Scale-factor storage.  Rows are split into tiles of `tile_rows` rows (128 for activations = one CTA's share of an M
// tile; 192 for weights = one N tile), every tile into atoms of 128 rows, K into blocks of 128 elements.  One
// (atom, K block) chunk is 128 32-bit words = 512 B, ordered the way tcgen05.cp.32x128b.warpx4 wants them:
//   word (r % 32) * 4 + (r / 32)  holds the four UE8M0 bytes (k-steps of 32 elements) of row r of the atom.
// Chunk index = ((tile * num_kb + kb) * atoms_per_tile + atom).
__host__ __device__ inline long long sf_chunks(long long rows, int tile_rows, int num_kb) {
    const long long tiles = (rows + tile_rows - 1) / tile_rows;
    return tiles * num_kb * ((tile_rows + 127) / 128);
}

__device__ __forceinline__ uint32_t e8m0_from_amax(float amax) {
    // smallest power of two s with amax / s <= 448 (the E4M3 maximum): no element saturates
    const float s = amax * (1.f / 448.f);
    uint32_t bits = __float_as_uint(s);
    uint32_t e = (bits >> 23) & 0xFFu;
    if (bits & 0x7FFFFFu) e += 1;
    return min(max(e, 1u), 253u);
}

// one thread per (row, 32-element block): 64 B (bf16) or 128 B (fp32) in, 32 B + one scale byte out
template <typename T>
__global__ void __launch_bounds__(256) quant_mxfp8_kernel(const T* __restrict__ in, long long ld_in,
                                                          uint8_t* __restrict__ out, long long ld_out,
                                                          uint8_t* __restrict__ sf, int rows_per_group, int groups, int K,
                                                          int tile_rows, const int* __restrict__ tile_group128,
                                                          const int* __restrict__ total_rows_dev) {
    const int blocks_per_row = K >> 5;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long grow = idx / blocks_per_row;  // row over all groups
    const int kb32 = static_cast<int>(idx - grow * blocks_per_row);
    if (grow >= static_cast<long long>(rows_per_group) * groups) return;
    if (total_rows_dev && grow >= *total_rows_dev) return;
    if (tile_group128 && __ldg(tile_group128 + (grow >> 7)) < 0) return;
    const int g = static_cast<int>(grow / rows_per_group);
    const int r = static_cast<int>(grow - static_cast<long long>(g) * rows_per_group);
    const T* src = in + grow * ld_in + kb32 * 32;
    float v[32];
    if (sizeof(T) == 2) {
        const int4* p4 = reinterpret_cast<const int4*>(src);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int4 q = __ldg(p4 + j);
            const uint32_t w[4] = {(uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_bf16x2(w[t]);
                v[8 * j + 2 * t] = f.x;
                v[8 * j + 2 * t + 1] = f.y;
            }
        }
    } else {
        const float4* p4 = reinterpret_cast<const float4*>(src);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 f = __ldg(p4 + j);
            v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
        }
    }
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) amax = fmaxf(amax, fabsf(v[j]));
    const uint32_t e = e8m0_from_amax(amax);
    const float inv = __uint_as_float((254u - e) << 23);
    uint32_t packed[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(v[4 * j] * inv, v[4 * j + 1] * inv), __NV_SATFINITE, __NV_E4M3);
        const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(v[4 * j + 2] * inv, v[4 * j + 3] * inv), __NV_SATFINITE, __NV_E4M3);
        packed[j] = lo | (hi << 16);
    }
    int4* dst = reinterpret_cast<int4*>(out + grow * ld_out + kb32 * 32);
    dst[0] = make_int4(packed[0], packed[1], packed[2], packed[3]);
    dst[1] = make_int4(packed[4], packed[5], packed[6], packed[7]);
    // scale byte
    const int num_kb = K >> 7;
    const int tiles_per_group = (rows_per_group + tile_rows - 1) / tile_rows;
    const int atoms = (tile_rows + 127) >> 7;
    const int tile = r / tile_rows;
    const int rt = r - tile * tile_rows;
    const int atom = rt >> 7, ra = rt & 127;
    const long long chunk = ((static_cast<long long>(g) * tiles_per_group + tile) * num_kb + (kb32 >> 2)) * atoms + atom;
    sf[chunk * 512 + ((ra & 31) * 4 + (ra >> 5)) * 4 + (kb32 & 3)] = static_cast<uint8_t>(e);
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

static int tmap(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* ptr, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw) {
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
    (void)elem_bytes;
    CUresult r = fn(tm, dt, rank, const_cast<void*>(ptr), gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

template <bool OUT_F32>
static int launch_fp8(const Params& p, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmSFA,
                      const CUtensorMap& tmSFB, int max_ctas, cudaStream_t st) {
    auto kern = gemm_fp8_kernel<OUT_F32>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
        if (e != cudaSuccess) return -static_cast<int>(e);
        configured = true;
    }
    const long long total = 1ll * p.num_m_tiles * p.n_tiles;
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
    kern<<<pairs * 2, NUM_THREADS, SMEM_TOTAL, st>>>(p, tmA, tmB, tmSFA, tmSFB);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

}  // namespace f8
}  // namespace lah

using namespace lah;
using namespace lah::f8;

extern "C" const int* lah_get_epoch_base();

extern "C" {

// bytes of the scale-factor buffer for `rows` rows per group (tile_rows = 128: activations, 192: weights)
long long lah_mxfp8_sf_bytes(long long rows_per_group, int groups, int K, int tile_rows) {
    return sf_chunks(rows_per_group, tile_rows, K / 128) * groups * 512;
}

// in [groups * rows_per_group, K] (bf16: in_f32 = 0, fp32: 1) -> out e4m3 (same shape, ld_out bytes per row) + scales
int lah_quant_mxfp8(const void* in, long long ld_in, int in_f32, void* out, long long ld_out, void* sf, int rows_per_group,
                    int groups, int K, int tile_rows, const int* tile_group128, const int* total_rows_dev,
                    cudaStream_t stream) {
    if ((K % 128) || (ld_out % 16) || (tile_rows != 128 && tile_rows != 192)) return -2;
    const long long threads = static_cast<long long>(rows_per_group) * groups * (K / 32);
    if (threads == 0) return 0;
    const int blocks = static_cast<int>((threads + 255) / 256);
    if (in_f32)
        quant_mxfp8_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(in), ld_in,
                                                             reinterpret_cast<uint8_t*>(out), ld_out,
                                                             reinterpret_cast<uint8_t*>(sf), rows_per_group, groups, K,
                                                             tile_rows, tile_group128, total_rows_dev);
    else
        quant_mxfp8_kernel<bf16><<<blocks, 256, 0, stream>>>(reinterpret_cast<const bf16*>(in), ld_in,
                                                            reinterpret_cast<uint8_t*>(out), ld_out,
                                                            reinterpret_cast<uint8_t*>(sf), rows_per_group, groups, K,
                                                            tile_rows, tile_group128, total_rows_dev);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

// A: e4m3 [a_rows, K] (lda bytes), sfa: activation scales (tile_rows = 128); B: e4m3 [G, N, K], sfb: weight scales
// (tile_rows = 192).  Expert groups padded to 256 rows; tile_group has one entry per 128 rows (like lah_gemm_mgroup2).
int lah_gemm_mgroup_fp8(const void* A, long long lda, int a_rows, const void* sfa, const void* B, const void* sfb, int G,
                        int N, int K, void* C, long long ldc, int out_f32, int m_valid, int num_m_tiles128,
                        const int* tile_group, const float* bias, const void* residual, long long ldr, int max_ctas,
                        const int* wait_flags, int wait_count, int wait_epoch, int* status, int act, cudaStream_t stream) {
    if ((K % 128) || (N % 64) || (lda % 16)) return -2;
    const int num_kb = K / 128;
    const int n_tiles = (N + TILE_N - 1) / TILE_N;
    CUtensorMap tmA, tmB, tmSFA, tmSFB;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)a_rows};
        uint64_t str[1] = {(uint64_t)lda};
        uint32_t box[2] = {BLOCK_K, CTA_M};
        int r = tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r) return r;
    }
    {
        uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)G};
        uint64_t str[2] = {(uint64_t)K, (uint64_t)N * K};
        uint32_t box[3] = {BLOCK_K, CTA_N, 1};
        int r = tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, B, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r) return r;
    }
    {
        const uint64_t chunks = (uint64_t)((a_rows + 127) / 128) * num_kb;
        uint64_t dims[2] = {128, chunks};
        uint64_t str[1] = {512};
        uint32_t box[2] = {128, 1};
        int r = tmap(&tmSFA, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, sfa, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (r) return r;
    }
    {
        const uint64_t chunks = (uint64_t)G * n_tiles * num_kb * 2;
        uint64_t dims[2] = {128, chunks};
        uint64_t str[1] = {512};
        uint32_t box[2] = {128, 2};
        int r = tmap(&tmSFB, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, sfb, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (r) return r;
    }
    Params p;
    p.N = N; p.K = K; p.M = m_valid; p.num_groups = G; p.num_m_tiles = (num_m_tiles128 + 1) / 2; p.n_tiles = n_tiles;
    p.num_kb = num_kb; p.tile_group = tile_group; p.C = C; p.ldc = ldc; p.bias = bias;
    p.residual = reinterpret_cast<const bf16*>(residual); p.ldr = ldr;
    p.wait_flags = wait_flags; p.wait_count = wait_count; p.wait_epoch = wait_epoch; p.epoch_base = lah_get_epoch_base(); p.status = status; p.act = act;
    return out_f32 ? launch_fp8<true>(p, tmA, tmB, tmSFA, tmSFB, max_ctas, stream)
                   : launch_fp8<false>(p, tmA, tmB, tmSFA, tmSFB, max_ctas, stream);
}

}  // extern "C"
