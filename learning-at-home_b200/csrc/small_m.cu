// small_m.cu — the "many small requests" regime of the reference (64 trainers x batch 4 -> O(1..64) rows per expert and
// step; /root/reference/experiments/convergence notebooks, lib/runtime/task_pool.py:105-125) on sm_100a.
//
// With a handful of rows per expert every expert GEMM is a WEIGHT-STREAMING problem (12.6 MB of bf16 weights per
// active expert and forward at hid 512) and the optimizer is a STATE-STREAMING problem (fp32 p, m, v, vmax).  Two
// kernels, both persistent / warp-specialised / TMA-fed with accumulators in TMEM:
//
//   swapab_kernel      D^T[out_features, tokens] = W[out_features, K] * X^T[K, tokens]   ("swap-AB": the weights sit on
//                      the 128-wide MMA-M side, the expert's 16..128 tokens on the MMA-N side, so a group of 16 rows
//                      costs one N=16 instruction instead of a 256-row padded tile).  A tile = (expert, 128-token
//                      chunk, 128-row slice of the weight matrix) streamed through a 6-stage TMA ring; experts with
//                      more than 128 rows are spread over several CTAs (hot experts).  A_MN selects dgrad
//                      (W^T read straight from the same [out, in] tensor as an MN-major operand).
//   wgrad_adam_kernel  dW tile = dY^T X on tcgen05 (both operands MN-major, two 32-token stages, reduction over the
//                      expert's tokens = the gradient reduction over all trainers that routed to it) with the per-expert AMSGrad step FUSED
//                      INTO THE EPILOGUE: p / m / v / vmax tiles arrive through TMA (128B-swizzled smem, per-warp
//                      3-slot ring), are updated in place from the TMEM accumulator and leave through TMA stores; the
//                      bf16 mirror is written from registers.  The weight gradient never exists in HBM: 34 B / parameter
//                      instead of 46 (wgrad write 4 + Adam 38 + re-read 4).
//                      Reference semantics: one torch.optim.Adam(amsgrad=True) step per expert right after its backward
//                      (/root/reference/lib/runtime/expert_backend.py:90-97).
#include "sm100.cuh"

namespace lah {
namespace smallm {

constexpr int BM = 128;      // MMA M: weight rows (swapab) / dY features (wgrad)
constexpr int BK = 64;       // k elements per stage row (128 B of bf16 = one swizzle row)
constexpr int UMMA_K = 16;
constexpr int BN_MAX = 128;  // max tokens per MMA (swapab) / X features per tile (wgrad)
constexpr int NUM_THREADS = 192;

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

static int make_tmap(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* ptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
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
    (void)esize;
    CUresult r = fn(tm, dt, rank, const_cast<void*>(ptr), gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

static int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

// =====================================================================================================================
// swap-AB grouped linear
// =====================================================================================================================
namespace sab {

constexpr int STAGES = 6;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES = BN_MAX * BK * 2;      // 16 KB (only box_rows * 128 B are filled)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
constexpr int QD = 4;                         // depth of the tile queue (dynamic scheduler)
constexpr int MAX_G = 1023;                   // groups per launch (prefix table of 128-token chunks lives in smem)
constexpr int CUM_OFFSET = BAR_OFFSET + (2 * STAGES + 4 + 2 * QD) * 8 + QD * 4 + 16;
constexpr int SMEM_TOTAL = CUM_OFFSET + (MAX_G + 1) * 4 + 1024;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

struct Params {
    int G, M_out, K;          // groups, output features (rows of the weight operand), reduction length
    const int* group_off;     // [G] first row of the group inside the token-major buffers
    const int* group_rows;    // [G] valid rows (0: skip the group)
    bf16* out;                // [rows, M_out]  token-major
    long long ldo;
    const float* bias;        // [G, M_out] or nullptr
    const bf16* residual;     // [rows, ldr] or nullptr
    long long ldr;
    const int* wait_flags;    // receive-side fusion: rows pushed by the peers must have landed (see grouped_gemm.cu)
    int wait_count, wait_epoch;
    const int* epoch_base;    // optional device-side epoch base added to wait_epoch (CUDA-graph replay)
    int* status;
    int* tile_counter;        // work-stealing tile counter (zeroed before the launch)
};

__device__ __forceinline__ int box_rows_of(int nn) { return nn <= 16 ? 16 : (nn <= 32 ? 32 : (nn <= 64 ? 64 : 128)); }

// Tiles are handed out DYNAMICALLY: SMs do not get equal shares of HBM bandwidth (ncu on the static version: SMs idle 15-20 %
// of the kernel waiting for the slowest one), so the TMA-producer thread of every CTA draws the next tile from a global
// counter and publishes it to the other roles of its CTA through a small smem queue.
// A tile is (group, 128-token chunk, 128-feature slice): a HOT expert (routing collapses while training; with 8 experts per
// rank one of them can hold > 1000 of the rank's rows) is split over as many CTAs as it has chunks instead of one CTA
// walking all of its chunks back to back - the slowest rank of a step is the one that owns the hottest expert, and every
// other rank waits for it in the combine.  Every CTA builds the prefix table of chunks per group in shared memory.
__device__ __forceinline__ int find_group(const int* cum, int G, int gc) {   // largest g with cum[g] <= gc
    int lo = 0, hi = G - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (cum[mid] <= gc) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <bool A_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
swapab_kernel(const Params p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB16,
              const __grid_constant__ CUtensorMap tmB32, const __grid_constant__ CUtensorMap tmB64,
              const __grid_constant__ CUtensorMap tmB128) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* q_full = tmem_empty + 2;
    uint64_t* q_empty = q_full + QD;
    volatile int* q_tile = reinterpret_cast<volatile int*>(q_empty + QD);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(const_cast<int*>(q_tile) + QD);
    int* cum = reinterpret_cast<int*>(smem + CUM_OFFSET);   // cum[g] = 128-token chunks of the groups before g

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 2) {   // prefix table of chunks (warp scan, 32 groups per round)
        int carry = 0;
        for (int base = 0; base < p.G; base += 32) {
            const int g = base + lane;
            int x = g < p.G ? (__ldg(p.group_rows + g) + BN_MAX - 1) / BN_MAX : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            if (g < p.G) cum[g + 1] = carry + x;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) cum[0] = 0;
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB16);
        tma_prefetch_desc(&tmB32);
        tma_prefetch_desc(&tmB64);
        tma_prefetch_desc(&tmB128);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        for (int i = 0; i < QD; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 5);   // MMA thread + 4 epilogue warps
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 2 * BN_MAX);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int m_slices = p.M_out / BM;
    const int total = cum[p.G] * m_slices;
    const int num_kb = p.K / BK;

    if (warp == 0 && lane == 0) {
        // =============================================================== scheduler + TMA producer
        if (p.wait_flags) {
            const int epoch = p.wait_epoch + (p.epoch_base ? *p.epoch_base : 0);
            const unsigned long long t_wait = globaltimer_ns();
            for (int sidx = 0; sidx < p.wait_count; ++sidx)
                spin_flag_ft(p.wait_flags + sidx, epoch, p.status, sidx, p.epoch_base ? p.epoch_base[1] : 0);
            if (blockIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.status + 2), globaltimer_ns() - t_wait);
            fence_proxy_async_global();
        }
        int stage = 0, qi = 0;
        uint32_t phase = 0, qphase = 0;
        while (true) {
            const int tile = atomicAdd(p.tile_counter, 1);
            mbar_wait(&q_empty[qi], qphase ^ 1);
            q_tile[qi] = tile;
            mbar_arrive(&q_full[qi]);
            if (++qi == QD) {
                qi = 0;
                qphase ^= 1;
            }
            if (tile >= total) break;
            const int gc = tile / m_slices, ms = tile - gc * m_slices;
            const int g = find_group(cum, p.G, gc);
            const int n0 = (gc - cum[g]) * BN_MAX;
            const int rows = __ldg(p.group_rows + g);
            const int off = __ldg(p.group_off + g);
            const int box = box_rows_of(min(rows - n0, BN_MAX));
            const CUtensorMap* tmB = box == 16 ? &tmB16 : (box == 32 ? &tmB32 : (box == 64 ? &tmB64 : &tmB128));
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                uint8_t* sb = sa + A_BYTES;
                const int k = kb * BK;
                mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + box * 128);
                if (!A_MN) {
                    tma_load_3d(sa, &tmA, &full_bar[stage], k, ms * BM, g);
                } else {
#pragma unroll
                    for (int i = 0; i < BM / 64; ++i)
                        tma_load_3d(sa + i * (BK * 128), &tmA, &full_bar[stage], ms * BM + i * 64, k, g);
                }
                tma_load_2d(sb, tmB, &full_bar[stage], k, off + n0);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // =============================================================== MMA issuer
        constexpr uint32_t A_LBO = A_MN ? BK * 128 : 0;
        constexpr uint32_t A_KSTEP = A_MN ? UMMA_K * 128 : UMMA_K * 2;
        int stage = 0, qi = 0;
        uint32_t phase = 0, qphase = 0;
        int iter = 0;
        while (true) {
            mbar_wait(&q_full[qi], qphase);
            const int tile = q_tile[qi];
            mbar_arrive(&q_empty[qi]);
            if (++qi == QD) {
                qi = 0;
                qphase ^= 1;
            }
            if (tile >= total) break;
            const int gc = tile / m_slices;
            const int g = find_group(cum, p.G, gc);
            const int nn = min(__ldg(p.group_rows + g) - (gc - cum[g]) * BN_MAX, BN_MAX);
            const uint32_t n16 = static_cast<uint32_t>((nn + 15) & ~15);
            const uint32_t idesc = make_idesc_bf16_f32(BM, n16, A_MN ? 1u : 0u, 0u);
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + as * BN_MAX;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t sb = sa + A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t da = make_smem_desc_sw128(sa + k * A_KSTEP, A_LBO, 1024);
                    const uint64_t db = make_smem_desc_sw128(sb + k * (UMMA_K * 2), 0, 1024);
                    umma_bf16_ss(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (kb == num_kb - 1) umma_commit(&tmem_full[as]);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            ++iter;
        }
    } else if (warp >= 2) {
        // =============================================================== epilogue: thread = output feature, column = token
        const int lane_group = warp & 3;
        int iter = 0, qi = 0;
        uint32_t qphase = 0;
        while (true) {
            mbar_wait(&q_full[qi], qphase);
            const int tile = q_tile[qi];
            __syncwarp();
            if (lane == 0) mbar_arrive(&q_empty[qi]);
            if (++qi == QD) {
                qi = 0;
                qphase ^= 1;
            }
            if (tile >= total) break;
            const int gc = tile / m_slices, ms = tile - gc * m_slices;
            const int g = find_group(cum, p.G, gc);
            const int n0 = (gc - cum[g]) * BN_MAX;
            const int nn = min(__ldg(p.group_rows + g) - n0, BN_MAX);
            const int off = __ldg(p.group_off + g);
            const int feat = ms * BM + lane_group * 32 + lane;
            const float bias = p.bias ? __ldg(p.bias + static_cast<long long>(g) * p.M_out + feat) : 0.f;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tmem_full[as], aphase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + as * BN_MAX;
            // the padding rows of the group's last 16-row block are WRITTEN too (their inputs are zero rows, so they get
            // bias / zero): every later kernel may then read whole 16-row blocks without meeting stale memory, and the
            // k-step masked wgrad sees exact zeros there
            const int n16 = (nn + 15) & ~15;
#pragma unroll 1
            for (int c0 = 0; c0 < n16; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c0, r);
                tmem_ld_wait();
                const int cn = min(32, n16 - c0);
                const long long row0 = off + n0 + c0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j < cn) {
                        float v = __uint_as_float(r[j]) + bias;
                        if (p.residual) v += __bfloat162float(p.residual[(row0 + j) * p.ldr + feat]);
                        p.out[(row0 + j) * p.ldo + feat] = __float2bfloat16(v);   // warp: 32 consecutive features = 64 B
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
        tmem_dealloc(tmem_base, 2 * BN_MAX);
    }
}

}  // namespace sab

// =====================================================================================================================
// fused wgrad + AMSGrad
// =====================================================================================================================
namespace wa {

constexpr int WBK = 32;                                  // tokens per operand stage
constexpr int OP_STAGES = 2;                             // a hot expert has tens of k-blocks: loads run ahead of the MMAs
constexpr int OP_STAGE_BYTES = 2 * (WBK * BM * 2);       // A [32 t][128 n] + B [32 t][128 k] (two 64-wide MN atoms each)
constexpr int OP_BYTES = OP_STAGES * OP_STAGE_BYTES;
constexpr int SLOTS = 3;                                 // state ring per epilogue warp
constexpr int STATE_TILE = 32 * 128;                     // 32 rows x 32 fp32 (one swizzle-128B atom group)
constexpr int SLOT_BYTES = 4 * STATE_TILE;               // p, m, v, vmax
constexpr int STATE_OFFSET = OP_BYTES;
constexpr int BAR_OFFSET = STATE_OFFSET + 4 * SLOTS * SLOT_BYTES;
constexpr int QD = 4;                                    // depth of the tile queue (dynamic scheduler)
constexpr int CHUNKS = BN_MAX / 32;                      // 32-column chunks per tile
constexpr int SMEM_TOTAL = BAR_OFFSET + (2 * OP_STAGES + 4 + 4 * SLOTS + 2 * QD) * 8 + QD * 4 + 16 + 1024;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

struct Params {
    int G, N, K;              // groups (slots of the segment), rows (= features of dY) and columns (= features of X) of W
    const int* group_off;     // [G] first token row of the group in dy / x
    const int* group_rows;    // [G] valid rows (0: the expert is not stepped)
    const int* skip;          // optional [G, 2]: skip[2g] >= 0 -> handled by the unfused path (shadowed experts)
    const int* step;          // [G] per-expert step count AFTER this update
    bf16* p_bf16;             // [G, N, K] bf16 mirror consumed by the GEMMs
    float lr, beta1, beta2, eps;
    int amsgrad;
    int* tile_counter;        // work-stealing tile counter (zeroed before the launch)
    const int* poison;        // optional status word: a step that timed out on a peer must not update anything
};

struct Tile {
    int g, mt, nt;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_adam_kernel(const Params p, const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmM,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmVM) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* op_full = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
    uint64_t* op_empty = op_full + OP_STAGES;
    uint64_t* tmem_full = op_empty + OP_STAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;   // [2]
    uint64_t* st_full = tmem_empty + 2;     // [4 warps][SLOTS]
    uint64_t* q_full = st_full + 4 * SLOTS;
    uint64_t* q_empty = q_full + QD;
    volatile int* q_tile = reinterpret_cast<volatile int*>(q_empty + QD);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(const_cast<int*>(q_tile) + QD);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (p.poison && (*p.poison & 1)) return;   // degraded step (a peer timed out): no optimizer step from partial data

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmP);
        tma_prefetch_desc(&tmM);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmVM);
        for (int i = 0; i < OP_STAGES; ++i) {
            mbar_init(&op_full[i], 1);
            mbar_init(&op_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        for (int i = 0; i < QD; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 5);   // MMA thread + 4 epilogue warps
        }
        fence_mbar_init();
    }
    if (warp >= 2 && lane == 0) {   // every epilogue warp owns (initialises, arms, waits on) the barriers of its state ring
        for (int i = 0; i < SLOTS; ++i) mbar_init(&st_full[(warp - 2) * SLOTS + i], 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 2 * BN_MAX);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int n_tiles = p.K / BN_MAX;
    const int tiles_per_group = (p.N / BM) * n_tiles;
    const int total = p.G * tiles_per_group;
    auto decode = [&](int tile) {
        Tile t;
        t.g = tile / tiles_per_group;
        const int r = tile - t.g * tiles_per_group;
        t.mt = r / n_tiles;
        t.nt = r - t.mt * n_tiles;
        return t;
    };

    if (warp == 0 && lane == 0) {
        // =============================================================== scheduler + operand producer (dY^T / X^T k-blocks)
        // tiles are drawn from a global counter (work stealing: SMs see different HBM bandwidth, a static split leaves the
        // fast ones idle for ~15 % of the kernel) and published to the MMA thread and the epilogue warps through a queue
        uint32_t phase = 0, qphase = 0;
        int qi = 0, stage = 0;
        while (true) {
            int tile;
            while (true) {
                tile = atomicAdd(p.tile_counter, 1);
                if (tile >= total) break;
                const int g = tile / tiles_per_group;
                if (__ldg(p.group_rows + g) > 0 && !(p.skip && __ldg(p.skip + 2 * g) >= 0)) break;
                atomicMax(p.tile_counter, (g + 1) * tiles_per_group);   // skip the rest of an inactive group
            }
            mbar_wait(&q_empty[qi], qphase ^ 1);
            q_tile[qi] = tile;
            mbar_arrive(&q_full[qi]);
            if (++qi == QD) {
                qi = 0;
                qphase ^= 1;
            }
            if (tile >= total) break;
            const Tile t = decode(tile);
            const int off = __ldg(p.group_off + t.g), rows = __ldg(p.group_rows + t.g);
            for (int t0 = 0; t0 < rows; t0 += WBK) {
                mbar_wait(&op_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&op_full[stage], OP_STAGE_BYTES);
                uint8_t* sa = smem + stage * OP_STAGE_BYTES;
                uint8_t* sb = sa + WBK * BM * 2;
#pragma unroll
                for (int i = 0; i < BM / 64; ++i)
                    tma_load_2d(sa + i * (WBK * 128), &tmDY, &op_full[stage], t.mt * BM + i * 64, off + t0);
#pragma unroll
                for (int i = 0; i < BN_MAX / 64; ++i)
                    tma_load_2d(sb + i * (WBK * 128), &tmX, &op_full[stage], t.nt * BN_MAX + i * 64, off + t0);
                if (++stage == OP_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // =============================================================== MMA issuer: only ceil(rows/16) k-steps of a block
        constexpr uint32_t idesc = make_idesc_bf16_f32(BM, BN_MAX, 1u, 1u);
        uint32_t phase = 0, qphase = 0;
        int iter = 0, qi = 0, stage = 0;
        while (true) {
            mbar_wait(&q_full[qi], qphase);
            const int tile = q_tile[qi];
            mbar_arrive(&q_empty[qi]);
            if (++qi == QD) {
                qi = 0;
                qphase ^= 1;
            }
            if (tile >= total) break;
            const Tile t = decode(tile);
            const int rows = __ldg(p.group_rows + t.g);
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + as * BN_MAX;
            for (int t0 = 0; t0 < rows; t0 += WBK) {
                mbar_wait(&op_full[stage], phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + stage * OP_STAGE_BYTES);
                const uint32_t sb = sa + WBK * BM * 2;
                const int ksteps = min(WBK / UMMA_K, (rows - t0 + UMMA_K - 1) / UMMA_K);
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t da = make_smem_desc_sw128(sa + k * (UMMA_K * 128), WBK * 128, 1024);
                    const uint64_t db = make_smem_desc_sw128(sb + k * (UMMA_K * 128), WBK * 128, 1024);
                    umma_bf16_ss(tmem_d, da, db, idesc, (t0 > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&op_empty[stage]);
                if (t0 + WBK >= rows) umma_commit(&tmem_full[as]);
                if (++stage == OP_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            ++iter;
        }
    } else if (warp >= 2) {
        // =============================================================== epilogue: state streaming + AMSGrad
        const int q = warp & 3;                     // TMEM lane quadrant = rows [32q, 32q+32) of the tile
        uint8_t* ring = smem + STATE_OFFSET + (warp - 2) * SLOTS * SLOT_BYTES;
        uint64_t* full = st_full + (warp - 2) * SLOTS;
        const int n_state = p.amsgrad ? 4 : 3;
        // chunk n of this CTA = 32-column chunk (n % CHUNKS) of the (n / CHUNKS)-th tile drawn from the queue; the loads run
        // two chunks ahead of the update, so at most two tiles are live: fifo[(n / CHUNKS) & 1]
        Tile fifo[2];
        int fetched = 0, qi = 0;
        uint32_t qphase = 0;
        bool drained = false;
        auto have_chunk = [&](int n) -> bool {
            while (!drained && n / CHUNKS >= fetched) {
                mbar_wait(&q_full[qi], qphase);
                const int tile = q_tile[qi];
                __syncwarp();
                if (lane == 0) mbar_arrive(&q_empty[qi]);
                if (++qi == QD) {
                    qi = 0;
                    qphase ^= 1;
                }
                if (tile >= total) {
                    drained = true;
                } else {
                    fifo[fetched & 1] = decode(tile);
                    ++fetched;
                }
            }
            return n / CHUNKS < fetched;
        };
        auto issue_load = [&](int n) {   // lane 0 only
            const Tile& t = fifo[(n / CHUNKS) & 1];
            const int cc = n % CHUNKS, slot = n % SLOTS;
            uint8_t* s = ring + slot * SLOT_BYTES;
            const int col = t.nt * BN_MAX + cc * 32;
            const int row = t.g * p.N + t.mt * BM + q * 32;
            mbar_arrive_expect_tx(&full[slot], n_state * STATE_TILE);
            tma_load_2d(s, &tmP, &full[slot], col, row);
            tma_load_2d(s + STATE_TILE, &tmM, &full[slot], col, row);
            tma_load_2d(s + 2 * STATE_TILE, &tmV, &full[slot], col, row);
            if (p.amsgrad) tma_load_2d(s + 3 * STATE_TILE, &tmVM, &full[slot], col, row);
        };
        for (int n = 0; n < SLOTS - 1; ++n)
            if (have_chunk(n) && lane == 0) issue_load(n);
        float step_size = 0.f, inv_sqrt_bc2 = 0.f;
        for (int c = 0; have_chunk(c); ++c) {
            const Tile t = fifo[(c / CHUNKS) & 1];
            const int cc = c % CHUNKS, slot = c % SLOTS, iter = c / CHUNKS, as = iter & 1;
            const uint32_t sphase = (c / SLOTS) & 1;
            if (cc == 0) {
                const float st = static_cast<float>(__ldg(p.step + t.g));
                step_size = p.lr / (1.f - powf(p.beta1, st));
                inv_sqrt_bc2 = rsqrtf(1.f - powf(p.beta2, st));
                mbar_wait(&tmem_full[as], (iter >> 1) & 1);
                tcgen05_fence_after();
            }
            // slot (c-1) % SLOTS is refilled with chunk c + SLOTS - 1: its previous contents (chunk c-1) must have been read by
            // the TMA store engine
            if (have_chunk(c + SLOTS - 1) && lane == 0) {
                tma_store_wait_read<0>();
                issue_load(c + SLOTS - 1);
            }
            mbar_wait(&full[slot], sphase);
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN_MAX + cc * 32, r);
            tmem_ld_wait();
            uint8_t* s = ring + slot * SLOT_BYTES + lane * 128;
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int o = (j ^ (lane & 7)) << 4;
                float4 pw = *reinterpret_cast<float4*>(s + o);
                float4 m = *reinterpret_cast<float4*>(s + STATE_TILE + o);
                float4 v = *reinterpret_cast<float4*>(s + 2 * STATE_TILE + o);
                float4 vm = p.amsgrad ? *reinterpret_cast<float4*>(s + 3 * STATE_TILE + o) : make_float4(0.f, 0.f, 0.f, 0.f);
                float* pp = &pw.x; float* mp = &m.x; float* vp = &v.x; float* vmp = &vm.x;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float grad = __uint_as_float(r[4 * j + e]);
                    mp[e] = mp[e] + (1.f - p.beta1) * (grad - mp[e]);
                    vp[e] = vp[e] * p.beta2 + (1.f - p.beta2) * grad * grad;
                    float denom;
                    if (p.amsgrad) {
                        vmp[e] = fmaxf(vmp[e], vp[e]);
                        denom = sqrtf(vmp[e]) * inv_sqrt_bc2 + p.eps;
                    } else {
                        denom = sqrtf(vp[e]) * inv_sqrt_bc2 + p.eps;
                    }
                    pp[e] -= step_size * (mp[e] / denom);
                }
                *reinterpret_cast<float4*>(s + o) = pw;
                *reinterpret_cast<float4*>(s + STATE_TILE + o) = m;
                *reinterpret_cast<float4*>(s + 2 * STATE_TILE + o) = v;
                if (p.amsgrad) *reinterpret_cast<float4*>(s + 3 * STATE_TILE + o) = vm;
                packed[2 * j] = pack_bf16x2(pw.x, pw.y);
                packed[2 * j + 1] = pack_bf16x2(pw.z, pw.w);
            }
            {   // bf16 mirror: 64 B per thread straight from registers
                const long long row = static_cast<long long>(t.g) * p.N + t.mt * BM + q * 32 + lane;
                int4* dst = reinterpret_cast<int4*>(p.p_bf16 + row * p.K + t.nt * BN_MAX + cc * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dst[j] = make_int4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                uint8_t* sb = ring + slot * SLOT_BYTES;
                const int col = t.nt * BN_MAX + cc * 32;
                const int row = t.g * p.N + t.mt * BM + q * 32;
                tma_store_2d(&tmP, sb, col, row);
                tma_store_2d(&tmM, sb + STATE_TILE, col, row);
                tma_store_2d(&tmV, sb + 2 * STATE_TILE, col, row);
                if (p.amsgrad) tma_store_2d(&tmVM, sb + 3 * STATE_TILE, col, row);
                tma_store_commit();
            }
            if (cc == CHUNKS - 1) {   // accumulator fully drained
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[as]);
            }
        }
        if (lane == 0) tma_store_wait<0>();   // all state tiles are in global memory before the CTA exits
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 2 * BN_MAX);
    }
}

}  // namespace wa
}  // namespace smallm
}  // namespace lah

using namespace lah;
using namespace lah::smallm;

extern "C" const int* lah_get_epoch_base();

// work-stealing tile counter shared by the launches of this file (stream-ordered: memset -> kernel); allocated on first use
static int* tile_counter() {
    static int* ctr = nullptr;
    if (!ctr) {
        if (cudaMalloc(&ctr, 256) != cudaSuccess) return nullptr;
        cudaMemset(ctr, 0, 256);
    }
    return ctr;
}
static const int* g_poison = nullptr;

extern "C" {

// status word whose bit 0 (a peer-flag wait timed out in this step) turns the optimizer kernels into no-ops; NULL disables
int lah_set_poison_word(const int* status) {
    g_poison = status;
    return 0;
}
const int* lah_get_poison_word() { return g_poison; }

// out[row, :] = act_rows[row, :] @ W[g]^T (+bias[g]) (+residual[row, :]) for the rows of every group, swap-AB tiles.
//   a_mn == 0: W is [G, M_out, K] (forward);  a_mn == 1: W is [G, K, M_out] and the product is x @ W (dgrad)
int lah_swapab_linear(const void* x, long long ldx, int x_rows, const void* W, int G, int M_out, int K, int a_mn,
                      void* out, long long ldo, const int* group_off, const int* group_rows, const float* bias,
                      const void* residual, long long ldr, const int* wait_flags, int wait_count, int wait_epoch,
                      const int* epoch_base, int* status, int max_ctas, cudaStream_t st) {
    if ((K % BK) || (M_out % BM) || (ldx % 8) || G > sab::MAX_G) return -2;
    CUtensorMap tmA, tmB[4];
    if (!a_mn) {
        uint64_t dims[3] = {(uint64_t)K, (uint64_t)M_out, (uint64_t)G};
        uint64_t str[2] = {(uint64_t)K * 2, (uint64_t)M_out * K * 2};
        uint32_t box[3] = {BK, BM, 1};
        int r = make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, 3, dims, str, box);
        if (r) return r;
    } else {
        uint64_t dims[3] = {(uint64_t)M_out, (uint64_t)K, (uint64_t)G};
        uint64_t str[2] = {(uint64_t)M_out * 2, (uint64_t)M_out * K * 2};
        uint32_t box[3] = {64, BK, 1};
        int r = make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, 3, dims, str, box);
        if (r) return r;
    }
    const uint32_t boxes[4] = {16, 32, 64, 128};
    for (int i = 0; i < 4; ++i) {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)x_rows};
        uint64_t str[1] = {(uint64_t)ldx * 2};
        uint32_t box[2] = {BK, boxes[i]};
        int r = make_tmap(&tmB[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, 2, dims, str, box);
        if (r) return r;
    }
    sab::Params p;
    p.G = G; p.M_out = M_out; p.K = K; p.group_off = group_off; p.group_rows = group_rows; p.out = (bf16*)out; p.ldo = ldo;
    p.bias = bias; p.residual = (const bf16*)residual; p.ldr = ldr; p.wait_flags = wait_flags; p.wait_count = wait_count;
    p.wait_epoch = wait_epoch; p.epoch_base = epoch_base ? epoch_base : lah_get_epoch_base(); p.status = status;
    p.tile_counter = tile_counter();
    if (!p.tile_counter) return -3;
    // upper bound of the tile count (the real one depends on the device-side row counts): every group has at least one
    // 128-token chunk per 128-feature slice, and all groups together at most x_rows / 128 + G chunks
    const long long total = 1ll * (M_out / BM) * (G + x_rows / BN_MAX);
    if (total <= 0 || G <= 0) return 0;
    if (cudaMemsetAsync(p.tile_counter, 0, sizeof(int), st) != cudaSuccess) return -4;
    int ctas = num_sms();
    if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
    if (total < ctas) ctas = static_cast<int>(total);
    static bool configured[2] = {false, false};
    if (!a_mn) {
        if (!configured[0]) {
            cudaError_t e = cudaFuncSetAttribute(sab::swapab_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sab::SMEM_TOTAL);
            if (e != cudaSuccess) return -(int)e;
            configured[0] = true;
        }
        sab::swapab_kernel<false><<<ctas, NUM_THREADS, sab::SMEM_TOTAL, st>>>(p, tmA, tmB[0], tmB[1], tmB[2], tmB[3]);
    } else {
        if (!configured[1]) {
            cudaError_t e = cudaFuncSetAttribute(sab::swapab_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sab::SMEM_TOTAL);
            if (e != cudaSuccess) return -(int)e;
            configured[1] = true;
        }
        sab::swapab_kernel<true><<<ctas, NUM_THREADS, sab::SMEM_TOTAL, st>>>(p, tmA, tmB[0], tmB[1], tmB[2], tmB[3]);
    }
    return -(int)cudaGetLastError();
}

// W[g] -= AMSGrad(dW[g] = dy_g^T x_g) for every group with rows > 0; p / m / v / vmax are [G, N, K] fp32, p_bf16 the mirror
int lah_wgrad_adam(const void* dy, long long lddy, const void* x, long long ldx, int total_rows, int G, int N, int K,
                   const int* group_off, const int* group_rows, const int* skip, const int* step, float* p, float* m,
                   float* v, float* vmax, void* p_bf16, float lr, float beta1, float beta2, float eps, int amsgrad,
                   int max_ctas, cudaStream_t st) {
    if ((N % BM) || (K % BN_MAX) || (lddy % 8) || (ldx % 8)) return -2;
    CUtensorMap tmDY, tmX, tmS[4];
    {
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)total_rows};
        uint64_t str[1] = {(uint64_t)lddy * 2};
        uint32_t box[2] = {64, wa::WBK};
        int r = make_tmap(&tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dy, 2, dims, str, box);
        if (r) return r;
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)total_rows};
        uint64_t str[1] = {(uint64_t)ldx * 2};
        uint32_t box[2] = {64, wa::WBK};
        int r = make_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, 2, dims, str, box);
        if (r) return r;
    }
    float* states[4] = {p, m, v, amsgrad ? vmax : p};
    for (int i = 0; i < 4; ++i) {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)G * N};
        uint64_t str[1] = {(uint64_t)K * 4};
        uint32_t box[2] = {32, 32};
        int r = make_tmap(&tmS[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, states[i], 2, dims, str, box);
        if (r) return r;
    }
    wa::Params a;
    a.G = G; a.N = N; a.K = K; a.group_off = group_off; a.group_rows = group_rows; a.skip = skip; a.step = step;
    a.p_bf16 = (bf16*)p_bf16; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.amsgrad = amsgrad;
    a.tile_counter = tile_counter() ? tile_counter() + 16 : nullptr;   // own word (64 B apart from the GEMM's)
    a.poison = g_poison;
    if (!a.tile_counter) return -3;
    const long long total = 1ll * G * (N / BM) * (K / BN_MAX);
    if (total <= 0) return 0;
    if (cudaMemsetAsync(a.tile_counter, 0, sizeof(int), st) != cudaSuccess) return -4;
    int ctas = num_sms();
    if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
    if (total < ctas) ctas = (int)total;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wa::wgrad_adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wa::SMEM_TOTAL);
        if (e != cudaSuccess) return -(int)e;
        configured = true;
    }
    wa::wgrad_adam_kernel<<<ctas, NUM_THREADS, wa::SMEM_TOTAL, st>>>(a, tmDY, tmX, tmS[0], tmS[1], tmS[2], tmS[3]);
    return -(int)cudaGetLastError();
}

}  // extern "C"
