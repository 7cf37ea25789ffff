// K3 on the 5th-generation tensor cores: out[i] = min_j (max_j) fl(fl(xn_i + yn_j) - 2 <x_i, y_j>)
// for fp32 rows, as a tcgen05 "3xTF32" contraction with the min-epilogue fused on the TMEM
// accumulators.  Same contract as the SIMT kernel in alq_mindist.cu (coreset_sampler.py:59-64,79,100
// under /root/reference/src/query_strategies); this is the one dense contraction on the query path.
//
// fp32 fidelity.  kind::tf32 keeps 10 mantissa bits, so each operand is split once, ahead of time,
// into two valid TF32 numbers  x = hi + lo + r,  hi = rna_tf32(x), lo = rna_tf32(x - hi), |r| <= 2^-22|x|,
// and every k-step issues three MMAs into the same fp32 accumulator:  hi*hi' + hi*lo' + lo*hi'.
// The dropped lo*lo' and r terms are <= 2^-21 |x||y| per product.  Small integers (the exact-arithmetic
// parity fixtures) have lo == 0 and give bit-identical results to fp32.
//
// Structure (one CTA per SM, persistent over a host-built work list):
//   warp 0   TMA producer: 4 tensor maps (X_hi, X_lo, Y_hi, Y_lo), boxes of 32 fp32 (=128 B, SWIZZLE_128B)
//            x 128 / 256 rows into a 2-stage shared-memory ring, mbarrier complete_tx.
//   warp 1   MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::tf32, M=128, N=256, K=8,
//            12 MMAs per stage, tcgen05.commit frees the stage / publishes the accumulator.
//   warps 2-5 epilogue: tcgen05.ld 32x32b.x32 (one accumulator row per thread), distance + running
//            row min in registers across all column tiles of a work item, one float atomic per row at the end.
//   TMEM: 512 columns = two 128x256 fp32 accumulators (double buffered; the factored/BADGE form uses
//   them as the <a,a'> and <h,h'> accumulators of one tile and multiplies them in the epilogue).
// Every CTA sweeps the column tiles in the same order, so a Y tile is fetched from HBM once and
// served to the other 147 CTAs from L2.
#include <cuda.h>
#include <stdlib.h>

#include "alq_common.cuh"

namespace tc {

constexpr int BM = 128;          // accumulator rows  (UMMA M)
constexpr int BN = 256;          // accumulator cols  (UMMA N)
constexpr int BK = 32;           // fp32 per stage row = 128 bytes = one swizzle span
constexpr int UK = 8;            // UMMA K for tf32
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 4;              // 16 KB
constexpr int B_BYTES = BN * BK * 4;              // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // hi + lo of both operands = 96 KB
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;

struct WorkItem {
    int m_blk, n_begin, n_end, pad;   // row block, [first, last) column tile
};

struct Params {
    const float* xn; const float* yn;      // row norms (already multiplied by the a-norms when factored)
    float* out;
    const WorkItem* items;                 // [grid][items_per_cta]
    const int* item_count;                 // [grid]
    int items_per_cta;
    int n, m;                              // rows of X / Y
    int kblocks_h, kblocks_a;              // ceil(d / 32), ceil(c / 32) (0 when dense)
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (= 1, unused for swizzled K-major),
//   [32,46) stride byte offset >> 4 (= 1024 B: 8 rows x 128 B), [46,48) version = 1, [61,64) layout = 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format TF32 (2) @7/@10, K-major both, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
           (static_cast<uint32_t>(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <bool FACTORED, bool RED_MAX>
__global__ void __launch_bounds__(THREADS, 1)
min_dist_tc_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                   const __grid_constant__ CUtensorMap map_yh, const __grid_constant__ CUtensorMap map_yl,
                   const __grid_constant__ CUtensorMap map_xah, const __grid_constant__ CUtensorMap map_xal,
                   const __grid_constant__ CUtensorMap map_yah, const __grid_constant__ CUtensorMap map_yal,
                   Params P) {
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries: align by hand (1 KB of slack is allocated)
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* stage_base = smem_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* cn_s = reinterpret_cast<float*>(tmem_slot + 4);                  // [2][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = P.item_count[blockIdx.x];
    const WorkItem* items = P.items + static_cast<size_t>(blockIdx.x) * P.items_per_cta;
    const int kb_total = P.kblocks_h + (FACTORED ? P.kblocks_a : 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation is warp-collective; the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int w = 0; w < n_items; ++w) {
                const WorkItem wi = items[w];
                for (int nt = wi.n_begin; nt < wi.n_end; ++nt) {
                    for (int kb = 0; kb < kb_total; ++kb, ++it) {
                        const int s = it % STAGES;
                        const uint32_t round = it / STAGES;
                        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1u);
                        unsigned char* st = stage_base + static_cast<size_t>(s) * STAGE_BYTES;
                        mbar_expect_tx(&full[s], STAGE_BYTES);
                        const bool apart = FACTORED && kb >= P.kblocks_h;
                        const int k0 = (apart ? kb - P.kblocks_h : kb) * BK;
                        tma_load_2d(st, apart ? &map_xah : &map_xh, k0, wi.m_blk * BM, &full[s]);
                        tma_load_2d(st + A_BYTES, apart ? &map_xal : &map_xl, k0, wi.m_blk * BM, &full[s]);
                        tma_load_2d(st + 2 * A_BYTES, apart ? &map_yah : &map_yh, k0, nt * BN, &full[s]);
                        tma_load_2d(st + 2 * A_BYTES + B_BYTES, apart ? &map_yal : &map_yl, k0, nt * BN, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc();
            uint32_t it = 0, tile = 0;
            for (int w = 0; w < n_items; ++w) {
                const WorkItem wi = items[w];
                for (int nt = wi.n_begin; nt < wi.n_end; ++nt, ++tile) {
                    // dense: accumulator buffer alternates per tile; factored: buffer 0 = <h,h'>, 1 = <a,a'>
                    if (FACTORED) {
                        if (tile > 0) { mbar_wait(&tmem_empty[0], (tile - 1) & 1u); }
                    } else {
                        const uint32_t use = tile >> 1;
                        if (use > 0) mbar_wait(&tmem_empty[tile & 1], (use - 1) & 1u);
                    }
                    tc_fence_after();
                    for (int kb = 0; kb < kb_total; ++kb, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(&full[s], (it / STAGES) & 1u);
                        tc_fence_after();
                        const bool apart = FACTORED && kb >= P.kblocks_h;
                        const uint32_t buf = FACTORED ? (apart ? 1u : 0u) : (tile & 1u);
                        const uint32_t d_tmem = tmem_base + buf * BN;
                        const bool first_kb = apart ? (kb == P.kblocks_h) : (kb == 0);
                        const uint32_t a_hi = smem_u32(stage_base + static_cast<size_t>(s) * STAGE_BYTES);
                        const uint32_t a_lo = a_hi + A_BYTES, b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            const uint32_t off = k * UK * 4;    // bytes along K inside the 128-byte swizzle span
                            const uint64_t dah = make_desc(a_hi + off), dal = make_desc(a_lo + off);
                            const uint64_t dbh = make_desc(b_hi + off), dbl = make_desc(b_lo + off);
                            umma_tf32(d_tmem, dah, dbh, idesc, (first_kb && k == 0) ? 0u : 1u);
                            umma_tf32(d_tmem, dah, dbl, idesc, 1u);
                            umma_tf32(d_tmem, dal, dbh, idesc, 1u);
                        }
                        umma_commit(&empty[s]);              // stage reusable once these MMAs retire
                    }
                    if (FACTORED) { umma_commit(&tmem_full[0]); }
                    else umma_commit(&tmem_full[tile & 1]);
                }
            }
        }
    } else {
        // ================= epilogue: 4 warps, one accumulator row per thread =================
        const int q = warp & 3;                         // TMEM lane quadrant this warp may access
        const int row_in_tile = q * 32 + lane;
        const int et = threadIdx.x - 64;                // 0..127
        uint32_t tile = 0;
        for (int w = 0; w < n_items; ++w) {
            const WorkItem wi = items[w];
            const int row = wi.m_blk * BM + row_in_tile;
            const float rn = row < P.n ? P.xn[row] : 0.f;
            float best = RED_MAX ? ALQ_NEG_INF : ALQ_POS_INF;
            for (int nt = wi.n_begin; nt < wi.n_end; ++nt, ++tile) {
                float* cn = cn_s + (tile & 1) * BN;
                for (int j = et; j < BN; j += 128) {
                    const int col = nt * BN + j;
                    cn[j] = col < P.m ? P.yn[col] : (RED_MAX ? ALQ_NEG_INF : ALQ_POS_INF);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const uint32_t buf = FACTORED ? 0u : (tile & 1u);
                const uint32_t par = FACTORED ? (tile & 1u) : ((tile >> 1) & 1u);
                mbar_wait(&tmem_full[buf], par);
                tc_fence_after();
                const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld32(lane_addr + buf * BN + c0, acc);
                    if (FACTORED) {
                        uint32_t acc2[32];
                        tmem_ld32(lane_addr + BN + c0, acc2);
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) * __uint_as_float(acc2[j]));
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float dist = (rn + cn[c0 + j]) - 2.0f * __uint_as_float(acc[j]);
                        best = RED_MAX ? fmaxf(best, dist) : fminf(best, dist);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            }
            if (row < P.n && wi.n_end > wi.n_begin) {
                if (RED_MAX) atomic_max_float(P.out + row, best);
                else atomic_min_float(P.out + row, best);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// hi = rna_tf32(x), lo = rna_tf32(x - hi): two valid TF32 operands whose sum is x to 2^-22 relative
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, int64_t rows, int d, int64_t ld, float* __restrict__ hi,
                  float* __restrict__ lo) {
    const int64_t total = rows * (d >> 2);
    const int dv = d >> 2;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / dv;
        const int k = static_cast<int>(i - r * dv);
        const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(x + r * ld) + k);
        const float e[4] = {v.x, v.y, v.z, v.w};
        float h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t hb, lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(e[j]));
            h[j] = __uint_as_float(hb);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(e[j] - h[j]));
            l[j] = __uint_as_float(lb);
        }
        reinterpret_cast<float4*>(hi + r * d)[k] = make_float4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<float4*>(lo + r * d)[k] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

__global__ void mul_norms_kernel(const float* a, const float* b, float* o, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] * b[i];
}
__global__ void fill_kernel(float* p, int64_t n, float v) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    }
    return fn;
}

// 2-D fp32 [rows, cols] row-major (contiguous: pitch == cols), box = [32 cols, box_rows], 128-byte swizzle
bool make_map(CUtensorMap* map, const float* base, int64_t rows, int cols, int box_rows) {
    EncodeFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 4};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc

// Called by alq_min_dist (alq_mindist.cu).  Returns ALQ_OK, or ALQ_ERR_STATE if the tensor-core path
// cannot be used here (caller then runs the SIMT kernel); other codes are real failures.
int alq_min_dist_tc(alq_ctx* ctx, const float* x, int64_t ldx, const float* xn, int64_t n, const float* y,
                    int64_t ldy, const float* yn, int64_t m, int32_t d, const float* xa, int64_t ldxa,
                    const float* xan, const float* ya, int64_t ldya, const float* yan, int32_t c,
                    int32_t reduce_max, int32_t accumulate, float* out, cudaStream_t st) {
    using namespace tc;
    const bool factored = xa != nullptr;
    if (!get_encode()) return ALQ_ERR_STATE;
    if (n >= (1LL << 31) || m >= (1LL << 31)) return ALQ_ERR_STATE;
    const int cpad = factored ? c : 0;
    // scratch: split operands (contiguous rows) + fused norms + work list
    const size_t xe = static_cast<size_t>(n) * d, ye = static_cast<size_t>(m) * d;
    const size_t xae = static_cast<size_t>(n) * cpad, yae = static_cast<size_t>(m) * cpad;
    const int grid = ctx->sm_count;
    const int m_blocks = static_cast<int>((n + BM - 1) / BM), n_tiles = static_cast<int>((m + BN - 1) / BN);
    // Work list.  A CTA re-reads its row block (hi + lo = 2 MB at d = 2048) for every column tile, so the row
    // blocks that are live at the same time must fit L2 next to the streamed Y tiles: with one row block per CTA
    // (148 x 2 MB) they do not, and ncu showed 258 GB of DRAM reads for 5 GB of operands.  Instead `split` CTAs
    // share a row block and divide its column tiles, so only grid / split (~37) row blocks are live per round.
    const size_t blk_bytes = static_cast<size_t>(BM) * (d + cpad) * 8;
    size_t live_mb = 48;                                                                              // sweep: 32-48 MB best
    if (const char* e = getenv("ALQ_K3_LIVE_MB")) live_mb = static_cast<size_t>(std::max(1, atoi(e)));   // tuning aid
    int live = static_cast<int>(std::max<size_t>(1, (live_mb << 20) / std::max<size_t>(blk_bytes, 1)));
    live = std::max(1, std::min(live, grid));
    const int rounds = (m_blocks + live - 1) / live;
    const int items_per_cta = rounds + 1;
    int rc = alq_scratch_reserve(ctx, scratch_need({xe * 4, xe * 4, ye * 4, ye * 4, xae * 4, xae * 4, yae * 4, yae * 4,
                                                    static_cast<size_t>(n) * 4, static_cast<size_t>(m) * 4,
                                                    static_cast<size_t>(grid) * items_per_cta * sizeof(WorkItem),
                                                    static_cast<size_t>(grid) * 4}));
    if (rc) return rc;
    ScratchCursor cur(ctx->scratch);
    float* xh = cur.take<float>(xe); float* xl = cur.take<float>(xe);
    float* yh = cur.take<float>(ye); float* yl = cur.take<float>(ye);
    float* xah = cur.take<float>(xae); float* xal = cur.take<float>(xae);
    float* yah = cur.take<float>(yae); float* yal = cur.take<float>(yae);
    float* xnn = cur.take<float>(n); float* ynn = cur.take<float>(m);
    WorkItem* d_items = cur.take<WorkItem>(static_cast<size_t>(grid) * items_per_cta);
    int* d_counts = cur.take<int>(grid);

    // ---- work list (see above): round r covers row blocks [r*live, r*live + mbs); CTA c takes part c % split of
    //      the column tiles of row block c / split
    std::vector<WorkItem> items(static_cast<size_t>(grid) * items_per_cta);
    std::vector<int> counts(grid, 0);
    for (int r = 0; r < rounds; ++r) {
        const int mb0 = r * live, mbs = std::min(live, m_blocks - mb0);
        const int split = std::max(1, grid / mbs);
        for (int cta = 0; cta < grid; ++cta) {
            const int which = cta / split, part = cta % split;
            if (which >= mbs) continue;
            const int nb = static_cast<int>(static_cast<int64_t>(n_tiles) * part / split);
            const int ne = static_cast<int>(static_cast<int64_t>(n_tiles) * (part + 1) / split);
            if (ne > nb) items[static_cast<size_t>(cta) * items_per_cta + counts[cta]++] = WorkItem{mb0 + which, nb, ne, 0};
        }
    }
    ALQ_CUDA(ctx, cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, st));
    ALQ_CUDA(ctx, cudaMemcpyAsync(d_counts, counts.data(), counts.size() * sizeof(int), cudaMemcpyHostToDevice, st));

    // ---- split operands
    auto split = [&](const float* src, int64_t rows, int dd, int64_t ld, float* hi, float* lo) {
        if (rows == 0 || dd == 0) return;
        int blocks = static_cast<int>(std::min<int64_t>((rows * (dd / 4) + 255) / 256, static_cast<int64_t>(ctx->sm_count) * 16));
        split_tf32_kernel<<<blocks, 256, 0, st>>>(src, rows, dd, ld, hi, lo);
        ctx->launches++;
    };
    split(x, n, d, ldx, xh, xl);
    split(y, m, d, ldy, yh, yl);
    const float* xn_use = xn; const float* yn_use = yn;
    if (factored) {
        split(xa, n, c, ldxa, xah, xal);
        split(ya, m, c, ldya, yah, yal);
        mul_norms_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(xn, xan, xnn, n);
        mul_norms_kernel<<<static_cast<int>((m + 255) / 256), 256, 0, st>>>(yn, yan, ynn, m);
        ctx->launches += 2;
        xn_use = xnn; yn_use = ynn;
    }
    if (!accumulate) {
        fill_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(out, n, reduce_max ? -INFINITY : INFINITY);
        ctx->launches++;
    }
    CUtensorMap mxh, mxl, myh, myl, mxah, mxal, myah, myal;
    bool ok = make_map(&mxh, xh, n, d, BM) && make_map(&mxl, xl, n, d, BM) && make_map(&myh, yh, m, d, BN) &&
              make_map(&myl, yl, m, d, BN);
    if (factored)
        ok = ok && make_map(&mxah, xah, n, c, BM) && make_map(&mxal, xal, n, c, BM) && make_map(&myah, yah, m, c, BN) &&
             make_map(&myal, yal, m, c, BN);
    else { mxah = mxh; mxal = mxl; myah = myh; myal = myl; }
    if (!ok) ALQ_FAIL(ctx, ALQ_ERR_CUDA, "alq_min_dist: cuTensorMapEncodeTiled failed");

    Params P{};
    P.xn = xn_use; P.yn = yn_use; P.out = out; P.items = d_items; P.item_count = d_counts; P.items_per_cta = items_per_cta;
    P.n = static_cast<int>(n); P.m = static_cast<int>(m);
    P.kblocks_h = (d + BK - 1) / BK; P.kblocks_a = factored ? (c + BK - 1) / BK : 0;
    const size_t smem = static_cast<size_t>(STAGES) * STAGE_BYTES + 256 + 2 * BN * sizeof(float) + 1024;
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        kern<<<grid, THREADS, smem, st>>>(mxh, mxl, myh, myl, mxah, mxal, myah, myal, P);
        return cudaGetLastError();
    };
    cudaError_t e;
    if (factored) e = reduce_max ? launch(min_dist_tc_kernel<true, true>) : launch(min_dist_tc_kernel<true, false>);
    else e = reduce_max ? launch(min_dist_tc_kernel<false, true>) : launch(min_dist_tc_kernel<false, false>);
    ctx->launches++;
    if (e != cudaSuccess) ALQ_FAIL(ctx, ALQ_ERR_CUDA, "min_dist_tc_kernel launch failed: %s", cudaGetErrorString(e));
    return ALQ_OK;
}
