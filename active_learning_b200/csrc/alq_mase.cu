// K6: MASE / BASE -- distance of every pool embedding to the pairwise decision
// boundaries of the linear head, and the class-balanced selection built on it.
//
// Reference semantics (under /root/reference/src/query_strategies): mase_sampler.py:52-80 broadcasts
// (B, C, M) tensors per loader batch to get  radius[i, c] = | -(w_p - w_c) * lam / 2 |,
// lam = 2 (h_i.(w_p - w_c) + b_p - b_c) / |w_p - w_c|^2,  p = argmax_c z_i.  With the algebra carried out,
// h_i.(w_p - w_c) + b_p - b_c is the logit gap z_ip - z_ic and the radius is |z_ip - z_ic| / |w_p - w_c|:
// a C x C table of head geometry (computed once per query) and one streaming pass over the logits slab, the same
// [N, C] input K1 reads.  NaN (c == p, duplicated class rows: 0/0, x/0 * 0) becomes +inf like :77.
// base_sampler.py:22-38 then selects class by class; alq_base_select runs that loop on the stream with K1b.
#include <limits.h>

#include "alq_common.cuh"
#include "alq_mase_rows.cuh"

namespace {

constexpr int kMaseThreads = 256;
constexpr int kGapTile = 64;
constexpr int kGapK = 32;
constexpr int kGapLd = kGapK + 4;     // row pitch 36 floats: 16-byte aligned, 128-bit loads of 16 consecutive rows hit 8 bank groups x 2

// ginv[a, c] = 1 / |w_a - w_c|  (IEEE sqrt and division; +inf where the rows coincide, incl. the diagonal).
// den is the direct sum of squared differences like mase_sampler.py:71 (not the |a|^2+|c|^2-2ac expansion, which
// cancels for nearby class rows).  64 x 64 output tile per CTA, 4 x 4 per thread (rows ty + 16 i, columns tx + 16 j:
// conflict-free 128-bit shared loads along k), operand chunks of 32 k fetched with coalesced 128-bit loads one chunk
// ahead of the math.  The table is symmetric: only tiles on or above the diagonal are computed, and mirrored on the way out.
__device__ __forceinline__ float4 gap_load4(const float* __restrict__ w, int row, int c, int k, int m, int64_t ldw, bool vec) {
    if (row >= c || k >= m) return make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p = w + static_cast<int64_t>(row) * ldw + k;
    if (vec && k + 3 < m) return *reinterpret_cast<const float4*>(p);
    float4 r = make_float4(p[0], 0.f, 0.f, 0.f);
    if (k + 1 < m) r.y = p[1];
    if (k + 2 < m) r.z = p[2];
    if (k + 3 < m) r.w = p[3];
    return r;
}

__global__ void __launch_bounds__(256)
class_gap_inv_kernel(const float* __restrict__ w, int c, int m, int64_t ldw, int vec, float* __restrict__ ginv, int64_t ldg) {
    // linear block index -> tile (by, bx) with bx >= by: row by holds tiles - by entries
    int by = 0, rem = blockIdx.x;
    const int tiles = (c + kGapTile - 1) / kGapTile;
    while (rem >= tiles - by) { rem -= tiles - by; ++by; }
    const int bx = by + rem;
    __shared__ __align__(16) float sa[kGapTile][kGapLd];
    __shared__ __align__(16) float sb[kGapTile][kGapLd];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int a0 = by * kGapTile, c0 = bx * kGapTile;
    // fill mapping: 512 float4 per operand chunk, two per thread; 8 consecutive threads cover one row's 128 bytes
    const int fr = threadIdx.x >> 3, fq = (threadIdx.x & 7) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float4 pa[2], pb[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        pa[h] = gap_load4(w, a0 + fr + 32 * h, c, fq, m, ldw, vec);
        pb[h] = gap_load4(w, c0 + fr + 32 * h, c, fq, m, ldw, vec);
    }
    for (int k0 = 0; k0 < m; k0 += kGapK) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            *reinterpret_cast<float4*>(&sa[fr + 32 * h][fq]) = pa[h];
            *reinterpret_cast<float4*>(&sb[fr + 32 * h][fq]) = pb[h];
        }
        __syncthreads();
        if (k0 + kGapK < m) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                pa[h] = gap_load4(w, a0 + fr + 32 * h, c, k0 + kGapK + fq, m, ldw, vec);
                pb[h] = gap_load4(w, c0 + fr + 32 * h, c, k0 + kGapK + fq, m, ldw, vec);
            }
        }
#pragma unroll
        for (int k4 = 0; k4 < kGapK; k4 += 4) {
            float4 av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                av[i] = *reinterpret_cast<const float4*>(&sa[ty + 16 * i][k4]);
                bv[i] = *reinterpret_cast<const float4*>(&sb[tx + 16 * i][k4]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float d = av[i].x - bv[j].x;
                    acc[i][j] = fmaf(d, d, acc[i][j]);
                    d = av[i].y - bv[j].y;
                    acc[i][j] = fmaf(d, d, acc[i][j]);
                    d = av[i].z - bv[j].z;
                    acc[i][j] = fmaf(d, d, acc[i][j]);
                    d = av[i].w - bv[j].w;
                    acc[i][j] = fmaf(d, d, acc[i][j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int a = a0 + ty + 16 * i, cc = c0 + tx + 16 * j;
            if (a < c && cc < c) {
                const float g = (a != cc) ? __fdiv_rn(1.0f, __fsqrt_rn(acc[i][j])) : ALQ_POS_INF;
                ginv[static_cast<int64_t>(a) * ldg + cc] = g;
                ginv[static_cast<int64_t>(cc) * ldg + a] = g;
            }
        }
}

// gmin[a] = min_c ginv[a, c]: the reciprocal of the largest distance from class a to any other class row, and
// gmin[C] = max_a (largest finite ginv[a, :]) / gmin[a], how unevenly the class rows are spread (>= 1; the pruning
// bound of mase_row_min_smem).  Also fills the pad columns c..ldg-1 of row a with +inf.  One warp per class row.
__global__ void __launch_bounds__(256)
class_gap_rowmin_kernel(float* __restrict__ ginv, int c, int64_t ldg, float* __restrict__ gmin) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= c) return;
    float* g = ginv + static_cast<int64_t>(row) * ldg;
    float mn = ALQ_POS_INF, mx = 0.f;
    for (int j = lane; j < c; j += 32) {
        const float v = g[j];
        mn = fminf(mn, v);
        if (v < ALQ_POS_INF) mx = fmaxf(mx, v);
    }
    for (int j = c + lane; j < ldg; j += 32) g[j] = ALQ_POS_INF;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0 && gmin) {
        gmin[row] = mn;
        const float ratio = __fdiv_ru(mx, mn);            // mn = 0 or inf, mx = 0: inf / NaN / 0 -> handled below
        if (ratio >= 1.0f)                                // positive floats order like their bit patterns
            atomicMax(reinterpret_cast<int*>(gmin + c), ratio < ALQ_POS_INF ? __float_as_int(ratio) : 0x7f800000);
    }
}

__global__ void gap_ratio_init_kernel(float* __restrict__ gmin, int c) { gmin[c] = 1.0f; }

// One warp per row, the row held as NV float4 per lane between the arg-max pass and the radius pass
// (row code: alq_mase_rows.cuh).
template <int NV>
__device__ __forceinline__ void load_row_regs(const float* logits, int64_t row, int64_t ld, int lane, int nvec,
                                              MaseRowRegs<NV>& r) {
    const float4* p = reinterpret_cast<const float4*>(logits + row * ld);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int idx = lane + 32 * k;
        r.v[k] = idx < nvec ? ld_stream_f4(p + idx) : make_float4(ALQ_NEG_INF, ALQ_NEG_INF, ALQ_NEG_INF, ALQ_NEG_INF);
    }
}

template <int NV, bool WRITE_R>
__global__ void __launch_bounds__(kMaseThreads)
mase_rows_vec_kernel(const float* __restrict__ logits, int64_t n, int c, int64_t ld, const float* __restrict__ ginv,
                     int64_t ldg, float* __restrict__ minm, int32_t* __restrict__ pred, float* __restrict__ radius,
                     int64_t ldr) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int nvec = c >> 2;
    for (int64_t row = warp; row < n; row += nwarps) {
        MaseRowRegs<NV> r;
        load_row_regs<NV>(logits, row, ld, lane, nvec, r);
        float mn;
        int arg;
        mase_row_full<NV, WRITE_R>(r, lane, nvec, ginv, ldg, WRITE_R ? reinterpret_cast<float4*>(radius + row * ldr) : nullptr,
                                   mn, arg);
        if (lane == 0) { minm[row] = mn; pred[row] = arg; }
    }
}

template <int NV>
__global__ void __launch_bounds__(kMaseThreads)
mase_rows_min_kernel(const float* __restrict__ logits, int64_t n, int c, int64_t ld, const float* __restrict__ ginv,
                     int64_t ldg, const float* __restrict__ gmin, float* __restrict__ minm, int32_t* __restrict__ pred) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int nvec = c >> 2;
    for (int64_t row = warp; row < n; row += nwarps) {
        MaseRowRegs<NV> r;
        load_row_regs<NV>(logits, row, ld, lane, nvec, r);
        float mn;
        int arg;
        mase_row_min<NV>(r, lane, c, ginv, ldg, gmin, mn, arg);
        if (lane == 0) { minm[row] = mn; pred[row] = arg; }
    }
}

// Any c / alignment: two passes over the row, the second one hits L1/L2.
__global__ void __launch_bounds__(kMaseThreads)
mase_rows_generic_kernel(const float* __restrict__ logits, int64_t n, int c, int64_t ld, const float* __restrict__ ginv,
                         int64_t ldg, float* __restrict__ minm, int32_t* __restrict__ pred, float* __restrict__ radius,
                         int64_t ldr) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = warp; row < n; row += nwarps) {
        const float* p = logits + row * ld;
        float best = ALQ_NEG_INF;
        int arg = INT_MAX;
        for (int j = lane; j < c; j += 32) {
            const float z = p[j];
            if (z > best) { best = z; arg = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            mase_argmax_merge(best, arg, ob, oa);
        }
        if (arg == INT_MAX) arg = 0;
        const float* gi = ginv + static_cast<int64_t>(arg) * ldg;
        float mn = ALQ_POS_INF;
        for (int j = lane; j < c; j += 32) {
            const float r = mase_radius_of(best, p[j], __ldg(gi + j), j == arg);
            mn = fminf(mn, r);
            if (radius) radius[row * ldr + j] = r;
        }
        if (radius)
            for (int j = c + lane; j < ldr; j += 32) radius[row * ldr + j] = ALQ_POS_INF;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if (lane == 0) { minm[row] = mn; pred[row] = arg; }
    }
}

// base_sampler.py:29-33: the key class `cls` sorts by; rows taken by earlier classes are pushed to +inf.
__global__ void __launch_bounds__(256)
base_keys_kernel(const float* __restrict__ minm, const float* __restrict__ radius, int64_t ldr,
                 const int32_t* __restrict__ pred, const unsigned char* __restrict__ taken, int64_t n, int cls,
                 float* __restrict__ keys) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float k = (pred[i] == cls) ? minm[i] : radius[i * ldr + cls];
        keys[i] = taken[i] ? ALQ_POS_INF : k;
    }
}

// marks this class's picks; a row picked twice (only possible once every free key is +inf) is what the
// reference's `assert len(labeled_idxs) == len(set(labeled_idxs))` (:40) trips on -> counted in dup[0].
__global__ void base_mark_kernel(const int32_t* __restrict__ picks, int cnt, unsigned char* __restrict__ taken,
                                 unsigned int* __restrict__ dup) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) {
        const int32_t r = picks[i];
        if (taken[r]) atomicAdd(dup, 1u);
        taken[r] = 1;
    }
}

int rows_grid(const alq_ctx* ctx, int64_t n, int warps_per_block) {
    int64_t need = (n + warps_per_block - 1) / warps_per_block;
    const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 8;
    if (need < 1) need = 1;
    return static_cast<int>(need < cap ? need : cap);
}

template <int NV>
void launch_mase_vec(int grid, cudaStream_t st, const float* logits, int64_t n, int c, int64_t ld, const float* ginv,
                     int64_t ldg, float* minm, int32_t* pred, float* radius, int64_t ldr) {
    if (radius)
        mase_rows_vec_kernel<NV, true><<<grid, kMaseThreads, 0, st>>>(logits, n, c, ld, ginv, ldg, minm, pred, radius, ldr);
    else
        mase_rows_vec_kernel<NV, false><<<grid, kMaseThreads, 0, st>>>(logits, n, c, ld, ginv, ldg, minm, pred, radius, ldr);
}

template <int NV>
void launch_mase_min(int grid, cudaStream_t st, const float* logits, int64_t n, int c, int64_t ld, const float* ginv,
                     int64_t ldg, const float* gmin, float* minm, int32_t* pred) {
    mase_rows_min_kernel<NV><<<grid, kMaseThreads, 0, st>>>(logits, n, c, ld, ginv, ldg, gmin, minm, pred);
}

}  // namespace

extern "C" int alq_class_gap_inv(alq_ctx* ctx, const float* weight, int32_t c, int32_t m, int64_t ldw, float* ginv,
                                 int64_t ldg, float* gmin, void* stream) {
    if (!ctx) return ALQ_ERR_INVALID;
    if (c <= 0 || m <= 0 || ldw < m || ldg < c) ALQ_FAIL(ctx, ALQ_ERR_INVALID, "alq_class_gap_inv: bad shape (c=%d m=%d ldw=%lld ldg=%lld)", c, m, (long long)ldw, (long long)ldg);
    if (!weight || !ginv) ALQ_FAIL(ctx, ALQ_ERR_INVALID, "alq_class_gap_inv: null pointer");
    const unsigned tiles = static_cast<unsigned>((c + kGapTile - 1) / kGapTile);
    class_gap_inv_kernel<<<tiles * (tiles + 1) / 2, 256, 0, static_cast<cudaStream_t>(stream)>>>(weight, c, m, ldw, (ldw % 4 == 0) && aligned16(weight), ginv, ldg);
    ALQ_LAUNCH_CHECK(ctx);
    if (gmin) {
        gap_ratio_init_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(gmin, c);
        ALQ_LAUNCH_CHECK(ctx);
    }
    class_gap_rowmin_kernel<<<(c + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(ginv, c, ldg, gmin);
    ALQ_LAUNCH_CHECK(ctx);
    return ALQ_OK;
}

extern "C" int alq_mase_margins(alq_ctx* ctx, const float* logits, int64_t n, int32_t c, int64_t ld, const float* ginv,
                                int64_t ldg, const float* gmin, float* min_margin, int32_t* pred, float* radius, int64_t ldr,
                                void* stream) {
    if (!ctx) return ALQ_ERR_INVALID;
    if (n < 0 || c <= 0 || ld < c || ldg < c || (radius && ldr < c)) ALQ_FAIL(ctx, ALQ_ERR_INVALID, "alq_mase_margins: bad shape");
    if (n >= (1LL << 31)) ALQ_FAIL(ctx, ALQ_ERR_INVALID, "alq_mase_margins: n must be < 2^31");
    if (n == 0) return ALQ_OK;
    if (!logits || !ginv || !min_margin || !pred) ALQ_FAIL(ctx, ALQ_ERR_INVALID, "alq_mase_margins: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = rows_grid(ctx, n, kMaseThreads / 32);
    const bool vec = (c % 4 == 0) && (ld % 4 == 0) && (ldg % 4 == 0) && aligned16(logits) && aligned16(ginv) &&
                     (!radius || ((ldr % 4 == 0) && ldr == c && aligned16(radius))) && c <= 1024;
    if (vec && ld == c && (radius ? ldr == c : gmin != nullptr) && n >= 4096 && ctx->greedy_variant != 1) {
        // bulk-copy pipelined kernel (the K1 / K2 pipeline with the K6 row code)
        const MaseArgs margs{ginv, ldg, gmin, pred, radius, ldr};
        cudaError_t e = cudaSuccess;
        if (alq_mase_rows_pipe(ctx, st, logits, n, c, margs, min_margin, &e)) {
            ctx->launches++;
            if (e != cudaSuccess) ALQ_FAIL(ctx, ALQ_ERR_CUDA, "rows_pipe_kernel (K6) launch failed: %s", cudaGetErrorString(e));
            return ALQ_OK;
        }
    }
    if (vec && !radius && gmin) {       // MASE: exact minimum with table reads pruned by gmin
        const int nv = (c / 4 + 31) / 32;
        switch (nv) {
            case 1: launch_mase_min<1>(grid, st, logits, n, c, ld, ginv, ldg, gmin, min_margin, pred); break;
            case 2: launch_mase_min<2>(grid, st, logits, n, c, ld, ginv, ldg, gmin, min_margin, pred); break;
            case 3: case 4: launch_mase_min<4>(grid, st, logits, n, c, ld, ginv, ldg, gmin, min_margin, pred); break;
            default: launch_mase_min<8>(grid, st, logits, n, c, ld, ginv, ldg, gmin, min_margin, pred); break;
        }
    } else if (vec) {
        const int nv = (c / 4 + 31) / 32;
        switch (nv) {
            case 1: launch_mase_vec<1>(grid, st, logits, n, c, ld, ginv, ldg, min_margin, pred, radius, ldr); break;
            case 2: launch_mase_vec<2>(grid, st, logits, n, c, ld, ginv, ldg, min_margin, pred, radius, ldr); break;
            case 3: case 4: launch_mase_vec<4>(grid, st, logits, n, c, ld, ginv, ldg, min_margin, pred, radius, ldr); break;
            default: launch_mase_vec<8>(grid, st, logits, n, c, ld, ginv, ldg, min_margin, pred, radius, ldr); break;
        }
    } else {
        mase_rows_generic_kernel<<<grid, kMaseThreads, 0, st>>>(logits, n, c, ld, ginv, ldg, min_margin, pred, radius, ldr);
    }
    ALQ_LAUNCH_CHECK(ctx);
    return ALQ_OK;
}

extern "C" int alq_base_select(alq_ctx* ctx, const float* min_margin, const float* radius, int64_t ldr, const int32_t* pred,
                               int64_t n, int32_t c, int64_t budget, int32_t* out_pos, void* stream) {
    if (!ctx) return ALQ_ERR_INVALID;
    if (n < 0 || c <= 0 || ldr < c || budget < 0 || budget > n || n >= (1LL << 31))
        ALQ_FAIL(ctx, ALQ_ERR_INVALID, "alq_base_select: need 0 <= budget <= n < 2^31, ldr >= c (n=%lld budget=%lld)", (long long)n, (long long)budget);
    if (budget == 0) return ALQ_OK;
    if (!min_margin || !radius || !pred || !out_pos) ALQ_FAIL(ctx, ALQ_ERR_INVALID, "alq_base_select: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // private arena: alq_select_smallest re-carves the shared scratch on every call
    int rc = alq_arena2_reserve(ctx, scratch_need({static_cast<size_t>(n) * sizeof(float), static_cast<size_t>(n), sizeof(unsigned int)}));
    if (rc) return rc;
    ScratchCursor cur(ctx->arena2);
    float* keys = cur.take<float>(n);
    unsigned char* taken = cur.take<unsigned char>(n);
    unsigned int* dup = cur.take<unsigned int>(1);
    ALQ_CUDA(ctx, cudaMemsetAsync(taken, 0, static_cast<size_t>(n), st));
    ALQ_CUDA(ctx, cudaMemsetAsync(dup, 0, sizeof(unsigned int), st));
    int kgrid = static_cast<int>((n + 255) / 256);
    if (kgrid > ctx->sm_count * 8) kgrid = ctx->sm_count * 8;
    int64_t at = 0;
    for (int cls = 0; cls < c; ++cls) {
        const int64_t cnt = budget / c + (cls < budget % c ? 1 : 0);      // base_sampler.py:24-25
        if (cnt == 0) continue;
        base_keys_kernel<<<kgrid, 256, 0, st>>>(min_margin, radius, ldr, pred, taken, n, cls, keys);
        ALQ_LAUNCH_CHECK(ctx);
        rc = alq_select_smallest(ctx, keys, n, cnt, out_pos + at, stream);
        if (rc) return rc;
        base_mark_kernel<<<static_cast<int>((cnt + 255) / 256), 256, 0, st>>>(out_pos + at, static_cast<int>(cnt), taken, dup);
        ALQ_LAUNCH_CHECK(ctx);
        at += cnt;
    }
    unsigned int dup_host = 0;
    ALQ_CUDA(ctx, cudaMemcpyAsync(&dup_host, dup, sizeof(dup_host), cudaMemcpyDeviceToHost, st));
    ALQ_CUDA(ctx, cudaStreamSynchronize(st));
    if (dup_host) ALQ_FAIL(ctx, ALQ_ERR_NUMERIC, "alq_base_select: %u rows were selected twice (every free key is +inf; base_sampler.py:40 asserts here)", dup_host);
    return ALQ_OK;
}
