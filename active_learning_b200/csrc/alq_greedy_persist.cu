// K4 / K5 as ONE persistent cooperative launch: the whole B-step selection loop of
// `CoresetSampler.coreset` (coreset_sampler.py:66-105 under /root/reference/src/query_strategies) runs
// inside one kernel, one CTA per SM, on one GPU or on G GPUs that talk through peer-memory windows.
//
// Why: a step is a 100-150 us HBM stream on one GPU but only 12-19 us on eight, and the rows a step
// streams do not depend on the centre -- only the dot products do.  So the TMA producer of every CTA
// runs free across step boundaries: while the grid (and the other GPUs) agree on the next centre, the
// bulk copies of the next step's first tiles are already landing in the shared-memory ring (~200 KB per
// SM, ~30 MB per GPU in flight), and HBM never idles across the exchange.  With one launch per step that
// window was spent on launch latency, the prologue and a flag chain.
//
// Agreement on the centre never moves a row: every rank holds a replica of the candidate rows (a few
// hundred MB next to 180 GB), so a pick is announced as a row id.  All cross-CTA / cross-GPU traffic is
// "LL" words: 8 bytes = {32-bit tag, 32-bit payload}, written with one plain store (single-copy atomic, so
// no fence and no separate flag: one NVLink one-way latency) and polled at the destination until the tag
// matches.  On one GPU the "window" is local scratch and the very same code runs.
//
//   arg-max (K4)   per CTA best key -> atomicMax + ticket; the last CTA of the rank pushes the rank's key
//                  to every rank; everybody polls `world` keys, the largest wins (lowest row on ties).
//   D^2 draw (K5)  NumPy-exact: (1) rank-local grid barrier; (2) leaves of NumPy's float32 pairwise-sum
//                  tree, one 8-lane group per leaf, pushed to every rank; every CTA folds the combine
//                  tree itself (shared memory) so all agree on S bit for bit; (3) fp64 mass of
//                  prob = clip(mind,0)/S per leaf, pushed; every CTA scans the leaf masses and locates the
//                  leaf holding u * total; (4) the CTA owning that leaf searches inside it and pushes the
//                  picked row id.  Shards are aligned to leaf boundaries, so no leaf needs a peer's rows.
#include "alq_greedy.cuh"

namespace {

constexpr int kMaxLeaves = 4096;          // per partition: 2*K floats + K words of shared memory
constexpr int kConsWarps = 16;
constexpr int kConsThreads = kConsWarps * 32;

struct PGroup {                 // one partition = one group of CTAs
    int cta_lo, ncta;
    int row_lo;                 // first row of the partition (pick reported after a numeric failure)
    int budget, pick_off;
    int first_pick;             // >= 0: centre 0 chosen by the caller (nothing labeled), else -1
    int full_n, cfull_off;
    int n_leaves, leaf_base;    // leaf_off[leaf_base .. leaf_base + n_leaves]
    int n_levels, level_base;   // level_off[level_base .. level_base + n_levels]
    int sched_base;             // sched[sched_base .. sched_base + n_leaves - 1): internal node j = (l | r << 16)
    int leaf_lo, leaf_hi;       // leaves summed by THIS rank
    unsigned int keyw, leafw, massw, pickw;   // byte offsets of the LL regions inside a window
};

struct PersistArgs {
    const float* x;  long long ldx; int d;
    const float* a;  long long lda; int c;
    const float* xn; const float* an;
    float* mind;
    const int* vpos;
    const BlockSeg* segs;
    const PGroup* groups;
    float* cfull;
    const int* posinv;
    const int* leaf_off;
    const int* level_off;
    const unsigned int* sched;
    const double* uniforms;
    unsigned long long* best;      // [P * bmax]
    unsigned int* ticket;          // [P * bmax]
    unsigned int* bar;             // [P] monotonic rank-local barrier counter
    int* picks;
    int* status;
    unsigned long long* prof;      // [4] ns: streaming, selection, steps, -
    int bmax;
    int world, rank;
    char* peer[ALQ_MAX_WORLD];     // windows (world == 1: peer[0] = local scratch)
    int leaf_bound[ALQ_MAX_WORLD + 1];   // multi-GPU D^2: rank r sums leaves [leaf_bound[r], leaf_bound[r+1])
    unsigned int ready_off;        // u64 ready[world] at the head of every window
    unsigned long long ready_tag;
    unsigned int tag_base;         // (epoch & 0xff) << 24
    long long timeout_cycles;
};

// ---- LL words --------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(void* p, unsigned int tag, unsigned int payload) {
    const unsigned long long v = (static_cast<unsigned long long>(tag) << 32) | payload;
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const void* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Bounded spin; after a failure anywhere (sticky status) every wait returns at once, so the loop drains
// to its end in bounded time instead of hanging the GPU on a dead peer.
__device__ __forceinline__ unsigned int ll_wait(const void* p, unsigned int tag, const PersistArgs& A) {
    unsigned long long v = ll_load(p);
    if (static_cast<unsigned int>(v >> 32) == tag) return static_cast<unsigned int>(v);
    const long long t0 = clock64();
    for (int spin = 0;; ++spin) {
        v = ll_load(p);
        if (static_cast<unsigned int>(v >> 32) == tag) break;
        if ((spin & 63) == 63) {
            if (*reinterpret_cast<volatile int*>(A.status) != 0) break;
            if (clock64() - t0 > A.timeout_cycles) { atomicCAS(A.status, 0, ALQ_ERR_STATE); break; }
        }
    }
    return static_cast<unsigned int>(v);
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsThreads) : "memory"); }
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <bool FACTORED, bool SAMPLE>
__global__ void __launch_bounds__(32 * (1 + kConsWarps), 1)
greedy_persist_kernel(const PersistArgs A, const PipeCfg cfg, const int k_max) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int d = A.d, c = FACTORED ? A.c : 0;
    const int dv = d >> 2, cv = c >> 2;
    float* sq = reinterpret_cast<float*>(smem_raw);                          // centre: d + c floats
    float* tiles = sq + ((d + c + 31) & ~31);                                // stage ring
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(cfg.stages) * cfg.tile_floats);
    uint64_t* empty = full + cfg.stages;
    unsigned long long* sbest = reinterpret_cast<unsigned long long*>(empty + cfg.stages);   // [kConsWarps]
    double* sh_w = reinterpret_cast<double*>(sbest + kConsWarps);                             // [kConsWarps]
    float* val = reinterpret_cast<float*>(sh_w + kConsWarps);                                 // [2 * k_max] (SAMPLE)
    unsigned int* s_sched = reinterpret_cast<unsigned int*>(val + 2 * static_cast<size_t>(k_max));   // [k_max]
    __shared__ int s_level[40];
    __shared__ int sh_centre, sh_hit, sh_nz;
    __shared__ double sh_base, sh_base_nz;

    const BlockSeg seg = A.segs[blockIdx.x];
    const PGroup G = A.groups[seg.part];
    const int p = seg.part;
    const int grank = static_cast<int>(blockIdx.x) - G.cta_lo;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = cfg.rows_per_tile;
    const int nrows = seg.row_hi - seg.row_lo;
    const int ntiles = (nrows + R - 1) / R;
    const int nstream = G.budget - 1;                 // steps 1 .. budget-1 stream the rows

    if (threadIdx.x == 0) {
        for (int s = 0; s < cfg.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (SAMPLE) {
        for (int i = threadIdx.x; i < G.n_leaves - 1; i += blockDim.x) s_sched[i] = A.sched[G.sched_base + i];
        if (threadIdx.x <= G.n_levels && threadIdx.x < 40) s_level[threadIdx.x] = A.level_off[G.level_base + threadIdx.x];
    }
    __syncthreads();

    if (warp == 0) {
        // ===== producer: one lane issues every bulk copy of every step; it never needs the centre =====
        if (lane == 0 && ntiles > 0) {
            const bool contig_x = (A.ldx == d), contig_a = FACTORED && (A.lda == c);
            int s = 0;
            unsigned int par = 0;                     // parity of the `empty` phase to wait for (first round: none)
            bool first_round = true;
            for (int st = 0; st < nstream; ++st) {
                for (int i = 0; i < ntiles; ++i) {
                    if (!first_round) mbar_wait(&empty[s], par);
                    const int row0 = seg.row_lo + i * R;
                    const int rr = min(R, seg.row_hi - row0);
                    float* tx = tiles + static_cast<size_t>(s) * cfg.tile_floats;
                    float* ta = tx + static_cast<size_t>(R) * d;
                    mbar_expect_tx(&full[s], static_cast<uint32_t>(rr) * static_cast<uint32_t>(d + c) * 4u);
                    if (contig_x) {
                        bulk_g2s(tx, A.x + static_cast<long long>(row0) * A.ldx, static_cast<uint32_t>(rr) * d * 4u, &full[s]);
                    } else {
                        for (int r = 0; r < rr; ++r)
                            bulk_g2s(tx + static_cast<size_t>(r) * d, A.x + static_cast<long long>(row0 + r) * A.ldx, d * 4u, &full[s]);
                    }
                    if (FACTORED) {
                        if (contig_a) {
                            bulk_g2s(ta, A.a + static_cast<long long>(row0) * A.lda, static_cast<uint32_t>(rr) * c * 4u, &full[s]);
                        } else {
                            for (int r = 0; r < rr; ++r)
                                bulk_g2s(ta + static_cast<size_t>(r) * c, A.a + static_cast<long long>(row0 + r) * A.lda, c * 4u, &full[s]);
                        }
                    }
                    if (++s == cfg.stages) {
                        s = 0;
                        if (first_round) first_round = false; else par ^= 1u;
                    }
                }
            }
        }
        return;
    }

    // ===== consumers =====================================================================================
    const int ct = threadIdx.x - 32;          // 0 .. kConsThreads-1
    const int cw = warp - 1;
    const int C = cfg.consumers;              // warps that take tiles (<= kConsWarps)
    const int cf_off = SAMPLE ? G.cfull_off : 0;
    float* cf = A.cfull + cf_off;
    char* const win = A.peer[A.rank];
    const int K = G.n_leaves;
    const int root = K > 1 ? 2 * K - 2 : 0;
    const float4* q4 = reinterpret_cast<const float4*>(sq);

    if (A.world > 1 && ct < A.world)          // peers may still be clearing their windows for this call
        for (long long t0 = clock64();;) {
            if (ld_acquire_sys(reinterpret_cast<const unsigned long long*>(win + A.ready_off) + ct) == A.ready_tag) break;
            if (clock64() - t0 > A.timeout_cycles) { atomicCAS(A.status, 0, ALQ_ERR_STATE); break; }
            __nanosleep(200);
        }
    cons_bar();

    // per-warp position in the tile stream (global over all steps): tiles cw, cw + C, cw + 2C, ...
    long long my_it = cw;                     // next tile index (over all steps) this warp consumes
    int my_s = cw % cfg.stages;
    unsigned int my_par = 0;
    for (int q = cw / cfg.stages; q > 0; --q) my_par ^= 1u;
    long long step_base = 0;                  // tile index of the first tile of the current step

    unsigned int rnd_key = 0, rnd_leaf = 0, rnd_mass = 0, rnd_pick = 0, n_bar = 0;
    int centre = -1;
    unsigned long long acc_stream = 0, acc_select = 0, t_a = 0, t_b = 0;
    const bool prof = A.prof != nullptr && blockIdx.x == 0 && ct == 0;

    for (int t = 0; t < G.budget; ++t) {
        unsigned long long best_key = 0ull;
        if (prof) t_a = gtime_ns();
        if (t == 0) {
            if (G.first_pick >= 0) {          // centre 0 chosen by the caller: no selection
                centre = G.first_pick;
                if (grank == 0 && ct == 0) A.picks[G.pick_off] = centre;
                continue;
            }
            for (int row = seg.row_lo + ct; row < seg.row_hi; row += kConsThreads) {
                const float m = __ldcg(A.mind + row);
                if (SAMPLE) cf[A.vpos[row]] = m;
                else {
                    const unsigned long long k = alq_maxkey(m, static_cast<uint32_t>(row));
                    best_key = k > best_key ? k : best_key;
                }
            }
        } else {
            // ---- centre row -> shared memory (a replica of every candidate row is local) ----
            {
                float4* dst = reinterpret_cast<float4*>(sq);
                const float4* src = reinterpret_cast<const float4*>(A.x + static_cast<long long>(centre) * A.ldx);
                for (int k = ct; k < dv; k += kConsThreads) dst[k] = __ldg(src + k);
                if (FACTORED) {
                    const float4* sa = reinterpret_cast<const float4*>(A.a + static_cast<long long>(centre) * A.lda);
                    for (int k = ct; k < cv; k += kConsThreads) dst[dv + k] = __ldg(sa + k);
                }
            }
            const float qn = FACTORED ? __ldg(A.xn + centre) * __ldg(A.an + centre) : __ldg(A.xn + centre);
            cons_bar();
            // ---- this step's tiles ----
            const long long step_end = step_base + ntiles;
            if (cw < C) {
                for (; my_it < step_end; my_it += C) {
                    const int i = static_cast<int>(my_it - step_base);
                    const int row0 = seg.row_lo + i * R;
                    const int rr = min(R, seg.row_hi - row0);
                    float m_old = 0.f, n_i = 0.f;
                    if (lane < rr) {
                        m_old = __ldcg(A.mind + row0 + lane);
                        n_i = FACTORED ? __ldg(A.xn + row0 + lane) * __ldg(A.an + row0 + lane) : __ldg(A.xn + row0 + lane);
                    }
                    mbar_wait(&full[my_s], my_par);
                    const float* tx = tiles + static_cast<size_t>(my_s) * cfg.tile_floats;
                    const float* ta = tx + static_cast<size_t>(R) * d;
                    float my_d2 = 0.f;
                    for (int r = 0; r < rr; ++r) {
                        const float4* pr = reinterpret_cast<const float4*>(tx + static_cast<size_t>(r) * d);
                        float dot = 0.f;
#pragma unroll 4
                        for (int k = lane; k < dv; k += 32) {
                            const float4 v = pr[k];
                            const float4 w = q4[k];
                            dot = fmaf(v.x, w.x, dot);
                            dot = fmaf(v.y, w.y, dot);
                            dot = fmaf(v.z, w.z, dot);
                            dot = fmaf(v.w, w.w, dot);
                        }
                        dot = warp_sum(dot);
                        if (FACTORED) {
                            const float4* pa = reinterpret_cast<const float4*>(ta + static_cast<size_t>(r) * c);
                            float da = 0.f;
#pragma unroll 4
                            for (int k = lane; k < cv; k += 32) {
                                const float4 v = pa[k];
                                const float4 w = q4[dv + k];
                                da = fmaf(v.x, w.x, da);
                                da = fmaf(v.y, w.y, da);
                                da = fmaf(v.z, w.z, da);
                                da = fmaf(v.w, w.w, da);
                            }
                            dot *= warp_sum(da);
                        }
                        if (lane == r) my_d2 = dot;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[my_s]);
                    if (lane < rr) {
                        const int row = row0 + lane;
                        float m = fminf(m_old, dist_dense(n_i, qn, my_d2));
                        if (row == centre) m = ALQ_NEG_INF;         // a picked row is never a candidate again
                        __stcg(A.mind + row, m);
                        if (SAMPLE) cf[A.vpos[row]] = m;            // raw running min; the draw clips at 0
                        else {
                            const unsigned long long k = alq_maxkey(m, static_cast<uint32_t>(row));
                            best_key = k > best_key ? k : best_key;
                        }
                    }
                    my_s += C;
                    if (my_s >= cfg.stages) { my_s -= cfg.stages; my_par ^= 1u; }
                }
            }
            step_base = step_end;
        }

        // =================================== selection of pick t ===================================
        if (!SAMPLE) {
            best_key = warp_max_u64(best_key);
            if (lane == 0) sbest[cw] = best_key;
            cons_bar();
            if (prof) t_b = gtime_ns();
            if (cw == 0) {
                const unsigned int tag = A.tag_base | (++rnd_key & 0xffffffu);
                const unsigned int slot = rnd_key & 1u;
                int last = 0;
                unsigned long long key = 0ull;
                if (lane == 0) {
                    unsigned long long b = 0ull;
                    for (int w = 0; w < kConsWarps; ++w) b = sbest[w] > b ? sbest[w] : b;
                    unsigned long long* bs = A.best + static_cast<size_t>(p) * A.bmax + t;
                    if (b) atomicMax(bs, b);
                    __threadfence();
                    last = atomicAdd(A.ticket + static_cast<size_t>(p) * A.bmax + t, 1u) == static_cast<unsigned int>(G.ncta - 1);
                    if (last) {
                        __threadfence();
                        key = *reinterpret_cast<volatile unsigned long long*>(bs);
                    }
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                key = __shfl_sync(0xffffffffu, key, 0);
                if (last && lane < A.world) {     // this rank's best -> every rank (two LL words)
                    char* dst = A.peer[lane] + G.keyw + (static_cast<size_t>(slot) * A.world + A.rank) * 16;
                    ll_store(dst, tag, static_cast<unsigned int>(key >> 32));
                    ll_store(dst + 8, tag, static_cast<unsigned int>(key));
                }
                unsigned long long k = 0ull;
                if (lane < A.world) {
                    const char* src = win + G.keyw + (static_cast<size_t>(slot) * A.world + lane) * 16;
                    const unsigned int hi = ll_wait(src, tag, A);
                    const unsigned int lo = ll_wait(src + 8, tag, A);
                    k = (static_cast<unsigned long long>(hi) << 32) | lo;
                }
                k = warp_max_u64(k);
                if (lane == 0) sh_centre = static_cast<int>(alq_maxkey_row(k));
            }
            cons_bar();
            centre = sh_centre;
        } else {
            // ---- (1) every min-distance of this rank's rows is in cfull ----
            cons_bar();
            if (prof) t_b = gtime_ns();
            float total32 = 0.f;
            bool failed = false;
            for (int attempt = 0;; ++attempt) {
                if (ct == 0) {
                    __threadfence();
                    atomicAdd(A.bar + p, 1u);
                    const unsigned int target = static_cast<unsigned int>(G.ncta) * (++n_bar);
                    const long long t0 = clock64();
                    for (int spin = 0; static_cast<int>(ld_acquire_gpu_u32(A.bar + p) - target) < 0; ++spin)
                        if ((spin & 63) == 63 && (*reinterpret_cast<volatile int*>(A.status) != 0 || clock64() - t0 > A.timeout_cycles)) {
                            atomicCAS(A.status, 0, ALQ_ERR_STATE);
                            break;
                        }
                }
                cons_bar();
                // ---- (2) leaf sums of NumPy's pairwise tree: one 8-lane group per leaf ----
                const unsigned int tag = A.tag_base | (++rnd_leaf & 0xffffffu);
                const unsigned int slot = rnd_leaf & 1u;
                {
                    const int grp = ct >> 3, g_lane = ct & 7;
                    const unsigned gmask = 0xffu << ((lane >> 3) * 8);
                    for (int lf = grp;; lf += kConsThreads / 8) {
                        const int leaf = G.leaf_lo + grank + lf * G.ncta;
                        if (leaf >= G.leaf_hi) break;
                        const int lo = A.leaf_off[G.leaf_base + leaf], len = A.leaf_off[G.leaf_base + leaf + 1] - lo;
                        float v = leaf_sum_group(cf + lo, len, g_lane, gmask);
                        v = __shfl_sync(gmask, v, lane & ~7);
                        if (g_lane < A.world)
                            ll_store(A.peer[g_lane] + G.leafw + (static_cast<size_t>(slot) * K + leaf) * 8, tag, __float_as_uint(v));
                    }
                }
                for (int i = ct; i < K; i += kConsThreads)
                    val[i] = __uint_as_float(ll_wait(win + G.leafw + (static_cast<size_t>(slot) * K + i) * 8, tag, A));
                cons_bar();
                for (int h = 0; h < G.n_levels; ++h) {
                    const int lo = s_level[h], hi = s_level[h + 1];
                    for (int j = lo + ct; j < hi; j += kConsThreads) {
                        const unsigned int e = s_sched[j];
                        val[K + j] = val[e & 0xffffu] + val[e >> 16];
                    }
                    cons_bar();
                }
                total32 = val[root];
                cons_bar();                                   // val is reused below
                if (total32 > 0.f && total32 <= 3.4028234e38f) break;
                if (!(total32 == 0.f) || attempt > (1 << 20)) { failed = true; break; }   // NaN / inf mass
                // sum == 0 -> prob is NaN -> `min_dist_labeled += 0.00001` and retry (:87-90).  In place: the next
                // step rewrites every candidate slot from `mind`, so the bump never outlives this draw.
                for (int row = seg.row_lo + ct; row < seg.row_hi; row += kConsThreads) {
                    float* q = cf + A.vpos[row];
                    __stcg(q, __ldcg(q) + 0.00001f);
                }
                cons_bar();
            }
            if (failed) {                                     // the same decision in every CTA of every rank
                if (ct == 0) atomicCAS(A.status, 0, ALQ_ERR_NUMERIC);
                centre = G.row_lo;
            } else {
                // ---- (3) fp64 mass of prob = clip(mind, 0) / S per leaf ----
                const unsigned int tag = A.tag_base | (++rnd_mass & 0xffffffu);
                const unsigned int slot = rnd_mass & 1u;
                {
                    const int grp = ct >> 3, g_lane = ct & 7;
                    const unsigned gmask = 0xffu << ((lane >> 3) * 8);
                    for (int lf = grp;; lf += kConsThreads / 8) {
                        const int leaf = G.leaf_lo + grank + lf * G.ncta;
                        if (leaf >= G.leaf_hi) break;
                        const int lo = A.leaf_off[G.leaf_base + leaf], len = A.leaf_off[G.leaf_base + leaf + 1] - lo;
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = (8 * j + g_lane < len) ? __ldcg(cf + lo + 8 * j + g_lane) : 0.f;
                        double m = 0.0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) m += prob64(v[j], total32);
                        m += __shfl_down_sync(gmask, m, 1, 8);
                        m += __shfl_down_sync(gmask, m, 2, 8);
                        m += __shfl_down_sync(gmask, m, 4, 8);
                        m = __shfl_sync(gmask, m, lane & ~7);
                        if (g_lane < A.world) {
                            char* dst = A.peer[g_lane] + G.massw + (static_cast<size_t>(slot) * K + leaf) * 16;
                            ll_store(dst, tag, static_cast<unsigned int>(__double2hiint(m)));
                            ll_store(dst + 8, tag, static_cast<unsigned int>(__double2loint(m)));
                        }
                    }
                }
                double* M = reinterpret_cast<double*>(val);   // K doubles == 2K floats
                for (int i = ct; i < K; i += kConsThreads) {
                    const char* src = win + G.massw + (static_cast<size_t>(slot) * K + i) * 16;
                    const unsigned int hi = ll_wait(src, tag, A);
                    const unsigned int lo = ll_wait(src + 8, tag, A);
                    M[i] = __hiloint2double(static_cast<int>(hi), static_cast<int>(lo));
                }
                if (ct == 0) { sh_hit = 0x7fffffff; sh_nz = -1; }
                cons_bar();
                // ---- np.random.choice == first k with cumsum64(p)[k] / total > u: locate the leaf.  One fixed
                //      chain (thread chunks -> lanes -> warps), identical in every CTA of every rank. ----
                const double u = A.uniforms[G.pick_off + t];
                const int per = (K + kConsThreads - 1) / kConsThreads;
                const int l0 = min(K, ct * per), l1 = min(K, l0 + per);
                double loc = 0.0;
                for (int l = l0; l < l1; ++l) loc += M[l];
                double inc = loc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) sh_w[cw] = inc;
                const double prev = __shfl_up_sync(0xffffffffu, inc, 1);
                cons_bar();
                double woff = 0.0, total = 0.0;
                for (int w = 0; w < kConsWarps; ++w) {
                    if (w == cw) woff = total;
                    total += sh_w[w];
                }
                double run = woff + (lane ? prev : 0.0);
                int my_hit = 0x7fffffff, my_nz = -1;
                double hit_base = 0.0, nz_base = 0.0;
                for (int l = l0; l < l1; ++l) {
                    const double before = run;
                    run += M[l];
                    if (M[l] > 0.0) { my_nz = l; nz_base = before; }
                    if (my_hit == 0x7fffffff && (run / total) > u) { my_hit = l; hit_base = before; }
                }
                if (my_hit != 0x7fffffff) atomicMin(&sh_hit, my_hit);
                if (my_nz >= 0) atomicMax(&sh_nz, my_nz);
                cons_bar();
                if (my_hit != 0x7fffffff && my_hit == sh_hit) sh_base = hit_base;
                if (my_nz >= 0 && my_nz == sh_nz) sh_base_nz = nz_base;
                cons_bar();
                int leaf = sh_hit;
                double base = sh_base;
                if (leaf == 0x7fffffff) { leaf = sh_nz; base = sh_base_nz; }   // u beyond the last mass by an ulp
                // ---- (4) the CTA that owns the leaf searches inside it and announces the row ----
                const unsigned int ptag = A.tag_base | (++rnd_pick & 0xffffffu);
                const unsigned int pslot = rnd_pick & 1u;
                int owner_rank = 0;
                for (int r = 1; r < A.world; ++r)
                    if (leaf >= A.leaf_bound[r]) owner_rank = r;
                const bool mine = leaf >= 0 && owner_rank == A.rank && (leaf - G.leaf_lo) % G.ncta == grank;
                if (leaf < 0 && grank == 0 && A.rank == 0 && cw == 0) {       // no mass at all: cannot happen with S > 0
                    if (lane == 0) atomicCAS(A.status, 0, ALQ_ERR_NUMERIC);
                    if (lane < A.world) ll_store(A.peer[lane] + G.pickw + pslot * 8, ptag, static_cast<unsigned int>(G.row_lo));
                }
                if (mine && cw == 0) {
                    const int lo = A.leaf_off[G.leaf_base + leaf], len = A.leaf_off[G.leaf_base + leaf + 1] - lo;
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = (4 * lane + j < len) ? __ldcg(cf + lo + 4 * lane + j) : 0.f;
                    double pl[4], lsum = 0.0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { pl[j] = prob64(v[j], total32); lsum += pl[j]; }
                    double linc = lsum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double w = __shfl_up_sync(0xffffffffu, linc, o);
                        if (lane >= o) linc += w;
                    }
                    double r2 = base + (linc - lsum);
                    int hit = -1, nz = -1;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        r2 += pl[j];
                        if (pl[j] > 0.0) nz = 4 * lane + j;
                        if (hit < 0 && 4 * lane + j < len && (r2 / total) > u) hit = 4 * lane + j;
                    }
                    const unsigned hb = __ballot_sync(0xffffffffu, hit >= 0);
                    int k;
                    if (hb) k = __shfl_sync(0xffffffffu, hit, __ffs(hb) - 1);
                    else {                                   // re-association moved the crossing by an ulp
                        const unsigned nb = __ballot_sync(0xffffffffu, nz >= 0);
                        k = nb ? __shfl_sync(0xffffffffu, nz, 31 - __clz(nb)) : -1;
                    }
                    int row = k >= 0 ? A.posinv[cf_off + lo + k] : -1;
                    if (row < 0) {
                        if (lane == 0) atomicCAS(A.status, 0, ALQ_ERR_NUMERIC);
                        row = G.row_lo;
                    }
                    if (lane < A.world) ll_store(A.peer[lane] + G.pickw + pslot * 8, ptag, static_cast<unsigned int>(row));
                }
                if (ct == 0) sh_centre = static_cast<int>(ll_wait(win + G.pickw + pslot * 8, ptag, A));
                cons_bar();
                centre = sh_centre;
            }
        }
        if (grank == 0 && ct == 0) A.picks[G.pick_off + t] = centre;
        if (prof) {
            const unsigned long long t_c = gtime_ns();
            if (t > 0) { acc_stream += t_b - t_a; acc_select += t_c - t_b; }
        }
    }
    if (prof) {
        A.prof[0] = acc_stream;
        A.prof[1] = acc_select;
        A.prof[2] = static_cast<unsigned long long>(G.budget > 1 ? G.budget - 1 : 0);
    }
}

__global__ void persist_fill_kernel(float* p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// posinv[off[p] + vpos[row]] = row for every candidate row of every partition
__global__ void persist_posinv_kernel(const int* __restrict__ vpos, const int* __restrict__ part_off,
                                      const int* __restrict__ cfull_off, int P, int n, int* __restrict__ posinv) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    int p = 0;
    while (p + 1 < P && row >= part_off[p + 1]) ++p;
    posinv[cfull_off[p] + vpos[row]] = row;
}

__global__ void persist_ready_kernel(PersistArgs A) {
    if (blockIdx.x == 0 && threadIdx.x < A.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(A.peer[threadIdx.x] + A.ready_off) + A.rank, A.ready_tag);
    }
}

// ---- NumPy's pairwise-summation tree ---------------------------------------------------------------
struct PairTree {
    std::vector<int> leaf_off;                       // n_leaves + 1
    struct Node { int height, l, r; };               // children: >= 0 leaf id, < 0 internal #(-v - 1)
    std::vector<Node> internal;                      // creation (post-) order
    std::pair<int, int> rec(int lo, int m) {         // -> (ref, height)
        if (m <= 128) {
            leaf_off.push_back(lo);
            return {static_cast<int>(leaf_off.size()) - 1, 0};
        }
        int half = m / 2;
        half -= half % 8;
        const auto L = rec(lo, half);
        const auto Rr = rec(lo + half, m - half);
        const int h = std::max(L.second, Rr.second) + 1;
        internal.push_back({h, L.first, Rr.first});
        return {-static_cast<int>(internal.size()), h};
    }
};

template <bool FACTORED, bool SAMPLE>
cudaError_t launch_persist(int grid, size_t smem, cudaStream_t st, PersistArgs& A, PipeCfg& cfg, int& k_max) {
    auto* fn = greedy_persist_kernel<FACTORED, SAMPLE>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    void* args[] = {&A, &cfg, &k_max};
    return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(fn), dim3(grid), dim3(32 * (1 + kConsWarps)), args, smem, st);
}

}  // namespace

extern "C" int64_t alq_pairwise_leaf_bounds(int64_t n, int32_t* out_host, int64_t cap) {
    if (n <= 0 || n >= (1LL << 31) || !out_host) return -1;
    PairTree tb;
    tb.rec(0, static_cast<int>(n));
    const int64_t k = static_cast<int64_t>(tb.leaf_off.size());
    if (cap < k + 1) return -1;
    for (int64_t i = 0; i < k; ++i) out_host[i] = tb.leaf_off[i];
    out_host[k] = static_cast<int32_t>(n);
    return k;
}

int alq_greedy_persist(alq_ctx* ctx, const alq_greedy_desc* D, void* stream) {
    const int P = D->n_parts;
    const int64_t n = D->n;
    const bool factored = D->a != nullptr;
    const bool sample = D->uniforms_host != nullptr;
    const int d = D->d, c = factored ? D->c : 0;
    const AlqComm& Gc = ctx->comm;
    const bool comm = D->shard_off_host != nullptr && Gc.world > 1;
    const int world = comm ? Gc.world : 1, rank = comm ? Gc.rank : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    // ---- does the problem fit? --------------------------------------------------------------------------
    const size_t row_bytes = static_cast<size_t>(d + c) * 4;
    if ((row_bytes % 16) || (static_cast<size_t>(d) * 4 % 16)) return kPersistNotApplicable;
    int row_lo = 0, row_hi = static_cast<int>(n);
    if (comm) { row_lo = D->shard_off_host[rank]; row_hi = D->shard_off_host[rank + 1]; }
    std::vector<int> pick_off(P + 1, 0);
    int bmax = 0, active = 0;
    for (int p = 0; p < P; ++p) {
        pick_off[p + 1] = pick_off[p] + D->budget_host[p];
        bmax = std::max(bmax, D->budget_host[p]);
        if (D->budget_host[p] > 0) ++active;
    }
    if (active == 0) return ALQ_OK;
    if (active > ctx->sm_count) return kPersistNotApplicable;

    // ---- trees (D^2 sampling) ------------------------------------------------------------------------------
    std::vector<PGroup> groups(P);
    std::vector<int> leaf_off_all, level_off_all, cfull_off(P, 0);
    std::vector<unsigned int> sched_all;
    int cfull_total = 0, k_max = 1;
    int leaf_bound[ALQ_MAX_WORLD + 1] = {};
    for (int p = 0; p < P; ++p) {
        PGroup& g = groups[p];
        g = PGroup{};
        g.row_lo = comm ? 0 : D->part_off_host[p];
        g.budget = D->budget_host[p];
        g.pick_off = pick_off[p];
        g.first_pick = D->first_pick_host ? D->first_pick_host[p] : -1;
        g.n_leaves = 1;
        if (!sample) continue;
        g.full_n = D->full_n_host[p];
        g.cfull_off = cfull_total;
        cfull_off[p] = cfull_total;
        cfull_total += (g.full_n + 3) & ~3;
        PairTree tb;
        tb.rec(0, g.full_n);
        const int K = static_cast<int>(tb.leaf_off.size());
        if (K > kMaxLeaves) return kPersistNotApplicable;
        tb.leaf_off.push_back(g.full_n);
        g.n_leaves = K;
        g.leaf_base = static_cast<int>(leaf_off_all.size());
        leaf_off_all.insert(leaf_off_all.end(), tb.leaf_off.begin(), tb.leaf_off.end());
        // internal nodes in level order; node id = K + position in that order
        int hmax = 0;
        for (auto& nd : tb.internal) hmax = std::max(hmax, nd.height);
        std::vector<int> order, newid(tb.internal.size());
        g.level_base = static_cast<int>(level_off_all.size());
        for (int h = 1; h <= hmax; ++h) {
            level_off_all.push_back(static_cast<int>(order.size()));
            for (size_t k = 0; k < tb.internal.size(); ++k)
                if (tb.internal[k].height == h) { newid[k] = K + static_cast<int>(order.size()); order.push_back(static_cast<int>(k)); }
        }
        level_off_all.push_back(static_cast<int>(order.size()));
        g.n_levels = hmax;
        g.sched_base = static_cast<int>(sched_all.size());
        auto nid = [&](int v) { return v >= 0 ? v : newid[-v - 1]; };
        for (int k : order)
            sched_all.push_back(static_cast<unsigned int>(nid(tb.internal[k].l)) | (static_cast<unsigned int>(nid(tb.internal[k].r)) << 16));
        g.leaf_lo = 0;
        g.leaf_hi = K;
        k_max = std::max(k_max, K);
        if (comm) {     // shards must be aligned to leaf boundaries
            if (!D->shard_pos_host) ALQ_FAIL(ctx, ALQ_ERR_INVALID, "alq_greedy_select: multi-GPU D^2 sampling needs shard_pos (leaf-aligned shards)");
            for (int r = 0; r <= world; ++r) {
                const int pos = D->shard_pos_host[r];
                auto it = std::lower_bound(tb.leaf_off.begin(), tb.leaf_off.end(), pos);
                if (it == tb.leaf_off.end() || *it != pos)
                    ALQ_FAIL(ctx, ALQ_ERR_INVALID, "alq_greedy_select: shard_pos[%d] = %d is not a leaf boundary of the pairwise-sum tree", r, pos);
                leaf_bound[r] = static_cast<int>(it - tb.leaf_off.begin());
            }
            if (leaf_bound[0] != 0 || leaf_bound[world] != K)
                ALQ_FAIL(ctx, ALQ_ERR_INVALID, "alq_greedy_select: shard_pos must span [0, full_n]");
            g.leaf_lo = leaf_bound[rank];
            g.leaf_hi = leaf_bound[rank + 1];
        }
    }

    // ---- shared-memory plan: centre + ring + tree ---------------------------------------------------------
    PipeCfg cfg{};
    size_t smem = 0;
    {
        const size_t centre_bytes = static_cast<size_t>((d + c + 31) & ~31) * 4;
        const size_t tree_bytes = sample ? static_cast<size_t>(k_max) * 12 + 64 : 64;
        const size_t fixed = centre_bytes + tree_bytes + kConsWarps * 16 + 256;
        const size_t budget_bytes = ctx->smem_optin > 8192 ? ctx->smem_optin - 1536 : 0;
        size_t tile_target = 32 * 1024;
        int max_stages = 16;
        if (const char* e = getenv("ALQ_TILE_KB")) tile_target = std::max(1, atoi(e)) * 1024;
        if (const char* e = getenv("ALQ_MAX_STAGES")) max_stages = std::max(3, std::min(16, atoi(e)));
        int R = static_cast<int>(std::max<size_t>(1, tile_target / row_bytes));
        R = std::min(R, 32);
        const size_t tile_bytes = R * row_bytes;
        int stages = budget_bytes > fixed ? static_cast<int>((budget_bytes - fixed) / (tile_bytes + 16)) : 0;
        stages = std::min(stages, max_stages);
        if (stages < 3) return kPersistNotApplicable;
        cfg.rows_per_tile = R;
        cfg.stages = stages;
        cfg.tile_floats = static_cast<int>(tile_bytes / 4);
        cfg.consumers = std::max(1, std::min(kConsWarps, stages));
        smem = centre_bytes + stages * tile_bytes + 2 * stages * sizeof(uint64_t) + kConsWarps * 16 + tree_bytes;
    }

    // ---- CTAs: at most one per SM in total, split over the partitions by row count -------------------------
    std::vector<BlockSeg> segs;
    {
        std::vector<int64_t> rows(P, 0);
        int64_t total = 0;
        for (int p = 0; p < P; ++p) {
            if (D->budget_host[p] <= 0) continue;
            rows[p] = comm ? (row_hi - row_lo) : (D->part_off_host[p + 1] - D->part_off_host[p]);
            total += rows[p];
        }
        int left = ctx->sm_count - active;           // one CTA per active partition first, the rest by share
        std::vector<int> nb(P, 0);
        for (int p = 0; p < P; ++p) {
            if (D->budget_host[p] <= 0) continue;
            const int extra = total > 0 ? static_cast<int>(static_cast<int64_t>(ctx->sm_count - active) * rows[p] / total) : 0;
            nb[p] = 1 + std::min(extra, left);
            left -= nb[p] - 1;
        }
        for (int p = 0; p < P && left > 0; ++p)
            if (nb[p] > 0 && rows[p] > nb[p]) { ++nb[p]; --left; }
        int cta = 0;
        for (int p = 0; p < P; ++p) {
            if (nb[p] == 0) continue;
            const int lo = comm ? row_lo : D->part_off_host[p];
            const int64_t r = rows[p];
            nb[p] = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(nb[p], std::max<int64_t>(r, 1))));
            groups[p].cta_lo = cta;
            groups[p].ncta = nb[p];
            for (int b = 0; b < nb[p]; ++b) {
                BlockSeg s;
                s.row_lo = lo + static_cast<int>(r * b / nb[p]);
                s.row_hi = lo + static_cast<int>(r * (b + 1) / nb[p]);
                s.part = p;
                s.pad = 0;
                segs.push_back(s);
                ++cta;
            }
        }
    }
    const int grid = static_cast<int>(segs.size());
    if (grid > ctx->sm_count) return kPersistNotApplicable;

    // ---- LL regions: identical layout in every window ------------------------------------------------------
    auto up = [](size_t v) { return (v + 127) & ~size_t(127); };
    size_t woff = up(8 * ALQ_MAX_WORLD);        // u64 ready[world] at offset 0
    for (int p = 0; p < P; ++p) {
        PGroup& g = groups[p];
        g.keyw = static_cast<unsigned int>(woff);  woff = up(woff + static_cast<size_t>(2) * world * 16);
        g.pickw = static_cast<unsigned int>(woff); woff = up(woff + 2 * 8);
        if (sample) {
            g.leafw = static_cast<unsigned int>(woff); woff = up(woff + static_cast<size_t>(2) * g.n_leaves * 8);
            g.massw = static_cast<unsigned int>(woff); woff = up(woff + static_cast<size_t>(2) * g.n_leaves * 16);
        }
    }
    const size_t win_bytes = woff;
    if (comm && win_bytes > Gc.greedy_bytes())
        ALQ_FAIL(ctx, ALQ_ERR_NOMEM, "alq_greedy_select: peer window too small (%zu needed, %zu usable)", win_bytes, Gc.greedy_bytes());

    // ---- scratch ----------------------------------------------------------------------------------------------
    const int total_picks = pick_off[P];
    const size_t n_slots = static_cast<size_t>(P) * bmax;
    const size_t need = scratch_need({segs.size() * sizeof(BlockSeg), groups.size() * sizeof(PGroup),
                                      leaf_off_all.size() * 4 + 4, level_off_all.size() * 4 + 4, sched_all.size() * 4 + 4,
                                      static_cast<size_t>(cfull_total) * 4 + 16, static_cast<size_t>(cfull_total) * 4 + 16,
                                      static_cast<size_t>(total_picks) * 8 + 8, n_slots * 8, n_slots * 4, static_cast<size_t>(P) * 4,
                                      static_cast<size_t>(P + 1) * 4, static_cast<size_t>(P) * 4, 64, 64, comm ? 0 : win_bytes});
    int rc = alq_scratch_reserve(ctx, need);
    if (rc) return rc;
    ScratchCursor cur(ctx->scratch);
    BlockSeg* d_segs = cur.take<BlockSeg>(segs.size());
    PGroup* d_groups = cur.take<PGroup>(groups.size());
    int* d_leaf_off = cur.take<int>(leaf_off_all.size() + 1);
    int* d_level_off = cur.take<int>(level_off_all.size() + 1);
    unsigned int* d_sched = cur.take<unsigned int>(sched_all.size() + 1);
    float* d_cfull = cur.take<float>(cfull_total + 4);
    int* d_posinv = cur.take<int>(cfull_total + 4);
    double* d_unif = cur.take<double>(total_picks + 1);
    unsigned long long* d_best = cur.take<unsigned long long>(n_slots);
    unsigned int* d_ticket = cur.take<unsigned int>(n_slots);
    unsigned int* d_bar = cur.take<unsigned int>(P);
    int* d_part_off = cur.take<int>(P + 1);
    int* d_cfull_off = cur.take<int>(P);
    int* d_status = cur.take<int>(1);
    unsigned long long* d_prof = cur.take<unsigned long long>(4);
    char* d_win = comm ? Gc.window : cur.take<char>(win_bytes);

    ALQ_CUDA(ctx, cudaMemcpyAsync(d_segs, segs.data(), segs.size() * sizeof(BlockSeg), cudaMemcpyHostToDevice, st));
    ALQ_CUDA(ctx, cudaMemcpyAsync(d_groups, groups.data(), groups.size() * sizeof(PGroup), cudaMemcpyHostToDevice, st));
    ALQ_CUDA(ctx, cudaMemsetAsync(d_status, 0, sizeof(int), st));
    ALQ_CUDA(ctx, cudaMemsetAsync(d_prof, 0, 4 * sizeof(unsigned long long), st));
    ALQ_CUDA(ctx, cudaMemsetAsync(d_best, 0, n_slots * 8, st));
    ALQ_CUDA(ctx, cudaMemsetAsync(d_ticket, 0, n_slots * 4, st));
    ALQ_CUDA(ctx, cudaMemsetAsync(d_bar, 0, static_cast<size_t>(P) * 4, st));
    // LL regions start without any valid tag (after the ready[] words; peers only write them after our ready flag)
    ALQ_CUDA(ctx, cudaMemsetAsync(d_win + up(8 * ALQ_MAX_WORLD), 0, win_bytes - up(8 * ALQ_MAX_WORLD), st));
    if (sample) {
        ALQ_CUDA(ctx, cudaMemcpyAsync(d_leaf_off, leaf_off_all.data(), leaf_off_all.size() * 4, cudaMemcpyHostToDevice, st));
        ALQ_CUDA(ctx, cudaMemcpyAsync(d_level_off, level_off_all.data(), level_off_all.size() * 4, cudaMemcpyHostToDevice, st));
        if (!sched_all.empty())
            ALQ_CUDA(ctx, cudaMemcpyAsync(d_sched, sched_all.data(), sched_all.size() * 4, cudaMemcpyHostToDevice, st));
        ALQ_CUDA(ctx, cudaMemcpyAsync(d_unif, D->uniforms_host, static_cast<size_t>(total_picks) * 8, cudaMemcpyHostToDevice, st));
        ALQ_CUDA(ctx, cudaMemcpyAsync(d_part_off, D->part_off_host, static_cast<size_t>(P + 1) * 4, cudaMemcpyHostToDevice, st));
        ALQ_CUDA(ctx, cudaMemcpyAsync(d_cfull_off, cfull_off.data(), static_cast<size_t>(P) * 4, cudaMemcpyHostToDevice, st));
        persist_fill_kernel<<<(cfull_total + 4 + 255) / 256, 256, 0, st>>>(d_cfull, cfull_total + 4, -INFINITY);
        ALQ_LAUNCH_CHECK(ctx);
        ALQ_CUDA(ctx, cudaMemsetAsync(d_posinv, 0xff, static_cast<size_t>(cfull_total + 4) * 4, st));
        persist_posinv_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(D->vpos, d_part_off, d_cfull_off, P, static_cast<int>(n), d_posinv);
        ALQ_LAUNCH_CHECK(ctx);
    }

    PersistArgs A{};
    A.x = D->x; A.ldx = D->ldx; A.d = d;
    A.a = D->a; A.lda = D->lda; A.c = c;
    A.xn = D->xn; A.an = D->an;
    A.mind = D->mind; A.vpos = D->vpos;
    A.segs = d_segs; A.groups = d_groups;
    A.cfull = d_cfull; A.posinv = d_posinv;
    A.leaf_off = d_leaf_off; A.level_off = d_level_off; A.sched = d_sched; A.uniforms = d_unif;
    A.best = d_best; A.ticket = d_ticket; A.bar = d_bar;
    A.picks = D->picks; A.status = d_status;
    A.prof = D->step_kernel_ms_host ? d_prof : nullptr;
    A.bmax = bmax;
    A.world = world; A.rank = rank;
    for (int r = 0; r < ALQ_MAX_WORLD; ++r) A.peer[r] = nullptr;
    if (comm) for (int r = 0; r < world; ++r) A.peer[r] = Gc.peer[r];
    else A.peer[0] = d_win;
    for (int r = 0; r <= ALQ_MAX_WORLD; ++r) A.leaf_bound[r] = leaf_bound[std::min(r, world)];
    A.ready_off = 0;
    ctx->comm.epoch += 1;
    A.ready_tag = ctx->comm.epoch << 32;
    A.tag_base = static_cast<unsigned int>(ctx->comm.epoch & 0x7fu) << 24 | 0x80000000u;   // never 0: a zeroed word is invalid
    int clock_khz = 1900000;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, ctx->device);
    A.timeout_cycles = static_cast<long long>(ctx->spin_timeout_ms) * clock_khz;

    if (comm) {
        persist_ready_kernel<<<1, 32, 0, st>>>(A);       // after the clears above, in stream order
        ALQ_LAUNCH_CHECK(ctx);
    }
    cudaError_t le;
    if (factored) le = sample ? launch_persist<true, true>(grid, smem, st, A, cfg, k_max) : launch_persist<true, false>(grid, smem, st, A, cfg, k_max);
    else le = sample ? launch_persist<false, true>(grid, smem, st, A, cfg, k_max) : launch_persist<false, false>(grid, smem, st, A, cfg, k_max);
    if (le != cudaSuccess) {
        cudaGetLastError();
        ALQ_FAIL(ctx, ALQ_ERR_CUDA, "alq_greedy_select: cooperative launch failed: %s (grid %d, %zu B shared)", cudaGetErrorString(le), grid, smem);
    }
    ALQ_LAUNCH_CHECK(ctx);

    int status = 0;
    unsigned long long prof[4] = {};
    if (sample || comm || D->step_kernel_ms_host) {
        ALQ_CUDA(ctx, cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (D->step_kernel_ms_host) ALQ_CUDA(ctx, cudaMemcpyAsync(prof, d_prof, sizeof(prof), cudaMemcpyDeviceToHost, st));
        ALQ_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (D->step_kernel_ms_host) {
        const double steps = prof[2] ? static_cast<double>(prof[2]) : 1.0;
        D->step_kernel_ms_host[0] = static_cast<float>(prof[0] * 1e-6 / steps);
        D->step_kernel_ms_host[1] = static_cast<float>(prof[1] * 1e-6 / steps);
        D->step_kernel_ms_host[2] = static_cast<float>(prof[2]);
        D->step_kernel_ms_host[3] = 3.0f;
    }
    if (status == ALQ_ERR_STATE) ALQ_FAIL(ctx, status, "alq_greedy_select: timed out waiting for a peer GPU (or a CTA of this grid)");
    if (status != 0) ALQ_FAIL(ctx, status, "alq_greedy_select: non-finite or empty probability mass during D^2 sampling");
    return ALQ_OK;
}
