// K1b, cluster-resident: the whole stable top-B selection in ONE launch for pools that fit the shared
// memory of a thread-block cluster (n <= 8 * 32768 scores, b <= 16384).  Same contract and same result as
// the multi-kernel path in alq_select.cu (`torch.sort(scores).indices[:B]` of margin_sampler.py:42 under
// /root/reference/src/query_strategies, ties by lowest position).
//
// One cluster of 8 CTAs; every CTA keeps its slice of the order-preserving uint32 score keys in shared
// memory for the whole kernel.  Stages are separated by cluster barriers (~0.4 us) instead of kernel
// boundaries (~7 us each in the multi-kernel path, which is latency-bound: 320 KB of data):
//   1-3  three MSB radix levels (11+11+10 bits): per-CTA histogram (warp-aggregated shared atomics) ->
//        global partials -> barrier -> every CTA reduces the 8 partials and locates the bin itself;
//   4    per-CTA counts of {key < T} and {key == T} -> barrier -> exclusive offsets;
//   5    compaction of the B winners as 64-bit words (key << 32 | position); ties with T are taken in
//        position order by one warp walking the slice with ballots;
//   6    CTAs sort runs of 2048 words (bitonic, shared memory) -> barrier;
//   7    every CTA pulls all runs into shared memory and places its share of the words by rank counting.
#include "alq_common.cuh"

namespace selc {

constexpr int CL = 8;
constexpr int THREADS = 1024;
constexpr int MAX_SLICE = 32768;          // keys per CTA
constexpr int BINS = 2048;
constexpr int RUN = 2048;
constexpr int MAX_B = 16384;

struct Scratch {
    uint32_t* hist_part;        // [3][CL][BINS]
    uint32_t* counts;           // [CL][2]  (lt, eq)
    unsigned long long* words;  // [MAX_B]
};

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t score_key(float s) { return alq_ord(s + 0.0f); }

// histogram add with intra-warp aggregation: lanes hitting the same bin issue one shared atomic
__device__ __forceinline__ void hist_add(uint32_t* hist, bool active, uint32_t bin) {
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned peers = __match_any_sync(act, bin);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
}

__device__ __forceinline__ void cmpxchg(unsigned long long& a, unsigned long long& b, bool up) {
    if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
select_cluster_kernel(const float* __restrict__ scores, int n, int b, Scratch S, int32_t* __restrict__ out_pos) {
    extern __shared__ __align__(16) unsigned char smem_sel[];
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem_sel);                       // [slice]  (later: all runs)
    __shared__ uint32_t hist[BINS];
    __shared__ unsigned long long part[THREADS / 32];
    __shared__ uint32_t sh_prefix, sh_lt_cursor;
    __shared__ unsigned long long sh_k;
    __shared__ uint32_t sh_cnt[2];
    const int rank = static_cast<int>(cluster_rank());
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = ((n + CL - 1) / CL + 3) & ~3;
    const int lo = rank * per, hi = min(n, lo + per);
    const int cnt = max(0, hi - lo);

    // ---- stage 0: slice -> keys in shared memory ---------------------------------------------------------
    for (int i = threadIdx.x; i < cnt; i += THREADS) keys[i] = score_key(scores[lo + i]);
    if (threadIdx.x == 0) { sh_prefix = 0; sh_k = static_cast<unsigned long long>(b); }
    __syncthreads();

    // ---- stages 1-3: radix levels ---------------------------------------------------------------------------
#pragma unroll 1
    for (int level = 0; level < 3; ++level) {
        for (int i = threadIdx.x; i < BINS; i += THREADS) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = sh_prefix;
        const int iters = (cnt + THREADS - 1) / THREADS;
        for (int it = 0; it < iters; ++it) {
            const int i = it * THREADS + threadIdx.x;
            const uint32_t key = i < cnt ? keys[i] : 0u;
            bool act = i < cnt;
            uint32_t bin;
            if (level == 0) bin = key >> 21;
            else if (level == 1) { act = act && (key >> 21) == prefix; bin = (key >> 10) & 0x7ffu; }
            else { act = act && (key >> 10) == prefix; bin = key & 0x3ffu; }
            hist_add(hist, act, bin);
        }
        __syncthreads();
        uint32_t* mine = S.hist_part + (static_cast<size_t>(level) * CL + rank) * BINS;
        for (int i = threadIdx.x; i < BINS; i += THREADS) mine[i] = hist[i];
        cluster_sync_all();
        // every CTA reduces the CL partial histograms itself: 2 bins per thread
        unsigned long long c0 = 0, c1 = 0;
        for (int q = 0; q < CL; ++q) {
            const uint32_t* hp = S.hist_part + (static_cast<size_t>(level) * CL + q) * BINS;
            c0 += __ldcg(hp + 2 * threadIdx.x);
            c1 += __ldcg(hp + 2 * threadIdx.x + 1);
        }
        // block inclusive scan of (c0 + c1)
        unsigned long long inc = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) part[warp] = inc;
        __syncthreads();
        unsigned long long woff = 0;
        for (int w = 0; w < warp; ++w) woff += part[w];
        const unsigned long long incl = woff + inc, excl = incl - (c0 + c1);
        const unsigned long long k = sh_k;
        __syncthreads();
        if (excl < k && k <= incl) {
            const bool first = k <= excl + c0;
            const uint32_t bin = 2 * threadIdx.x + (first ? 0 : 1);
            sh_prefix = level == 2 ? ((prefix << 10) | bin) : ((prefix << 11) | bin);
            sh_k = first ? k - excl : k - excl - c0;
        }
        __syncthreads();
    }
    const uint32_t T = sh_prefix;
    const unsigned long long ties = sh_k;                       // how many keys equal to T are inside the budget
    const unsigned long long n_less = static_cast<unsigned long long>(b) - ties;

    // ---- stage 4: counts and offsets ----------------------------------------------------------------------------
    uint32_t lt = 0, eq = 0;
    for (int i = threadIdx.x; i < cnt; i += THREADS) {
        const uint32_t key = keys[i];
        lt += key < T;
        eq += key == T;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lt += __shfl_xor_sync(0xffffffffu, lt, o);
        eq += __shfl_xor_sync(0xffffffffu, eq, o);
    }
    if (threadIdx.x < 2) sh_cnt[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 0) { atomicAdd(&sh_cnt[0], lt); atomicAdd(&sh_cnt[1], eq); }
    __syncthreads();
    if (threadIdx.x < 2) S.counts[rank * 2 + threadIdx.x] = sh_cnt[threadIdx.x];
    cluster_sync_all();
    uint32_t lt_off = 0, eq_off = 0;
    for (int q = 0; q < rank; ++q) {
        lt_off += __ldcg(&S.counts[q * 2]);
        eq_off += __ldcg(&S.counts[q * 2 + 1]);
    }

    // ---- stage 5: compaction ---------------------------------------------------------------------------------------
    if (threadIdx.x == 0) sh_lt_cursor = 0;
    __syncthreads();
    if (warp == 0) {
        // ties: the first `ties` keys equal to T by global position; this CTA's share starts at rank eq_off
        unsigned long long rk = eq_off;
        for (int i0 = 0; i0 < cnt && rk < ties; i0 += 32) {
            const int i = i0 + lane;
            const bool is = i < cnt && keys[i] == T;
            const unsigned m = __ballot_sync(0xffffffffu, is);
            if (is) {
                const unsigned long long r = rk + __popc(m & ((1u << lane) - 1u));
                if (r < ties) S.words[n_less + r] = (static_cast<unsigned long long>(T) << 32) | static_cast<uint32_t>(lo + i);
            }
            rk += __popc(m);
        }
    } else {
        // keys below T: any order (they are sorted afterwards); slots from a block cursor, warp-aggregated
        for (int i0 = (warp - 1) * 32; i0 < cnt; i0 += (THREADS / 32 - 1) * 32) {
            const int i = i0 + lane;
            const uint32_t key = i < cnt ? keys[i] : 0xffffffffu;
            const bool is = i < cnt && key < T;
            const unsigned m = __ballot_sync(0xffffffffu, is);
            uint32_t base = 0;
            if (lane == 0 && m) base = atomicAdd(&sh_lt_cursor, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (is)
                S.words[lt_off + base + __popc(m & ((1u << lane) - 1u))] =
                    (static_cast<unsigned long long>(key) << 32) | static_cast<uint32_t>(lo + i);
        }
    }
    cluster_sync_all();

    // ---- stage 6: sort runs of RUN words ------------------------------------------------------------------------------
    unsigned long long* sk = reinterpret_cast<unsigned long long*>(smem_sel);      // the key slice is no longer needed
    const int runs = (b + RUN - 1) / RUN;
    for (int run = rank; run < runs; run += CL) {
        const int base = run * RUN;
        for (int i = threadIdx.x; i < RUN; i += THREADS) sk[i] = base + i < b ? __ldcg(&S.words[base + i]) : ~0ull;
        __syncthreads();
        const int t = threadIdx.x;
        for (int k = 2; k <= RUN; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int l = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                cmpxchg(sk[l], sk[l | j], (l & k) == 0);
                __syncthreads();
            }
        }
        for (int i = threadIdx.x; i < RUN; i += THREADS)
            if (base + i < b) S.words[base + i] = sk[i];
        __syncthreads();
    }
    cluster_sync_all();

    // ---- stage 7: all runs into shared memory, place this CTA's share by rank counting ----------------------------------
    for (int i = threadIdx.x; i < b; i += THREADS) sk[i] = __ldcg(&S.words[i]);
    __syncthreads();
    for (int i = rank * THREADS + threadIdx.x; i < b; i += CL * THREADS) {
        const unsigned long long key = sk[i];
        const int my_run = i / RUN;
        int rk = i - my_run * RUN;
        for (int r = 0; r < runs; ++r) {
            if (r == my_run) continue;
            const int len = min(RUN, b - r * RUN);
            int l = 0, h = len;
            while (l < h) {
                const int mid = (l + h) >> 1;
                if (sk[r * RUN + mid] < key) l = mid + 1; else h = mid;
            }
            rk += l;
        }
        out_pos[rk] = static_cast<int32_t>(key & 0xffffffffu);
    }
}

}  // namespace selc

// Returns ALQ_ERR_STATE when the problem does not fit this kernel (caller uses the multi-kernel path).
int alq_select_smallest_cluster(alq_ctx* ctx, const float* scores, int64_t n, int64_t b, int32_t* out_pos,
                                cudaStream_t st) {
    using namespace selc;
    if (n > static_cast<int64_t>(CL) * MAX_SLICE || b > MAX_B || n < 1 || b < 1) return ALQ_ERR_STATE;
    const size_t hist_bytes = static_cast<size_t>(3) * CL * BINS * sizeof(uint32_t);
    int rc = alq_scratch_reserve(ctx, scratch_need({hist_bytes, CL * 2 * sizeof(uint32_t), MAX_B * sizeof(unsigned long long)}));
    if (rc) return rc;
    ScratchCursor cur(ctx->scratch);
    Scratch S;
    S.hist_part = cur.take<uint32_t>(3 * CL * BINS);
    S.counts = cur.take<uint32_t>(CL * 2);
    S.words = cur.take<unsigned long long>(MAX_B);
    const int per = ((static_cast<int>(n) + CL - 1) / CL + 3) & ~3;
    size_t smem = std::max<size_t>(static_cast<size_t>(per) * 4, std::max<size_t>(static_cast<size_t>(b) * 8, RUN * 8)) + 64;
    if (smem > ctx->smem_optin - 24 * 1024) return ALQ_ERR_STATE;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        ALQ_CUDA(ctx, cudaFuncSetAttribute(select_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        attr_set = smem;
    }
    select_cluster_kernel<<<CL, THREADS, smem, st>>>(scores, static_cast<int>(n), static_cast<int>(b), S, out_pos);
    ALQ_LAUNCH_CHECK(ctx);
    return ALQ_OK;
}
