// Latency microbenchmarks behind the design of the persistent selection kernel (never a bench number):
//   1. LL-word ping-pong between two CTAs (scope sys / gpu),  2. all-to-all publish+poll over the grid,
//   3. a grid barrier by atomic counter,  4. fp64 division throughput.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/microbench_ll tools/microbench_ll.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ld_sys(const void* p) { unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_gpu(const void* p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_sys(void* p, unsigned long long v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_gpu(void* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <bool SYS>
__global__ void pingpong(unsigned long long* w, int iters, int partner, unsigned long long* out) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        const unsigned long long t0 = gt();
        for (int i = 1; i <= iters; ++i) {
            if (SYS) st_sys(w, i); else st_gpu(w, i);
            while ((SYS ? ld_sys(w + 16) : ld_gpu(w + 16)) != (unsigned long long)i) {}
        }
        out[0] = gt() - t0;
    } else if ((int)blockIdx.x == partner) {
        for (int i = 1; i <= iters; ++i) {
            while ((SYS ? ld_sys(w) : ld_gpu(w)) != (unsigned long long)i) {}
            if (SYS) st_sys(w + 16, i); else st_gpu(w + 16, i);
        }
    }
}

// every CTA publishes `per` words, every CTA polls all gridDim.x * per words with `nthreads` threads
template <bool SYS>
__global__ void alltoall(unsigned long long* w, int iters, int per, int nthreads, unsigned long long* out) {
    const int total = gridDim.x * per;
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) t0 = gt();
    for (int it = 1; it <= iters; ++it) {
        unsigned long long* slot = w + (it & 1) * total;
        if ((int)threadIdx.x < per) { if (SYS) st_sys(slot + blockIdx.x * per + threadIdx.x, it); else st_gpu(slot + blockIdx.x * per + threadIdx.x, it); }
        if ((int)threadIdx.x < nthreads)
            for (int i = threadIdx.x; i < total; i += nthreads)
                while ((SYS ? ld_sys(slot + i) : ld_gpu(slot + i)) != (unsigned long long)it) {}
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = gt() - t0;
}

__global__ void gridbar(unsigned int* ctr, int iters, unsigned long long* out) {
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) t0 = gt();
    for (int it = 1; it <= iters; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ctr, 1u);
            unsigned int v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while ((int)(v - (unsigned)it * gridDim.x) < 0);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = gt() - t0;
}

__global__ void f64div(double* o, int iters, unsigned long long* out) {
    double a = 1.0 + threadIdx.x * 1e-3, b = 3.0 + blockIdx.x;
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) t0 = gt();
    for (int i = 0; i < iters; ++i) a = a / b + 1.0;
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = gt() - t0;
    o[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

int main() {
    unsigned long long *w, *out, h;
    unsigned int* ctr;
    double* o;
    cudaMalloc(&w, 1 << 22); cudaMalloc(&out, 64); cudaMalloc(&ctr, 64); cudaMalloc(&o, 148 * 512 * 8);
    const int iters = 2000;
    for (int partner : {1, 2, 37, 74, 147}) {
        for (int sys = 0; sys < 2; ++sys) {
            cudaMemset(w, 0, 1 << 22);
            if (sys) pingpong<true><<<148, 32>>>(w, iters, partner, out); else pingpong<false><<<148, 32>>>(w, iters, partner, out);
            cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("pingpong partner %3d scope %s: %.0f ns per round trip (2 one-way hops)\n", partner, sys ? "sys" : "gpu", (double)h / iters);
        }
    }
    for (int per : {1, 7}) for (int nth : {32, 512}) for (int sys = 0; sys < 2; ++sys) {
        cudaMemset(w, 0, 1 << 22);
        void* args[] = {&w, (void*)&iters, &per, &nth, &out};
        cudaLaunchCooperativeKernel(sys ? (void*)alltoall<true> : (void*)alltoall<false>, dim3(148), dim3(512), args, 0, 0);
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("all-to-all 148 CTAs x %d words, %3d polling threads, scope %s: %.0f ns per exchange  (%s)\n", per, nth, sys ? "sys" : "gpu", (double)h / iters, cudaGetErrorString(cudaGetLastError()));
    }
    {
        cudaMemset(ctr, 0, 64);
        void* args[] = {&ctr, (void*)&iters, &out};
        cudaLaunchCooperativeKernel((void*)gridbar, dim3(148), dim3(512), args, 0, 0);
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("grid barrier (atomicAdd + acquire poll), 148 CTAs: %.0f ns\n", (double)h / iters);
    }
    f64div<<<148, 512>>>(o, 256, out);
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("fp64 a = a / b + 1, 256 dependent iterations, 512 threads per SM: %.1f ns per iteration per thread-wave (%.0f ns total)\n", (double)h / 256, (double)h);
    return 0;
}
