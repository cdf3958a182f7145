// Throughput microbenchmark for the packed-16-bit / logic instructions the min-sum kernel
// is built from.  Each kernel runs a long chain of independent instruction streams per thread
// (ILP 8) on every SM with 1024 threads/SM and reports warp-instructions per cycle per SM.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP> __device__ __forceinline__ unsigned op(unsigned x, unsigned y, unsigned z) {
    if (OP == 0) return __viaddmin_s16x2_relu(x, y, 0x00ff00ff);
    if (OP == 1) return __vimin3_s16x2(x, y, z);
    if (OP == 2) return __vmins2(x, y);
    if (OP == 3) return __vadd2(x, y);
    if (OP == 4) return __byte_perm(x, y, 0xbb99);
    if (OP == 5) { unsigned r; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; }
    if (OP == 6) return __funnelshift_l(x, x, y);
    if (OP == 7) return x * 3u + y;                                   // IMAD
    if (OP == 8) { __half2 a = *(__half2*)&x, b = *(__half2*)&y; __half2 r = __hmin2(a, b); return *(unsigned*)&r; }
    if (OP == 9) { __half2 a = *(__half2*)&x, b = *(__half2*)&y; return __heq2_mask(a, b) ^ x; }
    if (OP == 10) { __half2 a = *(__half2*)&x, b = *(__half2*)&y; __half2 r = __hadd2(a, b); return *(unsigned*)&r; }
    if (OP == 11) return x + y;                                       // IADD3
    if (OP == 12) return min((int)x, (int)y);                         // IMNMX / VIMNMX 32
    if (OP == 13) return __vimax_s16x2_relu(x, y);
    return x;
}

template <int OP> __global__ void bench(unsigned *out, unsigned seed, long long *cycles) {
    unsigned v[ILP];
    unsigned y = seed + threadIdx.x, z = seed * 3 + 1;
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 7 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = op<OP>(v[i], y, z);
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// mixes: two different ops interleaved to see whether they issue on different pipes
template <int OPA, int OPB> __global__ void bench2(unsigned *out, unsigned seed, long long *cycles) {
    unsigned v[ILP];
    unsigned y = seed + threadIdx.x, z = seed * 3 + 1;
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 7 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i += 2) { v[i] = op<OPA>(v[i], y, z); v[i + 1] = op<OPB>(v[i + 1], y, z); }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <class K> void run(const char *name, K kern, int nsm) {
    unsigned *out; long long *cyc;
    cudaMalloc(&out, (size_t)nsm * 1024 * 4); cudaMalloc(&cyc, nsm * 8);
    kern<<<nsm, 1024>>>(out, 12345u, cyc);
    kern<<<nsm, 1024>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm;
    double winst = 32.0 * ITERS * ILP;   // warp-instructions per SM (32 warps)
    printf("%-28s %8.0f cycles  %.2f warp-inst/cycle/SM\n", name, avg, winst / avg);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", nsm);
    run("VIADDMNMX.S16x2.RELU", bench<0>, nsm);
    run("VIMNMX3.S16x2", bench<1>, nsm);
    run("VIMNMX.S16x2", bench<2>, nsm);
    run("VIADD.16x2", bench<3>, nsm);
    run("PRMT", bench<4>, nsm);
    run("LOP3", bench<5>, nsm);
    run("SHF", bench<6>, nsm);
    run("IMAD", bench<7>, nsm);
    run("HMNMX2", bench<8>, nsm);
    run("HSET2+LOP3", bench<9>, nsm);
    run("HADD2", bench<10>, nsm);
    run("IADD3", bench<11>, nsm);
    run("IMNMX", bench<12>, nsm);
    run("VIMNMX.S16x2.RELU", bench<13>, nsm);
    run("mix VIADDMNMX + IMAD", bench2<0, 7>, nsm);
    run("mix VIADDMNMX + LOP3", bench2<0, 5>, nsm);
    run("mix VIADDMNMX + HADD2", bench2<0, 10>, nsm);
    run("mix VIADDMNMX + HMNMX2", bench2<0, 8>, nsm);
    run("mix LOP3 + IMAD", bench2<5, 7>, nsm);
    run("mix LOP3 + HADD2", bench2<5, 10>, nsm);
    run("mix LOP3 + PRMT", bench2<5, 4>, nsm);
    run("mix LOP3 + VIMNMX", bench2<5, 2>, nsm);
    run("mix LOP3 + HMNMX2", bench2<5, 8>, nsm);
    run("mix VIMNMX3 + HMNMX2", bench2<1, 8>, nsm);
    run("mix VIADD + IMAD", bench2<3, 7>, nsm);
    run("mix IADD3 + LOP3", bench2<11, 5>, nsm);
    run("mix IADD3 + IMAD", bench2<11, 7>, nsm);
    return 0;
}
