// Second pipe microbenchmark: true cost of 2-input VIMNMX, fp16 FMA-pipe min/max emulation
// (relu trick on denormal-coded integers), VABSDIFF4, HSET2; plus an exactness check of the
// fp16 trick over the value range the decoder uses.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

__device__ __forceinline__ unsigned h2u(__half2 h) { return *reinterpret_cast<unsigned *>(&h); }
__device__ __forceinline__ __half2 u2h(unsigned u) { return *reinterpret_cast<__half2 *>(&u); }

// min / max of non-negative integers < 1024 held as fp16 bit patterns (denormals), FMA pipe only
__device__ __forceinline__ unsigned fmin_f(unsigned a, unsigned b) {
    const __half2 one = __float2half2_rn(1.0f);
    __half2 r = __hfma2_relu(u2h(a), one, __hneg2(u2h(b)));      // relu(a - b)
    return h2u(__hsub2(u2h(a), r));                              // a - relu(a-b) = min(a,b)
}
__device__ __forceinline__ unsigned fmax_f(unsigned a, unsigned b) {
    const __half2 one = __float2half2_rn(1.0f);
    __half2 r = __hfma2_relu(u2h(a), one, __hneg2(u2h(b)));
    return h2u(__hadd2(u2h(b), r));                              // b + relu(a-b) = max(a,b)
}

template <int OP> __device__ __forceinline__ unsigned op(unsigned x, unsigned y, unsigned z) {
    if (OP == 0) return __vminu2(x, y) + z;            // VIMNMX + VIADD(FMA pipe): defeats min(min(x,y),y) folding
    if (OP == 1) return (x ^ y) + z;                   // LOP3 + VIADD reference
    if (OP == 2) return fmin_f(x, y) ^ z;              // 2 FMA-pipe ops + LOP3
    if (OP == 3) return __vabsdiffu4(x, y) + z;        // VABSDIFF4 + VIADD
    if (OP == 4) return __vabsdiffu4(x, y) ^ z;        // VABSDIFF4 + LOP3
    if (OP == 5) return __heq2_mask(u2h(x), u2h(y)) + z;   // HSET2 + VIADD
    if (OP == 6) return fmin_f(x, y) + z;              // 3 FMA-pipe ops
    if (OP == 7) return __vimin3_u16x2(x, y, z) + z;   // VIMNMX3 + VIADD
    if (OP == 8) return fmin_f(x, y);                  // 2 fp16 ops alone
    if (OP == 9) return fmin_f(x, y) * 3u + z;         // 2 fp16 ops + IMAD
    if (OP == 10) return __vadd2(fmin_f(x, y), z);     // 2 fp16 ops + VIADD.16x2
    if (OP == 11) return fmin_f(fmax_f(x, y), z);      // 4 fp16 ops
    if (OP == 12) return (fmin_f(x, y) * 3u + z) ^ y;  // 2 fp16 + IMAD + LOP3
    if (OP == 13) return __vadd2(x * 3u + z, y);       // IMAD + VIADD.16x2
    return x;
}

template <int OP> __global__ void bench(unsigned *out, unsigned seed, long long *cycles) {
    unsigned v[ILP];
    unsigned y = (seed + threadIdx.x) & 0x00ff00ff, z = (seed * 3 + 1) & 0x00010001;
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = (threadIdx.x * 7 + i) & 0x00ff00ff;
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

template <class K> void run(const char *name, K kern, int nsm) {
    unsigned *out; long long *cyc;
    cudaMalloc(&out, (size_t)nsm * 1024 * 4); cudaMalloc(&cyc, nsm * 8);
    kern<<<nsm, 1024>>>(out, 12345u, cyc);
    kern<<<nsm, 1024>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm;
    printf("%-34s %8.0f cycles  %.3f cycles per (op group) per warp per SM\n", name, avg, avg / (32.0 * ITERS * ILP));
    cudaFree(out); cudaFree(cyc);
}

__global__ void exactness(int *bad) {
    // all pairs a, b in [0, 512): fp16 relu trick must equal integer min / max, lane-wise
    const int a = blockIdx.x, b = threadIdx.x;
    const unsigned pa = a | ((511 - a) << 16), pb = b | ((b ^ 0x155) << 16);
    const unsigned mn = fmin_f(pa, pb), mx = fmax_f(pa, pb);
    const unsigned wmn = min(a, b) | (min(511 - a, b ^ 0x155) << 16);
    const unsigned wmx = max(a, b) | (max(511 - a, b ^ 0x155) << 16);
    if (mn != wmn || mx != wmx) atomicAdd(bad, 1);
    // signed accumulate: (a + (b - 256)) exact for small ints, with negative results in sign-magnitude
    __half2 s = __hadd2(u2h(pa & 0xffff), u2h((unsigned)b));
    if ((h2u(s) & 0xffff) != (unsigned)(a + b)) atomicAdd(bad + 1, 1);
}

int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    run("VIMNMX.U16x2 + VIADD", bench<0>, nsm);
    run("LOP3 + VIADD", bench<1>, nsm);
    run("fp16 min (HFMA2.RELU,HADD2) + LOP3", bench<2>, nsm);
    run("VABSDIFF4 + VIADD", bench<3>, nsm);
    run("VABSDIFF4 + LOP3", bench<4>, nsm);
    run("HSET2 + VIADD", bench<5>, nsm);
    run("fp16 min + VIADD (3 FMA ops)", bench<6>, nsm);
    run("VIMNMX3 + VIADD", bench<7>, nsm);
    run("2 fp16 ops alone", bench<8>, nsm);
    run("2 fp16 ops + IMAD", bench<9>, nsm);
    run("2 fp16 ops + VIADD.16x2", bench<10>, nsm);
    run("4 fp16 ops", bench<11>, nsm);
    run("2 fp16 + IMAD + LOP3", bench<12>, nsm);
    run("IMAD + VIADD.16x2", bench<13>, nsm);
    int *bad; cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
    exactness<<<512, 512>>>(bad);
    int h[2]; cudaMemcpy(h, bad, 8, cudaMemcpyDeviceToHost);
    printf("fp16 relu-trick min/max mismatches over [0,512)^2: %d ; hadd2 int-add mismatches: %d\n", h[0], h[1]);
    return 0;
}
