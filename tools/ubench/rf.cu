// Register-file bandwidth probe: ALU (LOP3) and FMA-pipe (IMAD) instructions with 1, 2 or 3 DISTINCT
// register source operands (no operand-reuse between consecutive instructions), alone and interleaved.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define N 12

template <int MODE> __global__ void k(unsigned *out, unsigned seed, long long *cycles) {
    unsigned a[N], b[N], c[N];
#pragma unroll
    for (int i = 0; i < N; i++) { a[i] = threadIdx.x * 3 + i + seed; b[i] = threadIdx.x * 5 + i * 7 + seed; c[i] = threadIdx.x + i * 11 + seed; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int j = (i + 1) % N, l = (i + 5) % N;
            if (MODE == 0) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[j]), "r"(c[l])); }                       // ALU, 3 regs
            if (MODE == 1) { asm volatile("lop3.b32 %0, %0, %1, 0x12345, 0x96;" : "+r"(a[i]) : "r"(b[j])); }                           // ALU, 2 regs
            if (MODE == 2) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[j]), "r"(c[l])); }                        // FMA, 3 regs
            if (MODE == 3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[j]), "r"(c[l]));
                             asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c[j]), "r"(a[l])); }                        // ALU3 + FMA3
            if (MODE == 4) { asm volatile("lop3.b32 %0, %0, %1, 0x12345, 0x96;" : "+r"(a[i]) : "r"(b[j]));
                             asm volatile("mad.lo.u32 %0, %0, %1, 0x77;" : "+r"(b[i]) : "r"(c[j])); }                                  // ALU2 + FMA2
            if (MODE == 5) { asm volatile("lop3.b32 %0, %0, 0x54321, 0x12345, 0x96;" : "+r"(a[i]));
                             asm volatile("mad.lo.u32 %0, %0, 0x11, 0x77;" : "+r"(b[i])); }                                           // ALU1 + FMA1
            if (MODE == 6) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[j]), "r"(c[l]));
                             asm volatile("mad.lo.u32 %0, %0, %1, 0x77;" : "+r"(b[i]) : "r"(c[j])); }                                  // ALU3 + FMA2
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) acc ^= a[i] ^ b[i] ^ c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <class K> void run(const char *name, K kern, int per_iter) {
    unsigned *out; long long *cyc; int nsm = 148;
    cudaMalloc(&out, (size_t)nsm * 1024 * 4); cudaMalloc(&cyc, nsm * 8);
    kern<<<nsm, 1024>>>(out, 1u, cyc); kern<<<nsm, 1024>>>(out, 1u, cyc);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm;
    printf("%-22s %.3f warp-instructions per cycle per SMSP\n", name, 8.0 * ITERS * N * per_iter / avg);
}
int main() {
    run("ALU 3 regs", k<0>, 1); run("ALU 2 regs", k<1>, 1); run("FMA 3 regs", k<2>, 1);
    run("ALU3 + FMA3", k<3>, 2); run("ALU2 + FMA2", k<4>, 2); run("ALU1 + FMA1", k<5>, 2); run("ALU3 + FMA2", k<6>, 2);
    return 0;
}
