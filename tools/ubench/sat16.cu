// Does VIADDMNMX.S16x2 add with a wider-than-16-bit intermediate?  (development probe)
#include <cstdio>
#include <cstdint>
__global__ void k(uint32_t *o) {
    o[0] = __viaddmin_s16x2(0x7fff7fffu, 0x00010001u, 0x7fff7fffu);      // 32767 + 1 -> min(.., 32767)
    o[1] = __viaddmax_s16x2(0x80008000u, 0xffffffffu, 0x80008000u);      // -32768 - 1 -> max(.., -32768)
    o[2] = __viaddmin_s16x2(0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu);
    o[3] = __viaddmin_u16x2(0xffffffffu, 0x00020002u, 0xffffffffu);
    o[4] = __viaddmin_s16x2_relu(0x7fff7fffu, 0x00010001u, 0x7fff7fffu);
}
int main() {
    uint32_t *d, h[5];
    cudaMalloc(&d, 20); k<<<1, 1>>>(d); cudaMemcpy(h, d, 20, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 5; i++) printf("%d: %08x\n", i, h[i]);
    return 0;
}
