// SASS probe of the packed SIMD intrinsics on sm_100a (profiles/r02_tm_variants.md, section (ii)):
//   nvcc -gencode arch=compute_100a,code=sm_100a -cubin -o /tmp/p.cubin tools/ubench/simd_probe.cu && cuobjdump -sass /tmp/p.cubin
// One intrinsic per kernel; count the instructions between the loads and the store.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#define K(name, expr) extern "C" __global__ void k_##name(uint32_t *o, const uint32_t *a, const uint32_t *b, const uint32_t *c){ uint32_t x=a[threadIdx.x], y=b[threadIdx.x], z=c[threadIdx.x]; o[threadIdx.x] = expr; }
K(vminu4, __vminu4(x,y))
K(vmaxu4, __vmaxu4(x,y))
K(vmins4, __vmins4(x,y))
K(vmaxs4, __vmaxs4(x,y))
K(vcmpeq4, __vcmpeq4(x,y))
K(vsetgeu4, __vsetgeu4(x,y))
K(vcmpgeu4, __vcmpgeu4(x,y))
K(vcmpgtu4, __vcmpgtu4(x,y))
K(vaddus4, __vaddus4(x,y))
K(vsubus4, __vsubus4(x,y))
K(vaddss4, __vaddss4(x,y))
K(vsubss4, __vsubss4(x,y))
K(vabsdiffu4, __vabsdiffu4(x,y))
K(vabsdiffs4, __vabsdiffs4(x,y))
K(vabsss4, __vabsss4(x))
K(vabs4, __vabs4(x))
K(vneg4, __vneg4(x))
K(vadd4, __vadd4(x,y))
K(vsub4, __vsub4(x,y))
K(vavgu4, __vavgu4(x,y))
K(vhaddu4, __vhaddu4(x,y))
K(vsadu4, __vsadu4(x,y))
K(vminu2, __vminu2(x,y))
K(vimin3u, __vimin3_u16x2(x,y,z))
K(vimax3u, __vimax3_u16x2(x,y,z))
K(viaddmin, __viaddmin_s16x2_relu(x,y,z))
K(viaddmax, __viaddmax_s16x2(x,y,z))
K(vibmin, ({bool p,q; uint32_t r=__vibmin_u16x2(x,y,&p,&q); r + (p?1:0) + (q?2:0);}))
K(vaddus2, __vaddus2(x,y))
K(vsubus2, __vsubus2(x,y))
K(vaddss2, __vaddss2(x,y))
K(vcmpgeu2, __vcmpgeu2(x,y))
K(vabsdiffu2, __vabsdiffu2(x,y))

__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
K(absmin, h2u(__hmin2(__habs2(u2h(x)), __habs2(u2h(y)))))
K(absmin3, h2u(__hmin2(__hmin2(__habs2(u2h(x)), __habs2(u2h(y))), __habs2(u2h(z)))))
K(subabsmin, h2u(__hmin2(__habs2(__hsub2(u2h(x), u2h(0x007f007fu))), __habs2(__hsub2(u2h(y), u2h(0x007f007fu))))))
K(split_satadd, __vimin_s16x2_relu(__vadd2(x, y), 0x00ff00ffu))
