// Device-side scalar semantics of the five LLR types.
// Replaces the DecodeFrom impls of reference src/decoder.rs:42-86:
//   integers: saturating abs/add/sub;  floats: mask-abs, plain add/sub;
//   hard_bit is `< 0` (so -0.0 is a 0 bit);  maxval is T::MAX.
#pragma once
#include <cfloat>
#include <climits>
#include <cstdint>

namespace ldpc {

template <class T> struct Arith;

// Shared-memory storage type of a message: sub-word types are widened to 32 bits so that consecutive
// lanes hit consecutive banks (byte / half-word accesses serialise four / two lanes per bank).
template <class T> struct MsgStore { typedef T type; };
template <> struct MsgStore<int8_t> { typedef int32_t type; };
template <> struct MsgStore<int16_t> { typedef int32_t type; };

// i8 / i16 compute in 32-bit registers (type C = int): values are sign-extended once when loaded and every
// operation re-clamps to the narrow range, so no per-operation sign-extension instructions are needed.
template <> struct Arith<int8_t> {
    typedef int C;
    static constexpr const char *name = "i8";
    __device__ static __forceinline__ int zero() { return 0; }
    __device__ static __forceinline__ int one() { return 1; }
    __device__ static __forceinline__ int maxval() { return INT8_MAX; }
    __device__ static __forceinline__ int abs(int x) { return min(::abs(x), 127); }
    __device__ static __forceinline__ int sat_add(int a, int b) { return max(__viaddmin_s32(a, b, 127), -128); }
    __device__ static __forceinline__ int sat_sub(int a, int b) { return max(__viaddmin_s32(a, -b, 127), -128); }
    __device__ static __forceinline__ int neg(int x) { return -x; }
    __device__ static __forceinline__ int min(int a, int b) { return a < b ? a : b; }
    __device__ static __forceinline__ bool hard_bit(int x) { return x < 0; }
};

template <> struct Arith<int16_t> {
    typedef int C;
    static constexpr const char *name = "i16";
    __device__ static __forceinline__ int zero() { return 0; }
    __device__ static __forceinline__ int one() { return 1; }
    __device__ static __forceinline__ int maxval() { return INT16_MAX; }
    __device__ static __forceinline__ int abs(int x) { return min(::abs(x), 32767); }
    __device__ static __forceinline__ int sat_add(int a, int b) { return max(__viaddmin_s32(a, b, 32767), -32768); }
    __device__ static __forceinline__ int sat_sub(int a, int b) { return max(__viaddmin_s32(a, -b, 32767), -32768); }
    __device__ static __forceinline__ int neg(int x) { return -x; }
    __device__ static __forceinline__ int min(int a, int b) { return a < b ? a : b; }
    __device__ static __forceinline__ bool hard_bit(int x) { return x < 0; }
};

template <> struct Arith<int32_t> {
    typedef int32_t C;
    static constexpr const char *name = "i32";
    __device__ static __forceinline__ int32_t zero() { return 0; }
    __device__ static __forceinline__ int32_t one() { return 1; }
    __device__ static __forceinline__ int32_t maxval() { return INT32_MAX; }
    // |INT32_MIN| wraps to 0x80000000, which the unsigned minimum brings back to INT32_MAX (saturating_abs)
    __device__ static __forceinline__ int32_t abs(int32_t x) { return (int32_t)::min((uint32_t)::abs(x), (uint32_t)INT32_MAX); }
    __device__ static __forceinline__ int32_t sat_add(int32_t a, int32_t b) {
        int32_t r;
        asm("add.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
        return r;
    }
    __device__ static __forceinline__ int32_t sat_sub(int32_t a, int32_t b) {
        int32_t r;
        asm("sub.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
        return r;
    }
    __device__ static __forceinline__ int32_t neg(int32_t x) { return -x; }
    __device__ static __forceinline__ int32_t min(int32_t a, int32_t b) { return a < b ? a : b; }
    __device__ static __forceinline__ bool hard_bit(int32_t x) { return x < 0; }
};

template <> struct Arith<float> {
    typedef float C;
    static constexpr const char *name = "f32";
    __device__ static __forceinline__ float zero() { return 0.0f; }
    __device__ static __forceinline__ float one() { return 1.0f; }
    __device__ static __forceinline__ float maxval() { return FLT_MAX; }
    // fabsf clears the sign bit exactly like the reference's mask (:60-63) and folds into an operand modifier
    __device__ static __forceinline__ float abs(float x) { return fabsf(x); }
    // a < b ? a : b for the minima over |v| (never -0; NaN LLRs are outside the contract, include/labrador_ldpc.h):
    // one FMNMX instead of FSETP + FSEL
    __device__ static __forceinline__ float min(float a, float b) { return fminf(a, b); }
    __device__ static __forceinline__ float sat_add(float a, float b) { return __fadd_rn(a, b); }
    __device__ static __forceinline__ float sat_sub(float a, float b) { return __fsub_rn(a, b); }
    __device__ static __forceinline__ float neg(float x) { return -x; }
    __device__ static __forceinline__ bool hard_bit(float x) { return x < 0.0f; }
};

template <> struct Arith<double> {
    typedef double C;
    static constexpr const char *name = "f64";
    __device__ static __forceinline__ double zero() { return 0.0; }
    __device__ static __forceinline__ double one() { return 1.0; }
    __device__ static __forceinline__ double maxval() { return DBL_MAX; }
    __device__ static __forceinline__ double abs(double x) { return fabs(x); }
    // no native f64 min/max instruction: the compare-and-select form is the shorter one
    __device__ static __forceinline__ double min(double a, double b) { return a < b ? a : b; }
    __device__ static __forceinline__ double sat_add(double a, double b) { return __dadd_rn(a, b); }
    __device__ static __forceinline__ double sat_sub(double a, double b) { return __dsub_rn(a, b); }
    __device__ static __forceinline__ double neg(double x) { return -x; }
    __device__ static __forceinline__ bool hard_bit(double x) { return x < 0.0; }
};

}  // namespace ldpc
