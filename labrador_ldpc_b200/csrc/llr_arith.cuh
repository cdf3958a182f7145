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

template <> struct Arith<int8_t> {
    static constexpr const char *name = "i8";
    __device__ static __forceinline__ int8_t zero() { return 0; }
    __device__ static __forceinline__ int8_t one() { return 1; }
    __device__ static __forceinline__ int8_t maxval() { return INT8_MAX; }
    __device__ static __forceinline__ int8_t abs(int8_t x) { return (int8_t)min(::abs((int)x), 127); }
    __device__ static __forceinline__ int8_t sat_add(int8_t a, int8_t b) { return (int8_t)max(-128, min(127, (int)a + (int)b)); }
    __device__ static __forceinline__ int8_t sat_sub(int8_t a, int8_t b) { return (int8_t)max(-128, min(127, (int)a - (int)b)); }
    __device__ static __forceinline__ int8_t neg(int8_t x) { return (int8_t)(-(int)x); }
    __device__ static __forceinline__ bool hard_bit(int8_t x) { return x < 0; }
};

template <> struct Arith<int16_t> {
    static constexpr const char *name = "i16";
    __device__ static __forceinline__ int16_t zero() { return 0; }
    __device__ static __forceinline__ int16_t one() { return 1; }
    __device__ static __forceinline__ int16_t maxval() { return INT16_MAX; }
    __device__ static __forceinline__ int16_t abs(int16_t x) { return (int16_t)min(::abs((int)x), 32767); }
    __device__ static __forceinline__ int16_t sat_add(int16_t a, int16_t b) { return (int16_t)max(-32768, min(32767, (int)a + (int)b)); }
    __device__ static __forceinline__ int16_t sat_sub(int16_t a, int16_t b) { return (int16_t)max(-32768, min(32767, (int)a - (int)b)); }
    __device__ static __forceinline__ int16_t neg(int16_t x) { return (int16_t)(-(int)x); }
    __device__ static __forceinline__ bool hard_bit(int16_t x) { return x < 0; }
};

template <> struct Arith<int32_t> {
    static constexpr const char *name = "i32";
    __device__ static __forceinline__ int32_t zero() { return 0; }
    __device__ static __forceinline__ int32_t one() { return 1; }
    __device__ static __forceinline__ int32_t maxval() { return INT32_MAX; }
    __device__ static __forceinline__ int32_t clamp64(long long x) {
        return (int32_t)(x < (long long)INT32_MIN ? (long long)INT32_MIN : (x > (long long)INT32_MAX ? (long long)INT32_MAX : x));
    }
    __device__ static __forceinline__ int32_t abs(int32_t x) { return x == INT32_MIN ? INT32_MAX : (x < 0 ? -x : x); }
    __device__ static __forceinline__ int32_t sat_add(int32_t a, int32_t b) { return clamp64((long long)a + (long long)b); }
    __device__ static __forceinline__ int32_t sat_sub(int32_t a, int32_t b) { return clamp64((long long)a - (long long)b); }
    __device__ static __forceinline__ int32_t neg(int32_t x) { return -x; }
    __device__ static __forceinline__ bool hard_bit(int32_t x) { return x < 0; }
};

template <> struct Arith<float> {
    static constexpr const char *name = "f32";
    __device__ static __forceinline__ float zero() { return 0.0f; }
    __device__ static __forceinline__ float one() { return 1.0f; }
    __device__ static __forceinline__ float maxval() { return FLT_MAX; }
    __device__ static __forceinline__ float abs(float x) { return __uint_as_float(__float_as_uint(x) & 0x7FFFFFFFu); }
    __device__ static __forceinline__ float sat_add(float a, float b) { return __fadd_rn(a, b); }
    __device__ static __forceinline__ float sat_sub(float a, float b) { return __fsub_rn(a, b); }
    __device__ static __forceinline__ float neg(float x) { return -x; }
    __device__ static __forceinline__ bool hard_bit(float x) { return x < 0.0f; }
};

template <> struct Arith<double> {
    static constexpr const char *name = "f64";
    __device__ static __forceinline__ double zero() { return 0.0; }
    __device__ static __forceinline__ double one() { return 1.0; }
    __device__ static __forceinline__ double maxval() { return DBL_MAX; }
    __device__ static __forceinline__ double abs(double x) {
        return __longlong_as_double(__double_as_longlong(x) & 0x7FFFFFFFFFFFFFFFll);
    }
    __device__ static __forceinline__ double sat_add(double a, double b) { return __dadd_rn(a, b); }
    __device__ static __forceinline__ double sat_sub(double a, double b) { return __dsub_rn(a, b); }
    __device__ static __forceinline__ double neg(double x) { return -x; }
    __device__ static __forceinline__ bool hard_bit(double x) { return x < 0.0; }
};

}  // namespace ldpc
