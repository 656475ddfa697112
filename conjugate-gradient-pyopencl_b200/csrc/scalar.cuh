// scalar.cuh -- value-type traits for the four instantiations of the CG path.
//
// The reference has two: float (kernel/real/*.cl) and cfloat = float2 with
// cadd/csub/cmul (kernel/complex/cmplx.h:4-25).  The engine adds the double twins.
// Complex products are NEVER conjugated (kernel/complex/vdot.cl:15): with a
// complex-symmetric matrix this is COCG, which is what the reference computes.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace cgb {

template <typename T> struct Sc;

template <> struct Sc<float> {
    using real = float;
    static constexpr bool cplx = false;
    static constexpr int dtype = 0;
    __host__ __device__ static inline float zero() { return 0.f; }
    __host__ __device__ static inline float add(float a, float b) { return a + b; }
    __host__ __device__ static inline float sub(float a, float b) { return a - b; }
    __host__ __device__ static inline float mul(float a, float b) { return a * b; }
    // c + a*b
    __host__ __device__ static inline float fma(float a, float b, float c) { return fmaf(a, b, c); }
    // c - a*b
    __host__ __device__ static inline float fnma(float a, float b, float c) { return fmaf(-a, b, c); }
    __host__ __device__ static inline float div(float a, float b) { return a / b; }
    __host__ __device__ static inline double abs(float a) { return fabs((double)a); }
    __host__ __device__ static inline bool is_zero(float a) { return a == 0.f; }
    __host__ __device__ static inline bool finite(float a) { return isfinite(a); }
    __host__ __device__ static inline void to_double2(float a, double *o) { o[0] = a; }
};

template <> struct Sc<double> {
    using real = double;
    static constexpr bool cplx = false;
    static constexpr int dtype = 1;
    __host__ __device__ static inline double zero() { return 0.0; }
    __host__ __device__ static inline double add(double a, double b) { return a + b; }
    __host__ __device__ static inline double sub(double a, double b) { return a - b; }
    __host__ __device__ static inline double mul(double a, double b) { return a * b; }
    __host__ __device__ static inline double fma(double a, double b, double c) { return ::fma(a, b, c); }
    __host__ __device__ static inline double fnma(double a, double b, double c) { return ::fma(-a, b, c); }
    __host__ __device__ static inline double div(double a, double b) { return a / b; }
    __host__ __device__ static inline double abs(double a) { return fabs(a); }
    __host__ __device__ static inline bool is_zero(double a) { return a == 0.0; }
    __host__ __device__ static inline bool finite(double a) { return isfinite(a); }
    __host__ __device__ static inline void to_double2(double a, double *o) { o[0] = a; }
};

// Complex types: float2 / double2 with x = re, y = im -- the memory layout of
// `cfloat` (cmplx.h:4), C's `float complex` and numpy's csingle/cdouble.
template <typename V, typename R, int DT> struct ScC {
    using real = R;
    static constexpr bool cplx = true;
    static constexpr int dtype = DT;
    __host__ __device__ static inline V make(R re, R im) { V v; v.x = re; v.y = im; return v; }
    __host__ __device__ static inline V zero() { return make(R(0), R(0)); }
    __host__ __device__ static inline V add(V a, V b) { return make(a.x + b.x, a.y + b.y); }
    __host__ __device__ static inline V sub(V a, V b) { return make(a.x - b.x, a.y - b.y); }
    __host__ __device__ static inline V mul(V a, V b) {           // cmul, cmplx.h:20-25
        return make(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
    __host__ __device__ static inline V fma(V a, V b, V c) {      // c + a*b
        R re = c.x + a.x * b.x;  re = re - a.y * b.y;
        R im = c.y + a.x * b.y;  im = im + a.y * b.x;
        return make(re, im);
    }
    __host__ __device__ static inline V fnma(V a, V b, V c) {     // c - a*b
        R re = c.x - a.x * b.x;  re = re + a.y * b.y;
        R im = c.y - a.x * b.y;  im = im - a.y * b.x;
        return make(re, im);
    }
    // alpha = delta / (d.q), beta = delta_new / delta_old: the host-side C99 complex
    // division of clcg.c:326,390.  Smith's algorithm: no overflow in c^2 + d^2.
    __host__ __device__ static inline V div(V a, V b) {
        if (fabs((double)b.x) >= fabs((double)b.y)) {
            R t = b.y / b.x, den = b.x + b.y * t;
            return make((a.x + a.y * t) / den, (a.y - a.x * t) / den);
        } else {
            R t = b.x / b.y, den = b.x * t + b.y;
            return make((a.x * t + a.y) / den, (a.y * t - a.x) / den);
        }
    }
    __host__ __device__ static inline double abs(V a) { return hypot((double)a.x, (double)a.y); }
    __host__ __device__ static inline bool is_zero(V a) { return a.x == R(0) && a.y == R(0); }
    __host__ __device__ static inline bool finite(V a) { return isfinite(a.x) && isfinite(a.y); }
    __host__ __device__ static inline void to_double2(V a, double *o) { o[0] = a.x; o[1] = a.y; }
};

template <> struct Sc<float2> : ScC<float2, float, 2> {};
template <> struct Sc<double2> : ScC<double2, double, 3> {};

// Number of T that fit one 128-bit access.
template <typename T> struct VecW { static constexpr int value = 16 / (int)sizeof(T); };

// A 16-byte bundle of V values of T, for LDG.128 / STG.128.
template <typename T, int V> struct alignas(sizeof(T) * V) Pack { T v[V]; };

}  // namespace cgb
