// oracle/clref/clref.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's OWN OpenCL kernels (kernel/real/*.cl, kernel/complex/*.cl, cmplx.h), executed on the CPU.
//
// Nothing in this image can run OpenCL (no ICD, no headers) and clcg.c cannot be compiled (CL/cl.h, BeBOP), but the
// ten kernels are plain OpenCL C 1.2 and small: this file gives g++ the few things OpenCL C has and C++ lacks and
// then #includes the kernel sources FROM /root/reference AT BUILD TIME (-I/root/reference; nothing is copied into
// the repository; the product of the build goes to oracle/_ref/, which is git-ignored):
//
//   __kernel / __global / __local / __constant      empty macros
//   get_global_id / get_local_id / get_group_id / get_global_size    read the state of the running work-item
//   barrier(CLK_LOCAL_MEM_FENCE)                    every work-item of a work-group is a ucontext fiber; a barrier
//                                                   switches back to a scheduler that resumes the items in order,
//                                                   so all of them reach the barrier before any of them passes it
//   float2 and the vector literal (float2)(a, b)    x and y are of a class type `Float` (one float, float
//                                                   arithmetic) whose comma operator builds a pair that float2
//                                                   converts from; (float2)(0.0f, 0.0f) -- a comma of built-in
//                                                   floats -- reaches the one-scalar constructor, which replicates
//                                                   the scalar as OpenCL does
//   -D N_RHS=k -D WAVE_SIZE=32 -D WG_SIZE=256       (clcg.c:81-84) every kernel is included once per k = 1..4
//
// Around the kernels, the launch sequence of clcg.c: geometry :124-135, initialisation :253-292, loop :296-419,
// the sequential host sums of the per-work-group partials and alpha = delta/dq, beta = delta_new/delta_old in
// float / float complex on the host.  That sequence is restated here (clcg.c itself needs the OpenCL host API);
// the ARITHMETIC INSIDE THE KERNELS is the reference's own source.
//
// Used by tests/test_oracle.py (marked `reference`: build container only) to check that oracle/cpu_ref.c computes
// bit for bit what these kernels compute, and by oracle/make_golden.py to store their results as fixtures.
#include <ucontext.h>

#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// ---------------------------------------------------------------------------------------------------------
// OpenCL C for g++
// ---------------------------------------------------------------------------------------------------------
#define __kernel
#define __global
#define __local
#define __constant
#define __private
#define CLK_LOCAL_MEM_FENCE 1
#define WAVE_SIZE 32
#define WG_SIZE 256

struct FloatPair;
struct Float {
    float v;
    Float() : v(0.0f) {}
    Float(float f) : v(f) {}
    operator float() const { return v; }
};
struct FloatPair {
    float a, b;
};
static inline Float operator+(Float p, Float q) { return Float(p.v + q.v); }
static inline Float operator-(Float p, Float q) { return Float(p.v - q.v); }
static inline Float operator*(Float p, Float q) { return Float(p.v * q.v); }
static inline FloatPair operator,(Float p, Float q) { return FloatPair{p.v, q.v}; }

struct float2 {
    Float x, y;
    float2() {}
    float2(float s) : x(s), y(s) {}                 // (float2)(s): the scalar replicated
    float2(FloatPair p) : x(p.a), y(p.b) {}         // (float2)(a, b)
};
static inline float2 operator+(float2 p, float2 q) { return float2(FloatPair{p.x.v + q.x.v, p.y.v + q.y.v}); }
static_assert(sizeof(float2) == 2 * sizeof(float), "float2 is two packed floats, like cfloat on the device");

// ---- work-items as fibers --------------------------------------------------------------------------------
namespace wi {
constexpr int MAXLOCAL = 256;
constexpr size_t STACK = 64 * 1024;
static ucontext_t sched, item[MAXLOCAL];
static char *stacks = nullptr;
static bool done[MAXLOCAL];
static int cur_local, cur_group, local_size, global_size;
static void (*body)();

static void trampoline() {
    body();
    done[cur_local] = true;
    swapcontext(&item[cur_local], &sched);
}

// one work-group: every item runs until its next barrier (or its end), in local-id order, round after round
static void run_group(int group) {
    if (!stacks) stacks = (char *)malloc(STACK * MAXLOCAL);
    cur_group = group;
    for (int l = 0; l < local_size; l++) {
        getcontext(&item[l]);
        item[l].uc_stack.ss_sp = stacks + STACK * l;
        item[l].uc_stack.ss_size = STACK;
        item[l].uc_link = nullptr;
        makecontext(&item[l], trampoline, 0);
        done[l] = false;
    }
    for (int live = local_size; live > 0;) {
        live = 0;
        for (int l = 0; l < local_size; l++) {
            if (done[l]) continue;
            cur_local = l;
            swapcontext(&sched, &item[l]);
            live += !done[l];
        }
    }
}

template <typename F> static void ndrange(size_t global, size_t local, F f) {
    static F *fp;
    fp = &f;
    body = [] { (*fp)(); };
    local_size = (int)local;
    global_size = (int)global;
    for (size_t g = 0; g < global / local; g++) run_group((int)g);
}
}  // namespace wi

static inline int get_local_id(int) { return wi::cur_local; }
static inline int get_group_id(int) { return wi::cur_group; }
static inline int get_global_id(int) { return wi::cur_group * wi::local_size + wi::cur_local; }
static inline int get_global_size(int) { return wi::global_size; }
static inline int get_local_size(int) { return wi::local_size; }
static inline void barrier(int) { swapcontext(&wi::item[wi::cur_local], &wi::sched); }

// ---------------------------------------------------------------------------------------------------------
// the reference's kernel sources, once per N_RHS
// ---------------------------------------------------------------------------------------------------------
#include "kernel/complex/cmplx.h"

// (the preprocessor cannot put #include inside a macro: the four instantiations are spelled out)
#define N_RHS 1
namespace real_k1 {
#include "kernel/real/spmv.cl"
#include "kernel/real/vdot.cl"
#include "kernel/real/axpy.cl"
#include "kernel/real/aypx.cl"
#include "kernel/real/sub.cl"
}
namespace cplx_k1 {
#include "kernel/complex/spmv.cl"
#include "kernel/complex/vdot.cl"
#include "kernel/complex/axpy.cl"
#include "kernel/complex/aypx.cl"
#include "kernel/complex/sub.cl"
}
#undef N_RHS
#define N_RHS 2
namespace real_k2 {
#include "kernel/real/spmv.cl"
#include "kernel/real/vdot.cl"
#include "kernel/real/axpy.cl"
#include "kernel/real/aypx.cl"
#include "kernel/real/sub.cl"
}
namespace cplx_k2 {
#include "kernel/complex/spmv.cl"
#include "kernel/complex/vdot.cl"
#include "kernel/complex/axpy.cl"
#include "kernel/complex/aypx.cl"
#include "kernel/complex/sub.cl"
}
#undef N_RHS
#define N_RHS 3
namespace real_k3 {
#include "kernel/real/spmv.cl"
#include "kernel/real/vdot.cl"
#include "kernel/real/axpy.cl"
#include "kernel/real/aypx.cl"
#include "kernel/real/sub.cl"
}
namespace cplx_k3 {
#include "kernel/complex/spmv.cl"
#include "kernel/complex/vdot.cl"
#include "kernel/complex/axpy.cl"
#include "kernel/complex/aypx.cl"
#include "kernel/complex/sub.cl"
}
#undef N_RHS
#define N_RHS 4
namespace real_k4 {
#include "kernel/real/spmv.cl"
#include "kernel/real/vdot.cl"
#include "kernel/real/axpy.cl"
#include "kernel/real/aypx.cl"
#include "kernel/real/sub.cl"
}
namespace cplx_k4 {
#include "kernel/complex/spmv.cl"
#include "kernel/complex/vdot.cl"
#include "kernel/complex/axpy.cl"
#include "kernel/complex/aypx.cl"
#include "kernel/complex/sub.cl"
}
#undef N_RHS

// a kernel set as a type, so that the launch sequence below is written once
#define CLREF_SET(NS, V)                                                                                          \
    struct NS##_set {                                                                                             \
        typedef V val;                                                                                            \
        static void spmv(int n, const V *a, const int *p, const int *c, const V *x, V *y, V *l) { NS::spmv(n, a, p, c, x, y, l); } \
        static void vdot(const V *a, const V *b, V *l, V *g, int n) { NS::vdot(a, b, l, g, n); }                   \
        static void axpy(V *x, V *y, V *a, int sign, int n) { NS::axpy(x, y, a, sign, n); }                        \
        static void aypx(V *x, V *y, V *a, int n) { NS::aypx(x, y, a, n); }                                        \
        static void sub(const V *a, const V *b, V *r, int n) { NS::sub(a, b, r, n); }                              \
    };
CLREF_SET(real_k1, float) CLREF_SET(real_k2, float) CLREF_SET(real_k3, float) CLREF_SET(real_k4, float)
CLREF_SET(cplx_k1, float2) CLREF_SET(cplx_k2, float2) CLREF_SET(cplx_k3, float2) CLREF_SET(cplx_k4, float2)

// host-side scalars: float, or C99 float complex (clcg.c:326, :390 divide `float complex` values; g++'s
// std::complex<float> division is the same libgcc routine, __divsc3)
template <typename V> struct Host;
template <> struct Host<float> {
    typedef float T;
    static T load(const float &v) { return v; }
    static void store(float &d, T v) { d = v; }
};
template <> struct Host<float2> {
    typedef std::complex<float> T;
    static T load(const float2 &v) { return T(v.x.v, v.y.v); }
    static void store(float2 &d, T v) { d = float2(FloatPair{v.real(), v.imag()}); }
};

// The launch sequence of cg(): clcg.c:124-135 (geometry), :253-292 (initialisation), :296-419 (loop).
template <typename K>
static int run_cg(int size, const void *aValues_, const void *b_, const int *aPointers, const int *aCols, void *x_, int nRHS,
                  int nIterations) {
    typedef typename K::val V;
    typedef typename Host<V>::T H;
    if (size < WG_SIZE) return -2;       // "size less than 256 NOT SUPPORTED" (clcg.c:123): one odd-sized work-group there
    const V *aValues = (const V *)aValues_, *b = (const V *)b_;
    V *x = (V *)x_;
    const int workGroups = 1 + (size - 1) / WG_SIZE;                       // clcg.c:124
    const size_t globalSize = (size_t)workGroups * WG_SIZE, localSize = WG_SIZE;
    const size_t spmvGlobal = (size_t)(1 + (size - 1) / (WG_SIZE / WAVE_SIZE)) * WG_SIZE;   // clcg.c:132-134
    const size_t len = (size_t)size * nRHS;
    std::vector<V> r(len), d(len), q(len), dotRes((size_t)workGroups * nRHS), local((size_t)nRHS * WG_SIZE), konst(nRHS);
    std::vector<H> deltaNew(nRHS, H(0)), deltaOld(nRHS, H(0)), dq(nRHS), alpha(nRHS), beta(nRHS);
    // one more row offset than the matrix has: spmv.cl:18-19 reads aPointers[waveId + 1] before it tests waveId
    std::vector<int> ptr(aPointers, aPointers + size + 1);
    ptr.resize(spmvGlobal / WAVE_SIZE + 2, aPointers[size]);

    auto spmv = [&](const V *in, V *out) {
        wi::ndrange(spmvGlobal, WG_SIZE, [&] { K::spmv(size, aValues, ptr.data(), aCols, in, out, local.data()); });
    };
    auto dot = [&](const V *u, const V *v, std::vector<H> &acc) {       // kernel + the sequential host sum
        wi::ndrange(globalSize, localSize, [&] { K::vdot(u, v, local.data(), dotRes.data(), size); });
        for (int c = 0; c < nRHS; c++)
            for (int w = 0; w < workGroups; w++) acc[c] += Host<V>::load(dotRes[(size_t)c * workGroups + w]);
    };

    spmv(x, q.data());                                                                              // :255
    wi::ndrange(globalSize, localSize, [&] { K::sub(b, q.data(), r.data(), size); });                // :260
    d = r;                                                                                          // :264
    dot(r.data(), r.data(), deltaNew);                                                              // :268-279
    deltaOld = deltaNew;                                                                            // :282, :289
    for (int it = 0; it < nIterations; it++) {
        spmv(d.data(), q.data());                                                                   // :299-305
        for (int c = 0; c < nRHS; c++) dq[c] = H(0);                                                // :318-319
        dot(d.data(), q.data(), dq);                                                                // :309-324
        for (int c = 0; c < nRHS; c++) {
            alpha[c] = deltaNew[c] / dq[c];                                                         // :326-327
            Host<V>::store(konst[c], alpha[c]);                                                     // :334
        }
        wi::ndrange(globalSize, localSize, [&] { K::axpy(d.data(), x, konst.data(), 1, size); });        // :338-342
        wi::ndrange(globalSize, localSize, [&] { K::axpy(q.data(), r.data(), konst.data(), 0, size); }); // :345-349
        for (int c = 0; c < nRHS; c++) {                                                            // :350-356
            deltaOld[c] = deltaNew[c];
            deltaNew[c] = H(0);
        }
        dot(r.data(), r.data(), deltaNew);                                                          // :369-387
        for (int c = 0; c < nRHS; c++) {
            beta[c] = deltaNew[c] / deltaOld[c];                                                    // :389-391
            Host<V>::store(konst[c], beta[c]);                                                      // :411
        }
        wi::ndrange(globalSize, localSize, [&] { K::aypx(r.data(), d.data(), konst.data(), size); });   // :415
    }
    return 0;
}

extern "C" {
// size must be >= 256 (clcg.c:123); when it is not a multiple of 8 the out-of-bounds row-offset read of
// spmv.cl:18-19 lands in the padding added above (empty rows, never stored).  nRHS 1..4.  Same argument meaning as cg() (clcg.h:3-5).  Returns 0, -1 (nRHS not instantiated), -2 (size).
int clref_cg(int size, int nonZeros, const float *aValues, const float *b, const int *aPointers, const int *aCols,
             float *x, int nRHS, int nIterations, int isComplex) {
    (void)nonZeros;
    switch (nRHS * 2 + (isComplex ? 1 : 0)) {
    case 2: return run_cg<real_k1_set>(size, aValues, b, aPointers, aCols, x, nRHS, nIterations);
    case 3: return run_cg<cplx_k1_set>(size, aValues, b, aPointers, aCols, x, nRHS, nIterations);
    case 4: return run_cg<real_k2_set>(size, aValues, b, aPointers, aCols, x, nRHS, nIterations);
    case 5: return run_cg<cplx_k2_set>(size, aValues, b, aPointers, aCols, x, nRHS, nIterations);
    case 6: return run_cg<real_k3_set>(size, aValues, b, aPointers, aCols, x, nRHS, nIterations);
    case 7: return run_cg<cplx_k3_set>(size, aValues, b, aPointers, aCols, x, nRHS, nIterations);
    case 8: return run_cg<real_k4_set>(size, aValues, b, aPointers, aCols, x, nRHS, nIterations);
    case 9: return run_cg<cplx_k4_set>(size, aValues, b, aPointers, aCols, x, nRHS, nIterations);
    }
    return -1;
}
}
