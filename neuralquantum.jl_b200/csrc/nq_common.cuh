// nq_common.cuh -- shared internals of libnqcuda (context, staging, complex math, activations).
// Hand-written CUDA for sm_100a.  Nothing here is part of the public ABI (see include/nqcuda.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>
#include <cmath>
#include "../../include/nqcuda.h"

// --------------------------------------------------------------------------------------
// context
// --------------------------------------------------------------------------------------
enum { NQ_NSLOTS = 24 };

struct nq_ctx_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    size_t smem_optin = 0;
    std::string err;
    int64_t info = 0;
    uint64_t launches = 0;
    struct Slot { void* p = nullptr; size_t cap = 0; } slots[NQ_NSLOTS];
    bool host_out_pending = false;
    // NCCL (nq_comm.cu)
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    int64_t ns_total = 0;          // global sample count override (nq_comm_set_global_samples), 0 = Ns * nranks
    // one-shot structural hint for the next nq_sr_setup (nq_sr_hint_row_planes): per 128-row tile, which component planes
    // of O may be non-zero (bit 0 = re, bit 1 = im); replaces the scan of O by tile_activity_kernel
    std::vector<unsigned> hint_tile_flags;
    int64_t hint_P = 0;
    // row maxima of the matrix the last centring pass wrote (one-shot, consumed by the next nq_sr_setup on the same matrix):
    // the Ozaki S assembly scales every row by its exponent and would otherwise read O once more to find it
    unsigned long long* rowmax = nullptr;      // device [rowmax_cap] bit patterns of max |x| per parameter row
    int64_t rowmax_cap = 0;
    const void* rowmax_ptr = nullptr;
    int64_t rowmax_P = 0, rowmax_Ns = 0, rowmax_ld = 0;
    // deferred centring (nq_center_lazy): the means were computed but NOT subtracted; rowmax then holds the maxima of the
    // UNcentred rows and `shift` the means (device, complex128 [P]); the next nq_sr_setup on the same matrix either lets the
    // Ozaki pre-pass subtract on the fly or subtracts in place first
    void* shift = nullptr;
    int64_t shift_cap = 0;
    bool shift_pending = false;            // stays set until the subtraction happened (nq_center_finish, an S path that centres
    const void* shift_ptr = nullptr;       // in place) or the matrix is re-centred / rewritten by a machine kernel
    int64_t shift_P = 0, shift_Ns = 0, shift_ld = 0;
    // side stream + events of the software-pipelined host entry (nq_logpsi_grad_local_host), created on first use
    cudaStream_t side_stream = nullptr;
    cudaEvent_t side_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool shift_matches(const void* X, int64_t ld, int64_t P, int64_t Ns) const {
        return shift_pending && shift_ptr == X && shift_P == P && shift_Ns == Ns && shift_ld == ld;
    }
};

int nq_fail(nq_ctx_t ctx, int code, const char* fmt, ...);
// grow-only device scratch, one buffer per slot id
void* nq_scratch(nq_ctx_t ctx, int slot, size_t bytes);
bool nq_is_device_ptr(const void* p);

#define NQ_CUDA(ctx, call)                                                                    \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return nq_fail((ctx), NQ_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,     \
                           cudaGetErrorString(e__));                                          \
    } while (0)

#define NQ_CHECK(expr)                \
    do {                              \
        int s__ = (expr);             \
        if (s__ != NQ_OK) return s__; \
    } while (0)

// every kernel launch goes through this: counts launches, checks the launch error
#define NQ_LAUNCH(ctx, kern, grid, block, smem, ...)                                          \
    do {                                                                                      \
        kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                        \
        (ctx)->launches++;                                                                    \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess)                                                               \
            return nq_fail((ctx), NQ_ERR_CUDA, "%s:%d launch %s: %s", __FILE__, __LINE__,     \
                           #kern, cudaGetErrorString(e__));                                   \
    } while (0)

// Staging of caller buffers that may live on the host.
struct NqStage {
    nq_ctx_t ctx;
    struct Out { void* user; void* dev; size_t bytes; size_t width, pitch, rows; };
    std::vector<Out> outs;
    int status = NQ_OK;
    explicit NqStage(nq_ctx_t c) : ctx(c) {}
    // returns a device pointer holding the input
    const void* in(int slot, const void* p, size_t bytes) {
        if (!p || bytes == 0) return p;
        if (nq_is_device_ptr(p)) return p;
        void* d = nq_scratch(ctx, slot, bytes);
        if (!d) { status = NQ_ERR_ALLOC; return nullptr; }
        cudaError_t e = cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) status = nq_fail(ctx, NQ_ERR_CUDA, "H2D copy: %s", cudaGetErrorString(e));
        return d;
    }
    // returns a device pointer to write the output to
    void* out(int slot, void* p, size_t bytes) {
        if (!p || bytes == 0) return p;
        if (nq_is_device_ptr(p)) return p;
        void* d = nq_scratch(ctx, slot, bytes);
        if (!d) { status = NQ_ERR_ALLOC; return nullptr; }
        outs.push_back({p, d, bytes, 0, 0, 0});
        return d;
    }
    // pitched output (leading dimension > row width): only the [width x rows] payload is written back
    void* out2d(int slot, void* p, size_t width, size_t pitch, size_t rows) {
        if (!p || rows == 0) return p;
        if (nq_is_device_ptr(p)) return p;
        if (width == pitch) return out(slot, p, pitch * rows);
        void* d = nq_scratch(ctx, slot, pitch * rows);
        if (!d) { status = NQ_ERR_ALLOC; return nullptr; }
        outs.push_back({p, d, pitch * rows, width, pitch, rows});
        return d;
    }
    // copy staged outputs back; synchronises only when something went to the host
    int finish() {
        if (status != NQ_OK) return status;
        if (outs.empty()) return NQ_OK;
        for (auto& o : outs) {
            cudaError_t e = o.rows ? cudaMemcpy2DAsync(o.user, o.pitch, o.dev, o.pitch, o.width, o.rows, cudaMemcpyDeviceToHost, ctx->stream)
                                   : cudaMemcpyAsync(o.user, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx->stream);
            if (e != cudaSuccess) return nq_fail(ctx, NQ_ERR_CUDA, "D2H copy: %s", cudaGetErrorString(e));
        }
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return nq_fail(ctx, NQ_ERR_CUDA, "sync: %s", cudaGetErrorString(e));
        return NQ_OK;
    }
};

static inline size_t nq_dtype_size(nq_dtype d) {
    switch (d) { case NQ_F32: return 4; case NQ_F64: return 8; case NQ_C64: return 8; default: return 16; }
}
static inline bool nq_dtype_is_complex(nq_dtype d) { return d == NQ_C64 || d == NQ_C128; }
static inline bool nq_dtype_is_double(nq_dtype d) { return d == NQ_F64 || d == NQ_C128; }
static inline nq_dtype nq_complex_of(nq_dtype d) { return nq_dtype_is_double(d) ? NQ_C128 : NQ_C64; }
static inline nq_dtype nq_real_of(nq_dtype d) { return nq_dtype_is_double(d) ? NQ_F64 : NQ_F32; }
static inline int nq_words(int N) { return (N + 63) / 64; }

// --------------------------------------------------------------------------------------
// complex numbers
// --------------------------------------------------------------------------------------
template <typename T> struct alignas(2 * sizeof(T)) cx {
    T re, im;
    __host__ __device__ cx() {}
    __host__ __device__ cx(T r, T i) : re(r), im(i) {}
    __host__ __device__ explicit cx(T r) : re(r), im(T(0)) {}
};
typedef cx<float> cxf;
typedef cx<double> cxd;

template <typename E> struct elem_traits;
template <> struct elem_traits<float> { typedef float real; static const bool is_complex = false; };
template <> struct elem_traits<double> { typedef double real; static const bool is_complex = false; };
template <> struct elem_traits<cxf> { typedef float real; static const bool is_complex = true; };
template <> struct elem_traits<cxd> { typedef double real; static const bool is_complex = true; };

#define NQ_HD __host__ __device__ __forceinline__
template <typename T> NQ_HD cx<T> operator+(cx<T> a, cx<T> b) { return cx<T>(a.re + b.re, a.im + b.im); }
template <typename T> NQ_HD cx<T> operator-(cx<T> a, cx<T> b) { return cx<T>(a.re - b.re, a.im - b.im); }
template <typename T> NQ_HD cx<T> operator-(cx<T> a) { return cx<T>(-a.re, -a.im); }
template <typename T> NQ_HD cx<T> operator*(cx<T> a, cx<T> b) {
    return cx<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename T> NQ_HD cx<T> operator*(T s, cx<T> a) { return cx<T>(s * a.re, s * a.im); }
template <typename T> NQ_HD cx<T> operator*(cx<T> a, T s) { return cx<T>(s * a.re, s * a.im); }
template <typename T> NQ_HD cx<T>& operator+=(cx<T>& a, cx<T> b) { a.re += b.re; a.im += b.im; return a; }
template <typename T> NQ_HD cx<T>& operator-=(cx<T>& a, cx<T> b) { a.re -= b.re; a.im -= b.im; return a; }
template <typename T> NQ_HD cx<T> conj(cx<T> a) { return cx<T>(a.re, -a.im); }
NQ_HD float conj(float a) { return a; }
NQ_HD double conj(double a) { return a; }
template <typename T> NQ_HD T real_part(cx<T> a) { return a.re; }
NQ_HD float real_part(float a) { return a; }
NQ_HD double real_part(double a) { return a; }
template <typename T> NQ_HD T imag_part(cx<T> a) { return a.im; }
NQ_HD float imag_part(float) { return 0.f; }
NQ_HD double imag_part(double) { return 0.0; }
template <typename E> NQ_HD E make_zero();
template <> NQ_HD float make_zero<float>() { return 0.f; }
template <> NQ_HD double make_zero<double>() { return 0.0; }
template <> NQ_HD cxf make_zero<cxf>() { return cxf(0.f, 0.f); }
template <> NQ_HD cxd make_zero<cxd>() { return cxd(0.0, 0.0); }
// E from a real scalar
template <typename E, typename T> NQ_HD E from_real(T r);
template <> NQ_HD float from_real<float, float>(float r) { return r; }
template <> NQ_HD double from_real<double, double>(double r) { return r; }
template <> NQ_HD cxf from_real<cxf, float>(float r) { return cxf(r, 0.f); }
template <> NQ_HD cxd from_real<cxd, double>(double r) { return cxd(r, 0.0); }
// to complex of the same precision
NQ_HD cxf to_cx(float a) { return cxf(a, 0.f); }
NQ_HD cxd to_cx(double a) { return cxd(a, 0.0); }
NQ_HD cxf to_cx(cxf a) { return a; }
NQ_HD cxd to_cx(cxd a) { return a; }
// scale by a real
NQ_HD float rscale(float s, float a) { return s * a; }
NQ_HD double rscale(double s, double a) { return s * a; }
template <typename T> NQ_HD cx<T> rscale(T s, cx<T> a) { return cx<T>(s * a.re, s * a.im); }

// --------------------------------------------------------------------------------------
// scalar math overloads (precise versions; no fast-math intrinsics: FP32 mode must hold 1e-5)
// --------------------------------------------------------------------------------------
__device__ __forceinline__ float m_exp(float x) { return expf(x); }
__device__ __forceinline__ double m_exp(double x) { return exp(x); }
__device__ __forceinline__ float m_log(float x) { return logf(x); }
__device__ __forceinline__ double m_log(double x) { return log(x); }
__device__ __forceinline__ float m_log1p(float x) { return log1pf(x); }
__device__ __forceinline__ double m_log1p(double x) { return log1p(x); }
__device__ __forceinline__ float m_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double m_tanh(double x) { return tanh(x); }
__device__ __forceinline__ float m_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double m_abs(double x) { return fabs(x); }
__device__ __forceinline__ float m_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double m_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ void m_sincos(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ void m_sincos(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ float m_rem2pi(float y) { return remainderf(y, 6.283185307179586f); }
__device__ __forceinline__ double m_rem2pi(double y) { return remainder(y, 6.283185307179586476925); }
template <typename T> struct big_x;   // above this, exp(-x) is below one ulp of 1
template <> struct big_x<float> { static constexpr float v = 18.f; };
template <> struct big_x<double> { static constexpr double v = 38.0; };

template <typename T> __device__ __forceinline__ cx<T> cx_exp(cx<T> z) {
    T e = m_exp(z.re), s, c;
    m_sincos(z.im, &s, &c);
    return cx<T>(e * c, e * s);
}

// --------------------------------------------------------------------------------------
// activations: value f and derivative f' in one call.  ref: Networks/activation.jl:5-29
//   softplus: f = log1p(exp(x)),  f' = 1/(1+exp(-x))          (real and complex x)
//   logcosh : f = log cosh x (|x|<=12, else |x|-log 2),  f' = tanh x;
//             complex: f = logcosh(Re x) + log(cos(Im x) + i tanh(Re x) sin(Im x))
// The device formulas are algebraically identical rearrangements chosen to avoid overflow and
// cancellation; they agree with the reference formulas to a few ulp (absolute).
// --------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void act_softplus(T x, T& f, T& d) {
    T e = m_exp(-m_abs(x));                       // in (0,1]
    T l = m_log1p(e);
    f = (x > T(0) ? x : T(0)) + l;
    d = x >= T(0) ? T(1) / (T(1) + e) : e / (T(1) + e);
}
template <typename T> __device__ __forceinline__ void act_softplus(cx<T> z, cx<T>& f, cx<T>& d) {
    T x = z.re, s, c;
    m_sincos(z.im, &s, &c);
    if (x > big_x<T>::v) {                        // log1p(e^z) = z + e^{-z} (principal branch), sigma = 1 - e^{-z}
        T e = m_exp(-x);
        f = cx<T>(x + e * c, m_rem2pi(z.im) - e * s);
        d = cx<T>(T(1) - e * c, e * s);
        return;
    }
    T ex = m_exp(x);
    T a = ex * c, b = ex * s;                     // w = e^z
    T den_re = T(1) + a;
    T mod2 = den_re * den_re + b * b;             // |1+w|^2
    T re = (a > T(-0.5)) ? T(0.5) * m_log1p(a * (T(2) + a) + b * b) : T(0.5) * m_log(mod2);
    f = cx<T>(re, m_atan2(b, den_re));
    // sigma(z) = w/(1+w) = w conj(1+w)/|1+w|^2 = (a(1+a)+b^2 + i b)/|1+w|^2
    T inv = T(1) / mod2;
    d = cx<T>((a * den_re + b * b) * inv, b * inv);
}
template <typename T> __device__ __forceinline__ T logcosh_real(T x) {
    T ax = m_abs(x);
    return ax + m_log1p(m_exp(T(-2) * ax)) - T(0.693147180559945309417232121458);
}
template <typename T> __device__ __forceinline__ void act_logcosh(T x, T& f, T& d) {
    f = logcosh_real(x);
    d = m_tanh(x);
}
template <typename T> __device__ __forceinline__ void act_logcosh(cx<T> z, cx<T>& f, cx<T>& d) {
    T t = m_tanh(z.re), s, c;
    m_sincos(z.im, &s, &c);
    T sech2 = (T(1) - t) * (T(1) + t);
    T m2 = c * c + t * t * s * s;                 // |cos y + i t sin y|^2 = 1 - s^2 sech^2
    f = cx<T>(logcosh_real(z.re) + T(0.5) * m_log1p(-s * s * sech2), m_atan2(t * s, c));
    T inv = T(1) / m2;                            // tanh z = (t + i s c sech^2)/m2
    d = cx<T>(t * inv, s * c * sech2 * inv);
}
template <int ACT, typename E> __device__ __forceinline__ void act_eval(E x, E& f, E& d) {
    if (ACT == NQ_SOFTPLUS) act_softplus(x, f, d); else act_logcosh(x, f, d);
}

// --------------------------------------------------------------------------------------
// flip-ratio algebra.  With s = f'(theta) cached for the sampled configuration and Ep = exp(+Delta),
// Em = exp(-Delta) for the change Delta of the pre-activation under a flip,
//   softplus: (1+e^{theta+Delta})/(1+e^theta) = 1 + s (Ep - 1),         s' = s Ep / factor
//   logcosh : cosh(theta+Delta)/cosh(theta)   = (Ep (1+t) + Em (1-t))/2, t' = (Ep(1+t) - Em(1-t)) / (2 factor)
// so a connected configuration costs a handful of multiply-adds and one division per hidden unit and
// NO transcendental (Ep, Em come from the per-parameter tables built by nq_machine_ensure_tables).
// --------------------------------------------------------------------------------------
__device__ __forceinline__ float e_exp(float x) { return expf(x); }
__device__ __forceinline__ double e_exp(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ cx<T> e_exp(cx<T> x) { return cx_exp(x); }
template <typename T> NQ_HD cx<T> cx_div(cx<T> a, cx<T> b) {
    T inv = T(1) / (b.re * b.re + b.im * b.im);
    return cx<T>((a.re * b.re + a.im * b.im) * inv, (a.im * b.re - a.re * b.im) * inv);
}
NQ_HD float e_div(float a, float b) { return a / b; }
NQ_HD double e_div(double a, double b) { return a / b; }
template <typename T> NQ_HD cx<T> e_div(cx<T> a, cx<T> b) { return cx_div(a, b); }
NQ_HD float e_mul(float a, float b) { return a * b; }
NQ_HD double e_mul(double a, double b) { return a * b; }
template <typename T> NQ_HD cx<T> e_mul(cx<T> a, cx<T> b) { return a * b; }
template <typename E> NQ_HD E e_one() { return from_real<E, typename elem_traits<E>::real>(typename elem_traits<E>::real(1)); }
template <int ACT, typename E>
__device__ __forceinline__ void ratio_step(E s, E Ep, E Em, E& fac, E& snew) {
    typedef typename elem_traits<E>::real T;
    if (ACT == NQ_SOFTPLUS) {
        fac = e_one<E>() + e_mul(s, Ep - e_one<E>());
        snew = e_div(e_mul(s, Ep), fac);
    } else {
        E a = e_mul(Ep, e_one<E>() + s), b = e_mul(Em, e_one<E>() - s);
        E sum = a + b;
        fac = rscale(T(0.5), sum);
        snew = e_div(a - b, sum);
    }
}
NQ_HD double abs2_d(float a) { return (double)a * (double)a; }
NQ_HD double abs2_d(double a) { return a * a; }
template <typename T> NQ_HD double abs2_d(cx<T> a) { return (double)a.re * (double)a.re + (double)a.im * (double)a.im; }
// same step with the division deferred: snew = e_div(num, den) (bit-identical to ratio_step); the sampler only
// divides for accepted proposals.  den follows from fac (softplus: fac; logcosh: 2 fac, exact).
template <int ACT, typename E>
__device__ __forceinline__ void ratio_parts(E s, E Ep, E Em, E& fac, E& num) {
    typedef typename elem_traits<E>::real T;
    if (ACT == NQ_SOFTPLUS) {
        fac = e_one<E>() + e_mul(s, Ep - e_one<E>());
        num = e_mul(s, Ep);
    } else {
        E a = e_mul(Ep, e_one<E>() + s), b = e_mul(Em, e_one<E>() - s);
        fac = rscale(T(0.5), a + b);
        num = a - b;
    }
}
template <int ACT, typename E>
__device__ __forceinline__ E ratio_finish(E fac, E num) {
    typedef typename elem_traits<E>::real T;
    return e_div(num, ACT == NQ_SOFTPLUS ? fac : rscale(T(2), fac));
}
// products are accumulated in double precision whatever the mode
NQ_HD double to_d(float a) { return (double)a; }
NQ_HD double to_d(double a) { return a; }
template <typename T> NQ_HD cxd to_d(cx<T> a) { return cxd((double)a.re, (double)a.im); }
__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ cxd warp_prod(cxd v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        cxd o(__shfl_xor_sync(0xffffffffu, v.re, m), __shfl_xor_sync(0xffffffffu, v.im, m));
        v = v * o;
    }
    return v;
}

// --------------------------------------------------------------------------------------
// warp / block reductions
// --------------------------------------------------------------------------------------
__device__ __forceinline__ float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <typename T> __device__ __forceinline__ cx<T> shfl_xor(cx<T> v, int m) {
    return cx<T>(shfl_xor(v.re, m), shfl_xor(v.im, m));
}
template <typename E> __device__ __forceinline__ E warp_sum(E v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = v + shfl_xor(v, m);
    return v;
}

// bit helpers for packed configurations
__device__ __forceinline__ int get_bit(const uint64_t* w, int j) { return (int)((w[j >> 6] >> (j & 63)) & 1ull); }
// value of a digit: spin -> 2d-1, fock -> d
template <typename T> __device__ __forceinline__ T digit_value(int hilb, int d) {
    return hilb == NQ_SPIN ? T(2 * d - 1) : T(d);
}
// change of the value when a digit flips from d to 1-d: spin -> -2v = 2-4d, fock -> 1-2d
template <typename T> __device__ __forceinline__ T flip_delta(int hilb, int d) {
    return hilb == NQ_SPIN ? T(2 - 4 * d) : T(1 - 2 * d);
}
