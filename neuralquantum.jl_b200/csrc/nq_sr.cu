// nq_sr.cu -- K6-K9: gradient centring, force vector, S-matrix assembly, solve, parameter update.
//
//   K6  nq_center / nq_force_*      column means and weighted column sums over the [P, Ns] matrices
//                                    (HBM-bound streaming passes, FP64 accumulation, fixed order)
//   K7  nq_sr_setup                  S = Oc Oc^H / Ns as a split-K SYRK/HERK on the FP64 tensor path
//                                    (mma.sync m8n8k4 f64 = DMMA; tcgen05 has no FP64 kind), lower
//                                    tile triangle only, deterministic two-stage reduction; FP32-mode
//                                    operands go to the tcgen05 3xTF32 kernel in nq_syrk_tf32.cu
//   K8  nq_sr_solve(_matfree)        blocked Cholesky (cuSOLVER-free) or CG (IterativeSolvers 0.8.1 rule)
//   K9  nq_update                    w <- w - eta dw
//
// ref: IterativeInterface/Samplers/BaseIterativeSampler.jl:19-26, CostFun/BatchedValSampler.jl:97-115,
//      CostFun/BatchedGradSampler.jl:99-118, Algorithms/SR/SRDirect.jl:26-90, SRIterative.jl:71-153,
//      SR_notfull.jl:47-148, Optimisers/rules.jl:11-17, utils/stats.jl:26-50.
#include "nq_internal.cuh"
#include <algorithm>
#include <cstring>
#include <cstdlib>

int nq_allreduce_device(nq_ctx_t ctx, void* buf, int64_t n, nq_dtype dtype, bool mean);

namespace {

// ======================================================================================
// K6: column sums over samples.  X is [P, Ns] (leading dimension ld) of E; thread = parameter
// (coalesced), gridDim.y slices the samples; partials are summed in a fixed order.
//   out[k] = scale * sum_s w[s] * (CONJ ? conj(X[k,s]) : X[k,s])        (w == nullptr: w = 1)
// ======================================================================================
// Compensated (double-double) accumulation: the force vector is a difference of O(1) averages that cancels to a few
// per cent of their size, and the parity bound is element-wise 1e-11 (FP64 mode) -- plain sequential sums over 10^4-10^5
// samples sit right at that bound.  TwoSum / TwoProduct (Knuth, Dekker with FMA) make the sums exact to ~1e-30 relative
// to sum |terms| at no cost in time (the pass is bound by reading X once).
struct dd { double hi, lo; };
__device__ __forceinline__ void dd_add(dd& a, double p) {          // a += p, error-free
    const double t = a.hi + p, bp = t - a.hi;
    a.lo += (a.hi - (t - bp)) + (p - bp);
    a.hi = t;
}
__device__ __forceinline__ void dd_fma(dd& a, double x, double y) { // a += x * y, error-free
    const double p = x * y;
    a.lo += fma(x, y, -p);
    dd_add(a, p);
}

template <typename E, bool CONJ>
__global__ void colsum_partial_kernel(const E* __restrict__ X, int64_t ld, int64_t P, int64_t Ns,
                                      const cx<typename elem_traits<E>::real>* __restrict__ w,
                                      cxd* __restrict__ partial, cxd* __restrict__ partial_lo, unsigned long long* __restrict__ mx) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= P) return;
    int64_t per = (Ns + gridDim.y - 1) / gridDim.y;
    int64_t s0 = blockIdx.y * per, s1 = s0 + per < Ns ? s0 + per : Ns;
    dd ar = {0.0, 0.0}, ai = {0.0, 0.0};
    double m = 0.0;
    for (int64_t s = s0; s < s1; s++) {
        E x = X[k + ld * s];
        double xr = (double)real_part(x), xi = (double)imag_part(x);
        m = fmax(m, fmax(fabs(xr), fabs(xi)));
        if (CONJ) xi = -xi;
        if (w) {
            double wr = (double)w[s].re, wi = (double)w[s].im;
            dd_fma(ar, wr, xr); dd_fma(ar, -wi, xi);
            dd_fma(ai, wr, xi); dd_fma(ai, wi, xr);
        } else { dd_add(ar, xr); dd_add(ai, xi); }
    }
    partial[blockIdx.y * P + k] = cxd(ar.hi, ai.hi);
    partial_lo[blockIdx.y * P + k] = cxd(ar.lo, ai.lo);
    if (mx) atomicMax(&mx[k], (unsigned long long)__double_as_longlong(m));
}

// fixed-order compensated reduction of the slices: 8 threads per row take the slices g, g + 8, ... and their partial
// double-doubles are combined in lane order (the first version walked all slices in one thread: 9 CTAs, 0.14 ms on cfg4)
template <typename E>
__global__ void colsum_final_kernel(const cxd* __restrict__ partial, const cxd* __restrict__ partial_lo, int nslice, int64_t P,
                                    double scale, cxd* __restrict__ out) {
    const int g = threadIdx.x & 7;
    const int64_t k = blockIdx.x * (int64_t)(blockDim.x >> 3) + (threadIdx.x >> 3);
    dd ar = {0.0, 0.0}, ai = {0.0, 0.0};
    if (k < P)
        for (int i = g; i < nslice; i += 8) {
            dd_add(ar, partial[i * P + k].re); dd_add(ai, partial[i * P + k].im);
            if (partial_lo) { ar.lo += partial_lo[i * P + k].re; ai.lo += partial_lo[i * P + k].im; }
        }
    // combine the 8 partial sums of the row in a fixed order (lanes g = 0..7 of the same 8-lane group)
    dd tr = {0.0, 0.0}, ti = {0.0, 0.0};
    for (int j = 0; j < 8; j++) {
        const int src = (threadIdx.x & 31 & ~7) | j;
        const double rh = __shfl_sync(0xffffffffu, ar.hi, src), rl = __shfl_sync(0xffffffffu, ar.lo, src);
        const double ih = __shfl_sync(0xffffffffu, ai.hi, src), il = __shfl_sync(0xffffffffu, ai.lo, src);
        dd_add(tr, rh); tr.lo += rl;
        dd_add(ti, ih); ti.lo += il;
    }
    if (k < P && g == 0) out[k] = cxd((tr.hi + tr.lo) * scale, (ti.hi + ti.lo) * scale);
}

template <typename E, bool CONJ>
int colsum(nq_ctx_t ctx, const void* X, int64_t ld, int64_t P, int64_t Ns, const void* w, double scale, cxd* out,
           unsigned long long* mx = nullptr) {
    int nslice = (int)std::min<int64_t>(std::max<int64_t>(1, Ns / 256), 4 * (int64_t)ctx->num_sms * 8 / std::max<int64_t>(1, (P + 127) / 128));
    nslice = std::max(1, std::min(nslice, 1024));
    cxd* partial = (cxd*)nq_scratch(ctx, SL_W5, (size_t)2 * nslice * P * sizeof(cxd));
    if (!partial) return NQ_ERR_ALLOC;
    cxd* partial_lo = partial + (size_t)nslice * P;
    dim3 grid((unsigned)((P + 127) / 128), (unsigned)nslice);
    NQ_LAUNCH(ctx, (colsum_partial_kernel<E, CONJ>), grid, 128, 0, (const E*)X, ld, P, Ns,
              (const cx<typename elem_traits<E>::real>*)w, partial, partial_lo, mx);
    NQ_LAUNCH(ctx, colsum_final_kernel<E>, (unsigned)((P + 31) / 32), 256, 0, (const cxd*)partial, (const cxd*)partial_lo, nslice, P, scale, out);
    return NQ_OK;
}

template <bool CONJ>
int colsum_dispatch(nq_ctx_t ctx, nq_dtype dtype, const void* X, int64_t ld, int64_t P, int64_t Ns, const void* w,
                    double scale, cxd* out, unsigned long long* mx = nullptr) {
    switch (dtype) {
        case NQ_F32: return colsum<float, CONJ>(ctx, X, ld, P, Ns, w, scale, out, mx);
        case NQ_F64: return colsum<double, CONJ>(ctx, X, ld, P, Ns, w, scale, out, mx);
        case NQ_C64: return colsum<cxf, CONJ>(ctx, X, ld, P, Ns, w, scale, out, mx);
        default: return colsum<cxd, CONJ>(ctx, X, ld, P, Ns, w, scale, out, mx);
    }
}

// X <- X - avg; mx (optional): running maximum of |centred value| per parameter row (both components), see nq_ctx_s::rowmax
template <typename E>
__global__ void subtract_avg_kernel(E* __restrict__ X, int64_t ld, int64_t P, int64_t Ns, const cxd* __restrict__ avg,
                                    unsigned long long* __restrict__ mx) {
    typedef typename elem_traits<E>::real T;
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= P) return;
    cxd a = avg[k];
    double m = 0.0;
    for (int64_t s = blockIdx.y; s < Ns; s += gridDim.y) {
        E x = X[k + ld * s];
        if (elem_traits<E>::is_complex) {
            cx<T>* px = (cx<T>*)&X[k + ld * s];
            const cx<T> y((T)((double)real_part(x) - a.re), (T)((double)imag_part(x) - a.im));
            *px = y;
            m = fmax(m, fmax(fabs((double)y.re), fabs((double)y.im)));
        } else {
            T* px = (T*)&X[k + ld * s];
            const T y = (T)((double)real_part(x) - a.re);
            *px = y;
            m = fmax(m, fabs((double)y));
        }
    }
    if (mx) atomicMax(&mx[k], (unsigned long long)__double_as_longlong(m));
}

// (re)arm the one-shot row-maximum record of the context for the matrix X about to be centred; returns the device buffer
static unsigned long long* rowmax_arm(nq_ctx_t ctx, const void* X, int64_t ld, int64_t P, int64_t Ns) {
    ctx->rowmax_ptr = nullptr;
    if (ctx->rowmax_cap < P) {
        if (ctx->rowmax) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->rowmax); ctx->rowmax = nullptr; ctx->rowmax_cap = 0; }
        if (cudaMalloc((void**)&ctx->rowmax, (size_t)(P + P / 4 + 64) * 8) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        ctx->rowmax_cap = P + P / 4 + 64;
    }
    if (cudaMemsetAsync(ctx->rowmax, 0, (size_t)P * 8, ctx->stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->rowmax_ptr = X; ctx->rowmax_P = P; ctx->rowmax_Ns = Ns; ctx->rowmax_ld = ld;
    return ctx->rowmax;
}

// convert a cxd vector to dtype (real dtypes take the real part)
__global__ void convert_from_cxd_kernel(const cxd* __restrict__ in, void* __restrict__ out, int64_t n, int dtype, int conj_in) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    cxd v = in[i];
    if (conj_in) v.im = -v.im;
    switch (dtype) {
        case NQ_F32: ((float*)out)[i] = (float)v.re; break;
        case NQ_F64: ((double*)out)[i] = v.re; break;
        case NQ_C64: ((cxf*)out)[i] = cxf((float)v.re, (float)v.im); break;
        default: ((cxd*)out)[i] = v; break;
    }
}
__global__ void convert_to_cxd_kernel(const void* __restrict__ in, cxd* __restrict__ out, int64_t n, int dtype) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (dtype) {
        case NQ_F32: out[i] = cxd((double)((const float*)in)[i], 0.0); break;
        case NQ_F64: out[i] = cxd(((const double*)in)[i], 0.0); break;
        case NQ_C64: { cxf v = ((const cxf*)in)[i]; out[i] = cxd((double)v.re, (double)v.im); } break;
        default: out[i] = ((const cxd*)in)[i]; break;
    }
}

// sum_s |L_s|^2 (single block, fixed order)
template <typename T>
__global__ void abs2_sum_kernel(const cx<T>* __restrict__ L, int64_t n, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { double r = (double)L[i].re, m = (double)L[i].im; acc += r * r + m * m; }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w]; *out = s; }
}

// gradC_k = conj(x_k) - cost * avg_k' ...: out = a - c*b (complex vectors, c real on device)
__global__ void force_liouv_finish_kernel(const cxd* __restrict__ lg, const cxd* __restrict__ avg, const double* __restrict__ sumabs2,
                                          double inv_ns, int64_t P, cxd* __restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= P) return;
    double C = *sumabs2 * inv_ns;
    // lg_k = (1/Ns) sum_s conj(L_s) gL_ks ; F_k = lg_k - C avg_k
    out[k] = cxd(lg[k].re - C * avg[k].re, lg[k].im - C * avg[k].im);
}

// ======================================================================================
// K7: SYRK / HERK on DMMA.  X viewed as reals: X[k, s, c] at Xr[(k*NC + c) + ldr*s].
//   mode 0: C[k,l] = sum_{s,c} X[k,s,c] X[l,s,c]                       (real part of O O^H)
//   mode 1: C[k,l] = sum_s X[k,s,0] X[l,s,1] - X[k,s,1] X[l,s,0]       (imag part of conj(O O^H))
// CTA tile 128x128, 8 warps as 2(m) x 4(n), warp tile 64x32 = 8x4 mma tiles, K chunk = KS samples.
// ======================================================================================
constexpr int TS = 128;       // tile size
constexpr int KS = 8;         // samples per chunk

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Which components (bit 0 = re, bit 1 = im) of each 128-row tile are not identically zero.  For the
// real-parameter NDM the rows of the lambda-type parameters are purely real and those of the mu-type purely
// imaginary (NDMBatched.jl:262-277), so most tile pairs need half of K and lambda x mu tiles vanish.
template <typename T, int NC>
__global__ void tile_activity_kernel(const T* __restrict__ Xr, int64_t ldr, int64_t P, int64_t Ns, unsigned* __restrict__ flags) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;      // real row index (k*NC + c)
    if (r >= P * NC) return;
    int64_t per = (Ns + gridDim.y - 1) / gridDim.y;
    int64_t s0 = blockIdx.y * per, s1 = s0 + per < Ns ? s0 + per : Ns;
    bool nz = false;
    for (int64_t s = s0; s < s1; s++) nz |= Xr[r + ldr * s] != T(0);
    if (nz) atomicOr(&flags[(r / NC) / TS], 1u << (r % NC));
}

// tiles of the lower triangle sorted by work (number of active component pairs, descending; stable): the
// last CTAs of the grid are then the cheap ones and the tail of the launch is short
__global__ void tile_order_kernel(const unsigned* __restrict__ flags, int ntile, int mode, int* __restrict__ order) {
    __shared__ int cnt[256][3];
    const int ntri = ntile * (ntile + 1) / 2, tid = threadIdx.x;
    const int per = (ntri + 255) / 256, lo = min(ntri, tid * per), hi = min(ntri, lo + per);
    auto weight = [&](int t) {
        int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
        while (ti * (ti + 1) / 2 > t) ti--;
        const int tj = t - ti * (ti + 1) / 2;
        const unsigned fa = flags[ti], fb = flags[tj];
        int w = 0;
        for (int c = 0; c < 2; c++) { const int bc = mode == 0 ? c : 1 - c; w += ((fa >> c) & 1u) && ((fb >> bc) & 1u); }
        return w;
    };
    int c0 = 0, c1 = 0, c2 = 0;
    for (int t = lo; t < hi; t++) { const int w = weight(t); c0 += w == 0; c1 += w == 1; c2 += w == 2; }
    cnt[tid][0] = c0; cnt[tid][1] = c1; cnt[tid][2] = c2;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int w = 2; w >= 0; w--)
            for (int i = 0; i < 256; i++) { const int c = cnt[i][w]; cnt[i][w] = run; run += c; }
    }
    __syncthreads();
    int p[3] = {cnt[tid][0], cnt[tid][1], cnt[tid][2]};
    for (int t = lo; t < hi; t++) order[p[weight(t)]++] = t;
}

constexpr int LDP = TS + 8;                 // one sample of one component plane, padded (bank-conflict free)
constexpr int PLANE = KS * LDP + 8;         // component plane

template <typename T, int NC>
__global__ void __launch_bounds__(256, 1)
syrk_dmma_kernel(const T* __restrict__ Xr, int64_t ldr, int64_t P, int64_t Ns, int ntile, int nsplit, int mode,
                 const unsigned* __restrict__ tflags,
                 double* __restrict__ Wk /* [nsplit][Ppad*Ppad] col-major, Ppad = ntile*TS */) {
    constexpr int PER = TS * NC * KS / 256;             // reals per thread per tile per chunk
    constexpr int TILE = NC * PLANE;                    // doubles per staged tile
    extern __shared__ __align__(16) double smem[];
    // decode lower-triangular tile index
    int t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
    while (ti * (ti + 1) / 2 > t) ti--;
    int tj = t - ti * (ti + 1) / 2;
    const bool diag = ti == tj;
    const int split = blockIdx.y;
    const int64_t nchunk_tot = (Ns + KS - 1) / KS;
    const int64_t cper = (nchunk_tot + nsplit - 1) / nsplit;
    const int64_t c_begin = split * cper, c_end = std::min<int64_t>(nchunk_tot, c_begin + cper);

    double* As[2] = {smem, smem + 2 * TILE};
    double* Bs[2] = {smem + TILE, smem + 3 * TILE};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int wm = warp & 1, wn = warp >> 1;

    // component pairs that contribute: mode 0 pairs (c, c); mode 1 pairs (c of A, 1-c of B)
    const unsigned fa = NC == 2 ? tflags[ti] : 1u, fb = NC == 2 ? tflags[tj] : 1u;
    unsigned act = 0;                       // bit c: component c of A is used
    for (int c = 0; c < NC; c++) {
        unsigned bc = mode == 0 ? c : 1 - c;
        if (((fa >> c) & 1u) && ((fb >> bc) & 1u)) act |= 1u << c;
    }
    unsigned needA = act, needB = 0;
    for (int c = 0; c < NC; c++) if ((act >> c) & 1u) needB |= 1u << (mode == 0 ? c : 1 - c);
    if (diag) needA |= needB;
    // a thread always loads the same component: r = (tid + 256 i) % (TS*NC) has the parity of tid
    const int myc = NC == 2 ? (tid & 1) : 0;
    const bool ldA = (needA >> myc) & 1u, ldB = !diag && ((needB >> myc) & 1u);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    T ra[PER], rbv[PER];
    const int64_t rowA = (int64_t)ti * TS * NC, rowB = (int64_t)tj * TS * NC, PR = P * NC;

    auto gload = [&](int64_t chunk) {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            int e = tid + 256 * i;               // e = s * (TS*NC) + r
            int s = e / (TS * NC), r = e - s * (TS * NC);
            int64_t smp = chunk * KS + s;
            bool okA = ldA && smp < Ns && rowA + r < PR, okB = ldB && smp < Ns && rowB + r < PR;
            ra[i] = okA ? Xr[rowA + r + ldr * smp] : T(0);
            rbv[i] = okB ? Xr[rowB + r + ldr * smp] : T(0);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            int e = tid + 256 * i;
            int s = e / (TS * NC), r = e - s * (TS * NC);
            int m = r / NC;
            if (ldA) As[buf][myc * PLANE + s * LDP + m] = (double)ra[i];
            if (ldB) Bs[buf][myc * PLANE + s * LDP + m] = (double)rbv[i];
        }
    };

    if (act && c_begin < c_end) {
        gload(c_begin);
        sstore(0);
    }
    __syncthreads();
    if (act) {
        for (int64_t c = c_begin; c < c_end; c++) {
            const int buf = (int)((c - c_begin) & 1);
            if (c + 1 < c_end) gload(c + 1);
            const double* A = As[buf];
            const double* Bm = diag ? As[buf] : Bs[buf];
#pragma unroll
            for (int comp = 0; comp < NC; comp++) {
                if (!((act >> comp) & 1u)) continue;
                const int bcomp = mode == 0 ? comp : 1 - comp;
                const double* Ap = A + comp * PLANE + wm * 64 + g;
                const double* Bp = Bm + bcomp * PLANE + wn * 32 + g;
#pragma unroll
                for (int k4 = 0; k4 < KS / 4; k4++) {
                    const int sidx = (4 * k4 + tq) * LDP;
                    double a[8], b[4];
#pragma unroll
                    for (int i = 0; i < 8; i++) a[i] = Ap[sidx + i * 8];
#pragma unroll
                    for (int j = 0; j < 4; j++) { double v = Bp[sidx + j * 8]; b[j] = (mode == 1 && comp == 1) ? -v : v; }
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            }
            if (c + 1 < c_end) sstore(buf ^ 1);
            __syncthreads();
        }
    }
    const int64_t Ppad = (int64_t)ntile * TS;
    double* W = Wk + (size_t)split * Ppad * Ppad;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t row = (int64_t)ti * TS + wm * 64 + i * 8 + g;
            int64_t col = (int64_t)tj * TS + wn * 32 + j * 8 + 2 * tq;
            W[row + Ppad * col] = acc[i][j][0];
            W[row + Ppad * (col + 1)] = acc[i][j][1];
        }
}

// ---- FP64 operands: cp.async pipeline (no register staging), 4 stages of 8 samples, one barrier per chunk.
// Plane row pitch LDQ = 132 doubles: the fragment loads (lanes g -> consecutive rows, tq -> consecutive
// samples) hit 4 x 8 distinct banks per half-warp.
constexpr int LDQ = TS + 4;
constexpr int PLQ = KS * LDQ;               // doubles per component plane
constexpr int NSTAGE = 4;

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int sz = valid ? 8 : 0;           // src-size 0: zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int NC, int MODE>
__global__ void __launch_bounds__(256, 1)
syrk_dmma2_kernel(const double* __restrict__ Xr, int64_t ldr, int64_t P, int64_t Ns, int ntile, int nsplit,
                  const unsigned* __restrict__ tflags, const int* __restrict__ order, double* __restrict__ Wk) {
    constexpr int PER = TS * NC * KS / 256;             // elements per thread per tile per chunk
    constexpr int TILE = NC * PLQ;                      // doubles per staged tile
    constexpr int STG = 2 * TILE;                       // A tile | B tile
    constexpr int SPI = 256 / (TS * NC);                // samples covered by one pass of the 256 threads (1 or 2)
    extern __shared__ __align__(16) double smem[];
    int t = order ? order[blockIdx.x] : (int)blockIdx.x;   // heavy (two-component) tiles first, empty ones last
    int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
    while (ti * (ti + 1) / 2 > t) ti--;
    const int tj = t - ti * (ti + 1) / 2;
    const bool diag = ti == tj;
    const int split = blockIdx.y;
    const int64_t nchunk_tot = (Ns + KS - 1) / KS;
    const int64_t cper = (nchunk_tot + nsplit - 1) / nsplit;
    const int64_t c_begin = split * cper, c_end = std::min<int64_t>(nchunk_tot, c_begin + cper);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int wm = warp & 1, wn = warp >> 1;

    const unsigned fa = NC == 2 ? tflags[ti] : 1u, fb = NC == 2 ? tflags[tj] : 1u;
    unsigned act = 0;
    for (int c = 0; c < NC; c++) {
        unsigned bc = MODE == 0 ? c : 1 - c;
        if (((fa >> c) & 1u) && ((fb >> bc) & 1u)) act |= 1u << c;
    }
    unsigned needA = act, needB = 0;
    for (int c = 0; c < NC; c++) if ((act >> c) & 1u) needB |= 1u << (MODE == 0 ? c : 1 - c);
    if (diag) needA |= needB;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // copy coordinates of this thread: element e = tid + 256 i  ->  sample SPI*i + tid / (TS*NC), real row r
    const int r = tid % (TS * NC), s_off = tid / (TS * NC);
    const int myc = NC == 2 ? (r & 1) : 0;               // a thread always copies the same component
    const int64_t rowA = (int64_t)ti * TS * NC, rowB = (int64_t)tj * TS * NC, PR = P * NC;
    // rows beyond P are never copied: their planes stay zero (zeroed below) and only feed padding outputs
    const bool ldA = ((needA >> myc) & 1u) && rowA + r < PR, ldB = !diag && ((needB >> myc) & 1u) && rowB + r < PR;
    const int dst0 = myc * PLQ + s_off * LDQ + r / NC;
    const double* srcA = Xr + rowA + r + ldr * (c_begin * KS + s_off);
    const double* srcB = Xr + rowB + r + ldr * (c_begin * KS + s_off);
    const int64_t step = ldr * SPI;                      // doubles between two copies of this thread

    auto issue = [&](int64_t chunk, int stage) {
        double* As = smem + stage * STG + dst0;
        double* Bs = As + TILE;
        const int64_t smp0 = chunk * KS + s_off;
        const double* pa = srcA + (chunk - c_begin) * KS * ldr;
        const double* pb = srcB + (chunk - c_begin) * KS * ldr;
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const bool ok = smp0 + SPI * i < Ns;
            if (ldA) cp_async8(As + SPI * i * LDQ, ok ? pa : Xr, ok);
            if (ldB) cp_async8(Bs + SPI * i * LDQ, ok ? pb : Xr, ok);
            pa += step; pb += step;
        }
    };

    // one active component (real-parameter NDM: purely real x purely real, imaginary x imaginary): the two
    // planes of a stage hold 16 samples of that component, so a barrier still covers 128 DMMAs per warp
    const bool single = NC == 2 && MODE == 0 && (act == 1u || act == 2u) && needA == act && (diag || needB == act);
    if (single) {
        const int comp = act == 1u ? 0 : 1;
        for (int i = tid; i < NSTAGE * STG; i += 256) smem[i] = 0.0;
        __syncthreads();
        const int64_t smp_begin = c_begin * KS, smp_end = std::min<int64_t>(Ns, c_end * KS);
        const int64_t nch = (smp_end - smp_begin + 2 * KS - 1) / (2 * KS);
        const int m = tid & (TS - 1), sh = tid >> 7;                      // row of the tile, sample parity
        const bool okA = rowA + 2 * m < PR, okB = !diag && rowB + 2 * m < PR;
        const double* sA = Xr + rowA + 2 * m + comp + ldr * (smp_begin + sh);
        const double* sB = Xr + rowB + 2 * m + comp + ldr * (smp_begin + sh);
        auto issue1 = [&](int64_t ch, int stage) {
            double* As = smem + stage * STG + sh * LDQ + m;
            double* Bs = As + TILE;
            const int64_t smp0 = smp_begin + ch * 2 * KS + sh;
            const double* pa = sA + ch * 2 * KS * ldr;
            const double* pb = sB + ch * 2 * KS * ldr;
#pragma unroll
            for (int i = 0; i < KS; i++) {
                const bool ok = smp0 + 2 * i < smp_end;
                if (okA) cp_async8(As + 2 * i * LDQ, ok ? pa : Xr, ok);
                if (okB) cp_async8(Bs + 2 * i * LDQ, ok ? pb : Xr, ok);
                pa += 2 * ldr; pb += 2 * ldr;
            }
        };
#pragma unroll
        for (int p = 0; p < NSTAGE - 1; p++) {
            if (p < nch) issue1(p, p);
            cp_async_commit();
        }
        int stage = 0;
        for (int64_t it = 0; it < nch; it++) {
            cp_async_wait<NSTAGE - 2>();
            __syncthreads();
            const int nstage = stage == 0 ? NSTAGE - 1 : stage - 1;
            if (it + NSTAGE - 1 < nch) issue1(it + NSTAGE - 1, nstage);
            cp_async_commit();
            const double* A = smem + stage * STG;
            const double* Ap = A + wm * 64 + g;
            const double* Bp = (diag ? A : A + TILE) + wn * 32 + g;
#pragma unroll
            for (int k4 = 0; k4 < 2 * KS / 4; k4++) {
                const int sidx = (4 * k4 + tq) * LDQ;
                double a[8], b[4];
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = Ap[sidx + i * 8];
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = Bp[sidx + j * 8];
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            stage = stage == NSTAGE - 1 ? 0 : stage + 1;
        }
        cp_async_wait<0>();
    } else if (NC == 2 && act == 3u) {
        // both component pairs active (complex rows x complex rows): (re, im) stay interleaved in shared memory --
        // 16-byte cp.async copies, 16-byte fragment loads that serve both components (row pitch LDI = 260 doubles:
        // a quarter-warp covers all 32 banks), half the copy and load instructions of the planar path
        constexpr int LDI = 2 * TS + 4;
        constexpr int TILEI = KS * LDI;              // doubles per staged tile (<= TILE)
        static_assert(TILEI <= 2 * PLQ, "interleaved tile must fit the planar stage");
        for (int i = tid; i < NSTAGE * STG; i += 256) smem[i] = 0.0;
        __syncthreads();
        const int64_t nch = c_end - c_begin;
        const int m = tid & (TS - 1), sh = tid >> 7;                      // row of the tile, sample parity
        const bool okA = rowA + 2 * m < PR, okB = !diag && rowB + 2 * m < PR;
        const double* sA = Xr + rowA + 2 * m + ldr * (c_begin * KS + sh);
        const double* sB = Xr + rowB + 2 * m + ldr * (c_begin * KS + sh);
        auto issue2 = [&](int64_t ch, int stage) {
            double* As = smem + stage * STG + sh * LDI + 2 * m;
            double* Bs = As + TILE;
            const int64_t smp0 = (c_begin + ch) * KS + sh;
            const double* pa = sA + ch * KS * ldr;
            const double* pb = sB + ch * KS * ldr;
#pragma unroll
            for (int i = 0; i < KS / 2; i++) {
                const bool ok = smp0 + 2 * i < Ns;
                const int sz = ok ? 16 : 0;
                if (okA) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"((unsigned)__cvta_generic_to_shared(As + 2 * i * LDI)), "l"(ok ? pa : Xr), "r"(sz) : "memory");
                if (okB) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"((unsigned)__cvta_generic_to_shared(Bs + 2 * i * LDI)), "l"(ok ? pb : Xr), "r"(sz) : "memory");
                pa += 2 * ldr; pb += 2 * ldr;
            }
        };
#pragma unroll
        for (int p = 0; p < NSTAGE - 1; p++) {
            if (p < nch) issue2(p, p);
            cp_async_commit();
        }
        int stage = 0;
        for (int64_t it = 0; it < nch; it++) {
            cp_async_wait<NSTAGE - 2>();
            __syncthreads();
            const int nstage = stage == 0 ? NSTAGE - 1 : stage - 1;
            if (it + NSTAGE - 1 < nch) issue2(it + NSTAGE - 1, nstage);
            cp_async_commit();
            const double* A = smem + stage * STG;
            const double2* Ap = reinterpret_cast<const double2*>(A) + wm * 64 + g;
            const double2* Bp = reinterpret_cast<const double2*>(diag ? A : A + TILE) + wn * 32 + g;
#pragma unroll
            for (int k4 = 0; k4 < KS / 4; k4++) {
                const int sidx = (4 * k4 + tq) * (LDI / 2);
                double2 a[8], b[4];
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = Ap[sidx + i * 8];
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = Bp[sidx + j * 8];
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (MODE == 0) {
                            dmma(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
                            dmma(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
                        } else {
                            dmma(acc[i][j][0], acc[i][j][1], a[i].x, b[j].y);
                            dmma(acc[i][j][0], acc[i][j][1], a[i].y, -b[j].x);
                        }
                    }
            }
            stage = stage == NSTAGE - 1 ? 0 : stage + 1;
        }
        cp_async_wait<0>();
    } else if (act) {
        for (int i = tid; i < NSTAGE * STG; i += 256) smem[i] = 0.0;
        __syncthreads();
        const int64_t nch = c_end - c_begin;
#pragma unroll
        for (int p = 0; p < NSTAGE - 1; p++) {
            if (p < nch) issue(c_begin + p, p);
            cp_async_commit();
        }
        int stage = 0;
        for (int64_t it = 0; it < nch; it++) {
            cp_async_wait<NSTAGE - 2>();                // chunk `it` has landed (this thread's copies)
            __syncthreads();                            // ... everyone's; and the stage of chunk it-1 is free again
            const int nstage = stage == 0 ? NSTAGE - 1 : stage - 1;     // (it + NSTAGE - 1) % NSTAGE
            if (it + NSTAGE - 1 < nch) issue(c_begin + it + NSTAGE - 1, nstage);
            cp_async_commit();
            const double* A = smem + stage * STG;
            const double* Bm = diag ? A : A + TILE;
#pragma unroll
            for (int comp = 0; comp < NC; comp++) {
                if (!((act >> comp) & 1u)) continue;
                const int bcomp = MODE == 0 ? comp : 1 - comp;
                const double* Ap = A + comp * PLQ + wm * 64 + g;
                const double* Bp = Bm + bcomp * PLQ + wn * 32 + g;
#pragma unroll
                for (int k4 = 0; k4 < KS / 4; k4++) {
                    const int sidx = (4 * k4 + tq) * LDQ;
                    double a[8], b[4];
#pragma unroll
                    for (int i = 0; i < 8; i++) a[i] = Ap[sidx + i * 8];
#pragma unroll
                    for (int j = 0; j < 4; j++) { double v = Bp[sidx + j * 8]; b[j] = (MODE == 1 && comp == 1) ? -v : v; }
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            }
            stage = stage == NSTAGE - 1 ? 0 : stage + 1;
        }
        cp_async_wait<0>();
    }
    const int64_t Ppad = (int64_t)ntile * TS;
    double* W = Wk + (size_t)split * Ppad * Ppad;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t row = (int64_t)ti * TS + wm * 64 + i * 8 + g;
            int64_t col = (int64_t)tj * TS + wn * 32 + j * 8 + 2 * tq;
            W[row + Ppad * col] = acc[i][j][0];
            W[row + Ppad * (col + 1)] = acc[i][j][1];
        }
}

// S[k,l] from the lower tile triangle: sum the splits in order, scale, mirror (Hermitian)
// out_complex: S complex (re from Wre, im from Wim); else real.
template <typename TS_>
__global__ void syrk_finalize_kernel(const double* __restrict__ Wre, const double* __restrict__ Wim, int nsplit,
                                     int64_t Ppad, int64_t P, double scale, int out_complex, TS_* __restrict__ S) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= P) return;
    for (int64_t l = blockIdx.y; l < P; l += gridDim.y) {        // grid.y is capped at 65535
        bool lower = (k / TS) >= (l / TS);
        int64_t r = lower ? k : l, c = lower ? l : k;
        double re = 0.0, im = 0.0;
        for (int s = 0; s < nsplit; s++) {
            re += Wre[(size_t)s * Ppad * Ppad + r + Ppad * c];
            if (Wim) im += Wim[(size_t)s * Ppad * Ppad + r + Ppad * c];
        }
        if (!lower) im = -im;
        if (out_complex) { S[2 * (k + P * l)] = (TS_)(re * scale); S[2 * (k + P * l) + 1] = (TS_)(im * scale); }
        else S[k + P * l] = (TS_)(re * scale);
    }
}

template <typename T, int NC>
int launch_syrk(nq_ctx_t ctx, const void* X, int64_t ldr, int64_t P, int64_t Ns, int ntile, int nsplit, int mode, double* W) {
    size_t smem = (size_t)4 * NC * PLANE * sizeof(double);
    const int ntri = ntile * (ntile + 1) / 2;
    unsigned* flags = (unsigned*)nq_scratch(ctx, SL_W3, ((size_t)ntile + ntri) * sizeof(unsigned) + 32);
    if (!flags) return NQ_ERR_ALLOC;
    int* order = NC == 2 ? (int*)(flags + ntile) : nullptr;
    if (NC == 2 && mode == 0) {       // mode 1 (same launch sequence) reuses the flags
        if (ctx->hint_P == P && (int)ctx->hint_tile_flags.size() == ntile) {
            // the caller knows the structure of the rows (nq_sr_hint_row_planes): no pass over O
            NQ_CUDA(ctx, cudaMemcpyAsync(flags, ctx->hint_tile_flags.data(), (size_t)ntile * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
        } else {
            NQ_CUDA(ctx, cudaMemsetAsync(flags, 0, (size_t)ntile * sizeof(unsigned), ctx->stream));
            dim3 g((unsigned)((P * NC + 255) / 256), (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, Ns / 64)));
            NQ_LAUNCH(ctx, (tile_activity_kernel<T, NC>), g, 256, 0, (const T*)X, ldr, P, Ns, flags);
        }
    }
    if (NC == 2) NQ_LAUNCH(ctx, tile_order_kernel, 1, 256, 0, (const unsigned*)flags, ntile, mode, order);
    dim3 grid((unsigned)ntri, (unsigned)nsplit);
    if (sizeof(T) == 8) {
        size_t smem2 = (size_t)NSTAGE * 2 * NC * PLQ * sizeof(double);
        if (mode == 0) {
            auto kern2 = syrk_dmma2_kernel<NC, 0>;
            NQ_CUDA(ctx, cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            NQ_LAUNCH(ctx, kern2, grid, 256, smem2, (const double*)X, ldr, P, Ns, ntile, nsplit, (const unsigned*)flags, (const int*)order, W);
        } else {
            auto kern2 = syrk_dmma2_kernel<NC, 1>;
            NQ_CUDA(ctx, cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            NQ_LAUNCH(ctx, kern2, grid, 256, smem2, (const double*)X, ldr, P, Ns, ntile, nsplit, (const unsigned*)flags, (const int*)order, W);
        }
        return NQ_OK;
    }
    auto kern = syrk_dmma_kernel<T, NC>;
    NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NQ_LAUNCH(ctx, kern, grid, 256, smem, (const T*)X, ldr, P, Ns, ntile, nsplit, mode, (const unsigned*)flags, W);
    return NQ_OK;
}

// ======================================================================================
// K8: blocked Cholesky  A = L L^H  (lower, column-major, in place) on double or complex double
// ======================================================================================
constexpr int NB = 32;

__device__ __forceinline__ double mulc(double a, double b) { return a * b; }            // a * conj(b)
__device__ __forceinline__ cxd mulc(cxd a, cxd b) { return cxd(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im); }
__device__ __forceinline__ double divr(double a, double r) { return a / r; }
__device__ __forceinline__ cxd divr(cxd a, double r) { return cxd(a.re / r, a.im / r); }

__device__ __forceinline__ double lane_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ cxd lane_bcast(cxd v, int src) { return cxd(__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src)); }

// factor the diagonal block (shared memory Lb[NB][NB+1]) with one warp: lane r keeps row r in registers, the
// pivot and the column below it travel by shuffles (right-looking, fully unrolled -- the shared-memory
// read-modify-write version spent ~20 us per block on load/store round trips)
template <typename E>
__device__ void potf2_warp(E (*Lb)[NB + 1], double* Dinv, int nb, int64_t j0, int* info) {
    const int r = threadIdx.x & 31;
    E row[NB];
#pragma unroll
    for (int c = 0; c < NB; c++) row[c] = Lb[r][c];
#pragma unroll
    for (int c = 0; c < NB; c++) {
        double d = __shfl_sync(0xffffffffu, real_part(row[c]), c);
        if (c >= nb) d = 1.0;
        else if (!(d > 0.0)) { if (r == 0) atomicCAS(info, -1, (int)(j0 + c)); d = 1.0; }
        // the pivot chain is the critical path of the whole factorisation: reciprocal square root + multiplies
        // instead of sqrt + divisions (1/l is kept for the TRSM rows and the triangular solves)
        const double inv = rsqrt(d);
        if (r == c) { row[c] = from_real<E, double>(d * inv); Dinv[c] = inv; }
        else if (r > c) row[c] = rscale(inv, row[c]);
        const E lrc = row[c];
#pragma unroll
        for (int c2 = c + 1; c2 < NB; c2++) {
            const E t = lane_bcast(lrc, c2);            // L[c2][c]
            if (r >= c2) row[c2] -= mulc(lrc, t);
        }
    }
#pragma unroll
    for (int c = 0; c < NB; c++) Lb[r][c] = row[c];
    __syncwarp();
}

// panel: every CTA factors the diagonal block redundantly in shared memory; CTA 0 writes it back;
// all CTAs apply X = A[i, j0:j0+nb] L_jj^{-H} to their rows below the block.
template <typename E>
__global__ void chol_panel_kernel(E* __restrict__ A, int64_t P, int64_t j0, int nb, int* __restrict__ info) {
    __shared__ E Lb[NB][NB + 1];
    __shared__ double Dinv[NB];
    const int tid = threadIdx.x;
    for (int i = tid; i < NB * NB; i += blockDim.x) {
        int r = i % NB, c = i / NB;
        Lb[r][c] = (r < nb && c < nb && c <= r) ? A[(j0 + r) + P * (j0 + c)] : make_zero<E>();
    }
    __syncthreads();
    if (tid < 32) potf2_warp<E>(Lb, Dinv, nb, j0, info);
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int i = tid; i < NB * NB; i += blockDim.x) {
            int r = i % NB, c = i / NB;
            if (r < nb && c < nb) A[(j0 + r) + P * (j0 + c)] = c <= r ? Lb[r][c] : make_zero<E>();
        }
    }
    int64_t row = j0 + nb + blockIdx.x * (int64_t)blockDim.x + tid;
    if (row >= P) return;
    E x[NB];
#pragma unroll
    for (int c = 0; c < NB; c++) x[c] = c < nb ? A[row + P * (j0 + c)] : make_zero<E>();
#pragma unroll
    for (int c = 0; c < NB; c++) {
        if (c < nb) {
            E v = x[c];
#pragma unroll
            for (int c2 = 0; c2 < NB; c2++) if (c2 < c) v -= mulc(x[c2], Lb[c][c2]);
            x[c] = rscale(Dinv[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < NB; c++) if (c < nb) A[row + P * (j0 + c)] = x[c];
}


// trailing update on the FP64 tensor pipe: A[i,l] -= sum_c X[i,c] conj(X[l,c]) for the lower 64x64 tiles, K = nb <= 32.
// Panel rows staged as planar (re | im) [c][row] with pitch 68 (conflict-free fragment loads); 8 warps as
// 2 (rows) x 4 (cols), warp tile 32x16 = 4x2 mma tiles; complex: re = rr + ii, im = ir - ri.
constexpr int LDU = 64 + 4;
template <typename E>
__global__ void __launch_bounds__(256) chol_update_dmma_kernel(E* __restrict__ A, int64_t P, int64_t j0, int nb) {
    constexpr bool CPLX = sizeof(E) == 16;
    constexpr int NPL = CPLX ? 2 : 1;
    extern __shared__ __align__(16) double usm[];
    double* Xi = usm;                          // [NPL][NB][LDU]
    double* Xl = usm + NPL * NB * LDU;
    int t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
    while (ti * (ti + 1) / 2 > t) ti--;
    const int tj = t - ti * (ti + 1) / 2;
    const bool diag = ti == tj;
    const int64_t base = j0 + nb;
    const int64_t i0 = base + (int64_t)ti * 64, l0 = base + (int64_t)tj * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3, wm = warp & 1, wn = warp >> 1;
    const double* Ad = reinterpret_cast<const double*>(A);
    for (int e = tid; e < NB * 64; e += 256) {
        const int r = e & 63, c = e >> 6;
        const bool okc = c < nb;
        {
            const bool ok = okc && i0 + r < P;
            const int64_t at = (i0 + r) + P * (j0 + c);
            Xi[c * LDU + r] = ok ? Ad[at * NPL] : 0.0;
            if (CPLX) Xi[(NB + c) * LDU + r] = ok ? Ad[at * NPL + 1] : 0.0;
        }
        if (!diag) {
            const bool ok = okc && l0 + r < P;
            const int64_t at = (l0 + r) + P * (j0 + c);
            Xl[c * LDU + r] = ok ? Ad[at * NPL] : 0.0;
            if (CPLX) Xl[(NB + c) * LDU + r] = ok ? Ad[at * NPL + 1] : 0.0;
        }
    }
    __syncthreads();
    const double* Bq = diag ? Xi : Xl;
    double cre[4][2][2], cim[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) { cre[i][j][0] = cre[i][j][1] = 0.0; cim[i][j][0] = cim[i][j][1] = 0.0; }
    const double* Ap = Xi + wm * 32 + g;
    const double* Bp = Bq + wn * 16 + g;
#pragma unroll
    for (int k4 = 0; k4 < NB / 4; k4++) {
        const int sidx = (4 * k4 + tq) * LDU;
        double ar[4], ai[4], br[2], bi[2], nbi[2];
#pragma unroll
        for (int i = 0; i < 4; i++) { ar[i] = Ap[sidx + i * 8]; if (CPLX) ai[i] = Ap[NB * LDU + sidx + i * 8]; }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            br[j] = Bp[sidx + j * 8];
            if (CPLX) { bi[j] = Bp[NB * LDU + sidx + j * 8]; nbi[j] = -bi[j]; }
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                dmma(cre[i][j][0], cre[i][j][1], ar[i], br[j]);
                if (CPLX) {
                    dmma(cre[i][j][0], cre[i][j][1], ai[i], bi[j]);
                    dmma(cim[i][j][0], cim[i][j][1], ai[i], br[j]);
                    dmma(cim[i][j][0], cim[i][j][1], ar[i], nbi[j]);
                }
            }
    }
    double* Aw = reinterpret_cast<double*>(A);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int64_t row = i0 + wm * 32 + i * 8 + g, col = l0 + wn * 16 + j * 8 + 2 * tq + h;
                if (row < P && col < P && row >= col) {
                    const int64_t at = (row + P * col) * NPL;
                    Aw[at] -= cre[i][j][h];
                    if (CPLX) Aw[at + 1] -= cim[i][j][h];
                }
            }
}

template <typename E> __device__ __forceinline__ E mk(double re, double im);
template <> __device__ __forceinline__ double mk<double>(double re, double) { return re; }
template <> __device__ __forceinline__ cxd mk<cxd>(double re, double im) { return cxd(re, im); }

// forward L y = b (right-looking, coalesced row updates) then backward L^H x = y (left-looking,
// coalesced column dot products); a single CTA walks the block columns.
template <typename E>
__global__ void __launch_bounds__(512) chol_solve_kernel(const E* __restrict__ L, int64_t P, E* __restrict__ x) {
    __shared__ E blk[NB];
    __shared__ E Lb[NB][NB + 1];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int64_t j0 = 0; j0 < P; j0 += NB) {
        const int nb = (int)((P - j0) < NB ? (P - j0) : NB);
        for (int i = tid; i < NB * NB; i += nt) {
            int r = i % NB, c = i / NB;
            Lb[r][c] = (r < nb && c < nb) ? L[(j0 + r) + P * (j0 + c)] : make_zero<E>();
        }
        __syncthreads();
        if (warp == 0) {
            E v = lane < nb ? x[j0 + lane] : make_zero<E>();
            const double dinv = lane < nb ? 1.0 / real_part(Lb[lane][lane]) : 1.0;     // off the serial chain
            for (int c = 0; c < nb; c++) {
                E yc = rscale(__shfl_sync(0xffffffffu, dinv, c), lane_bcast(v, c));
                if (lane == c) v = yc;
                if (lane > c && lane < nb) v -= Lb[lane][c] * yc;
            }
            if (lane < nb) { x[j0 + lane] = v; blk[lane] = v; }
        }
        __syncthreads();
        for (int64_t i = j0 + nb + tid; i < P; i += nt) {
            E v = x[i];
            for (int c = 0; c < nb; c++) v -= L[i + P * (j0 + c)] * blk[c];
            x[i] = v;
        }
        __syncthreads();
    }
    const int64_t nblk = (P + NB - 1) / NB;
    for (int64_t bi = nblk - 1; bi >= 0; bi--) {
        const int64_t j0 = bi * NB;
        const int nb = (int)((P - j0) < NB ? (P - j0) : NB);
        for (int i = tid; i < NB * NB; i += nt) {
            int r = i % NB, c = i / NB;
            Lb[r][c] = (r < nb && c < nb) ? L[(j0 + r) + P * (j0 + c)] : make_zero<E>();
        }
        // rhs_c = y_c - sum_{r >= j0+nb} conj(L[r, j0+c]) x_r
        for (int c = warp; c < nb; c += nw) {
            E acc = make_zero<E>();
            for (int64_t r = j0 + nb + lane; r < P; r += 32) acc += mulc(x[r], L[r + P * (j0 + c)]);
            acc = warp_sum(acc);
            if (lane == 0) blk[c] = x[j0 + c] - acc;
        }
        __syncthreads();
        if (warp == 0) {
            E v = lane < nb ? blk[lane] : make_zero<E>();
            const double dinv = lane < nb ? 1.0 / real_part(Lb[lane][lane]) : 1.0;
            for (int c = nb - 1; c >= 0; c--) {
                E xc = rscale(__shfl_sync(0xffffffffu, dinv, c), lane_bcast(v, c));
                if (lane == c) v = xc;
                if (lane < c) v -= mulc(xc, Lb[c][lane]);   // conj(L[c][lane]) x_c
            }
            if (lane < nb) x[j0 + lane] = v;
        }
        __syncthreads();
    }
}

// Multi-CTA triangular solves, one launch per block column.  Every CTA solves the 32x32 diagonal system
// redundantly (reciprocal diagonals, ~3 us) and then updates its share of the remaining right-hand side:
//   FWD  L y = b:    y_blk = L_jj^{-1} b_blk;   b[i] -= L[i, blk] y_blk        for the rows i below the block
//   BWD  L^H x = y:  x_blk = L_jj^{-H} y_blk;   y[k] -= conj(L[blk, k])^T x_blk  for the columns k before it
// The block result goes to a separate vector (other CTAs still read the block of the right-hand side).
template <typename E, bool FWD>
__global__ void __launch_bounds__(256) chol_trsv_step_kernel(const E* __restrict__ L, int64_t P, int64_t j0, int nb,
                                                             E* __restrict__ rhs, E* __restrict__ out) {
    __shared__ E Lb[NB][NB + 1];
    __shared__ E xb[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NB * NB; i += 256) {
        int r = i % NB, c = i / NB;
        Lb[r][c] = (r < nb && c < nb && c <= r) ? L[(j0 + r) + P * (j0 + c)] : make_zero<E>();
    }
    __syncthreads();
    if (warp == 0) {
        E v = lane < nb ? rhs[j0 + lane] : make_zero<E>();
        const double dinv = lane < nb ? 1.0 / real_part(Lb[lane][lane]) : 1.0;
        if (FWD) {
            for (int c = 0; c < nb; c++) {
                E yc = rscale(__shfl_sync(0xffffffffu, dinv, c), lane_bcast(v, c));
                if (lane == c) v = yc;
                if (lane > c && lane < nb) v -= Lb[lane][c] * yc;
            }
        } else {
            for (int c = nb - 1; c >= 0; c--) {
                E xc = rscale(__shfl_sync(0xffffffffu, dinv, c), lane_bcast(v, c));
                if (lane == c) v = xc;
                if (lane < c) v -= mulc(xc, Lb[c][lane]);   // conj(L[c][lane]) x_c
            }
        }
        xb[lane] = lane < nb ? v : make_zero<E>();
        if (blockIdx.x == 0 && lane < nb) out[j0 + lane] = v;
    }
    __syncthreads();
    if (FWD) {
        const int64_t i = j0 + nb + blockIdx.x * 256 + tid;
        if (i < P) {
            E v = rhs[i];
#pragma unroll 8
            for (int c = 0; c < NB; c++) if (c < nb) v -= L[i + P * (j0 + c)] * xb[c];
            rhs[i] = v;
        }
    } else {
        const int64_t k = blockIdx.x * 256 + tid;
        if (k < j0) {
            E v = rhs[k];
            const E* col = L + j0 + P * k;
#pragma unroll 8
            for (int r = 0; r < NB; r++) if (r < nb) v -= mulc(xb[r], col[r]);     // conj(L[j0+r, k]) x_r
            rhs[k] = v;
        }
    }
}

template <typename E>
__global__ void scale_diag_kernel(E* __restrict__ A, int64_t P, typename elem_traits<E>::real f) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < P) A[i + P * i] = rscale(f, A[i + P * i]);
}

template <typename E>
__global__ void add_diag_kernel(E* __restrict__ A, int64_t P, double eps) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < P) A[i + P * i] += from_real<E, double>(eps);
}

// --------------------------------------------------------------------------------------
// Fused Cholesky solve: ONE persistent cooperative launch for the whole factorisation and both triangular solves.
// Per block column (NB = 32):
//   [panel]  the CTAs that own rows below the block (and CTA 0) factor the diagonal block in one warp (left-looking
//            loop over shared memory: the fully unrolled register version is 64 KB of straight-line code that ran at
//            instruction-fetch speed, 24 us per block), solve the block of the right-hand side with it -- the forward
//            substitution rides on the panel, b is one more row of A -- and their threads apply the TRSM to their rows
//            (right-looking in registers: 31 independent FMAs per step instead of a 496-long dependent chain);
//   grid barrier;
//   [update] the 32x32 blocks of the trailing matrix are dealt to WARPS: DMMA fragments straight from L2, no shared
//            memory staging, no CTA barrier;
//   grid barrier.
// Backward solve L^H x = y: per block a barrier, the 32-step triangular solve (redundant per CTA) and one fused
// multiply-add pass of every thread over its own entry of y.  Loads of data produced by other CTAs go through L2.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ double ldcg_e(const double* p) { return __ldcg(p); }
__device__ __forceinline__ cxd ldcg_e(const cxd* p) { const double2 v = __ldcg(reinterpret_cast<const double2*>(p)); return cxd(v.x, v.y); }

__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& target, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblocks;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// 1/sqrt(d) on the serial pivot chain: MUFU.RSQ64H seed + two Newton steps (the library rsqrt is ~3x longer)
__device__ __forceinline__ double fast_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double hd = 0.5 * d;
    double e = fma(-hd * y, y, 0.5);        // (1 - d y^2) / 2
    y = fma(y, e, y);
    e = fma(-hd * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hd * y, y, 0.5);
    return fma(y, e, y);
}

// left-looking Cholesky of the block in shared memory, one warp, lane = row
template <typename E>
__device__ void potf2_warp_ll(E (*Lb)[NB + 1], double* Dinv, int nb, int64_t j0, int* info) {
    const int r = threadIdx.x & 31;
    for (int c = 0; c < nb; c++) {
        E a0 = make_zero<E>(), a1 = make_zero<E>();
        int k = 0;
        for (; k + 1 < c; k += 2) { a0 += mulc(Lb[r][k], Lb[c][k]); a1 += mulc(Lb[r][k + 1], Lb[c][k + 1]); }
        if (k < c) a0 += mulc(Lb[r][k], Lb[c][k]);
        const E sv = Lb[r][c] - (a0 + a1);
        double d = __shfl_sync(0xffffffffu, real_part(sv), c);
        if (!(d > 0.0)) { if (r == 0) atomicCAS(info, -1, (int)(j0 + c)); d = 1.0; }
        const double inv = (d > 1e-280 && d < 1e280) ? fast_rsqrt(d) : rsqrt(d);
        if (r == c) { Lb[c][c] = from_real<E, double>(d * inv); Dinv[c] = inv; }
        else if (r > c && r < nb) Lb[r][c] = rscale(inv, sv);
        __syncwarp();
    }
}

template <typename E>
__global__ void __launch_bounds__(256) chol_fused_kernel(E* __restrict__ A, int64_t P, E* __restrict__ x,
                                                         int* __restrict__ info, unsigned* __restrict__ bar,
                                                         unsigned long long* __restrict__ prof) {
    constexpr bool CPLX = sizeof(E) == 16;
    constexpr int NPL = CPLX ? 2 : 1;
    constexpr int KU = CPLX ? 2 : 8;           // k-steps of 4 whose fragments are in flight together
    unsigned long long t_prev = 0, t_acc[6] = {0, 0, 0, 0, 0, 0};
    auto tick = [&](int slot) {
        if (prof && blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (slot >= 0) t_acc[slot] += t - t_prev;
            t_prev = t;
        }
    };
    tick(-1);
    __shared__ E Lb[NB + 1][NB + 1];           // row NB: the block of the right-hand side (factor_block)
    __shared__ double Dinv[NB];
    __shared__ double dsv[NB];
    __shared__ E yb[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nblocks = gridDim.x;
    unsigned target = 0;
    const int64_t nblk = (P + NB - 1) / NB;
    const int g = lane >> 2, tq = lane & 3;
    const double* Ad = reinterpret_cast<const double*>(A);
    double* Aw = reinterpret_cast<double*>(A);
    // factor the diagonal block at j0 (already updated, in Lb) and solve its block of the right-hand side, CTA 0, all 256
    // threads: right-looking elimination in shared memory on the UNSCALED columns (A[r][c2] -= A[r][c] conj(A[c2][c]) / d_c,
    // one barrier per column, the pivot reciprocal is recomputed by every thread), the right-hand side rides along as
    // row NB (it holds conj(y) l_cc: the last row of the factor of the bordered matrix).  Results go straight to A and
    // x -- the other CTAs only read them after the next grid barrier.
    auto factor_block = [&](int64_t j0, int nb) {
        if (tid < NB) Lb[NB][tid] = tid < nb ? conj(ldcg_e(x + j0 + tid)) : make_zero<E>();
        int er[5], ec[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int e = tid + 256 * q;
            er[q] = e / NB; ec[q] = e % NB;
            const bool ok = e < (NB + 1) * NB && ec[q] < nb && ((er[q] >= ec[q] && er[q] < nb) || er[q] == NB);
            if (!ok) ec[q] = -1;
        }
        __syncthreads();
        for (int c = 0; c < nb; c++) {
            double d = real_part(Lb[c][c]);
            if (!(d > 0.0)) { if (tid == 0) atomicCAS(info, -1, (int)(j0 + c)); d = 1.0; }
            if (tid == 0) dsv[c] = d;
            const double rd = 1.0 / d;
#pragma unroll
            for (int q = 0; q < 5; q++)
                if (ec[q] > c) Lb[er[q]][ec[q]] -= rscale(rd, mulc(Lb[er[q]][c], Lb[ec[q]][c]));
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int r = er[q], c = ec[q];
            if (c < 0) continue;
            const double dd = dsv[c];
            const double inv = (dd > 1e-280 && dd < 1e280) ? fast_rsqrt(dd) : rsqrt(dd);
            if (r == NB) x[j0 + c] = conj(rscale(inv, Lb[NB][c]));
            else A[(j0 + r) + P * (j0 + c)] = r == c ? from_real<E, double>(dd * inv) : rscale(inv, Lb[r][c]);
        }
        // the strict upper triangle of the block is not referenced by the solves (the legacy path zeroes it)
        __syncthreads();
    };
    if (blockIdx.x == 0) {
        const int nb0 = (int)(P < NB ? P : NB);
        for (int i = tid; i < NB * NB; i += 256) {
            const int r = i % NB, c = i / NB;
            Lb[r][c] = (r < nb0 && c < nb0 && c <= r) ? A[r + P * c] : make_zero<E>();
        }
        factor_block(0, nb0);
    }
    grid_sync(bar, target, nblocks);
    for (int64_t kb = 0; kb < nblk; kb++) {
        const int64_t j0 = kb * NB;
        const int nb = (int)((P - j0) < NB ? (P - j0) : NB);
        const int64_t base = j0 + nb, below = P - base;
        if (below <= 0) break;
        // ---- panel: the CTAs that own rows below the block load the factored block and y_blk, then TRSM their rows
        // rows are dealt round-robin to ALL CTAs (row = base + CTA + nblocks * thread): with contiguous blocks of 256 rows only
        // below / 256 CTAs worked while the others waited at the barrier (cfg3: 21 of 148, 41 us per step)
        const bool has_rows = (int64_t)blockIdx.x < below;
        if (has_rows) {
            for (int i = tid; i < NB * NB; i += 256) {
                const int r = i % NB, c = i / NB;
                Lb[r][c] = (r < nb && c < nb && c <= r) ? ldcg_e(A + (j0 + r) + P * (j0 + c)) : make_zero<E>();
            }
            if (tid < NB) yb[tid] = tid < nb ? ldcg_e(x + j0 + tid) : make_zero<E>();
            __syncthreads();
            if (tid < NB) Dinv[tid] = tid < nb ? 1.0 / real_part(Lb[tid][tid]) : 1.0;
            __syncthreads();
            tick(0);
            const int64_t row = base + (int64_t)blockIdx.x + (int64_t)nblocks * tid;
            if (row < P) {
                E xr[NB];
#pragma unroll
                for (int c = 0; c < NB; c++) xr[c] = c < nb ? ldcg_e(A + row + P * (j0 + c)) : make_zero<E>();
                E bacc = ldcg_e(x + row);
#pragma unroll
                for (int c = 0; c < NB; c++) {
                    if (c < nb) {
                        const E xc = rscale(Dinv[c], xr[c]);
                        xr[c] = xc;
#pragma unroll
                        for (int c2 = c + 1; c2 < NB; c2++) xr[c2] -= mulc(xc, Lb[c2][c]);
                        bacc -= xc * yb[c];
                    }
                }
#pragma unroll
                for (int c = 0; c < NB; c++) if (c < nb) A[row + P * (j0 + c)] = xr[c];
                x[row] = bacc;
            }
            __syncthreads();
        }
        tick(1);
        grid_sync(bar, target, nblocks);
        tick(2);
        // ---- look-ahead on CTA 0: update the NEXT diagonal block with this panel, factor it and solve its block of b
        //      while the other CTAs run the trailing update (they skip item 0)
        if (blockIdx.x == 0) {
            const int nbn = (int)((P - base) < NB ? (P - base) : NB);
            __shared__ E Xn[NB][NB + 1];
            for (int i = tid; i < NB * NB; i += 256) {
                const int r = i % NB, c = i / NB;
                const bool ok = r < nbn && c < nb;
                Xn[r][c] = ok ? ldcg_e(A + (base + r) + P * (j0 + c)) : make_zero<E>();
            }
            __syncthreads();
            for (int i = tid; i < NB * NB; i += 256) {
                const int r = i % NB, c = i / NB;
                E v = make_zero<E>();
                if (r < nbn && c <= r) {
                    v = ldcg_e(A + (base + r) + P * (base + c));
                    for (int q = 0; q < NB; q++) v -= mulc(Xn[r][q], Xn[c][q]);
                }
                Lb[r][c] = v;
            }
            factor_block(base, nbn);
        }
        // ---- trailing update: 32x32 blocks dealt to warps, fragments straight from L2 (nb == NB here)
        if (blockIdx.x != 0 || nblocks == 1) {
            const int64_t n32 = (below + 31) / 32;
            const int64_t nitems = n32 * (n32 + 1) / 2;
            const int64_t wid = nblocks == 1 ? warp : (int64_t)(blockIdx.x - 1) * 8 + warp;
            const int64_t wstride = nblocks == 1 ? 8 : (int64_t)(nblocks - 1) * 8;
            for (int64_t t = 1 + wid; t < nitems; t += wstride) {
                int bi = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
                while ((int64_t)(bi + 1) * (bi + 2) / 2 <= t) bi++;
                while ((int64_t)bi * (bi + 1) / 2 > t) bi--;
                const int bj = (int)(t - (int64_t)bi * (bi + 1) / 2);
                const int64_t i0 = base + (int64_t)bi * 32, l0 = base + (int64_t)bj * 32;
                // C_new = C_old - A B^H: the accumulators START from the C tile (all its loads are issued here, in flight
                // together with the operand fragments) and the A fragments are negated, so the epilogue is a pure store.
                // The first version read-modified-wrote the tile after the k loop: 30 % of the kernel's samples waited on
                // those loads (profiles/r2n_chol_c128_lines_before.txt).
                double cre[4][4][2], cim[CPLX ? 4 : 1][CPLX ? 4 : 1][2];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int64_t row = i0 + i * 8 + g, col = l0 + j * 8 + 2 * tq + h;
                            const bool ok = row < P && col < P && row >= col;
                            const int64_t at = (row + P * col) * NPL;
                            cre[i][j][h] = ok ? __ldcg(Ad + at) : 0.0;
                            if (CPLX) cim[i][j][h] = ok ? __ldcg(Ad + at + 1) : 0.0;
                        }
#pragma unroll 1
                for (int k0 = 0; k0 < NB / 4; k0 += KU) {
                    double ar[KU][4], ai[CPLX ? KU : 1][4], br[KU][4], bi_[CPLX ? KU : 1][4];
#pragma unroll
                    for (int u = 0; u < KU; u++) {
                        const int64_t colo = P * (j0 + 4 * (k0 + u) + tq);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int64_t ra = i0 + 8 * i + g, rb = l0 + 8 * i + g;
                            if (CPLX) {
                                const double2 va = ra < P ? __ldcg(reinterpret_cast<const double2*>(Ad) + ra + colo) : make_double2(0.0, 0.0);
                                const double2 vb = rb < P ? __ldcg(reinterpret_cast<const double2*>(Ad) + rb + colo) : make_double2(0.0, 0.0);
                                ar[u][i] = -va.x; ai[u][i] = -va.y; br[u][i] = vb.x; bi_[u][i] = vb.y;
                            } else {
                                ar[u][i] = ra < P ? -__ldcg(Ad + ra + colo) : 0.0;
                                br[u][i] = rb < P ? __ldcg(Ad + rb + colo) : 0.0;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < KU; u++)
#pragma unroll
                        for (int i = 0; i < 4; i++)
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                dmma(cre[i][j][0], cre[i][j][1], ar[u][i], br[u][j]);
                                if (CPLX) {
                                    dmma(cre[i][j][0], cre[i][j][1], ai[u][i], bi_[u][j]);
                                    dmma(cim[i][j][0], cim[i][j][1], ai[u][i], br[u][j]);
                                    dmma(cim[i][j][0], cim[i][j][1], ar[u][i], -bi_[u][j]);
                                }
                            }
                }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int64_t row = i0 + i * 8 + g, col = l0 + j * 8 + 2 * tq + h;
                            if (row < P && col < P && row >= col) {
                                const int64_t at = (row + P * col) * NPL;
                                Aw[at] = cre[i][j][h];
                                if (CPLX) Aw[at + 1] = cim[i][j][h];
                            }
                        }
            }
        }
        __syncthreads();
        tick(3);
        grid_sync(bar, target, nblocks);
        tick(4);
    }
    // ---- backward solve L^H x = y: only the CTAs that own entries of y take part (their own barrier counter).
    // Thread k owns y[k] in a REGISTER for the whole phase (only it updates that entry); per block: the owners of the
    // block's entries publish them, barrier, every CTA solves the 32x32 triangular system redundantly (the diagonal block
    // and this thread's 32 entries of L for the step were prefetched before the barrier: they are final since the
    // factorisation), one fused multiply-add pass over the registers.
    const unsigned nb2 = (unsigned)((P + 255) / 256) < nblocks ? (unsigned)((P + 255) / 256) : nblocks;
    if (blockIdx.x >= nb2) return;
    unsigned* bar2 = bar + 1;
    unsigned target2 = 0;
    const int64_t myk = (int64_t)blockIdx.x * 256 + tid;            // nb2 * 256 >= P: one entry per thread
    E myy = myk < P ? ldcg_e(x + myk) : make_zero<E>();
    // software pipeline: the (static) diagonal block and column segment of block bi - 1 are fetched while block bi is solved
    E dnext[4], lseg[NB], lnext[NB];
    auto fetch = [&](int64_t bi, E* dblk, E* seg) {
        const int64_t j0 = bi * NB;
        const int nb = (int)((P - j0) < NB ? (P - j0) : NB);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = tid + 256 * q, r = i % NB, c = i / NB;
            dblk[q] = (r < nb && c < nb) ? ldcg_e(A + (j0 + r) + P * (j0 + c)) : make_zero<E>();
        }
        if (myk < j0) {
            const E* col = A + j0 + P * myk;
#pragma unroll
            for (int r = 0; r < NB; r++) seg[r] = r < nb ? ldcg_e(col + r) : make_zero<E>();
        }
    };
    fetch(nblk - 1, dnext, lnext);
    for (int64_t bi = nblk - 1; bi >= 0; bi--) {
        const int64_t j0 = bi * NB;
        const int nb = (int)((P - j0) < NB ? (P - j0) : NB);
#pragma unroll
        for (int q = 0; q < 4; q++) { const int i = tid + 256 * q; Lb[i % NB][i / NB] = dnext[q]; }
#pragma unroll
        for (int r = 0; r < NB; r++) lseg[r] = lnext[r];
        if (myk >= j0 && myk < j0 + nb) x[myk] = myy;                // publish y_blk (final value of these entries)
        if (bi > 0) fetch(bi - 1, dnext, lnext);
        grid_sync(bar2, target2, nb2);
        if (warp == 0) {
            E v = lane < nb ? ldcg_e(x + j0 + lane) : make_zero<E>();
            const double dinv = lane < nb ? 1.0 / real_part(Lb[lane][lane]) : 1.0;
            for (int c = nb - 1; c >= 0; c--) {
                const E xc = rscale(__shfl_sync(0xffffffffu, dinv, c), lane_bcast(v, c));
                if (lane == c) v = xc;
                if (lane < c) v -= mulc(xc, Lb[c][lane]);   // conj(L[c][lane]) x_c
            }
            yb[lane] = lane < nb ? v : make_zero<E>();
        }
        __syncthreads();
        if (myk < j0) {
            E a0 = make_zero<E>(), a1 = make_zero<E>();
#pragma unroll
            for (int r = 0; r < NB; r += 2) { a0 += mulc(yb[r], lseg[r]); a1 += mulc(yb[r + 1], lseg[r + 1]); }
            myy -= a0 + a1;
        } else if (myk < j0 + nb) {
            myy = yb[myk - j0];                                      // x of the block: final
        }
        __syncthreads();                                             // yb / Lb are reused by the next block
    }
    if (myk < P) x[myk] = myy;
    tick(5);
    if (prof && blockIdx.x == 0 && tid == 0) for (int i = 0; i < 6; i++) prof[i] = t_acc[i];
}

template <typename E>
int cholesky_solve_fused(nq_ctx_t ctx, E* A, int64_t P, E* x, int* dinfo, bool* used) {
    *used = false;
    int coop = 0;
    NQ_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    if (!coop) return NQ_OK;
    auto kern = chol_fused_kernel<E>;
    int per_sm = 0;
    NQ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
    if (per_sm < 1) return NQ_OK;
    // enough warps for the first trailing update, never more CTAs than are co-resident
    const int64_t n32 = (P + 31) / 32;
    int grid = (int)std::min<int64_t>((int64_t)ctx->num_sms, std::max<int64_t>(1, (n32 * (n32 + 1) / 2 + 7) / 8));
    unsigned* bar = (unsigned*)((char*)dinfo + 8);
    NQ_CUDA(ctx, cudaMemsetAsync(bar, 0, 8, ctx->stream));
    static const bool want_prof = getenv("NQ_CHOL_PROF") != nullptr;
    unsigned long long* prof = want_prof ? (unsigned long long*)((char*)dinfo + 16) : nullptr;
    void* args[] = {(void*)&A, (void*)&P, (void*)&x, (void*)&dinfo, (void*)&bar, (void*)&prof};
    NQ_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3(256), args, 0, ctx->stream));
    ctx->launches++;
    if (want_prof) {
        unsigned long long h[6];
        NQ_CUDA(ctx, cudaMemcpyAsync(h, prof, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        fprintf(stderr, "chol_fused P=%lld grid=%d us: potf2 %.1f trsm %.1f bar1 %.1f update %.1f bar2 %.1f backward %.1f\n",
                (long long)P, grid, h[0] * 1e-3, h[1] * 1e-3, h[2] * 1e-3, h[3] * 1e-3, h[4] * 1e-3, h[5] * 1e-3);
    }
    *used = true;
    return NQ_OK;
}

template <typename E>
int cholesky_solve(nq_ctx_t ctx, E* A, int64_t P, E* x, int* dinfo) {
    static const bool legacy = [] { const char* e = getenv("NQ_CHOL"); return e && !strcmp(e, "legacy"); }();
    if (!legacy) {
        bool used = false;
        NQ_CHECK(cholesky_solve_fused<E>(ctx, A, P, x, dinfo, &used));
        if (used) return NQ_OK;
    }
    for (int64_t j0 = 0; j0 < P; j0 += NB) {
        int nb = (int)std::min<int64_t>(NB, P - j0);
        int64_t below = P - j0 - nb;
        unsigned gp = (unsigned)std::max<int64_t>(1, (below + 127) / 128);
        NQ_LAUNCH(ctx, chol_panel_kernel<E>, gp, 128, 0, A, P, j0, nb, dinfo);
        if (below > 0) {
            int64_t nt = (below + 63) / 64;
            const size_t usmem = (size_t)2 * (sizeof(E) / 8) * NB * LDU * sizeof(double);
            auto ku = chol_update_dmma_kernel<E>;
            if (j0 == 0) NQ_CUDA(ctx, cudaFuncSetAttribute(ku, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usmem));
            NQ_LAUNCH(ctx, ku, (unsigned)(nt * (nt + 1) / 2), 256, usmem, A, P, j0, nb);
        }
    }
    static const bool single_cta = [] { const char* e = getenv("NQ_CHOL_TRSV"); return e && !strcmp(e, "single"); }();
    // one CTA streams L twice (time ~ P^2, latency-bound per block); per-block launches cost ~7 us each (time ~ P):
    // measured crossover near P = 2500 doubles (cfg4, P = 2176: 0.6 ms single vs 1.0 ms; cfg3, P = 5364 complex: 20.8 vs 6 ms)
    if (single_cta || P * (int64_t)(sizeof(E) / 8) <= 3000) {
        NQ_LAUNCH(ctx, chol_solve_kernel<E>, 1, 512, 0, (const E*)A, P, x);
        return NQ_OK;
    }
    E* W = (E*)nq_scratch(ctx, SL_W2, (size_t)2 * P * sizeof(E));
    if (!W) return NQ_ERR_ALLOC;
    E* y = W;
    E* xo = W + P;
    for (int64_t j0 = 0; j0 < P; j0 += NB) {
        int nb = (int)std::min<int64_t>(NB, P - j0);
        unsigned g = (unsigned)std::max<int64_t>(1, (P - j0 - nb + 255) / 256);
        NQ_LAUNCH(ctx, (chol_trsv_step_kernel<E, true>), g, 256, 0, (const E*)A, P, j0, nb, x, y);
    }
    for (int64_t j0 = ((P - 1) / NB) * NB; j0 >= 0; j0 -= NB) {
        int nb = (int)std::min<int64_t>(NB, P - j0);
        unsigned g = (unsigned)std::max<int64_t>(1, (j0 + 255) / 256);
        NQ_LAUNCH(ctx, (chol_trsv_step_kernel<E, false>), g, 256, 0, (const E*)A, P, j0, nb, y, xo);
    }
    NQ_CUDA(ctx, cudaMemcpyAsync(x, xo, (size_t)P * sizeof(E), cudaMemcpyDeviceToDevice, ctx->stream));
    return NQ_OK;
}

// ======================================================================================
// CG on an explicit S (+ eps I) or matrix-free; vectors are E = double | cxd
// ======================================================================================
template <typename E>
__global__ void gemv_partial_kernel(const E* __restrict__ S, int64_t P, const E* __restrict__ v, E* __restrict__ partial) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    int64_t per = (P + gridDim.y - 1) / gridDim.y;
    int64_t j0 = blockIdx.y * per, j1 = j0 + per < P ? j0 + per : P;
    E acc = make_zero<E>();
    for (int64_t j = j0; j < j1; j++) acc += S[i + P * j] * v[j];
    partial[blockIdx.y * P + i] = acc;
}
template <typename E>
__global__ void gemv_final_kernel(const E* __restrict__ partial, int nslice, int64_t P, const E* __restrict__ v, double eps,
                                  E* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    E acc = rscale(eps, v[i]);
    for (int s = 0; s < nslice; s++) acc += partial[s * P + i];
    out[i] = acc;
}
// out[0] = <a, b> = sum conj(a_i) b_i (single block, fixed order), as (re, im)
template <typename E>
__global__ void dot_kernel(const E* __restrict__ a, const E* __restrict__ b, int64_t n, double* __restrict__ out) {
    __shared__ double red[64];
    double ar = 0.0, ai = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        cxd p = mulc(to_cx(b[i]), to_cx(a[i]));
        ar += p.re; ai += p.im;
    }
    ar = warp_sum(ar); ai = warp_sum(ai);
    if ((threadIdx.x & 31) == 0) { red[2 * (threadIdx.x >> 5)] = ar; red[2 * (threadIdx.x >> 5) + 1] = ai; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sr = 0.0, si = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { sr += red[2 * w]; si += red[2 * w + 1]; }
        out[0] = sr; out[1] = si;
    }
}
// u = r + beta u
template <typename E>
__global__ void cg_update_u_kernel(E* __restrict__ u, const E* __restrict__ r, double beta, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) u[i] = r[i] + rscale(beta, u[i]);
}
// x += alpha u ; r -= alpha c    (alpha = res^2 / <u,c>, complex in general)
template <typename E>
__global__ void cg_update_xr_kernel(E* __restrict__ x, E* __restrict__ r, const E* __restrict__ u, const E* __restrict__ c,
                                    double res2, const double* __restrict__ uc, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dr = uc[0], di = uc[1], den = dr * dr + di * di;
    cxd alpha(res2 * dr / den, -res2 * di / den);
    cxd au = alpha * to_cx(u[i]), ac = alpha * to_cx(c[i]);
    x[i] += mk<E>(au.re, au.im);
    r[i] -= mk<E>(ac.re, ac.im);
}

// matrix-free pieces: t[s] = sum_k O[k,s] v_k   (one warp per sample)
template <typename EO, typename EV>
__global__ void ot_v_kernel(const EO* __restrict__ O, int64_t ld, int64_t P, int64_t Ns, const EV* __restrict__ v,
                            cx<typename elem_traits<EO>::real>* __restrict__ t) {
    typedef typename elem_traits<EO>::real T;
    int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (s >= Ns) return;
    double ar = 0.0, ai = 0.0;
    for (int64_t k = lane; k < P; k += 32) {
        cxd o = cxd((double)real_part(O[k + ld * s]), (double)imag_part(O[k + ld * s]));
        cxd p = o * to_cx(v[k]);
        ar += p.re; ai += p.im;
    }
    ar = warp_sum(ar); ai = warp_sum(ai);
    if (lane == 0) t[s] = cx<T>((T)ar, (T)ai);
}
// out = eps v + (real ? Re(y) : y)
template <typename E>
__global__ void matfree_finish_kernel(const cxd* __restrict__ y, const E* __restrict__ v, double eps, int64_t P, E* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < P) out[i] = rscale(eps, v[i]) + mk<E>(y[i].re, y[i].im);
}

template <typename E>
struct CgOps {
    nq_ctx_t ctx;
    int64_t P;
    // explicit
    const E* S = nullptr;
    double eps = 0.0;
    // matrix-free
    const void* O = nullptr; int64_t ld = 0, Ns = 0, Ns_total = 0; nq_dtype odtype = NQ_C128;
    int matvec(const E* v, E* out) {
        if (S) {
            int nslice = (int)std::max<int64_t>(1, std::min<int64_t>(64, P / 64));
            E* partial = (E*)nq_scratch(ctx, SL_W4, (size_t)nslice * P * sizeof(E));
            if (!partial) return NQ_ERR_ALLOC;
            dim3 grid((unsigned)((P + 127) / 128), (unsigned)nslice);
            NQ_LAUNCH(ctx, gemv_partial_kernel<E>, grid, 128, 0, S, P, v, partial);
            NQ_LAUNCH(ctx, gemv_final_kernel<E>, (unsigned)((P + 255) / 256), 256, 0, (const E*)partial, nslice, P, v, eps, out);
            return NQ_OK;
        }
        // t = O^T v (complex, precision of O), y = sum_s conj(O[:,s]) t[s] / Ns_total, out = eps v + y
        size_t tb = (size_t)Ns * nq_dtype_size(nq_complex_of(odtype));
        void* t = nq_scratch(ctx, SL_W3, tb ? tb : 16);
        cxd* y = (cxd*)nq_scratch(ctx, SL_W4, (size_t)P * sizeof(cxd));
        if (!t || !y) return NQ_ERR_ALLOC;
        unsigned g = (unsigned)((Ns * 32 + 255) / 256);
        switch (odtype) {
            case NQ_F32: NQ_LAUNCH(ctx, (ot_v_kernel<float, E>), g, 256, 0, (const float*)O, ld, P, Ns, v, (cxf*)t); break;
            case NQ_F64: NQ_LAUNCH(ctx, (ot_v_kernel<double, E>), g, 256, 0, (const double*)O, ld, P, Ns, v, (cxd*)t); break;
            case NQ_C64: NQ_LAUNCH(ctx, (ot_v_kernel<cxf, E>), g, 256, 0, (const cxf*)O, ld, P, Ns, v, (cxf*)t); break;
            default: NQ_LAUNCH(ctx, (ot_v_kernel<cxd, E>), g, 256, 0, (const cxd*)O, ld, P, Ns, v, (cxd*)t); break;
        }
        NQ_CHECK(colsum_dispatch<true>(ctx, odtype, O, ld, P, Ns, t, 1.0 / (double)Ns_total, y));
        if (ctx->nccl_comm) NQ_CHECK(nq_allreduce_device(ctx, y, P, NQ_C128, false));
        NQ_LAUNCH(ctx, matfree_finish_kernel<E>, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)y, v, eps, P, out);
        return NQ_OK;
    }
};

// IterativeSolvers 0.8.1 cg: x0 = 0, u = 0, r = b; converged when ||r|| <= tol ||b||
template <typename E>
int cg_solve(CgOps<E>& ops, const E* b, double tol, int64_t maxiter, E* x, int64_t* iters) {
    nq_ctx_t ctx = ops.ctx;
    const int64_t P = ops.P;
    E* work = (E*)nq_scratch(ctx, SL_W2, (size_t)3 * P * sizeof(E) + 64);
    if (!work) return NQ_ERR_ALLOC;
    E *u = work, *r = work + P, *c = work + 2 * P;
    double* dsc = (double*)nq_scratch(ctx, SL_W1, 64);
    if (!dsc) return NQ_ERR_ALLOC;
    unsigned gv = (unsigned)((P + 255) / 256);
    NQ_CUDA(ctx, cudaMemsetAsync(x, 0, (size_t)P * sizeof(E), ctx->stream));
    NQ_CUDA(ctx, cudaMemsetAsync(u, 0, (size_t)P * sizeof(E), ctx->stream));
    NQ_CUDA(ctx, cudaMemcpyAsync(r, b, (size_t)P * sizeof(E), cudaMemcpyDeviceToDevice, ctx->stream));
    double h[2];
    NQ_LAUNCH(ctx, dot_kernel<E>, 1, 1024, 0, (const E*)r, (const E*)r, P, dsc);
    NQ_CUDA(ctx, cudaMemcpyAsync(h, dsc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double residual = sqrt(h[0]), prev = 1.0;
    const double reltol = residual * tol;
    int64_t it = 0;
    while (it < maxiter && !(residual <= reltol)) {
        double beta = residual * residual / (prev * prev);
        NQ_LAUNCH(ctx, cg_update_u_kernel<E>, gv, 256, 0, u, (const E*)r, beta, P);
        NQ_CHECK(ops.matvec(u, c));
        NQ_LAUNCH(ctx, dot_kernel<E>, 1, 1024, 0, (const E*)u, (const E*)c, P, dsc);
        NQ_LAUNCH(ctx, cg_update_xr_kernel<E>, gv, 256, 0, x, r, (const E*)u, (const E*)c, residual * residual, (const double*)dsc, P);
        NQ_LAUNCH(ctx, dot_kernel<E>, 1, 1024, 0, (const E*)r, (const E*)r, P, dsc + 2);
        NQ_CUDA(ctx, cudaMemcpyAsync(h, dsc + 2, 16, cudaMemcpyDeviceToHost, ctx->stream));
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        prev = residual;
        residual = sqrt(h[0]);
        it++;
    }
    if (iters) *iters = it;
    return residual <= reltol ? NQ_OK : NQ_ERR_NOT_CONVERGED;
}

// ---- MINRES (Paige & Saunders) on the same operator: Hermitian A = S + eps I (explicit or matrix-free), x0 = 0,
// no preconditioner, stop when the recurrence residual ||r_k|| = phibar <= tol ||b||.  ref: the sr_minres branch of
// SRIterative.jl:101-125 (IterativeSolvers 0.8.1 minres); the reference pins neither iterates nor iteration counts,
// so parity is on the converged solution.
template <typename E> __global__ void scale_copy_kernel(E* __restrict__ out, const E* __restrict__ in, double s, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = rscale(s, in[i]);
}
template <typename E> __global__ void axpy_real_kernel(E* __restrict__ y, double a, const E* __restrict__ x, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] = y[i] + rscale(a, x[i]);
}
// wn = (v - oldeps w1 - delta w2) / gamma written over w1;  x += phi wn
template <typename E>
__global__ void minres_w_kernel(E* __restrict__ w1, const E* __restrict__ w2, const E* __restrict__ v, E* __restrict__ x,
                                double oldeps, double delta, double denom, double phi, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    E t = rscale(denom, v[i] - rscale(oldeps, w1[i]) - rscale(delta, w2[i]));
    w1[i] = t;
    x[i] = x[i] + rscale(phi, t);
}

template <typename E>
int minres_solve(CgOps<E>& ops, const E* b, double tol, int64_t maxiter, E* x, int64_t* iters) {
    nq_ctx_t ctx = ops.ctx;
    const int64_t P = ops.P;
    E* work = (E*)nq_scratch(ctx, SL_W2, (size_t)6 * P * sizeof(E) + 64);
    double* dsc = (double*)nq_scratch(ctx, SL_W1, 64);
    if (!work || !dsc) return NQ_ERR_ALLOC;
    E *r1 = work, *r2 = work + P, *y = work + 2 * P, *v = work + 3 * P, *wa = work + 4 * P, *wb = work + 5 * P;
    const unsigned gv = (unsigned)((P + 255) / 256);
    NQ_CUDA(ctx, cudaMemsetAsync(x, 0, (size_t)P * sizeof(E), ctx->stream));
    NQ_CUDA(ctx, cudaMemsetAsync(wa, 0, (size_t)2 * P * sizeof(E), ctx->stream));
    NQ_CUDA(ctx, cudaMemcpyAsync(r2, b, (size_t)P * sizeof(E), cudaMemcpyDeviceToDevice, ctx->stream));
    NQ_CUDA(ctx, cudaMemsetAsync(r1, 0, (size_t)P * sizeof(E), ctx->stream));
    double h[2];
    auto dot = [&](const E* a_, const E* b_, double* out) -> int {
        NQ_LAUNCH(ctx, dot_kernel<E>, 1, 1024, 0, a_, b_, P, dsc);
        NQ_CUDA(ctx, cudaMemcpyAsync(h, dsc, 16, cudaMemcpyDeviceToHost, ctx->stream));
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *out = h[0];
        return NQ_OK;
    };
    double bb;
    NQ_CHECK(dot(r2, r2, &bb));
    const double beta1 = sqrt(bb);
    int64_t it = 0;
    if (beta1 == 0.0) { if (iters) *iters = 0; return NQ_OK; }
    double oldb = 0.0, beta = beta1, dbar = 0.0, epsln = 0.0, phibar = beta1, cs = -1.0, sn = 0.0;
    while (it < maxiter && !(phibar <= tol * beta1)) {
        it++;
        NQ_LAUNCH(ctx, scale_copy_kernel<E>, gv, 256, 0, v, (const E*)r2, 1.0 / beta, P);          // v = r2 / beta
        NQ_CHECK(ops.matvec(v, y));                                                                // y = A v
        if (it >= 2) NQ_LAUNCH(ctx, axpy_real_kernel<E>, gv, 256, 0, y, -beta / oldb, (const E*)r1, P);
        double alfa;
        NQ_CHECK(dot(v, y, &alfa));                                                                // Re <v, y>
        NQ_LAUNCH(ctx, axpy_real_kernel<E>, gv, 256, 0, y, -alfa / beta, (const E*)r2, P);
        { E* t = r1; r1 = r2; r2 = y; y = t; }                                                     // r1 <- r2, r2 <- y
        oldb = beta;
        double b2;
        NQ_CHECK(dot(r2, r2, &b2));
        beta = sqrt(b2);
        const double oldeps = epsln, delta = cs * dbar + sn * alfa, gbar = sn * dbar - cs * alfa;
        epsln = sn * beta; dbar = -cs * beta;
        double gamma = sqrt(gbar * gbar + beta * beta);
        if (gamma < 1e-300) gamma = 1e-300;
        cs = gbar / gamma; sn = beta / gamma;
        const double phi = cs * phibar;
        phibar = sn * phibar;
        NQ_LAUNCH(ctx, minres_w_kernel<E>, gv, 256, 0, wa, (const E*)wb, (const E*)v, x, oldeps, delta, 1.0 / gamma, phi, P);
        { E* t = wa; wa = wb; wb = t; }                                                            // (w_{k-1}, w_k)
        if (beta == 0.0) break;                                                                    // exact solution reached
    }
    if (iters) *iters = it;
    return (phibar <= tol * beta1 || beta == 0.0) ? NQ_OK : NQ_ERR_NOT_CONVERGED;
}

// --------------------------------------------------------------------------------------
// MINRES-QLP as the reference runs it (External/IterativeSolvers/minresqlp.jl; restated with its quirks Q19-Q22 in
// oracle/minresqlp.py): Lanczos + left Givens rotations like MINRES, plus right rotations that keep the solution the
// minimum-length one on singular systems.  The reference's MINRES/QLP switch never takes the MINRES branch (Q19), so
// every iteration is a QLP update.  Scalars live on the host (two dot products per iteration come back), vectors on
// the device; x0 != nullptr: warm start (iterate on F - A x0, add x0 at the end) for the restart ladder.
// flag: 1 relres <= tol, 2 relAres <= tol, 3/4 machine-precision variants, 5 eigenvector, 6 xnorm limit, 7 Acond limit,
// 8 iteration limit, 9 singular last rotation.
// --------------------------------------------------------------------------------------
template <typename E>
__global__ void qlp_update_kernel(int mode, const E* __restrict__ v, E* __restrict__ w, E* __restrict__ wl, E* __restrict__ wl2,
                                  E* __restrict__ xl2, E* __restrict__ x, double cr1, double sr1, double cr2, double sr2,
                                  double ul2, double ul, double u, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const E vi = v[i], wi = w[i], wli = wl[i];
    E nwl2, nwl, nw;
    if (mode == 1) { nwl2 = wli; nwl = rscale(sr1, vi); nw = rscale(-cr1, vi); }
    else if (mode == 2) { nwl2 = wli; nwl = rscale(cr1, wi) + rscale(sr1, vi); nw = rscale(sr1, wi) - rscale(cr1, vi); }
    else {
        const E t2 = wli, t1 = wi;
        const E wp = rscale(sr2, t2) - rscale(cr2, vi);
        nwl2 = rscale(cr2, t2) + rscale(sr2, vi);
        nwl = rscale(cr1, t1) + rscale(sr1, wp);
        nw = rscale(sr1, t1) - rscale(cr1, wp);
    }
    const E nxl2 = xl2[i] + rscale(ul2, nwl2);
    wl2[i] = nwl2; wl[i] = nwl; w[i] = nw; xl2[i] = nxl2;
    x[i] = nxl2 + rscale(ul, nwl) + rscale(u, nw);
}

static void sym_givens(double a, double b, double& c, double& s, double& r) {
    auto sgn = [](double t) { return (double)((t > 0.0) - (t < 0.0)); };
    if (b == 0.0) { c = a == 0.0 ? 1.0 : sgn(a); s = 0.0; r = fabs(a); }
    else if (a == 0.0) { c = 0.0; s = sgn(b); r = fabs(b); }
    else if (fabs(b) > fabs(a)) { const double t = a / b; s = sgn(b) / sqrt(1.0 + t * t); c = s * t; r = b / s; }
    else { const double t = b / a; c = sgn(a) / sqrt(1.0 + t * t); s = c * t; r = a / c; }
}

template <typename E>
int qlp_solve(CgOps<E>& ops, const E* b, const E* x0, double tol, int64_t maxiter, E* x, int64_t* iters, int* flag_out) {
    nq_ctx_t ctx = ops.ctx;
    const int64_t P = ops.P;
    E* work = (E*)nq_scratch(ctx, SL_W2, (size_t)9 * P * sizeof(E) + 64);
    double* dsc = (double*)nq_scratch(ctx, SL_W1, 128);
    if (!work || !dsc) return NQ_ERR_ALLOC;
    E *ra = work, *rb = work + P, *rc = work + 2 * P, *v = work + 3 * P, *w = work + 4 * P, *wl = work + 5 * P,
      *wl2 = work + 6 * P, *xl2 = work + 7 * P, *xs = work + 8 * P;
    const unsigned gv = (unsigned)((P + 255) / 256);
    NQ_CUDA(ctx, cudaMemsetAsync(w, 0, (size_t)4 * P * sizeof(E), ctx->stream));          // w, wl, wl2, xl2
    NQ_CUDA(ctx, cudaMemsetAsync(ra, 0, (size_t)P * sizeof(E), ctx->stream));             // r1 = 0
    NQ_CUDA(ctx, cudaMemcpyAsync(rb, b, (size_t)P * sizeof(E), cudaMemcpyDeviceToDevice, ctx->stream));   // r2 = b
    if (x0) {                                                                                               // r2 = b - A x0
        NQ_CUDA(ctx, cudaMemcpyAsync(xs, x0, (size_t)P * sizeof(E), cudaMemcpyDeviceToDevice, ctx->stream));
        NQ_CHECK(ops.matvec(xs, rc));
        NQ_LAUNCH(ctx, axpy_real_kernel<E>, gv, 256, 0, rb, -1.0, (const E*)rc, P);
    }
    NQ_CUDA(ctx, cudaMemsetAsync(x, 0, (size_t)P * sizeof(E), ctx->stream));
    double h[2];
    auto dot = [&](const E* a_, const E* b_, double* out) -> int {
        NQ_LAUNCH(ctx, dot_kernel<E>, 1, 1024, 0, a_, b_, P, dsc);
        NQ_CUDA(ctx, cudaMemcpyAsync(h, dsc, 16, cudaMemcpyDeviceToHost, ctx->stream));
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *out = h[0];
        return NQ_OK;
    };
    double bb;
    NQ_CHECK(dot(rb, rb, &bb));
    const double beta1 = sqrt(bb);
    int flag0 = -2, flag = flag0;
    int64_t it = 0, qlp_it = 0;
    if (beta1 == 0.0) {
        if (x0) NQ_LAUNCH(ctx, axpy_real_kernel<E>, gv, 256, 0, x, 1.0, (const E*)xs, P);
        if (iters) *iters = 0;
        if (flag_out) *flag_out = 0;
        return NQ_OK;
    }
    const double TranCond = 10e6, maxxnorm = 10e6, Acondlim = TranCond;      // minresqlp.jl:98-100, :150 (Q21)
    const double realmin = 2.2250738585072014e-308, dbl_eps = 2.220446049250313e-16;
    E *r1 = ra, *r2 = rb, *r3 = rc;                 // r3 is the free buffer; the reference keeps r3 == r2 between iterations
    double beta = 0.0, tau = 0.0, taul = 0.0, phi = beta1, betan = beta1, gmin = 0.0;
    double cs = -1.0, sn = 0.0, cr1 = -1.0, sr1 = 0.0, cr2 = -1.0, sr2 = 0.0;
    double dltan = 0.0, eplnn = 0.0, gama = 0.0, gamal = 0.0, gamal2 = 0.0;
    double eta = 0.0, etal = 0.0, etal2 = 0.0, vepln = 0.0, veplnl = 0.0, veplnl2 = 0.0;
    double ul3 = 0.0, ul2 = 0.0, ul = 0.0, u = 0.0;
    double rnorm = beta1, xnorm = 0.0, xl2norm = 0.0, Anorm = 0.0, Acond = 1.0;
    double relres = beta1 / (beta1 + 1e-50), relAres = 0.0, resnorm = 123.0;
    int64_t iteration = 1;
    while (!(iteration > maxiter || resnorm <= tol || flag != flag0)) {
        iteration++;
        it++;
        const double betal = beta;
        beta = betan;
        NQ_LAUNCH(ctx, scale_copy_kernel<E>, gv, 256, 0, v, (const E*)r2, 1.0 / beta, P);          // v = r3 / beta
        NQ_CHECK(ops.matvec(v, r3));                                                               // r3 = A v (shift inside)
        if (it > 1) NQ_LAUNCH(ctx, axpy_real_kernel<E>, gv, 256, 0, r3, -beta / betal, (const E*)r1, P);
        double alfa;
        NQ_CHECK(dot(v, r3, &alfa));                                                               // Re <v, r3>
        NQ_LAUNCH(ctx, axpy_real_kernel<E>, gv, 256, 0, r3, -alfa / beta, (const E*)r2, P);
        { E* t = r1; r1 = r2; r2 = r3; r3 = t; }
        double b2;
        NQ_CHECK(dot(r2, r2, &b2));
        betan = sqrt(b2);
        if (it == 1 && betan == 0.0) {
            if (alfa == 0.0) flag = 0;
            else { flag = -1; NQ_LAUNCH(ctx, scale_copy_kernel<E>, gv, 256, 0, x, b, 1.0 / alfa, P); }
            break;
        }
        const double pnorm = sqrt(betal * betal + alfa * alfa + betan * betan);
        const double dbar = dltan;
        double dlta = cs * dbar + sn * alfa;
        const double gbar = sn * dbar - cs * alfa;
        eplnn = sn * betan;
        dltan = -cs * betan;
        gamal2 = gamal; gamal = gama;
        sym_givens(gbar, betan, cs, sn, gama);
        const double taul2 = taul;
        taul = tau;
        tau = cs * phi;
        phi = sn * phi;
        if (it > 2) {
            veplnl2 = veplnl; etal2 = etal; etal = eta;
            const double dlta_tmp = sr2 * vepln - cr2 * dlta;
            veplnl = cr2 * vepln + sr2 * dlta;
            dlta = dlta_tmp;
            eta = sr2 * gama;
            gama = -cr2 * gama;
        }
        if (it > 1) {
            sym_givens(gamal, dlta, cr1, sr1, gamal);
            vepln = sr1 * gama;
            gama = -cr1 * gama;
        }
        const double ul4 = ul3;
        ul3 = ul2;
        if (it > 2) ul2 = (taul2 - etal2 * ul4 - veplnl2 * ul3) / gamal2;
        if (it > 1) ul = (taul - etal * ul3 - veplnl * ul2) / gamal;
        const double xnorm_tmp = sqrt(xl2norm * xl2norm + ul2 * ul2 + ul * ul);
        if (fabs(gama) > realmin && xnorm_tmp < maxxnorm) {
            u = (tau - eta * ul2 - vepln * ul) / gama;
            if (sqrt(xnorm_tmp * xnorm_tmp + u * u) > maxxnorm) { u = 0.0; flag = 6; }
        } else { u = 0.0; flag = 9; }
        xl2norm = sqrt(xl2norm * xl2norm + ul2 * ul2);
        xnorm = sqrt(xl2norm * xl2norm + ul * ul + u * u);
        qlp_it++;
        NQ_LAUNCH(ctx, qlp_update_kernel<E>, gv, 256, 0, it == 1 ? 1 : (it == 2 ? 2 : 3), (const E*)v, w, wl, wl2, xl2, x,
                  cr1, sr1, cr2, sr2, ul2, ul, u, P);
        sym_givens(gamal, eplnn, cr2, sr2, gamal);
        const double abs_gama = fabs(gama);
        Anorm = std::max(std::max(Anorm, pnorm), std::max(gamal, abs_gama));
        if (it == 1) gmin = gama;                                                                  // never updated again (Q20)
        const double Acondl = Acond, rnorml = rnorm, relresl = relres;
        Acond = Anorm / gmin;
        if (flag != 9) rnorm = phi;
        relres = rnorm / (Anorm * xnorm + beta1);
        const double rootl = sqrt(gbar * gbar + dltan * dltan);
        relAres = rootl / Anorm;
        const double epsx = Anorm * xnorm * dbl_eps;
        if (flag == flag0 || flag == 9) {
            const double t1 = 1.0 + relres, t2 = 1.0 + relAres;
            if (it >= maxiter) flag = 8;
            if (Acond >= Acondlim) flag = 7;
            if (xnorm >= maxxnorm) flag = 6;
            if (epsx >= beta1) flag = 5;
            if (t2 <= 1.0) flag = 4;
            if (t1 <= 1.0) flag = 3;
            if (relAres <= tol) flag = 2;
            if (relres <= tol) flag = 1;
        }
        if (flag == 2 || flag == 4 || flag == 6 || flag == 7) { it--; Acond = Acondl; rnorm = rnorml; relres = relresl; }
        resnorm = std::min(relres, relAres);
    }
    if (x0) NQ_LAUNCH(ctx, axpy_real_kernel<E>, gv, 256, 0, x, 1.0, (const E*)xs, P);
    if (iters) *iters = it;
    if (flag_out) *flag_out = flag;
    ctx->info = flag;
    (void)qlp_it;
    // the reference counts every exit as converged (Q22); the iteration limit is reported so that callers can tell
    return flag == 8 ? NQ_ERR_NOT_CONVERGED : NQ_OK;
}

template <typename E>
__global__ void update_kernel(E* __restrict__ w, const E* __restrict__ dw, typename elem_traits<E>::real eta, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) w[i] = w[i] - rscale(eta, dw[i]);
}

// per-chain mean and centred second moment of vals[B, L] (column-major: chain fastest)
template <typename E>
__global__ void chain_stats_kernel(const E* __restrict__ vals, int64_t B, int64_t L, double* __restrict__ out /* [B][3] */) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double mr = 0.0, mi = 0.0;
    for (int64_t l = 0; l < L; l++) { mr += (double)real_part(vals[b + B * l]); mi += (double)imag_part(vals[b + B * l]); }
    mr /= (double)L; mi /= (double)L;
    double m2 = 0.0;
    for (int64_t l = 0; l < L; l++) {
        double dr = (double)real_part(vals[b + B * l]) - mr, di = (double)imag_part(vals[b + B * l]) - mi;
        m2 += dr * dr + di * di;
    }
    out[3 * b] = mr; out[3 * b + 1] = mi; out[3 * b + 2] = m2;
}

template <typename T>
__global__ void abs2_kernel(const cx<T>* __restrict__ in, T* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i].re * in[i].re + in[i].im * in[i].im;
}

}  // namespace

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" int nq_abs2(nq_ctx_t ctx, const void* vals, int64_t n, nq_dtype vdtype, void* out) {
    if (!ctx || !vals || !out || n < 0) return NQ_ERR_ARG;
    if (!nq_dtype_is_complex(vdtype)) return nq_fail(ctx, NQ_ERR_ARG, "nq_abs2 takes complex values");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const void* dv = st.in(SL_IN0, vals, (size_t)n * nq_dtype_size(vdtype));
    void* dout = st.out(SL_OUT0, out, (size_t)n * nq_dtype_size(nq_real_of(vdtype)));
    if (st.status != NQ_OK) return st.status;
    if (n > 0) {
        unsigned g = (unsigned)((n + 255) / 256);
        if (vdtype == NQ_C64) NQ_LAUNCH(ctx, abs2_kernel<float>, g, 256, 0, (const cxf*)dv, (float*)dout, n);
        else NQ_LAUNCH(ctx, abs2_kernel<double>, g, 256, 0, (const cxd*)dv, (double*)dout, n);
    }
    return st.finish();
}

// The SR entry points take the [P, Ns] matrices where the caller keeps them.  Device-resident (the iteration driver):
// used in place.  Host-resident (a Julia Array / C buffer): copied to a device scratch slot first.
static const void* sr_matrix_on_device(nq_ctx_t ctx, int slot, const void* X, int64_t ld, int64_t Ns, nq_dtype dtype, int* status) {
    *status = NQ_OK;
    if (nq_is_device_ptr(X)) return X;
    const size_t bytes = (size_t)ld * Ns * nq_dtype_size(dtype);
    void* d = nq_scratch(ctx, slot, bytes);
    if (!d) { *status = NQ_ERR_ALLOC; return nullptr; }
    cudaError_t e = cudaMemcpyAsync(d, X, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { *status = nq_fail(ctx, NQ_ERR_CUDA, "H2D copy of O: %s", cudaGetErrorString(e)); return nullptr; }
    return d;
}

static int launch_subtract(nq_ctx_t ctx, void* O, int64_t ldO, int64_t P, int64_t Ns, nq_dtype dtype, const cxd* a, unsigned long long* mx) {
    dim3 grid((unsigned)((P + 127) / 128), (unsigned)std::min<int64_t>(Ns, 4096));
    switch (dtype) {
        case NQ_F32: NQ_LAUNCH(ctx, subtract_avg_kernel<float>, grid, 128, 0, (float*)O, ldO, P, Ns, a, mx); break;
        case NQ_F64: NQ_LAUNCH(ctx, subtract_avg_kernel<double>, grid, 128, 0, (double*)O, ldO, P, Ns, a, mx); break;
        case NQ_C64: NQ_LAUNCH(ctx, subtract_avg_kernel<cxf>, grid, 128, 0, (cxf*)O, ldO, P, Ns, a, mx); break;
        default: NQ_LAUNCH(ctx, subtract_avg_kernel<cxd>, grid, 128, 0, (cxd*)O, ldO, P, Ns, a, mx); break;
    }
    return NQ_OK;
}

// lazy != NULL: the subtraction may be deferred (*lazy = 1): the means are returned, O is left as it is, and the context keeps
// (maxima of the uncentred rows, means) for the next nq_sr_setup on this matrix; nq_center_finish applies the subtraction
static int center_impl(nq_ctx_t ctx, void* O_user, int64_t ldO, int64_t P, int64_t Ns, nq_dtype dtype, void* avg, int* lazy) {
    if (!ctx || !O_user || !avg || P <= 0 || Ns <= 0 || ldO < P) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (lazy) *lazy = 0;
    int hs = NQ_OK;
    void* O = const_cast<void*>(sr_matrix_on_device(ctx, SL_HOSTO, O_user, ldO, Ns, dtype, &hs));
    if (hs != NQ_OK) return hs;
    const bool o_on_host = O != O_user;
    NqStage st(ctx);
    void* davg = st.out(SL_OUT0, avg, (size_t)P * nq_dtype_size(dtype));
    cxd* a = (cxd*)nq_scratch(ctx, SL_W0, (size_t)P * sizeof(cxd));
    if (!a || st.status != NQ_OK) return NQ_ERR_ALLOC;
    // deferral only pays when the next consumer is the Ozaki S assembly (FP64, device-resident, >= 4096 samples)
    const bool defer = lazy && nq_dtype_is_double(dtype) && !o_on_host && Ns >= 4096;
    unsigned long long* mx_raw = defer ? rowmax_arm(ctx, O, ldO, P, Ns) : nullptr;
    ctx->shift_pending = false;
    NQ_CHECK(colsum_dispatch<false>(ctx, dtype, O, ldO, P, Ns, nullptr, 1.0, a, mx_raw));
    // under sharding: <O> is the global mean (C1, BaseIterativeSampler.jl:23)
    if (ctx->nccl_comm) {
        NQ_CHECK(nq_allreduce_device(ctx, a, P, NQ_C128, false));
    }
    // a currently holds the (global) sum; divide by the global sample count
    {
        int64_t ns_tot = ctx->ns_total > 0 ? ctx->ns_total : Ns * ctx->nranks;
        cxd* tmp = (cxd*)nq_scratch(ctx, SL_W1, (size_t)P * sizeof(cxd));
        if (!tmp) return NQ_ERR_ALLOC;
        NQ_LAUNCH(ctx, colsum_final_kernel<double>, (unsigned)((P + 31) / 32), 256, 0, (const cxd*)a, (const cxd*)nullptr, 1, P, 1.0 / (double)ns_tot, tmp);
        a = tmp;
    }
    bool deferred = false;
    if (defer && mx_raw) {
        if (ctx->shift_cap < P) {
            if (ctx->shift) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->shift); ctx->shift = nullptr; ctx->shift_cap = 0; }
            if (cudaMalloc(&ctx->shift, (size_t)(P + P / 4 + 64) * sizeof(cxd)) == cudaSuccess) ctx->shift_cap = P + P / 4 + 64;
            else cudaGetLastError();
        }
        if (ctx->shift_cap >= P) {
            NQ_CUDA(ctx, cudaMemcpyAsync(ctx->shift, a, (size_t)P * sizeof(cxd), cudaMemcpyDeviceToDevice, ctx->stream));
            ctx->shift_pending = true;
            ctx->shift_ptr = O; ctx->shift_P = P; ctx->shift_Ns = Ns; ctx->shift_ld = ldO;
            deferred = true;
        }
    }
    if (!deferred) {
        unsigned long long* mx = (nq_dtype_is_double(dtype) && !o_on_host) ? rowmax_arm(ctx, O, ldO, P, Ns) : nullptr;
        NQ_CHECK(launch_subtract(ctx, O, ldO, P, Ns, dtype, (const cxd*)a, mx));
    }
    if (lazy) *lazy = deferred ? 1 : 0;
    NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)a, davg, P, (int)dtype, 0);
    if (o_on_host) {      // centred O goes back to the caller's array (the reference centres in place)
        NQ_CUDA(ctx, cudaMemcpyAsync(O_user, O, (size_t)ldO * Ns * nq_dtype_size(dtype), cudaMemcpyDeviceToHost, ctx->stream));
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return st.finish();
}

extern "C" int nq_center(nq_ctx_t ctx, void* O_user, int64_t ldO, int64_t P, int64_t Ns, nq_dtype dtype, void* avg) {
    return center_impl(ctx, O_user, ldO, P, Ns, dtype, avg, nullptr);
}

extern "C" int nq_center_lazy(nq_ctx_t ctx, void* O, int64_t ldO, int64_t P, int64_t Ns, nq_dtype dtype, void* avg, int* deferred) {
    if (!deferred) return NQ_ERR_ARG;
    return center_impl(ctx, O, ldO, P, Ns, dtype, avg, deferred);
}

// the subtraction a deferred nq_center_lazy left out (no-op when nothing is pending for this matrix)
extern "C" int nq_center_finish(nq_ctx_t ctx, void* O, int64_t ldO, int64_t P, int64_t Ns, nq_dtype dtype) {
    if (!ctx || !O || P <= 0 || Ns <= 0 || ldO < P) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->shift_matches(O, ldO, P, Ns)) return NQ_OK;
    ctx->shift_pending = false;
    ctx->rowmax_ptr = nullptr;             // the recorded maxima belong to the uncentred rows
    return launch_subtract(ctx, O, ldO, P, Ns, dtype, (const cxd*)ctx->shift, nullptr);
}

extern "C" int nq_force_ket(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, nq_dtype dtype,
                            const void* Eloc, void* gradC) {
    if (!ctx || !Oc || !Eloc || !gradC || P <= 0 || Ns <= 0 || ldO < P) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    int hs = NQ_OK;
    Oc = sr_matrix_on_device(ctx, SL_HOSTO, Oc, ldO, Ns, dtype, &hs);
    if (hs != NQ_OK) return hs;
    NqStage st(ctx);
    nq_dtype cdt = nq_complex_of(dtype);
    const void* dE = st.in(SL_IN0, Eloc, (size_t)Ns * nq_dtype_size(cdt));
    void* dg = st.out(SL_OUT0, gradC, (size_t)P * nq_dtype_size(cdt));
    cxd* a = (cxd*)nq_scratch(ctx, SL_W0, (size_t)P * sizeof(cxd));
    if (!a || st.status != NQ_OK) return NQ_ERR_ALLOC;
    int64_t ns_tot = ctx->ns_total > 0 ? ctx->ns_total : Ns * ctx->nranks;
    NQ_CHECK(colsum_dispatch<true>(ctx, dtype, Oc, ldO, P, Ns, dE, 1.0 / (double)ns_tot, a));
    if (ctx->nccl_comm) NQ_CHECK(nq_allreduce_device(ctx, a, P, NQ_C128, false));   // C3
    NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)a, dg, P, (int)cdt, 0);
    return st.finish();
}

extern "C" int nq_force_liouvillian(nq_ctx_t ctx, const void* Lloc, const void* gLloc, int64_t ld, int64_t P, int64_t Ns,
                                    nq_dtype dtype, const void* avg, void* gradC, double* cost) {
    if (!ctx || !Lloc || !gLloc || !avg || !gradC || P <= 0 || Ns <= 0 || ld < P) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!nq_dtype_is_complex(dtype)) return nq_fail(ctx, NQ_ERR_ARG, "L_loc / grad L_loc are complex");
    int hs = NQ_OK;
    gLloc = sr_matrix_on_device(ctx, SL_HOSTG, gLloc, ld, Ns, dtype, &hs);
    if (hs != NQ_OK) return hs;
    NqStage st(ctx);
    size_t cs = nq_dtype_size(dtype);
    const void* dL = st.in(SL_IN0, Lloc, (size_t)Ns * cs);
    const void* davg = st.in(SL_IN1, avg, (size_t)P * cs);
    void* dg = st.out(SL_OUT0, gradC, (size_t)P * cs);
    if (st.status != NQ_OK) return st.status;
    // packed reduction buffer: [ sum_s L_s conj(gL_ks) (P complex) | sum_s |L_s|^2 ]  -> one all-reduce
    cxd* a = (cxd*)nq_scratch(ctx, SL_W0, (size_t)(P + 1) * sizeof(cxd));
    cxd* avgd = (cxd*)nq_scratch(ctx, SL_W1, (size_t)P * sizeof(cxd));
    cxd* outd = (cxd*)nq_scratch(ctx, SL_W2, (size_t)P * sizeof(cxd));
    if (!a || !avgd || !outd) return NQ_ERR_ALLOC;
    int64_t ns_tot = ctx->ns_total > 0 ? ctx->ns_total : Ns * ctx->nranks;
    // (1/Ns) sum_s L_s conj(gL_ks), conjugated at the end: F_k = conj(.) - C avg_k
    NQ_CHECK(colsum_dispatch<true>(ctx, dtype, gLloc, ld, P, Ns, dL, 1.0 / (double)ns_tot, a));
    NQ_CUDA(ctx, cudaMemsetAsync(a + P, 0, sizeof(cxd), ctx->stream));
    if (dtype == NQ_C64) NQ_LAUNCH(ctx, abs2_sum_kernel<float>, 1, 1024, 0, (const cxf*)dL, Ns, (double*)(a + P));
    else NQ_LAUNCH(ctx, abs2_sum_kernel<double>, 1, 1024, 0, (const cxd*)dL, Ns, (double*)(a + P));
    if (ctx->nccl_comm) NQ_CHECK(nq_allreduce_device(ctx, a, P + 1, NQ_C128, false));   // C2 + C3 packed
    NQ_LAUNCH(ctx, convert_to_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, davg, avgd, P, (int)dtype);
    // conj of the column sum, then subtract C avg
    cxd* lg = (cxd*)nq_scratch(ctx, SL_W3, (size_t)P * sizeof(cxd));
    if (!lg) return NQ_ERR_ALLOC;
    NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)a, (void*)lg, P, (int)NQ_C128, 1);
    NQ_LAUNCH(ctx, force_liouv_finish_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)lg, (const cxd*)avgd,
              (const double*)(a + P), 1.0 / (double)ns_tot, P, outd);
    NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)outd, dg, P, (int)dtype, 0);
    if (cost) {
        double s = 0.0;
        NQ_CUDA(ctx, cudaMemcpyAsync(&s, a + P, 8, cudaMemcpyDeviceToHost, ctx->stream));
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *cost = s / (double)ns_tot;
    }
    return st.finish();
}

extern "C" int nq_sr_hint_row_planes(nq_ctx_t ctx, const uint8_t* row_planes, int64_t P) {
    if (!ctx || P < 0) return NQ_ERR_ARG;
    ctx->hint_P = 0;
    ctx->hint_tile_flags.clear();
    if (!row_planes || P == 0) return NQ_OK;
    const int64_t ntile = (P + TS - 1) / TS;
    ctx->hint_tile_flags.assign((size_t)ntile, 0u);
    for (int64_t k = 0; k < P; k++) ctx->hint_tile_flags[(size_t)(k / TS)] |= (unsigned)(row_planes[k] & 3);
    ctx->hint_P = P;
    return NQ_OK;
}

extern "C" int nq_sr_setup(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, int64_t Ns_total,
                           nq_dtype dtype, const void* gradC, int real_params, void* S, void* F) {
    if (!ctx || !Oc || !gradC || !S || !F || P <= 0 || Ns <= 0 || ldO < P || Ns_total < Ns) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    int hs = NQ_OK;
    Oc = sr_matrix_on_device(ctx, SL_HOSTO, Oc, ldO, Ns, dtype, &hs);
    if (hs != NQ_OK) return hs;
    const bool ocx = nq_dtype_is_complex(dtype);
    const bool out_complex = ocx && !real_params;
    // S/F dtype: real nets -> real of the same precision; complex nets -> complex
    nq_dtype sdt = out_complex ? dtype : nq_real_of(dtype);
    NqStage st(ctx);
    nq_dtype cdt = nq_complex_of(dtype);
    const void* dgc = st.in(SL_IN0, gradC, (size_t)P * nq_dtype_size(cdt));
    void* dS = st.out(SL_OUT0, S, (size_t)P * P * nq_dtype_size(sdt));
    void* dF = st.out(SL_OUT1, F, (size_t)P * nq_dtype_size(sdt));
    if (st.status != NQ_OK) return st.status;
    const int ntile = (int)((P + TS - 1) / TS);
    const int64_t Ppad = (int64_t)ntile * TS;
    const int64_t ntri = (int64_t)ntile * (ntile + 1) / 2;
    // split K so that the grid is close to a whole number of waves (one 256-thread CTA per SM): with
    // 153 tiles a 4-way split runs 4.13 waves, i.e. a 5th wave that is 87% idle
    const size_t plane = (size_t)Ppad * Ppad * sizeof(double);
    const int64_t nchunk = (Ns + KS - 1) / KS;
    int nsplit = 1;
    {
        double best = 1e30;
        for (int ns = 1; ns <= 64; ns++) {
            if (ns > 1 && nchunk / ns < 16) break;
            if (plane * ns * (out_complex ? 2 : 1) > ((size_t)3 << 30)) break;
            double waves = (double)ntri * ns / ctx->num_sms;
            double cost = std::ceil(waves) / waves + 0.002 * ns;       // tail inefficiency + partial-sum traffic
            if (waves >= 1.0 || ns == 1) { if (cost < best) { best = cost; nsplit = ns; } }
        }
    }
    if (const char* e = getenv("NQ_SYRK_NSPLIT")) { int v = atoi(e); if (v >= 1 && v <= 64 && nchunk / v >= 1) nsplit = v; }   // tuning knob
    const int64_t ldr = ldO * (ocx ? 2 : 1);
    // FP32 mode: tcgen05 3xTF32 path (nq_syrk_tf32.cu) when the rows allow 16-byte loads
    static const bool fp32_dmma = [] { const char* e = getenv("NQ_SR_FP32_PATH"); return e && !strcmp(e, "dmma"); }();
    const bool use_tf32 = !nq_dtype_is_double(dtype) && !fp32_dmma && (ldr % 4 == 0) && ((uintptr_t)Oc % 16 == 0);
    if (use_tf32) {
        NQ_CHECK(nq_syrk_tf32_device(ctx, Oc, ldO, P, Ns, Ns_total, ocx, out_complex, dS));
    } else {
    double* Wre = (double*)nq_scratch(ctx, SL_W0, plane * nsplit);
    double* Wim = out_complex ? (double*)nq_scratch(ctx, SL_W1, plane * nsplit) : nullptr;
    if (!Wre || (out_complex && !Wim)) return NQ_ERR_ALLOC;
    // FP64: S runs on the integer tensor cores (Ozaki scheme, nq_syrk_ozaki.cu) from 4096 samples on; NQ_SR_FP64=dmma
    // forces the DMMA kernel (which is also the fallback when the digit planes cannot be allocated)
    static const int want_ozaki = [] { const char* e = getenv("NQ_SR_FP64"); return e ? (!strcmp(e, "dmma") ? 0 : 1) : 1; }();
    bool oz = false;
    // row maxima left by the centring pass of this very matrix (one-shot): the Ozaki pre-pass skips its own read of O
    const bool rec = ctx->rowmax_ptr == Oc && ctx->rowmax_P == P && ctx->rowmax_Ns == Ns && ctx->rowmax_ld == ldO;
    const unsigned long long* known_max = rec ? ctx->rowmax : nullptr;
    // deferred centring of this matrix (nq_center_lazy): stays pending while the rows stay uncentred
    const cxd* shift = ctx->shift_matches(Oc, ldO, P, Ns) ? (const cxd*)ctx->shift : nullptr;
    ctx->rowmax_ptr = nullptr;
    const bool ozaki_ok = want_ozaki && nq_dtype_is_double(dtype) && Ns >= 4096;
    if (shift && !ozaki_ok) {            // nobody to subtract on the fly: centre in place now
        NQ_CHECK(launch_subtract(ctx, const_cast<void*>(Oc), ldO, P, Ns, dtype, shift, nullptr));
        ctx->shift_pending = false;
        shift = nullptr; known_max = nullptr;
    }
    if (ozaki_ok) {        // below: fixed costs of the pre-pass win (cfg2: 0.34 vs 0.47 ms)
        // its own split of K: every CTA drains its accumulators into the FP64 partial once per pass, so a split should cover
        // >= 2048 samples (the DMMA split above goes down to 128, which at 8 192 samples per GPU made the Ozaki path slower)
        int ns_oz = 1;
        double best = 1e30;
        for (int ns = 1; ns <= nsplit && ns <= 32; ns++) {
            if (ns > 1 && Ns / ns < 2048) break;
            const double waves = (double)ntri * ns / ctx->num_sms;
            const double cost = std::ceil(waves) / waves + 0.01 * ns;
            if ((waves >= 1.0 || ns == 1) && cost < best) { best = cost; ns_oz = ns; }
        }
        NQ_CHECK(nq_syrk_ozaki_device(ctx, (const double*)Oc, ldr, P, Ns, ocx ? 2 : 1, ntile, ns_oz, Wre, out_complex ? Wim : nullptr,
                                      known_max, (const double*)shift, &oz));
        if (oz) nsplit = ns_oz;
        else if (shift) {                                                                                           // fallback path
            NQ_CHECK(launch_subtract(ctx, const_cast<void*>(Oc), ldO, P, Ns, dtype, shift, nullptr));
            ctx->shift_pending = false;
        }
    }
    if (ocx) {
        if (nq_dtype_is_double(dtype)) {
            if (!oz) NQ_CHECK((launch_syrk<double, 2>(ctx, Oc, ldr, P, Ns, ntile, nsplit, 0, Wre)));
            if (out_complex && !oz) NQ_CHECK((launch_syrk<double, 2>(ctx, Oc, ldr, P, Ns, ntile, nsplit, 1, Wim)));
        } else {
            NQ_CHECK((launch_syrk<float, 2>(ctx, Oc, ldr, P, Ns, ntile, nsplit, 0, Wre)));
            if (out_complex) NQ_CHECK((launch_syrk<float, 2>(ctx, Oc, ldr, P, Ns, ntile, nsplit, 1, Wim)));
        }
    } else {
        if (nq_dtype_is_double(dtype)) { if (!oz) NQ_CHECK((launch_syrk<double, 1>(ctx, Oc, ldr, P, Ns, ntile, nsplit, 0, Wre))); }
        else NQ_CHECK((launch_syrk<float, 1>(ctx, Oc, ldr, P, Ns, ntile, nsplit, 0, Wre)));
    }
    dim3 grid((unsigned)((P + 127) / 128), (unsigned)std::min<int64_t>(P, 65535));
    const double scale = 1.0 / (double)Ns_total;
    if (nq_dtype_is_double(dtype))
        NQ_LAUNCH(ctx, syrk_finalize_kernel<double>, grid, 128, 0, (const double*)Wre, (const double*)Wim, nsplit, Ppad, P, scale, (int)out_complex, (double*)dS);
    else
        NQ_LAUNCH(ctx, syrk_finalize_kernel<float>, grid, 128, 0, (const double*)Wre, (const double*)Wim, nsplit, Ppad, P, scale, (int)out_complex, (float*)dS);
    }
    ctx->hint_P = 0;                       // the structural hint is one-shot
    // F = gradC (complex nets) or Re(gradC)
    cxd* tmp = (cxd*)nq_scratch(ctx, SL_W2, (size_t)P * sizeof(cxd));
    if (!tmp) return NQ_ERR_ALLOC;
    NQ_LAUNCH(ctx, convert_to_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, dgc, tmp, P, (int)cdt);
    NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)tmp, dF, P, (int)sdt, 0);
    return st.finish();
}

template <typename E>
static int solve_typed(nq_ctx_t ctx, const cxd* Sd, const cxd* Fd, int64_t P, double eps, nq_solver algo, double tol,
                       int64_t maxiter, cxd* xd, int64_t* iters);

// convert (possibly complex) cxd arrays to E arrays and back
__global__ void cxd_to_real_kernel(const cxd* __restrict__ in, double* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i].re;
}

extern "C" int nq_sr_scale_diagonal(nq_ctx_t ctx, void* S, int64_t P, nq_dtype sdtype, double lambda) {
    if (!ctx || !S || P <= 0) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t es = nq_dtype_size(sdtype);
    const bool on_dev = nq_is_device_ptr(S);
    void* d = S;
    if (!on_dev) {
        d = nq_scratch(ctx, SL_W0, (size_t)P * P * es);
        if (!d) return NQ_ERR_ALLOC;
        NQ_CUDA(ctx, cudaMemcpyAsync(d, S, (size_t)P * P * es, cudaMemcpyHostToDevice, ctx->stream));
    }
    const unsigned g = (unsigned)((P + 255) / 256);
    switch (sdtype) {
        case NQ_F32: NQ_LAUNCH(ctx, scale_diag_kernel<float>, g, 256, 0, (float*)d, P, (float)(1.0 + lambda)); break;
        case NQ_F64: NQ_LAUNCH(ctx, scale_diag_kernel<double>, g, 256, 0, (double*)d, P, 1.0 + lambda); break;
        case NQ_C64: NQ_LAUNCH(ctx, scale_diag_kernel<cxf>, g, 256, 0, (cxf*)d, P, (float)(1.0 + lambda)); break;
        default: NQ_LAUNCH(ctx, scale_diag_kernel<cxd>, g, 256, 0, (cxd*)d, P, 1.0 + lambda); break;
    }
    if (!on_dev) {
        NQ_CUDA(ctx, cudaMemcpyAsync(S, d, (size_t)P * P * es, cudaMemcpyDeviceToHost, ctx->stream));
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return NQ_OK;
}

extern "C" int nq_sr_solve(nq_ctx_t ctx, void* S, const void* F, int64_t P, nq_dtype sdtype, double eps, nq_solver algo,
                           double tol, int64_t maxiter, void* dw, int64_t* iters) {
    if (!ctx || !S || !F || !dw || P <= 0) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const size_t es = nq_dtype_size(sdtype);
    const void* dS = st.in(SL_IN0, S, (size_t)P * P * es);
    const void* dF = st.in(SL_IN1, F, (size_t)P * es);
    const bool warm = algo == NQ_SOLVE_QLP_WARM;
    const void* dx0 = warm ? st.in(SL_IN2, dw, (size_t)P * es) : nullptr;     // warm start: dw holds x0 on entry
    void* ddw = st.out(SL_OUT0, dw, (size_t)P * es);
    if (st.status != NQ_OK) return st.status;
    const bool cplx = nq_dtype_is_complex(sdtype);
    const size_t ws = cplx ? sizeof(cxd) : sizeof(double);
    // double-precision working copies (S itself if it already is double and on the device)
    void* A = nullptr;
    nq_dtype wdt = cplx ? NQ_C128 : NQ_F64;
    const int64_t nn = P * P;
    if (sdtype == wdt && nq_is_device_ptr(S)) A = S;
    else {
        A = nq_scratch(ctx, SL_W0, (size_t)nn * ws);
        if (!A) return NQ_ERR_ALLOC;
        if (sdtype == wdt) NQ_CUDA(ctx, cudaMemcpyAsync(A, dS, (size_t)nn * ws, cudaMemcpyDeviceToDevice, ctx->stream));
        else if (cplx) NQ_LAUNCH(ctx, convert_to_cxd_kernel, (unsigned)((nn + 255) / 256), 256, 0, dS, (cxd*)A, nn, (int)sdtype);
        else {
            cxd* t = (cxd*)nq_scratch(ctx, SL_W5, (size_t)nn * sizeof(cxd));
            if (!t) return NQ_ERR_ALLOC;
            NQ_LAUNCH(ctx, convert_to_cxd_kernel, (unsigned)((nn + 255) / 256), 256, 0, dS, t, nn, (int)sdtype);
            NQ_LAUNCH(ctx, cxd_to_real_kernel, (unsigned)((nn + 255) / 256), 256, 0, (const cxd*)t, (double*)A, nn);
        }
    }
    void* b = nq_scratch(ctx, SL_W3, (size_t)3 * P * ws);
    if (!b) return NQ_ERR_ALLOC;
    void* x = (char*)b + (size_t)P * ws;
    void* x0w = (char*)b + (size_t)2 * P * ws;
    {
        cxd* t = (cxd*)nq_scratch(ctx, SL_W4, (size_t)P * sizeof(cxd));
        if (!t) return NQ_ERR_ALLOC;
        NQ_LAUNCH(ctx, convert_to_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, dF, t, P, (int)sdtype);
        NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)t, b, P, (int)wdt, 0);
        if (warm) {
            NQ_LAUNCH(ctx, convert_to_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, dx0, t, P, (int)sdtype);
            NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)t, x0w, P, (int)wdt, 0);
        }
    }
    int status = NQ_OK;
    int64_t its = 0;
    if (algo == NQ_SOLVE_QLP || algo == NQ_SOLVE_QLP_WARM) {
        if (maxiter <= 0) maxiter = 10 * P;
        int flag = 0;
        if (cplx) {
            CgOps<cxd> ops; ops.ctx = ctx; ops.P = P; ops.S = (const cxd*)A; ops.eps = eps;
            status = qlp_solve<cxd>(ops, (const cxd*)b, warm ? (const cxd*)x0w : nullptr, tol, maxiter, (cxd*)x, &its, &flag);
        } else {
            CgOps<double> ops; ops.ctx = ctx; ops.P = P; ops.S = (const double*)A; ops.eps = eps;
            status = qlp_solve<double>(ops, (const double*)b, warm ? (const double*)x0w : nullptr, tol, maxiter, (double*)x, &its, &flag);
        }
        if (status != NQ_OK && status != NQ_ERR_NOT_CONVERGED) return status;
        if (status == NQ_ERR_NOT_CONVERGED) nq_fail(ctx, status, "MINRES-QLP: iteration limit after %lld iterations", (long long)its);
    } else if (algo == NQ_SOLVE_CHOLESKY) {
        int* dinfo = (int*)nq_scratch(ctx, SL_W1, 128);
        if (!dinfo) return NQ_ERR_ALLOC;
        NQ_CUDA(ctx, cudaMemsetAsync(dinfo, 0xff, 4, ctx->stream));
        NQ_CUDA(ctx, cudaMemcpyAsync(x, b, (size_t)P * ws, cudaMemcpyDeviceToDevice, ctx->stream));
        if (cplx) {
            NQ_LAUNCH(ctx, add_diag_kernel<cxd>, (unsigned)((P + 255) / 256), 256, 0, (cxd*)A, P, eps);
            NQ_CHECK(cholesky_solve<cxd>(ctx, (cxd*)A, P, (cxd*)x, dinfo));
        } else {
            NQ_LAUNCH(ctx, add_diag_kernel<double>, (unsigned)((P + 255) / 256), 256, 0, (double*)A, P, eps);
            NQ_CHECK(cholesky_solve<double>(ctx, (double*)A, P, (double*)x, dinfo));
        }
        int hinfo = -1;
        NQ_CUDA(ctx, cudaMemcpyAsync(&hinfo, dinfo, 4, cudaMemcpyDeviceToHost, ctx->stream));
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (hinfo >= 0) {
            ctx->info = hinfo;
            status = nq_fail(ctx, NQ_ERR_NOT_POSDEF, "Cholesky: non-positive pivot at index %d", hinfo);
        }
    } else if (algo == NQ_SOLVE_CG || algo == NQ_SOLVE_MINRES) {
        if (maxiter <= 0) maxiter = 10 * P;
        const bool mr = algo == NQ_SOLVE_MINRES;
        if (cplx) {
            CgOps<cxd> ops; ops.ctx = ctx; ops.P = P; ops.S = (const cxd*)A; ops.eps = eps;
            status = mr ? minres_solve<cxd>(ops, (const cxd*)b, tol, maxiter, (cxd*)x, &its)
                        : cg_solve<cxd>(ops, (const cxd*)b, tol, maxiter, (cxd*)x, &its);
        } else {
            CgOps<double> ops; ops.ctx = ctx; ops.P = P; ops.S = (const double*)A; ops.eps = eps;
            status = mr ? minres_solve<double>(ops, (const double*)b, tol, maxiter, (double*)x, &its)
                        : cg_solve<double>(ops, (const double*)b, tol, maxiter, (double*)x, &its);
        }
        if (status != NQ_OK && status != NQ_ERR_NOT_CONVERGED) return status;
        if (status == NQ_ERR_NOT_CONVERGED) nq_fail(ctx, status, "%s: not converged after %lld iterations", algo == NQ_SOLVE_MINRES ? "MINRES" : "CG", (long long)its);
    } else return nq_fail(ctx, NQ_ERR_ARG, "unknown solver");
    if (iters) *iters = its;
    {
        cxd* t = (cxd*)nq_scratch(ctx, SL_W4, (size_t)P * sizeof(cxd));
        NQ_LAUNCH(ctx, convert_to_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const void*)x, t, P, (int)wdt);
        NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)t, ddw, P, (int)sdtype, 0);
    }
    int fs = st.finish();
    return fs != NQ_OK ? fs : status;
}

static int sr_solve_matfree_impl(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, int64_t Ns_total,
                                 nq_dtype dtype, const void* F, int real_params, double eps, nq_solver algo, double tol,
                                 int64_t maxiter, void* dw, int64_t* iters) {
    if (!ctx || !Oc || !F || !dw || P <= 0 || Ns <= 0 || ldO < P || Ns_total < Ns) return NQ_ERR_ARG;
    if (algo != NQ_SOLVE_CG && algo != NQ_SOLVE_MINRES && algo != NQ_SOLVE_QLP && algo != NQ_SOLVE_QLP_WARM)
        return nq_fail(ctx, NQ_ERR_ARG, "matrix-free SR needs an iterative solver");
    const bool mr = algo == NQ_SOLVE_MINRES, qlp = algo == NQ_SOLVE_QLP || algo == NQ_SOLVE_QLP_WARM, warm = algo == NQ_SOLVE_QLP_WARM;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    int hs = NQ_OK;
    Oc = sr_matrix_on_device(ctx, SL_HOSTO, Oc, ldO, Ns, dtype, &hs);
    if (hs != NQ_OK) return hs;
    const bool out_complex = nq_dtype_is_complex(dtype) && !real_params;
    nq_dtype sdt = out_complex ? dtype : nq_real_of(dtype);
    nq_dtype wdt = out_complex ? NQ_C128 : NQ_F64;
    const size_t ws = nq_dtype_size(wdt);
    NqStage st(ctx);
    const void* dF = st.in(SL_IN0, F, (size_t)P * nq_dtype_size(sdt));
    const void* dx0 = warm ? st.in(SL_IN2, dw, (size_t)P * nq_dtype_size(sdt)) : nullptr;
    void* ddw = st.out(SL_OUT0, dw, (size_t)P * nq_dtype_size(sdt));
    if (st.status != NQ_OK) return st.status;
    void* b = nq_scratch(ctx, SL_W0, (size_t)3 * P * ws);
    cxd* t = (cxd*)nq_scratch(ctx, SL_IN4, (size_t)P * sizeof(cxd));
    if (!b || !t) return NQ_ERR_ALLOC;
    void* x = (char*)b + (size_t)P * ws;
    void* x0w = (char*)b + (size_t)2 * P * ws;
    NQ_LAUNCH(ctx, convert_to_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, dF, t, P, (int)sdt);
    NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)t, b, P, (int)wdt, 0);
    if (warm) {
        NQ_LAUNCH(ctx, convert_to_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, dx0, t, P, (int)sdt);
        NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)t, x0w, P, (int)wdt, 0);
    }
    if (maxiter <= 0) maxiter = 10 * P;
    int status;
    int64_t its = 0;
    if (qlp) {
        int flag = 0;
        if (out_complex) {
            CgOps<cxd> ops; ops.ctx = ctx; ops.P = P; ops.eps = eps; ops.O = Oc; ops.ld = ldO; ops.Ns = Ns; ops.Ns_total = Ns_total; ops.odtype = dtype;
            status = qlp_solve<cxd>(ops, (const cxd*)b, warm ? (const cxd*)x0w : nullptr, tol, maxiter, (cxd*)x, &its, &flag);
        } else {
            CgOps<double> ops; ops.ctx = ctx; ops.P = P; ops.eps = eps; ops.O = Oc; ops.ld = ldO; ops.Ns = Ns; ops.Ns_total = Ns_total; ops.odtype = dtype;
            status = qlp_solve<double>(ops, (const double*)b, warm ? (const double*)x0w : nullptr, tol, maxiter, (double*)x, &its, &flag);
        }
    } else if (out_complex) {
        CgOps<cxd> ops; ops.ctx = ctx; ops.P = P; ops.eps = eps; ops.O = Oc; ops.ld = ldO; ops.Ns = Ns; ops.Ns_total = Ns_total; ops.odtype = dtype;
        status = mr ? minres_solve<cxd>(ops, (const cxd*)b, tol, maxiter, (cxd*)x, &its)
                    : cg_solve<cxd>(ops, (const cxd*)b, tol, maxiter, (cxd*)x, &its);
    } else {
        CgOps<double> ops; ops.ctx = ctx; ops.P = P; ops.eps = eps; ops.O = Oc; ops.ld = ldO; ops.Ns = Ns; ops.Ns_total = Ns_total; ops.odtype = dtype;
        status = mr ? minres_solve<double>(ops, (const double*)b, tol, maxiter, (double*)x, &its)
                    : cg_solve<double>(ops, (const double*)b, tol, maxiter, (double*)x, &its);
    }
    if (status != NQ_OK && status != NQ_ERR_NOT_CONVERGED) return status;
    if (status == NQ_ERR_NOT_CONVERGED) nq_fail(ctx, status, "%s: not converged after %lld iterations", qlp ? "MINRES-QLP" : (algo == NQ_SOLVE_MINRES ? "MINRES" : "CG"), (long long)its);
    if (iters) *iters = its;
    NQ_LAUNCH(ctx, convert_to_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const void*)x, t, P, (int)wdt);
    NQ_LAUNCH(ctx, convert_from_cxd_kernel, (unsigned)((P + 255) / 256), 256, 0, (const cxd*)t, ddw, P, (int)sdt, 0);
    int fs = st.finish();
    return fs != NQ_OK ? fs : status;
}

extern "C" int nq_sr_solve_matfree(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, int64_t Ns_total,
                                   nq_dtype dtype, const void* F, int real_params, double eps, double tol, int64_t maxiter,
                                   void* dw, int64_t* iters) {
    return sr_solve_matfree_impl(ctx, Oc, ldO, P, Ns, Ns_total, dtype, F, real_params, eps, NQ_SOLVE_CG, tol, maxiter, dw, iters);
}

extern "C" int nq_sr_solve_matfree_algo(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, int64_t Ns_total,
                                        nq_dtype dtype, const void* F, int real_params, double eps, nq_solver algo, double tol,
                                        int64_t maxiter, void* dw, int64_t* iters) {
    return sr_solve_matfree_impl(ctx, Oc, ldO, P, Ns, Ns_total, dtype, F, real_params, eps, algo, tol, maxiter, dw, iters);
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming S assembly: batches whose gradient matrix does not fit in memory (BASELINE cfg5: O > 60 GB) are produced
// chunk by chunk (nq_logpsi_grad_packed into ONE reused [P, Nc] buffer) and consumed at once.  The rows are shifted by
// a PROVISIONAL mean c (the mean of the first chunk) before they enter the Gram matrix, so the final centring
//     S = sum_chunks (O_c - c)(O_c - c)^H / Ns  -  (a - c)(a - c)^H,        a = <O> = c + sum (O - c) / Ns
// subtracts a term of the order of the statistical error of c instead of |<O>|^2: no cancellation in either precision.
// ref: the same S as SRDirect.jl:26-49 on the centred rows of BaseIterativeSampler.jl:19-26.
// ---------------------------------------------------------------------------------------------------------------
namespace {
template <typename T> __global__ void acc_add_kernel(T* __restrict__ acc, const T* __restrict__ x, int64_t n, int first) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i < n; i += (int64_t)gridDim.x * blockDim.x) acc[i] = first ? x[i] : acc[i] + x[i];
}
// state = [running sum of (O - c) | shift c]; cs = column sums of the UNshifted chunk
// first chunk: c = csg / count with csg the column sums of the first chunks of ALL ranks (= cs on one GPU)
__global__ void stream_shift_kernel(cxd* __restrict__ state, const cxd* __restrict__ cs, const cxd* __restrict__ csg,
                                    int64_t P, int64_t Nc, int first) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    if (first) {
        const double cnt = csg[P].re;
        state[P + i] = cxd(csg[i].re / cnt, csg[i].im / cnt);
        state[i] = cxd(0.0, 0.0);
    }
    const cxd c = state[P + i];
    state[i] = cxd(state[i].re + (cs[i].re - (double)Nc * c.re), state[i].im + (cs[i].im - (double)Nc * c.im));
}
// S(i,j) -= Re(d_i conj(d_j)) (real S) or conj(d_i) d_j (complex S), d = sum (O - c) / Ns_total; avg = c + d
template <typename T> __global__ void rank1_center_kernel(T* __restrict__ S, int64_t P, const cxd* __restrict__ state, double inv, int out_complex) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    const double ar = state[i].re * inv, ai = state[i].im * inv;
    for (int64_t j = blockIdx.y; j < P; j += gridDim.y) {
        const double br = state[j].re * inv, bi = state[j].im * inv;
        if (out_complex) {
            S[2 * (i + P * j)] = (T)((double)S[2 * (i + P * j)] - (ar * br + ai * bi));
            S[2 * (i + P * j) + 1] = (T)((double)S[2 * (i + P * j) + 1] - (ar * bi - ai * br));
        } else {
            S[i + P * j] = (T)((double)S[i + P * j] - (ar * br + ai * bi));
        }
    }
}
__global__ void stream_avg_kernel(cxd* __restrict__ state, int64_t P, double inv) {       // state[0..P) <- <O>
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < P) state[i] = cxd(state[P + i].re + state[i].re * inv, state[P + i].im + state[i].im * inv);
}
}  // namespace

extern "C" int nq_sr_accumulate(nq_ctx_t ctx, void* O, int64_t ldO, int64_t P, int64_t Nc, int64_t Ns_total, nq_dtype dtype,
                                int real_params, void* Sacc, void* state, int first) {
    if (!ctx || !O || !Sacc || !state || P <= 0 || Nc <= 0 || ldO < P || Ns_total < Nc) return NQ_ERR_ARG;
    if (!nq_is_device_ptr(O) || !nq_is_device_ptr(Sacc) || !nq_is_device_ptr(state))
        return nq_fail(ctx, NQ_ERR_ARG, "nq_sr_accumulate works on device-resident buffers");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool ocx = nq_dtype_is_complex(dtype), out_complex = ocx && !real_params;
    const nq_dtype sdt = out_complex ? dtype : nq_real_of(dtype);
    const size_t sbytes = (size_t)P * P * nq_dtype_size(sdt);
    cxd* cs = (cxd*)nq_scratch(ctx, SL_W2, (size_t)(2 * P + 2) * sizeof(cxd));
    if (!cs) return NQ_ERR_ALLOC;
    cxd* csg = cs + P + 1;                 // [P] column sums + [1] sample count of the first chunks of all ranks
    NQ_CHECK(colsum_dispatch<false>(ctx, dtype, O, ldO, P, Nc, nullptr, 1.0, cs));
    if (first) {
        const cxd cnt((double)Nc, 0.0);
        NQ_CUDA(ctx, cudaMemcpyAsync(csg, cs, (size_t)P * sizeof(cxd), cudaMemcpyDeviceToDevice, ctx->stream));
        NQ_CUDA(ctx, cudaMemcpyAsync(csg + P, &cnt, sizeof(cxd), cudaMemcpyHostToDevice, ctx->stream));
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // cnt lives on the host stack
        // shards: the SAME shift on every rank, so that the partial sums of nq_sr_finish are plain sums
        if (ctx->nccl_comm) NQ_CHECK(nq_allreduce_device(ctx, csg, P + 1, NQ_C128, false));
    }
    NQ_LAUNCH(ctx, stream_shift_kernel, (unsigned)((P + 255) / 256), 256, 0, (cxd*)state, (const cxd*)cs, (const cxd*)csg, P, Nc, first);
    {   // O_c <- O_c - c in place (the chunk buffer is the caller's scratch)
        dim3 grid((unsigned)((P + 127) / 128), (unsigned)std::min<int64_t>(Nc, 4096));
        const cxd* c = (const cxd*)state + P;
        unsigned long long* mx = nq_dtype_is_double(dtype) ? rowmax_arm(ctx, O, ldO, P, Nc) : nullptr;
        switch (dtype) {
            case NQ_F32: NQ_LAUNCH(ctx, subtract_avg_kernel<float>, grid, 128, 0, (float*)O, ldO, P, Nc, c, mx); break;
            case NQ_F64: NQ_LAUNCH(ctx, subtract_avg_kernel<double>, grid, 128, 0, (double*)O, ldO, P, Nc, c, mx); break;
            case NQ_C64: NQ_LAUNCH(ctx, subtract_avg_kernel<cxf>, grid, 128, 0, (cxf*)O, ldO, P, Nc, c, mx); break;
            default: NQ_LAUNCH(ctx, subtract_avg_kernel<cxd>, grid, 128, 0, (cxd*)O, ldO, P, Nc, c, mx); break;
        }
    }
    // chunk Gram matrix with the global normalisation into a scratch S (the first chunk goes straight to Sacc)
    void* St = first ? Sacc : nq_scratch(ctx, SL_HOSTG, sbytes);
    void* zg = nq_scratch(ctx, SL_IN4, (size_t)2 * P * sizeof(cxd));
    if (!St || !zg) return NQ_ERR_ALLOC;
    NQ_CUDA(ctx, cudaMemsetAsync(zg, 0, (size_t)P * sizeof(cxd), ctx->stream));
    NQ_CHECK(nq_sr_setup(ctx, O, ldO, P, Nc, Ns_total, dtype, zg, real_params, St, (char*)zg + (size_t)P * sizeof(cxd)));
    if (!first) {
        const int64_t n = (int64_t)(sbytes / (nq_dtype_is_double(dtype) ? 8 : 4));
        const unsigned g = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 16);
        if (nq_dtype_is_double(dtype)) NQ_LAUNCH(ctx, acc_add_kernel<double>, g, 256, 0, (double*)Sacc, (const double*)St, n, 0);
        else NQ_LAUNCH(ctx, acc_add_kernel<float>, g, 256, 0, (float*)Sacc, (const float*)St, n, 0);
    }
    return NQ_OK;
}

extern "C" int nq_sr_finish(nq_ctx_t ctx, void* Sacc, void* state, int64_t P, int64_t Ns_total, nq_dtype dtype, int real_params) {
    if (!ctx || !Sacc || !state || P <= 0 || Ns_total <= 0) return NQ_ERR_ARG;
    if (!nq_is_device_ptr(Sacc) || !nq_is_device_ptr(state)) return nq_fail(ctx, NQ_ERR_ARG, "nq_sr_finish works on device-resident buffers");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool out_complex = nq_dtype_is_complex(dtype) && !real_params;
    if (ctx->nccl_comm) {
        // shards: every rank used the same shift (all-reduced in the first nq_sr_accumulate), so partial S and partial
        // sums of (O - c) are plain sums
        const nq_dtype sdt = out_complex ? dtype : nq_real_of(dtype);
        NQ_CHECK(nq_allreduce_device(ctx, Sacc, P * P, sdt, false));
        NQ_CHECK(nq_allreduce_device(ctx, state, P, NQ_C128, false));
    }
    dim3 grid((unsigned)((P + 127) / 128), (unsigned)std::min<int64_t>(P, 65535));
    const double inv = 1.0 / (double)Ns_total;
    if (nq_dtype_is_double(dtype)) NQ_LAUNCH(ctx, rank1_center_kernel<double>, grid, 128, 0, (double*)Sacc, P, (const cxd*)state, inv, (int)out_complex);
    else NQ_LAUNCH(ctx, rank1_center_kernel<float>, grid, 128, 0, (float*)Sacc, P, (const cxd*)state, inv, (int)out_complex);
    NQ_LAUNCH(ctx, stream_avg_kernel, (unsigned)((P + 255) / 256), 256, 0, (cxd*)state, P, inv);
    return NQ_OK;
}

// Nesterov(lr, mu) (Optimisers/rules.jl:36-55) on device vectors of the machine dtype: with the velocity v of the
// parameter vector,  d = mu^2 v - (1 + mu) lr dw;  v <- mu v - lr dw;  delta = -d  (the caller applies w <- w - delta).
namespace {
template <typename E> __global__ void nesterov_kernel(E* __restrict__ v, const E* __restrict__ dw, E* __restrict__ delta,
                                                      typename elem_traits<E>::real lr, typename elem_traits<E>::real mu, int64_t n) {
    typedef typename elem_traits<E>::real T;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const E vi = v[i], gi = dw[i];
    delta[i] = rscale((T(1) + mu) * lr, gi) - rscale(mu * mu, vi);
    v[i] = rscale(mu, vi) - rscale(lr, gi);
}
}  // namespace

extern "C" int nq_nesterov(nq_ctx_t ctx, void* velocity, const void* dw, int64_t n, nq_dtype dtype, double lr, double mu, void* delta) {
    if (!ctx || !velocity || !dw || !delta || n <= 0) return NQ_ERR_ARG;
    if (!nq_is_device_ptr(velocity) || !nq_is_device_ptr(dw) || !nq_is_device_ptr(delta)) return nq_fail(ctx, NQ_ERR_ARG, "nq_nesterov works on device vectors");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned g = (unsigned)((n + 255) / 256);
    switch (dtype) {
        case NQ_F32: NQ_LAUNCH(ctx, nesterov_kernel<float>, g, 256, 0, (float*)velocity, (const float*)dw, (float*)delta, (float)lr, (float)mu, n); break;
        case NQ_F64: NQ_LAUNCH(ctx, nesterov_kernel<double>, g, 256, 0, (double*)velocity, (const double*)dw, (double*)delta, lr, mu, n); break;
        case NQ_C64: NQ_LAUNCH(ctx, nesterov_kernel<cxf>, g, 256, 0, (cxf*)velocity, (const cxf*)dw, (cxf*)delta, (float)lr, (float)mu, n); break;
        default: NQ_LAUNCH(ctx, nesterov_kernel<cxd>, g, 256, 0, (cxd*)velocity, (const cxd*)dw, (cxd*)delta, lr, mu, n); break;
    }
    return NQ_OK;
}

extern "C" int nq_update(nq_machine_t m, const void* dw, double eta) {
    if (!m || !dw) return NQ_ERR_ARG;
    nq_ctx_t ctx = m->ctx;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const void* d = st.in(SL_IN0, dw, (size_t)m->P * nq_dtype_size(m->dtype));
    if (st.status != NQ_OK) return st.status;
    unsigned g = (unsigned)((m->P + 255) / 256);
    switch (m->dtype) {
        case NQ_F32: NQ_LAUNCH(ctx, update_kernel<float>, g, 256, 0, (float*)m->params, (const float*)d, (float)eta, m->P); break;
        case NQ_F64: NQ_LAUNCH(ctx, update_kernel<double>, g, 256, 0, (double*)m->params, (const double*)d, eta, m->P); break;
        case NQ_C64: NQ_LAUNCH(ctx, update_kernel<cxf>, g, 256, 0, (cxf*)m->params, (const cxf*)d, (float)eta, m->P); break;
        default: NQ_LAUNCH(ctx, update_kernel<cxd>, g, 256, 0, (cxd*)m->params, (const cxd*)d, eta, m->P); break;
    }
    m->etab_valid = false;
    if (!nq_is_device_ptr(dw)) NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NQ_OK;
}

extern "C" int nq_stat_analysis(nq_ctx_t ctx, const void* vals, int64_t B, int64_t L, nq_dtype vdtype, double out[6]) {
    if (!ctx || !vals || !out || B <= 0 || L <= 0) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const void* dv = st.in(SL_IN0, vals, (size_t)B * L * nq_dtype_size(vdtype));
    double* dst = (double*)nq_scratch(ctx, SL_W0, (size_t)B * 3 * sizeof(double));
    if (st.status != NQ_OK || !dst) return NQ_ERR_ALLOC;
    unsigned g = (unsigned)((B + 127) / 128);
    switch (vdtype) {
        case NQ_F32: NQ_LAUNCH(ctx, chain_stats_kernel<float>, g, 128, 0, (const float*)dv, B, L, dst); break;
        case NQ_F64: NQ_LAUNCH(ctx, chain_stats_kernel<double>, g, 128, 0, (const double*)dv, B, L, dst); break;
        case NQ_C64: NQ_LAUNCH(ctx, chain_stats_kernel<cxf>, g, 128, 0, (const cxf*)dv, B, L, dst); break;
        default: NQ_LAUNCH(ctx, chain_stats_kernel<cxd>, g, 128, 0, (const cxd*)dv, B, L, dst); break;
    }
    std::vector<double> h((size_t)B * 3);
    NQ_CUDA(ctx, cudaMemcpyAsync(h.data(), dst, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // combine the per-chain moments exactly as utils/stats.jl:26-50 defines them; under sharding over the union of
    // the ranks' chains (two tiny all-reduces: counts and sums, then the spread around the global mean)
    double mr = 0, mi = 0, m2sum = 0;
    for (int64_t b = 0; b < B; b++) { mr += h[3 * b]; mi += h[3 * b + 1]; m2sum += h[3 * b + 2]; }
    double Bt = (double)B;
    auto host_allreduce = [&](double* v, int n) -> int {
        double* d = (double*)nq_scratch(ctx, SL_W1, 8 * sizeof(double));
        if (!d) return NQ_ERR_ALLOC;
        NQ_CUDA(ctx, cudaMemcpyAsync(d, v, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        NQ_CHECK(nq_allreduce_device(ctx, d, n, NQ_F64, false));
        NQ_CUDA(ctx, cudaMemcpyAsync(v, d, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return NQ_OK;
    };
    if (ctx->nccl_comm) {
        double v[4] = {Bt, mr, mi, m2sum};
        NQ_CHECK(host_allreduce(v, 4));
        Bt = v[0]; mr = v[1]; mi = v[2]; m2sum = v[3];
    }
    mr /= Bt; mi /= Bt;
    double between = 0;
    for (int64_t b = 0; b < B; b++) { double dr = h[3 * b] - mr, di = h[3 * b + 1] - mi; between += dr * dr + di * di; }
    if (ctx->nccl_comm) NQ_CHECK(host_allreduce(&between, 1));
    double var_chains_mean = L > 1 ? (m2sum / (double)(L - 1)) / Bt : NAN;
    double var_mu_ch = Bt > 1 ? between / (Bt - 1.0) : NAN;
    double var_mu = (Bt * L > 1) ? (m2sum + (double)L * between) / (Bt * (double)L - 1.0) : NAN;
    double t = var_mu_ch / var_mu;
    out[0] = mr; out[1] = mi;
    out[2] = sqrt(var_mu_ch / Bt);
    out[3] = var_chains_mean;
    out[4] = std::max(0.0, 0.5 * (t * (double)L - 1.0));
    out[5] = sqrt((double)(L - 1) / (double)L + t);
    return NQ_OK;
}
