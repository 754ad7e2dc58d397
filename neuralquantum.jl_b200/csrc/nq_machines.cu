// nq_machines.cu -- K1-K3: fused evaluation (+ gradient rows O) of RBM / RBMSplit / NDM.
//
// One CTA owns TB configurations.  Their site values sit in shared memory (decoded once from the
// bit-packed words); each thread owns one hidden unit of the current chunk, accumulates
// theta = bias + W v for the TB configurations in registers (W is read once per CTA, coalesced along
// the hidden index), applies the activation and its derivative, and the CTA then streams the
// gradient rows [bias-grad | d (x) v] to HBM with fully coalesced parameter-fastest stores.
// With O materialised the kernel is HBM-write bound: P*sizeof(elem) bytes per configuration.
//
// ref: Networks/ClosedSystems/RBMBatched.jl:37-91, Networks/MixedDensityMatrix/RBMSplitBatched.jl:35-101,
//      Networks/MixedDensityMatrix/NDMBatched.jl:94-280, utils/math.jl:23-77 (outer products),
//      tuple_logic.jl:82-118 (flat gradient layout).
#include "nq_internal.cuh"

namespace {

template <typename E, int TB, int NT>
__device__ __forceinline__ void block_reduce_store(E (&lsum)[TB], E* red, E* out, int64_t s0, int nb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int s = 0; s < TB; s++) {
        E v = warp_sum(lsum[s]);
        if (lane == 0) red[warp * TB + s] = v;
    }
    __syncthreads();
    if (tid < nb) {
        E v = red[tid];
        for (int w = 1; w < NT / 32; w++) v = v + red[w * TB + tid];
        out[s0 + tid] = v;
    }
}

// ---------------------------------------------------------------------------------------
// RBM (DOUBLED=false) and RBMSplit (DOUBLED=true)
// ---------------------------------------------------------------------------------------
template <typename E, int ACT, bool DOUBLED, bool GRAD, int TB, int NT>
__global__ void __launch_bounds__(NT)
rbm_evalgrad_kernel(const E* __restrict__ par, const uint64_t* __restrict__ prow,
                    const uint64_t* __restrict__ pcol, int64_t B, int N, int M, int hilb,
                    E* __restrict__ out, E* __restrict__ O, int64_t ldO) {
    typedef typename elem_traits<E>::real T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* vs = (T*)smem_raw;                                   // [2][N][TB]
    E* ds = (E*)(smem_raw + (size_t)2 * N * TB * sizeof(T)); // [TB][NT]
    E* red = ds + TB * NT;                                  // [NT/32][TB]
    const int tid = threadIdx.x;
    const int W64 = (N + 63) >> 6;
    const int64_t s0 = blockIdx.x * (int64_t)TB;
    const int nb = (int)((B - s0) < TB ? (B - s0) : TB);

    for (int i = tid; i < 2 * N * TB; i += NT) {
        int s = i % TB, j = (i / TB) % N, c = i / (TB * N);
        T v = T(0);
        if (s < nb && (c == 0 || DOUBLED)) {
            const uint64_t* p = (c ? pcol : prow) + (s0 + s) * W64;
            v = digit_value<T>(hilb, get_bit(p, j));
        }
        vs[i] = v;
    }
    __syncthreads();
    const T* vr = vs;
    const T* vc = vs + N * TB;

    const int64_t off_b = DOUBLED ? 2 * N : N;
    const int64_t off_Wr = off_b + M;
    const int64_t off_Wc = off_Wr + (int64_t)M * N;
    const E* __restrict__ bb = par + off_b;
    const E* __restrict__ Wr = par + off_Wr;
    const E* __restrict__ Wc = par + off_Wc;

    E lsum[TB];
#pragma unroll
    for (int s = 0; s < TB; s++) lsum[s] = make_zero<E>();

    // visible biases: a . v   (+ O rows of a)
    for (int j = tid; j < N; j += NT) {
        E ar = par[j];
#pragma unroll
        for (int s = 0; s < TB; s++) lsum[s] += rscale(vr[j * TB + s], ar);
        if (DOUBLED) {
            E ac = par[N + j];
#pragma unroll
            for (int s = 0; s < TB; s++) lsum[s] += rscale(vc[j * TB + s], ac);
        }
    }
    if (GRAD) {
        for (int i = tid; i < nb * N; i += NT) {
            int s = i / N, j = i - s * N;
            E* Os = O + (s0 + s) * ldO;
            Os[j] = from_real<E, T>(vr[j * TB + s]);
            if (DOUBLED) Os[N + j] = from_real<E, T>(vc[j * TB + s]);
        }
    }

    for (int k0 = 0; k0 < M; k0 += NT) {
        const int k = k0 + tid;
        const int kc = (M - k0) < NT ? (M - k0) : NT;
        if (k < M) {
            E acc[TB];
            E b = bb[k];
#pragma unroll
            for (int s = 0; s < TB; s++) acc[s] = b;
            for (int j = 0; j < N; j++) {
                E w = Wr[k + (int64_t)M * j];
#pragma unroll
                for (int s = 0; s < TB; s++) acc[s] += rscale(vr[j * TB + s], w);
                if (DOUBLED) {
                    E w2 = Wc[k + (int64_t)M * j];
#pragma unroll
                    for (int s = 0; s < TB; s++) acc[s] += rscale(vc[j * TB + s], w2);
                }
            }
#pragma unroll
            for (int s = 0; s < TB; s++) {
                E f, d;
                act_eval<ACT>(acc[s], f, d);
                lsum[s] += f;
                if (GRAD) ds[s * NT + tid] = d;
            }
        }
        if (GRAD) {
            __syncthreads();
            for (int s = 0; s < nb; s++) {
                E* Os = O + (s0 + s) * ldO;
                const E* dss = ds + s * NT;
                if (k < M) Os[off_b + k] = dss[tid];
                if (kc == NT) {
                    E d = dss[tid];
                    for (int j = 0; j < N; j++) {
                        Os[off_Wr + (int64_t)M * j + k] = rscale(vr[j * TB + s], d);
                        if (DOUBLED) Os[off_Wc + (int64_t)M * j + k] = rscale(vc[j * TB + s], d);
                    }
                } else {
                    for (int i = tid; i < N * kc; i += NT) {
                        int j = i / kc, kk = i - j * kc;
                        E d = dss[kk];
                        Os[off_Wr + (int64_t)M * j + k0 + kk] = rscale(vr[j * TB + s], d);
                        if (DOUBLED) Os[off_Wc + (int64_t)M * j + k0 + kk] = rscale(vc[j * TB + s], d);
                    }
                }
            }
            __syncthreads();
        }
    }
    block_reduce_store<E, TB, NT>(lsum, red, out, s0, nb);
}

// ---------------------------------------------------------------------------------------
// NDM: real parameters, complex output.  Appendix A.3 of SURVEY.md / NDMBatched.jl:177-280
// ---------------------------------------------------------------------------------------
template <typename T, int ACT, bool GRAD, int TB, int NT>
__global__ void __launch_bounds__(NT)
ndm_evalgrad_kernel(const T* __restrict__ par, const uint64_t* __restrict__ prow,
                    const uint64_t* __restrict__ pcol, int64_t B, int N, int M, int A, int hilb,
                    cx<T>* __restrict__ out, cx<T>* __restrict__ O, int64_t ldO) {
    typedef cx<T> C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* vs = (T*)smem_raw;                                    // [2][N][TB]
    T* ds = (T*)(smem_raw + (size_t)2 * N * TB * sizeof(T));  // [4][TB][NT] reals, or [TB][NT] complex
    C* dpi = (C*)ds;
    C* red = (C*)(ds + 4 * TB * NT);                         // [NT/32][TB]
    const int tid = threadIdx.x;
    const int W64 = (N + 63) >> 6;
    const int64_t s0 = blockIdx.x * (int64_t)TB;
    const int nb = (int)((B - s0) < TB ? (B - s0) : TB);

    for (int i = tid; i < 2 * N * TB; i += NT) {
        int s = i % TB, j = (i / TB) % N, c = i / (TB * N);
        T v = T(0);
        if (s < nb) {
            const uint64_t* p = (c ? pcol : prow) + (s0 + s) * W64;
            v = digit_value<T>(hilb, get_bit(p, j));
        }
        vs[i] = v;
    }
    __syncthreads();
    const T* vr = vs;
    const T* vc = vs + N * TB;

    const int64_t MN = (int64_t)M * N, AN = (int64_t)A * N;
    const int64_t o_bmu = 0, o_hmu = N, o_wmu = N + M, o_umu = o_wmu + MN, o_blam = o_umu + AN,
                  o_hlam = o_blam + N, o_dlam = o_hlam + M, o_wlam = o_dlam + A, o_ulam = o_wlam + MN;
    const T half = T(0.5);

    C lsum[TB];
#pragma unroll
    for (int s = 0; s < TB; s++) lsum[s] = C(T(0), T(0));

    for (int j = tid; j < N; j += NT) {
        T bl = par[o_blam + j], bm = par[o_bmu + j];
#pragma unroll
        for (int s = 0; s < TB; s++) {
            T a = vr[j * TB + s], b = vc[j * TB + s];
            lsum[s].re += half * bl * (a + b);
            lsum[s].im += half * bm * (a - b);
        }
    }
    if (GRAD) {
        for (int i = tid; i < nb * N; i += NT) {
            int s = i / N, j = i - s * N;
            C* Os = O + (s0 + s) * ldO;
            T a = vr[j * TB + s], b = vc[j * TB + s];
            Os[o_bmu + j] = C(T(0), half * (a - b));
            Os[o_blam + j] = C(half * (a + b), T(0));
        }
    }

    // hidden layers lambda / mu on sigma and sigma'
    for (int k0 = 0; k0 < M; k0 += NT) {
        const int k = k0 + tid;
        const int kc = (M - k0) < NT ? (M - k0) : NT;
        if (k < M) {
            T tl[TB], tlp[TB], tm[TB], tmp[TB];
            T hl = par[o_hlam + k], hm = par[o_hmu + k];
#pragma unroll
            for (int s = 0; s < TB; s++) { tl[s] = hl; tlp[s] = hl; tm[s] = hm; tmp[s] = hm; }
            const T* __restrict__ wl = par + o_wlam + k;
            const T* __restrict__ wm = par + o_wmu + k;
            for (int j = 0; j < N; j++) {
                T a = wl[(int64_t)M * j], b = wm[(int64_t)M * j];
#pragma unroll
                for (int s = 0; s < TB; s++) {
                    T x = vr[j * TB + s], y = vc[j * TB + s];
                    tl[s] += a * x; tlp[s] += a * y; tm[s] += b * x; tmp[s] += b * y;
                }
            }
#pragma unroll
            for (int s = 0; s < TB; s++) {
                T fl, dl, flp, dlp, fm, dm, fmp, dmp;
                act_eval<ACT>(tl[s], fl, dl);
                act_eval<ACT>(tlp[s], flp, dlp);
                act_eval<ACT>(tm[s], fm, dm);
                act_eval<ACT>(tmp[s], fmp, dmp);
                lsum[s].re += half * (fl + flp);
                lsum[s].im += half * (fm - fmp);
                if (GRAD) {
                    ds[(0 * TB + s) * NT + tid] = dl;
                    ds[(1 * TB + s) * NT + tid] = dlp;
                    ds[(2 * TB + s) * NT + tid] = dm;
                    ds[(3 * TB + s) * NT + tid] = dmp;
                }
            }
        }
        if (GRAD) {
            __syncthreads();
            for (int s = 0; s < nb; s++) {
                C* Os = O + (s0 + s) * ldO;
                const T* Dl = ds + (0 * TB + s) * NT;
                const T* Dlp = ds + (1 * TB + s) * NT;
                const T* Dm = ds + (2 * TB + s) * NT;
                const T* Dmp = ds + (3 * TB + s) * NT;
                if (k < M) {
                    Os[o_hmu + k] = C(T(0), half * (Dm[tid] - Dmp[tid]));
                    Os[o_hlam + k] = C(half * (Dl[tid] + Dlp[tid]), T(0));
                }
                for (int i = tid; i < N * kc; i += NT) {
                    int j = i / kc, kk = i - j * kc;
                    T x = vr[j * TB + s], y = vc[j * TB + s];
                    Os[o_wmu + (int64_t)M * j + k0 + kk] = C(T(0), half * (Dm[kk] * x - Dmp[kk] * y));
                    Os[o_wlam + (int64_t)M * j + k0 + kk] = C(half * (Dl[kk] * x + Dlp[kk] * y), T(0));
                }
            }
            __syncthreads();
        }
    }

    // ancilla layer Pi (complex pre-activation)
    for (int a0 = 0; a0 < A; a0 += NT) {
        const int a = a0 + tid;
        const int ac = (A - a0) < NT ? (A - a0) : NT;
        if (a < A) {
            T pr[TB], pim[TB];
            T dl0 = par[o_dlam + a];
#pragma unroll
            for (int s = 0; s < TB; s++) { pr[s] = dl0; pim[s] = T(0); }
            const T* __restrict__ ul = par + o_ulam + a;
            const T* __restrict__ um = par + o_umu + a;
            for (int j = 0; j < N; j++) {
                T p = half * ul[(int64_t)A * j], q = half * um[(int64_t)A * j];
#pragma unroll
                for (int s = 0; s < TB; s++) {
                    T x = vr[j * TB + s], y = vc[j * TB + s];
                    pr[s] += p * (x + y);
                    pim[s] += q * (x - y);
                }
            }
#pragma unroll
            for (int s = 0; s < TB; s++) {
                C f, d;
                act_eval<ACT>(C(pr[s], pim[s]), f, d);
                lsum[s] += f;
                if (GRAD) dpi[s * NT + tid] = d;
            }
        }
        if (GRAD) {
            __syncthreads();
            for (int s = 0; s < nb; s++) {
                C* Os = O + (s0 + s) * ldO;
                const C* D = dpi + s * NT;
                if (a < A) Os[o_dlam + a] = D[tid];
                for (int i = tid; i < N * ac; i += NT) {
                    int j = i / ac, kk = i - j * ac;
                    T x = vr[j * TB + s], y = vc[j * TB + s];
                    C d = D[kk];
                    T hs = half * (x + y), hd = half * (x - y);
                    Os[o_ulam + (int64_t)A * j + a0 + kk] = C(hs * d.re, hs * d.im);
                    Os[o_umu + (int64_t)A * j + a0 + kk] = C(-hd * d.im, hd * d.re);
                }
            }
            __syncthreads();
        }
    }
    block_reduce_store<C, TB, NT>(lsum, red, out, s0, nb);
}

// exp(+-c W) tables: out[sign][i] = exp((sign ? -c : c) * w[i])
template <typename E>
__global__ void etab_kernel(const E* __restrict__ w, E* __restrict__ out, int64_t n, typename elem_traits<E>::real c) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    E x = rscale(c, w[i]);
    out[i] = e_exp(x);
    out[n + i] = e_exp(-x);
}
// Tables of the register-resident RBM sampler (nq_sampler.cu, sampler_rbm_reg_kernel).  log cosh(theta) = softplus(2 theta)
// - theta - ln 2, so both activations run the softplus recurrence on q = sigmoid(g theta) (g = 2 for logcosh, 1 for
// softplus):  out[sign][i] = exp(+-c g w_i) - 1 (expm1-accurate), followed by wsum[mat][j] = sum_k W_kj (the -theta term).
__device__ __forceinline__ float em1(float x) { return expm1f(x); }
__device__ __forceinline__ double em1(double x) { return expm1(x); }
template <typename T> __device__ __forceinline__ cx<T> em1(cx<T> z) {
    T sy, cy, sh, ch;
    m_sincos(z.im, &sy, &cy);
    m_sincos(T(0.5) * z.im, &sh, &ch);
    const T e = em1(z.re);
    return cx<T>(e * cy - T(2) * sh * sh, (e + T(1)) * sy);     // e^x cos y - 1 = expm1(x) cos y - 2 sin^2(y/2)
}
template <typename E>
__global__ void qtab_kernel(const E* __restrict__ w, E* __restrict__ out, int64_t n, typename elem_traits<E>::real c) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    E x = rscale(c, w[i]);
    out[i] = em1(x);
    out[n + i] = em1(-x);
}
template <typename E>
__global__ void wsum_kernel(const E* __restrict__ w, E* __restrict__ out, int M, int ncol) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncol) return;
    E s = make_zero<E>();
    for (int k = 0; k < M; k++) s += w[k + (int64_t)M * j];
    out[j] = s;
}
// NDM ancilla tables: out[sign][a + A j] = exp(+-(c/2) (u_lam + i u_mu))
template <typename T>
__global__ void etab_pi_kernel(const T* __restrict__ ul, const T* __restrict__ um, cx<T>* __restrict__ out, int64_t n, T c) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    cx<T> x(T(0.5) * c * ul[i], T(0.5) * c * um[i]);
    out[i] = cx_exp(x);
    out[n + i] = cx_exp(-x);
}

template <typename T>
__global__ void etab_strided_kernel(const T* __restrict__ w, T* __restrict__ out, int64_t n, int64_t sign_stride, T c) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    T x = c * w[i];
    out[i] = m_exp(x);
    out[sign_stride + i] = m_exp(-x);
}

template <typename E>
__global__ void log_prob_kernel(const E* __restrict__ in, typename elem_traits<E>::real* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = typename elem_traits<E>::real(2) * real_part(in[i]);
}

constexpr int TB_RBM = 8;
constexpr int TB_NDM = 4;

// threads per CTA follow the number of hidden units (one thread per unit of the current chunk): a
// machine with 32 units runs one-warp CTAs (no barrier, 16 CTAs per SM) instead of idling 7 of 8 warps
inline int pick_threads(int units) { return units <= 32 ? 32 : units <= 64 ? 64 : units <= 128 ? 128 : 256; }

template <typename E, int ACT, bool DOUBLED, int NT>
int launch_rbm_nt(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B, void* out, void* O, int64_t ldO) {
    typedef typename elem_traits<E>::real T;
    nq_ctx_t ctx = m->ctx;
    size_t smem = (size_t)2 * m->N * TB_RBM * sizeof(T) + (size_t)TB_RBM * NT * sizeof(E) + (NT / 32) * TB_RBM * sizeof(E);
    unsigned grid = (unsigned)((B + TB_RBM - 1) / TB_RBM);
    if (O) {
        auto kern = rbm_evalgrad_kernel<E, ACT, DOUBLED, true, TB_RBM, NT>;
        NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NQ_LAUNCH(ctx, kern, grid, NT, smem, (const E*)m->params, prow, pcol, B, m->N, m->M, (int)m->hilb, (E*)out, (E*)O, ldO);
    } else {
        auto kern = rbm_evalgrad_kernel<E, ACT, DOUBLED, false, TB_RBM, NT>;
        NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NQ_LAUNCH(ctx, kern, grid, NT, smem, (const E*)m->params, prow, pcol, B, m->N, m->M, (int)m->hilb, (E*)out, (E*)nullptr, ldO);
    }
    return NQ_OK;
}

template <typename E, int ACT, bool DOUBLED>
int launch_rbm(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B, void* out, void* O, int64_t ldO) {
    switch (pick_threads(m->M)) {
        case 32: return launch_rbm_nt<E, ACT, DOUBLED, 32>(m, prow, pcol, B, out, O, ldO);
        case 64: return launch_rbm_nt<E, ACT, DOUBLED, 64>(m, prow, pcol, B, out, O, ldO);
        case 128: return launch_rbm_nt<E, ACT, DOUBLED, 128>(m, prow, pcol, B, out, O, ldO);
        default: return launch_rbm_nt<E, ACT, DOUBLED, 256>(m, prow, pcol, B, out, O, ldO);
    }
}

template <typename T, int ACT, int NT>
int launch_ndm_nt(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B, void* out, void* O, int64_t ldO) {
    nq_ctx_t ctx = m->ctx;
    size_t smem = (size_t)2 * m->N * TB_NDM * sizeof(T) + (size_t)4 * TB_NDM * NT * sizeof(T) + (NT / 32) * TB_NDM * sizeof(cx<T>);
    unsigned grid = (unsigned)((B + TB_NDM - 1) / TB_NDM);
    if (O) {
        auto kern = ndm_evalgrad_kernel<T, ACT, true, TB_NDM, NT>;
        NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NQ_LAUNCH(ctx, kern, grid, NT, smem, (const T*)m->params, prow, pcol, B, m->N, m->M, m->A, (int)m->hilb, (cx<T>*)out, (cx<T>*)O, ldO);
    } else {
        auto kern = ndm_evalgrad_kernel<T, ACT, false, TB_NDM, NT>;
        NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NQ_LAUNCH(ctx, kern, grid, NT, smem, (const T*)m->params, prow, pcol, B, m->N, m->M, m->A, (int)m->hilb, (cx<T>*)out, (cx<T>*)nullptr, ldO);
    }
    return NQ_OK;
}

template <typename T, int ACT>
int launch_ndm(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B, void* out, void* O, int64_t ldO) {
    switch (pick_threads(m->M > m->A ? m->M : m->A)) {
        case 32: return launch_ndm_nt<T, ACT, 32>(m, prow, pcol, B, out, O, ldO);
        case 64: return launch_ndm_nt<T, ACT, 64>(m, prow, pcol, B, out, O, ldO);
        case 128: return launch_ndm_nt<T, ACT, 128>(m, prow, pcol, B, out, O, ldO);
        default: return launch_ndm_nt<T, ACT, 256>(m, prow, pcol, B, out, O, ldO);
    }
}

template <typename E>
int dispatch_rbm(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B, void* out, void* O, int64_t ldO) {
    if (m->kind == NQ_RBMSPLIT) return launch_rbm<E, NQ_SOFTPLUS, true>(m, prow, pcol, B, out, O, ldO);
    if (m->act == NQ_SOFTPLUS) return launch_rbm<E, NQ_SOFTPLUS, false>(m, prow, pcol, B, out, O, ldO);
    return launch_rbm<E, NQ_LOGCOSH, false>(m, prow, pcol, B, out, O, ldO);
}

}  // namespace

int nq_machine_eval_device(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B,
                           void* out, void* O, int64_t ldO) {
    if (B == 0) return NQ_OK;
    if (O) { m->ctx->shift_pending = false; m->ctx->rowmax_ptr = nullptr; }     // new rows: records of a centring pass are stale
    if (m->kind == NQ_NDM) {
        if (m->dtype == NQ_F64)
            return m->act == NQ_SOFTPLUS ? launch_ndm<double, NQ_SOFTPLUS>(m, prow, pcol, B, out, O, ldO)
                                         : launch_ndm<double, NQ_LOGCOSH>(m, prow, pcol, B, out, O, ldO);
        return m->act == NQ_SOFTPLUS ? launch_ndm<float, NQ_SOFTPLUS>(m, prow, pcol, B, out, O, ldO)
                                     : launch_ndm<float, NQ_LOGCOSH>(m, prow, pcol, B, out, O, ldO);
    }
    switch (m->dtype) {
        case NQ_F32: return dispatch_rbm<float>(m, prow, pcol, B, out, O, ldO);
        case NQ_F64: return dispatch_rbm<double>(m, prow, pcol, B, out, O, ldO);
        case NQ_C64: return dispatch_rbm<cxf>(m, prow, pcol, B, out, O, ldO);
        default: return dispatch_rbm<cxd>(m, prow, pcol, B, out, O, ldO);
    }
}

// Layouts.  RBM/RBMSplit (element E): etab[sign][mat][k + M j], mat 0 = W / Wr, 1 = Wc.
// NDM (real T): etab[sign][lay][k + M j], lay 0 = w_lam, 1 = w_mu, followed (16-byte aligned) by the complex
// ancilla table [sign][a + A j].  sign 0: the site value increases, 1: it decreases.
static size_t etab_bytes(nq_machine_t m) {
    const int64_t MN = (int64_t)m->M * m->N, AN = (int64_t)m->A * m->N;
    const size_t es = nq_dtype_size(m->dtype);
    if (m->kind == NQ_NDM) return ((size_t)4 * MN * es + 15) / 16 * 16 + (size_t)2 * AN * 2 * es;
    // RBM / RBMSplit: [exp tables 2 nmat MN][q tables 2 nmat MN][wsum nmat N]  (nq_machine_qtab)
    const size_t nmat = m->kind == NQ_RBMSPLIT ? 2 : 1;
    return ((size_t)4 * nmat * MN + nmat * m->N) * es;
}
const void* nq_machine_qtab(nq_machine_t m) {
    const size_t nmat = m->kind == NQ_RBMSPLIT ? 2 : 1;
    return (const char*)m->etab + (size_t)2 * nmat * m->M * m->N * nq_dtype_size(m->dtype);
}

template <typename E>
static int build_tables_rbm(nq_machine_t m) {
    typedef typename elem_traits<E>::real T;
    nq_ctx_t ctx = m->ctx;
    const int64_t MN = (int64_t)m->M * m->N;
    const int nmat = m->kind == NQ_RBMSPLIT ? 2 : 1;
    const int64_t n = nmat * MN;       // Wr and Wc are contiguous in the parameter vector
    const E* W = (const E*)m->params + (m->kind == NQ_RBMSPLIT ? 2 * m->N : m->N) + m->M;
    T c = m->hilb == NQ_SPIN ? T(2) : T(1);
    NQ_LAUNCH(ctx, etab_kernel<E>, (unsigned)((n + 255) / 256), 256, 0, W, (E*)m->etab, n, c);
    const T g = m->act == NQ_LOGCOSH ? T(2) : T(1);
    NQ_LAUNCH(ctx, qtab_kernel<E>, (unsigned)((n + 255) / 256), 256, 0, W, (E*)m->etab + 2 * n, n, c * g);
    NQ_LAUNCH(ctx, wsum_kernel<E>, (unsigned)((nmat * m->N + 127) / 128), 128, 0, W, (E*)m->etab + 4 * n, m->M, nmat * m->N);
    return NQ_OK;
}

template <typename T>
static int build_tables_ndm(nq_machine_t m) {
    nq_ctx_t ctx = m->ctx;
    const int64_t N = m->N, M = m->M, A = m->A, MN = M * N, AN = A * N;
    const T* par = (const T*)m->params;
    const int64_t o_wmu = N + M, o_umu = o_wmu + MN, o_wlam = o_umu + AN + N + M + A, o_ulam = o_wlam + MN;
    T c = m->hilb == NQ_SPIN ? T(2) : T(1);
    T* tr = (T*)m->etab;
    // [sign][lay][MN]: write lam and mu with sign-major layout through two launches of n = MN each
    unsigned g = (unsigned)((MN + 255) / 256);
    // etab_kernel writes out[i] (sign 0) and out[n+i] (sign 1) with n = stride between signs = 2 MN
    NQ_LAUNCH(ctx, etab_strided_kernel<T>, g, 256, 0, par + o_wlam, tr, MN, (int64_t)2 * MN, c);
    NQ_LAUNCH(ctx, etab_strided_kernel<T>, g, 256, 0, par + o_wmu, tr + MN, MN, (int64_t)2 * MN, c);
    cx<T>* tc = (cx<T>*)((char*)m->etab + ((size_t)4 * MN * sizeof(T) + 15) / 16 * 16);
    NQ_LAUNCH(ctx, etab_pi_kernel<T>, (unsigned)((AN + 255) / 256), 256, 0, par + o_ulam, par + o_umu, tc, AN, c);
    return NQ_OK;
}

int nq_machine_ensure_tables(nq_machine_t m) {
    if (m->etab_valid) return NQ_OK;
    nq_ctx_t ctx = m->ctx;
    if (!m->etab) {
        if (cudaMalloc(&m->etab, etab_bytes(m)) != cudaSuccess) { cudaGetLastError(); return nq_fail(ctx, NQ_ERR_ALLOC, "ratio table allocation failed"); }
    }
    int s;
    if (m->kind == NQ_NDM) s = m->dtype == NQ_F64 ? build_tables_ndm<double>(m) : build_tables_ndm<float>(m);
    else switch (m->dtype) {
        case NQ_F32: s = build_tables_rbm<float>(m); break;
        case NQ_F64: s = build_tables_rbm<double>(m); break;
        case NQ_C64: s = build_tables_rbm<cxf>(m); break;
        default: s = build_tables_rbm<cxd>(m); break;
    }
    if (s == NQ_OK) m->etab_valid = true;
    return s;
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" int nq_machine_create(nq_ctx_t ctx, nq_machine_kind kind, nq_hilbert h, int N, int M, int A,
                                 nq_activation act, nq_dtype dtype, nq_machine_t* out) {
    if (!ctx || !out) return NQ_ERR_ARG;
    *out = nullptr;
    if (N <= 0 || M <= 0 || A < 0) return nq_fail(ctx, NQ_ERR_SHAPE, "N, M must be positive");
    if (kind == NQ_NDM && (nq_dtype_is_complex(dtype) || A <= 0))
        return nq_fail(ctx, NQ_ERR_ARG, "NDM takes real parameters and A > 0 ancillas");
    if (kind != NQ_NDM && A != 0) return nq_fail(ctx, NQ_ERR_ARG, "A must be 0 for RBM/RBMSplit");
    if (kind != NQ_RBM && kind != NQ_RBMSPLIT && kind != NQ_NDM) return nq_fail(ctx, NQ_ERR_ARG, "unknown machine kind");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    nq_machine_t m = new nq_machine_s();
    m->ctx = ctx; m->kind = kind; m->hilb = h; m->N = N; m->M = M; m->A = A;
    m->act = kind == NQ_RBMSPLIT ? NQ_SOFTPLUS : act;
    m->dtype = dtype;
    m->etab = nullptr; m->etab_valid = false;
    m->out_dtype = kind == NQ_NDM ? nq_complex_of(dtype) : dtype;
    int64_t MN = (int64_t)M * N;
    m->P = kind == NQ_RBM ? N + M + MN : kind == NQ_RBMSPLIT ? 2 * N + M + 2 * MN
                                                            : 2 * N + 2 * M + A + 2 * MN + 2 * (int64_t)A * N;
    if (cudaMalloc(&m->params, (size_t)m->P * nq_dtype_size(dtype)) != cudaSuccess) {
        cudaGetLastError();
        delete m;
        return nq_fail(ctx, NQ_ERR_ALLOC, "parameter allocation failed");
    }
    cudaMemsetAsync(m->params, 0, (size_t)m->P * nq_dtype_size(dtype), ctx->stream);
    *out = m;
    return NQ_OK;
}

extern "C" int nq_machine_destroy(nq_machine_t m) {
    if (!m) return NQ_ERR_ARG;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->params);
    cudaFree(m->etab);
    delete m;
    return NQ_OK;
}

extern "C" int nq_machine_nparams(nq_machine_t m, int64_t* P) {
    if (!m || !P) return NQ_ERR_ARG;
    *P = m->P;
    return NQ_OK;
}

extern "C" int nq_machine_out_dtype(nq_machine_t m, nq_dtype* out) {
    if (!m || !out) return NQ_ERR_ARG;
    *out = m->out_dtype;
    return NQ_OK;
}

extern "C" int nq_machine_set_params(nq_machine_t m, const void* params, int64_t P) {
    if (!m || !params) return NQ_ERR_ARG;
    if (P != m->P) return nq_fail(m->ctx, NQ_ERR_SHAPE, "expected %lld parameters, got %lld", (long long)m->P, (long long)P);
    NQ_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
    NQ_CUDA(m->ctx, cudaMemcpyAsync(m->params, params, (size_t)P * nq_dtype_size(m->dtype), cudaMemcpyDefault, m->ctx->stream));
    m->etab_valid = false;
    if (!nq_is_device_ptr(params)) NQ_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
    return NQ_OK;
}

extern "C" int nq_machine_get_params(nq_machine_t m, void* params, int64_t P) {
    if (!m || !params) return NQ_ERR_ARG;
    if (P != m->P) return nq_fail(m->ctx, NQ_ERR_SHAPE, "expected %lld parameters, got %lld", (long long)m->P, (long long)P);
    NQ_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
    NQ_CUDA(m->ctx, cudaMemcpyAsync(params, m->params, (size_t)P * nq_dtype_size(m->dtype), cudaMemcpyDefault, m->ctx->stream));
    if (!nq_is_device_ptr(params)) NQ_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
    return NQ_OK;
}

static int check_states(nq_machine_t m, const void* srow, const void* scol, int64_t B) {
    if (!m || !srow || B < 0) return NQ_ERR_ARG;
    if (m->doubled() && !scol) return nq_fail(m->ctx, NQ_ERR_ARG, "density-matrix machine needs sigma and sigma'");
    if (!m->doubled() && scol) return nq_fail(m->ctx, NQ_ERR_ARG, "ket machine takes a single configuration array");
    return NQ_OK;
}

// stage + pack float configurations; returns device packed pointers
int nq_stage_pack(nq_machine_t m, NqStage& st, const void* srow, const void* scol, nq_dtype sdtype, int64_t B,
                  const uint64_t** prow, const uint64_t** pcol) {
    nq_ctx_t ctx = m->ctx;
    size_t fbytes = (size_t)B * m->N * nq_dtype_size(sdtype);
    size_t pbytes = (size_t)B * nq_words(m->N) * 8;
    const void* dr = st.in(SL_IN0, srow, fbytes);
    const void* dc = scol ? st.in(SL_IN1, scol, fbytes) : nullptr;
    if (st.status != NQ_OK) return st.status;
    uint64_t* pr = (uint64_t*)nq_scratch(ctx, SL_PROW, pbytes ? pbytes : 8);
    uint64_t* pc = scol ? (uint64_t*)nq_scratch(ctx, SL_PCOL, pbytes ? pbytes : 8) : nullptr;
    if (!pr || (scol && !pc)) return NQ_ERR_ALLOC;
    NQ_CHECK(nq_pack_device(ctx, m->hilb, m->N, B, dr, sdtype, pr));
    if (scol) NQ_CHECK(nq_pack_device(ctx, m->hilb, m->N, B, dc, sdtype, pc));
    *prow = pr;
    *pcol = pc;
    return NQ_OK;
}

extern "C" int nq_logpsi_grad(nq_machine_t m, const void* srow, const void* scol, nq_dtype sdtype, int64_t B,
                              void* out, void* O, int64_t ldO) {
    NQ_CHECK(check_states(m, srow, scol, B));
    if (!out) return NQ_ERR_ARG;
    if (O && ldO < m->P) return nq_fail(m->ctx, NQ_ERR_SHAPE, "ldO < P");
    nq_ctx_t ctx = m->ctx;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const uint64_t *pr, *pc;
    NQ_CHECK(nq_stage_pack(m, st, srow, scol, sdtype, B, &pr, &pc));
    size_t es = nq_dtype_size(m->out_dtype);
    void* dout = st.out(SL_OUT0, out, (size_t)B * es);
    void* dO = O ? st.out2d(SL_OUT1, O, (size_t)m->P * es, (size_t)ldO * es, (size_t)B) : nullptr;
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(nq_machine_eval_device(m, pr, pc, B, dout, dO, ldO));
    return st.finish();
}

extern "C" int nq_logpsi(nq_machine_t m, const void* srow, const void* scol, nq_dtype sdtype, int64_t B, void* out) {
    return nq_logpsi_grad(m, srow, scol, sdtype, B, out, nullptr, 0);
}

extern "C" int nq_log_prob(nq_machine_t m, const void* srow, const void* scol, nq_dtype sdtype, int64_t B, void* out) {
    NQ_CHECK(check_states(m, srow, scol, B));
    if (!out) return NQ_ERR_ARG;
    nq_ctx_t ctx = m->ctx;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const uint64_t *pr, *pc;
    NQ_CHECK(nq_stage_pack(m, st, srow, scol, sdtype, B, &pr, &pc));
    void* lp = nq_scratch(ctx, SL_LOGPSI, (size_t)(B ? B : 1) * nq_dtype_size(m->out_dtype));
    if (!lp) return NQ_ERR_ALLOC;
    void* dout = st.out(SL_OUT0, out, (size_t)B * nq_dtype_size(nq_real_of(m->out_dtype)));
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(nq_machine_eval_device(m, pr, pc, B, lp, nullptr, 0));
    if (B > 0) {
        unsigned grid = (unsigned)((B + 255) / 256);
        switch (m->out_dtype) {
            case NQ_F32: NQ_LAUNCH(ctx, log_prob_kernel<float>, grid, 256, 0, (const float*)lp, (float*)dout, B); break;
            case NQ_F64: NQ_LAUNCH(ctx, log_prob_kernel<double>, grid, 256, 0, (const double*)lp, (double*)dout, B); break;
            case NQ_C64: NQ_LAUNCH(ctx, log_prob_kernel<cxf>, grid, 256, 0, (const cxf*)lp, (float*)dout, B); break;
            default: NQ_LAUNCH(ctx, log_prob_kernel<cxd>, grid, 256, 0, (const cxd*)lp, (double*)dout, B); break;
        }
    }
    return st.finish();
}

extern "C" int nq_logpsi_grad_packed(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B,
                                     void* out, void* O, int64_t ldO) {
    if (!m || !prow || !out || B < 0) return NQ_ERR_ARG;
    if (m->doubled() != (pcol != nullptr)) return nq_fail(m->ctx, NQ_ERR_ARG, "row/col configuration mismatch");
    if (O && ldO < m->P) return nq_fail(m->ctx, NQ_ERR_SHAPE, "ldO < P");
    nq_ctx_t ctx = m->ctx;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    size_t pbytes = (size_t)B * nq_words(m->N) * 8, es = nq_dtype_size(m->out_dtype);
    const uint64_t* pr = (const uint64_t*)st.in(SL_PROW, prow, pbytes);
    const uint64_t* pc = pcol ? (const uint64_t*)st.in(SL_PCOL, pcol, pbytes) : nullptr;
    void* dout = st.out(SL_OUT0, out, (size_t)B * es);
    void* dO = O ? st.out2d(SL_OUT1, O, (size_t)m->P * es, (size_t)ldO * es, (size_t)B) : nullptr;
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(nq_machine_eval_device(m, pr, pc, B, dout, dO, ldO));
    return st.finish();
}

extern "C" int nq_logpsi_packed(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B, void* out) {
    return nq_logpsi_grad_packed(m, prow, pcol, B, out, nullptr, 0);
}
