// nq_operator.cu -- K5: operator connection tables on the device, connection enumeration and the
// local estimators (E_loc; L_loc and grad L_loc) evaluated with incremental (flip-delta) updates.
//
// ref: Operators/Operators/KLocalOperator.jl:54-112,183-199; KLocalOperatorSum.jl:65-72;
//      KLocalOperatorTensor.jl:129-157; KLocalLiouvillian.jl:46-52;
//      IterativeInterface/Accumulators/AccumulatorObsScalar.jl:52-137, AccumulatorObsGrad.jl:39-127.
//
// Estimator kernels: one CTA per configuration.  Each thread owns hidden units ("items") of the
// machine; the pre-activations theta of the sampled configuration are computed once, every connected
// configuration eta_c differs by <= 4+4 flipped sites so theta(eta_c) = theta + sum_j W[:,j] dv_j.
// All diagonal connections (eta = sigma, ratio 1) are folded into one coefficient.  For the gradient
// estimator the rows sum_c w_c grad log rho(eta_c) are assembled in shared memory from
//   (sum_c w_c d^c_k) v_j  +  corrections on the columns j flipped by c,
// then streamed out with coalesced stores.
#include "nq_internal.cuh"

namespace {

struct OpDev {
    int n_terms, super;
    const int32_t* part_nsites;
    const int32_t* part_site_ptr;
    const int32_t* part_sites;
    const int64_t* part_row0;
    const int64_t* row_ptr;
    const double* entry_mel;
    const uint32_t* entry_flip;
    const int32_t* term_left;
    const int32_t* term_right;
};

OpDev op_dev(nq_operator_t op) {
    OpDev d;
    d.n_terms = op->n_terms; d.super = op->space == NQ_SUPER;
    d.part_nsites = op->part_nsites; d.part_site_ptr = op->part_site_ptr; d.part_sites = op->part_sites;
    d.part_row0 = op->part_row0; d.row_ptr = op->row_ptr; d.entry_mel = op->entry_mel;
    d.entry_flip = op->entry_flip; d.term_left = op->term_left; d.term_right = op->term_right;
    return d;
}

// entry range of the local row selected by `bits` in part p
__device__ __forceinline__ void part_row_range(const OpDev& op, int p, const uint64_t* bits, int64_t& e0, int64_t& e1) {
    int k = op.part_nsites[p];
    const int32_t* s = op.part_sites + op.part_site_ptr[p];
    int r = 0;
    for (int i = 0; i < k; i++) r |= get_bit(bits, s[i]) << i;
    int64_t row = op.part_row0[p] + r;
    e0 = op.row_ptr[row];
    e1 = op.row_ptr[row + 1];
}

// Visit the connections of term t in reference order.  f(mel_re, mel_im, partL, flipL, partR, flipR)
template <typename F>
__device__ __forceinline__ void visit_term(const OpDev& op, int t, const uint64_t* rb, const uint64_t* cb, F&& f) {
    int L = op.term_left[t], R = op.super ? op.term_right[t] : -1;
    if (L >= 0 && R < 0) {
        int64_t e0, e1;
        part_row_range(op, L, rb, e0, e1);
        for (int64_t e = e0; e < e1; e++) f(op.entry_mel[2 * e], op.entry_mel[2 * e + 1], L, op.entry_flip[e], -1, 0u);
    } else if (L < 0 && R >= 0) {
        int64_t e0, e1;
        part_row_range(op, R, cb, e0, e1);
        for (int64_t e = e0; e < e1; e++) f(op.entry_mel[2 * e], op.entry_mel[2 * e + 1], -1, 0u, R, op.entry_flip[e]);
    } else if (L >= 0 && R >= 0) {
        int64_t l0, l1, r0, r1;
        part_row_range(op, L, rb, l0, l1);
        part_row_range(op, R, cb, r0, r1);
        for (int64_t el = l0; el < l1; el++) {
            double ar = op.entry_mel[2 * el], ai = op.entry_mel[2 * el + 1];
            uint32_t fl = op.entry_flip[el];
            for (int64_t er = r0; er < r1; er++) {
                double br = op.entry_mel[2 * er], bi = op.entry_mel[2 * er + 1];
                // no FMA contraction: the product must carry the same bits as the host's complex multiply
                f(__dsub_rn(__dmul_rn(ar, br), __dmul_rn(ai, bi)), __dadd_rn(__dmul_rn(ar, bi), __dmul_rn(ai, br)), L, fl, R,
                  op.entry_flip[er]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// debug / integer-parity path: row_valdiff! over a batch, one thread per configuration
// ---------------------------------------------------------------------------------------
__global__ void connections_kernel(OpDev op, const uint64_t* __restrict__ prow, const uint64_t* __restrict__ pcol,
                                   int64_t B, int W64, int64_t max_conn, int32_t* __restrict__ counts,
                                   double* __restrict__ mels, uint64_t* __restrict__ frow, uint64_t* __restrict__ fcol) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint64_t* rb = prow + b * W64;
    const uint64_t* cb = pcol ? pcol + b * W64 : nullptr;
    int n = 0;
    for (int t = 0; t < op.n_terms; t++) {
        visit_term(op, t, rb, cb, [&](double mr, double mi, int L, uint32_t fl, int R, uint32_t fr) {
            if (n < max_conn) {
                int64_t c = b * max_conn + n;
                mels[2 * c] = mr; mels[2 * c + 1] = mi;
                for (int w = 0; w < W64; w++) { frow[c * W64 + w] = 0; if (fcol) fcol[c * W64 + w] = 0; }
                if (L >= 0) {
                    const int32_t* s = op.part_sites + op.part_site_ptr[L];
                    for (int i = 0; fl >> i; i++) if ((fl >> i) & 1u) frow[c * W64 + (s[i] >> 6)] |= 1ull << (s[i] & 63);
                }
                if (R >= 0 && fcol) {
                    const int32_t* s = op.part_sites + op.part_site_ptr[R];
                    for (int i = 0; fr >> i; i++) if ((fr >> i) & 1u) fcol[c * W64 + (s[i] >> 6)] |= 1ull << (s[i] & 63);
                }
            }
            n++;
        });
    }
    counts[b] = n;
}

// ---------------------------------------------------------------------------------------
// connection list of one configuration in shared memory (non-diagonal, non-zero mel only);
// diagonal mels are summed into *wdiag.  Deterministic order (term order).  All threads call.
// ---------------------------------------------------------------------------------------
constexpr int MAXF = 4;   // max flipped sites per side of one connection

template <typename T>
struct ConnList {
    cx<T>* mel;        // [cap]
    uint8_t* nflip;    // [cap] nr | nc << 4
    uint8_t* sites;    // [cap][2*MAXF] row sites then col sites
    int* counts;       // [n_terms + 1]
    double* red;       // [2 * 32] block-reduction scratch
};

template <typename T>
__device__ int build_conn_list(const OpDev& op, const uint64_t* rb, const uint64_t* cb, ConnList<T>& cl,
                               cx<T>* wdiag_out) {
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = NT >> 5;
    double dr = 0.0, di = 0.0;
    for (int t = tid; t < op.n_terms; t += NT) {
        int cnt = 0;
        visit_term(op, t, rb, cb, [&](double mr, double mi, int, uint32_t fl, int, uint32_t fr) {
            if (mr == 0.0 && mi == 0.0) return;
            if (fl == 0u && fr == 0u) { dr += mr; di += mi; } else cnt++;
        });
        cl.counts[t] = cnt;
    }
    dr = warp_sum(dr); di = warp_sum(di);
    if (lane == 0) { cl.red[2 * warp] = dr; cl.red[2 * warp + 1] = di; }
    __syncthreads();
    // exclusive scan of counts by warp 0 (chunked)
    if (warp == 0) {
        int n = op.n_terms;
        int chunk = (n + 31) / 32;
        int lo = lane * chunk, hi = min(n, lo + chunk);
        int s = 0;
        for (int t = lo; t < hi; t++) s += cl.counts[t];
        int incl = s;
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int run = incl - s;
        for (int t = lo; t < hi; t++) { int c = cl.counts[t]; cl.counts[t] = run; run += c; }
        if (lane == 31) cl.counts[n] = incl;
    }
    __syncthreads();
    double sr = 0.0, si = 0.0;
    for (int w = 0; w < nw; w++) { sr += cl.red[2 * w]; si += cl.red[2 * w + 1]; }
    *wdiag_out = cx<T>((T)sr, (T)si);
    for (int t = tid; t < op.n_terms; t += NT) {
        int pos = cl.counts[t];
        visit_term(op, t, rb, cb, [&](double mr, double mi, int L, uint32_t fl, int R, uint32_t fr) {
            if (mr == 0.0 && mi == 0.0) return;
            if (fl == 0u && fr == 0u) return;
            cl.mel[pos] = cx<T>((T)mr, (T)mi);
            int nr = 0, nc = 0;
            uint8_t* s = cl.sites + pos * 2 * MAXF;
            if (L >= 0) {
                const int32_t* ps = op.part_sites + op.part_site_ptr[L];
                for (int i = 0; fl >> i; i++) if ((fl >> i) & 1u) s[nr++] = (uint8_t)ps[i];
            }
            if (R >= 0) {
                const int32_t* ps = op.part_sites + op.part_site_ptr[R];
                for (int i = 0; fr >> i; i++) if ((fr >> i) & 1u) s[MAXF + nc++] = (uint8_t)ps[i];
            }
            cl.nflip[pos] = (uint8_t)(nr | (nc << 4));
            pos++;
        });
    }
    __syncthreads();
    return cl.counts[op.n_terms];
}

template <typename T>
__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// carve shared memory; returns bytes used
template <typename T>
__host__ __device__ inline size_t conn_list_bytes(int cap, int n_terms) {
    return align16<T>((size_t)cap * sizeof(cx<T>)) + align16<T>((size_t)cap) + align16<T>((size_t)cap * 2 * MAXF) +
           align16<T>((size_t)(n_terms + 1) * sizeof(int)) + 64 * sizeof(double);
}
template <typename T>
__device__ inline unsigned char* conn_list_carve(unsigned char* p, int cap, int n_terms, ConnList<T>& cl) {
    cl.mel = (cx<T>*)p; p += align16<T>((size_t)cap * sizeof(cx<T>));
    cl.nflip = (uint8_t*)p; p += align16<T>((size_t)cap);
    cl.sites = (uint8_t*)p; p += align16<T>((size_t)cap * 2 * MAXF);
    cl.counts = (int*)p; p += align16<T>((size_t)(n_terms + 1) * sizeof(int));
    cl.red = (double*)p; p += 64 * sizeof(double);
    return p;
}

constexpr int CB = 4;   // connections processed per batch (ILP + one block reduction per batch)

// block-wide sum of CB complex numbers; result broadcast to every thread (2 barriers)
template <typename T>
__device__ __forceinline__ void block_sum_cb(cx<T> (&acc)[CB], cx<T>* red /* [32][CB] */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < CB; i++) {
        cx<T> v = warp_sum(acc[i]);
        if (lane == 0) red[warp * CB + i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CB; i++) {
        cx<T> v = red[i];
        for (int w = 1; w < nw; w++) v += red[w * CB + i];
        acc[i] = v;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// RBM (ket, scalar only) and RBMSplit (super; scalar and gradient)
// ---------------------------------------------------------------------------------------
template <typename E, int ACT, bool DOUBLED, bool GRAD>
__global__ void local_rbm_kernel(OpDev op, const E* __restrict__ par, const uint64_t* __restrict__ prow,
                                 const uint64_t* __restrict__ pcol, int64_t B, int N, int M, int hilb,
                                 int cap, cx<typename elem_traits<E>::real>* __restrict__ out_loc,
                                 cx<typename elem_traits<E>::real>* __restrict__ out_g, int64_t ld) {
    typedef typename elem_traits<E>::real T;
    typedef cx<T> C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int W64 = (N + 63) >> 6;
    const int64_t b = blockIdx.x;
    const uint64_t* rb = prow + b * W64;
    const uint64_t* cb = DOUBLED ? pcol + b * W64 : nullptr;

    ConnList<T> cl;
    unsigned char* p = conn_list_carve<T>(smem_raw, cap, op.n_terms, cl);
    C* red = (C*)p; p += 32 * CB * sizeof(C);
    E* th = (E*)p; p += align16<T>((size_t)M * sizeof(E));       // theta_k
    E* fk = (E*)p; p += align16<T>((size_t)M * sizeof(E));       // f(theta_k)
    E* dc = (E*)p; p += align16<T>((size_t)CB * M * sizeof(E));  // d^c_k of the current batch
    C* Ak = (C*)p; p += align16<T>((size_t)(GRAD ? M : 0) * sizeof(C));  // sum_c w_c d^c_k
    C* vis = (C*)p; p += align16<T>((size_t)(GRAD ? 2 * N : 0) * sizeof(C)); // sum_{c flips j} w_c dv_j
    C* g = (C*)p;                                                // staged output row [P]

    const int64_t off_b = DOUBLED ? 2 * N : N;
    const int64_t off_Wr = off_b + M, off_Wc = off_Wr + (int64_t)M * N;
    const int64_t P = DOUBLED ? off_Wc + (int64_t)M * N : off_Wc;
    const E* __restrict__ Wr = par + off_Wr;
    const E* __restrict__ Wc = par + off_Wc;

    C wdiag;
    const int nconn = build_conn_list<T>(op, rb, cb, cl, &wdiag);

    // base pre-activations
    for (int k = tid; k < M; k += NT) {
        E t = par[off_b + k];
        for (int j = 0; j < N; j++) {
            t += rscale(digit_value<T>(hilb, get_bit(rb, j)), Wr[k + (int64_t)M * j]);
            if (DOUBLED) t += rscale(digit_value<T>(hilb, get_bit(cb, j)), Wc[k + (int64_t)M * j]);
        }
        E f, d;
        act_eval<ACT>(t, f, d);
        th[k] = t; fk[k] = f;
        if (GRAD) Ak[k] = wdiag * to_cx(d);
    }
    if (GRAD) {
        for (int64_t i = tid; i < P; i += NT) g[i] = C(T(0), T(0));
        for (int i = tid; i < 2 * N; i += NT) vis[i] = C(T(0), T(0));
    }
    __syncthreads();

    C wtot = wdiag;
    for (int c0 = 0; c0 < nconn; c0 += CB) {
        C acc[CB];
#pragma unroll
        for (int i = 0; i < CB; i++) acc[i] = C(T(0), T(0));
        for (int k = tid; k < M; k += NT) {
            E t0 = th[k], f0 = fk[k];
#pragma unroll
            for (int i = 0; i < CB; i++) {
                int c = c0 + i;
                if (c < nconn) {
                    E t = t0;
                    int nf = cl.nflip[c];
                    const uint8_t* s = cl.sites + c * 2 * MAXF;
                    for (int q = 0; q < (nf & 15); q++) {
                        int j = s[q];
                        t += rscale(flip_delta<T>(hilb, get_bit(rb, j)), Wr[k + (int64_t)M * j]);
                    }
                    if (DOUBLED) for (int q = 0; q < (nf >> 4); q++) {
                        int j = s[MAXF + q];
                        t += rscale(flip_delta<T>(hilb, get_bit(cb, j)), Wc[k + (int64_t)M * j]);
                    }
                    E f, d;
                    act_eval<ACT>(t, f, d);
                    acc[i] += to_cx(f - f0);
                    if (GRAD) dc[i * M + k] = d;
                }
            }
        }
        if (tid == 0) {   // visible-bias part of log psi(eta) - log psi(sigma)
#pragma unroll
            for (int i = 0; i < CB; i++) {
                int c = c0 + i;
                if (c < nconn) {
                    int nf = cl.nflip[c];
                    const uint8_t* s = cl.sites + c * 2 * MAXF;
                    E lin = make_zero<E>();
                    for (int q = 0; q < (nf & 15); q++) { int j = s[q]; lin += rscale(flip_delta<T>(hilb, get_bit(rb, j)), par[j]); }
                    if (DOUBLED) for (int q = 0; q < (nf >> 4); q++) { int j = s[MAXF + q]; lin += rscale(flip_delta<T>(hilb, get_bit(cb, j)), par[N + j]); }
                    acc[i] += to_cx(lin);
                }
            }
        }
        block_sum_cb<T>(acc, red);
        C w[CB];
#pragma unroll
        for (int i = 0; i < CB; i++) {
            w[i] = C(T(0), T(0));
            if (c0 + i < nconn) { w[i] = cl.mel[c0 + i] * cx_exp(acc[i]); wtot += w[i]; }
        }
        if (GRAD) {
            for (int k = tid; k < M; k += NT) {
                C a = Ak[k];
#pragma unroll
                for (int i = 0; i < CB; i++) {
                    int c = c0 + i;
                    if (c < nconn) {
                        C wd = w[i] * to_cx(dc[i * M + k]);
                        a += wd;
                        int nf = cl.nflip[c];
                        const uint8_t* s = cl.sites + c * 2 * MAXF;
                        for (int q = 0; q < (nf & 15); q++) { int j = s[q]; g[off_Wr + (int64_t)M * j + k] += rscale(flip_delta<T>(hilb, get_bit(rb, j)), wd); }
                        for (int q = 0; q < (nf >> 4); q++) { int j = s[MAXF + q]; g[off_Wc + (int64_t)M * j + k] += rscale(flip_delta<T>(hilb, get_bit(cb, j)), wd); }
                    }
                }
                Ak[k] = a;
            }
            if (tid == 0) {
#pragma unroll
                for (int i = 0; i < CB; i++) {
                    int c = c0 + i;
                    if (c < nconn) {
                        int nf = cl.nflip[c];
                        const uint8_t* s = cl.sites + c * 2 * MAXF;
                        for (int q = 0; q < (nf & 15); q++) { int j = s[q]; vis[j] += rscale(flip_delta<T>(hilb, get_bit(rb, j)), w[i]); }
                        for (int q = 0; q < (nf >> 4); q++) { int j = s[MAXF + q]; vis[N + j] += rscale(flip_delta<T>(hilb, get_bit(cb, j)), w[i]); }
                    }
                }
            }
            __syncthreads();   // dc is rewritten by the next batch
        }
    }
    if (tid == 0) out_loc[b] = wtot;
    if (GRAD) {
        // base terms: (sum_c w_c d^c_k) v_j on every column, visible rows
        for (int k = tid; k < M; k += NT) {
            C a = Ak[k];
            g[off_b + k] = a;
            for (int j = 0; j < N; j++) {
                g[off_Wr + (int64_t)M * j + k] += rscale(digit_value<T>(hilb, get_bit(rb, j)), a);
                g[off_Wc + (int64_t)M * j + k] += rscale(digit_value<T>(hilb, get_bit(cb, j)), a);
            }
        }
        __syncthreads();
        for (int j = tid; j < N; j += NT) {
            g[j] = rscale(digit_value<T>(hilb, get_bit(rb, j)), wtot) + vis[j];
            g[N + j] = rscale(digit_value<T>(hilb, get_bit(cb, j)), wtot) + vis[N + j];
        }
        __syncthreads();
        C* og = out_g + b * ld;
        for (int64_t i = tid; i < P; i += NT) og[i] = g[i];
    }
}

// ---------------------------------------------------------------------------------------
// NDM (super).  items: [0,M) lambda units, [M,2M) mu units (real, each on sigma and sigma'),
// [2M, 2M+A) ancilla units (complex pre-activation)
// ---------------------------------------------------------------------------------------
template <typename T, int ACT, bool GRAD>
__global__ void local_ndm_kernel(OpDev op, const T* __restrict__ par, const uint64_t* __restrict__ prow,
                                 const uint64_t* __restrict__ pcol, int64_t B, int N, int M, int A, int hilb,
                                 int cap, cx<T>* __restrict__ out_loc, cx<T>* __restrict__ out_g, int64_t ld) {
    typedef cx<T> C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int W64 = (N + 63) >> 6;
    const int64_t b = blockIdx.x;
    const uint64_t* rb = prow + b * W64;
    const uint64_t* cb = pcol + b * W64;
    const int R2 = 2 * M, NI = 2 * M + A;
    const T half = T(0.5);

    ConnList<T> cl;
    unsigned char* p = conn_list_carve<T>(smem_raw, cap, op.n_terms, cl);
    C* red = (C*)p; p += 32 * CB * sizeof(C);
    T* th = (T*)p; p += align16<T>((size_t)2 * R2 * sizeof(T));      // theta (sigma) [2M], theta' [2M]
    T* fk = (T*)p; p += align16<T>((size_t)2 * R2 * sizeof(T));
    C* pi0 = (C*)p; p += align16<T>((size_t)A * sizeof(C));           // Pi pre-activation
    C* fpi = (C*)p; p += align16<T>((size_t)A * sizeof(C));
    T* dcr = (T*)p; p += align16<T>((size_t)(GRAD ? CB * 2 * R2 : 0) * sizeof(T));  // d^c (sigma | sigma')
    C* dcp = (C*)p; p += align16<T>((size_t)(GRAD ? CB * A : 0) * sizeof(C));
    T* dbase = (T*)p; p += align16<T>((size_t)(GRAD ? 2 * R2 : 0) * sizeof(T));
    C* Ak = (C*)p; p += align16<T>((size_t)(GRAD ? 2 * R2 + A : 0) * sizeof(C));     // A (sigma)[2M], A'(sigma')[2M], A_pi[A]
    C* vis = (C*)p; p += align16<T>((size_t)(GRAD ? 2 * N : 0) * sizeof(C));
    C* g = (C*)p;

    const int64_t MN = (int64_t)M * N, AN = (int64_t)A * N;
    const int64_t o_bmu = 0, o_hmu = N, o_wmu = N + M, o_umu = o_wmu + MN, o_blam = o_umu + AN,
                  o_hlam = o_blam + N, o_dlam = o_hlam + M, o_wlam = o_dlam + A, o_ulam = o_wlam + MN;
    const int64_t P = o_ulam + AN;

    C wdiag;
    const int nconn = build_conn_list<T>(op, rb, cb, cl, &wdiag);

    for (int it = tid; it < NI; it += NT) {
        if (it < R2) {
            int lay = it >= M, k = it - lay * M;
            const T* __restrict__ w = par + (lay ? o_wmu : o_wlam) + k;
            T t = par[(lay ? o_hmu : o_hlam) + k], tp = t;
            for (int j = 0; j < N; j++) {
                T wv = w[(int64_t)M * j];
                t += wv * digit_value<T>(hilb, get_bit(rb, j));
                tp += wv * digit_value<T>(hilb, get_bit(cb, j));
            }
            T f, d, fp, dp;
            act_eval<ACT>(t, f, d);
            act_eval<ACT>(tp, fp, dp);
            th[it] = t; th[R2 + it] = tp; fk[it] = f; fk[R2 + it] = fp;
            if (GRAD) { dbase[it] = d; dbase[R2 + it] = dp; Ak[it] = rscale(d, wdiag); Ak[R2 + it] = rscale(dp, wdiag); }
        } else {
            int a = it - R2;
            T pr = par[o_dlam + a], pim = T(0);
            for (int j = 0; j < N; j++) {
                T x = digit_value<T>(hilb, get_bit(rb, j)), y = digit_value<T>(hilb, get_bit(cb, j));
                pr += half * par[o_ulam + a + (int64_t)A * j] * (x + y);
                pim += half * par[o_umu + a + (int64_t)A * j] * (x - y);
            }
            C f, d;
            act_eval<ACT>(C(pr, pim), f, d);
            pi0[a] = C(pr, pim); fpi[a] = f;
            if (GRAD) Ak[2 * R2 + a] = wdiag * d;
        }
    }
    if (GRAD) {
        for (int64_t i = tid; i < P; i += NT) g[i] = C(T(0), T(0));
        for (int i = tid; i < 2 * N; i += NT) vis[i] = C(T(0), T(0));
    }
    __syncthreads();

    C wtot = wdiag;
    for (int c0 = 0; c0 < nconn; c0 += CB) {
        C acc[CB];
#pragma unroll
        for (int i = 0; i < CB; i++) acc[i] = C(T(0), T(0));
        for (int it = tid; it < NI; it += NT) {
            if (it < R2) {
                int lay = it >= M, k = it - lay * M;
                const T* __restrict__ w = par + (lay ? o_wmu : o_wlam) + k;
                T t0 = th[it], tp0 = th[R2 + it], f0 = fk[it], fp0 = fk[R2 + it];
#pragma unroll
                for (int i = 0; i < CB; i++) {
                    int c = c0 + i;
                    if (c < nconn) {
                        int nf = cl.nflip[c];
                        const uint8_t* s = cl.sites + c * 2 * MAXF;
                        T df = T(0), dfp = T(0);
                        if (nf & 15) {
                            T t = t0;
                            for (int q = 0; q < (nf & 15); q++) { int j = s[q]; t += w[(int64_t)M * j] * flip_delta<T>(hilb, get_bit(rb, j)); }
                            T f, d;
                            act_eval<ACT>(t, f, d);
                            df = f - f0;
                            if (GRAD) dcr[i * 2 * R2 + it] = d;
                        } else if (GRAD) dcr[i * 2 * R2 + it] = dbase[it];
                        if (nf >> 4) {
                            T t = tp0;
                            for (int q = 0; q < (nf >> 4); q++) { int j = s[MAXF + q]; t += w[(int64_t)M * j] * flip_delta<T>(hilb, get_bit(cb, j)); }
                            T f, d;
                            act_eval<ACT>(t, f, d);
                            dfp = f - fp0;
                            if (GRAD) dcr[i * 2 * R2 + R2 + it] = d;
                        } else if (GRAD) dcr[i * 2 * R2 + R2 + it] = dbase[R2 + it];
                        if (lay == 0) acc[i].re += half * (df + dfp); else acc[i].im += half * (df - dfp);
                    }
                }
            } else {
                int a = it - R2;
                const T* __restrict__ ul = par + o_ulam + a;
                const T* __restrict__ um = par + o_umu + a;
                C p0 = pi0[a], f0 = fpi[a];
#pragma unroll
                for (int i = 0; i < CB; i++) {
                    int c = c0 + i;
                    if (c < nconn) {
                        int nf = cl.nflip[c];
                        const uint8_t* s = cl.sites + c * 2 * MAXF;
                        C t = p0;
                        for (int q = 0; q < (nf & 15); q++) {
                            int j = s[q]; T dv = half * flip_delta<T>(hilb, get_bit(rb, j));
                            t.re += ul[(int64_t)A * j] * dv; t.im += um[(int64_t)A * j] * dv;
                        }
                        for (int q = 0; q < (nf >> 4); q++) {
                            int j = s[MAXF + q]; T dv = half * flip_delta<T>(hilb, get_bit(cb, j));
                            t.re += ul[(int64_t)A * j] * dv; t.im -= um[(int64_t)A * j] * dv;
                        }
                        C f, d;
                        act_eval<ACT>(t, f, d);
                        acc[i] += f - f0;
                        if (GRAD) dcp[i * A + a] = d;
                    }
                }
            }
        }
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < CB; i++) {
                int c = c0 + i;
                if (c < nconn) {
                    int nf = cl.nflip[c];
                    const uint8_t* s = cl.sites + c * 2 * MAXF;
                    for (int q = 0; q < (nf & 15); q++) {
                        int j = s[q]; T dv = half * flip_delta<T>(hilb, get_bit(rb, j));
                        acc[i].re += par[o_blam + j] * dv; acc[i].im += par[o_bmu + j] * dv;
                    }
                    for (int q = 0; q < (nf >> 4); q++) {
                        int j = s[MAXF + q]; T dv = half * flip_delta<T>(hilb, get_bit(cb, j));
                        acc[i].re += par[o_blam + j] * dv; acc[i].im -= par[o_bmu + j] * dv;
                    }
                }
            }
        }
        block_sum_cb<T>(acc, red);
        C w[CB];
#pragma unroll
        for (int i = 0; i < CB; i++) {
            w[i] = C(T(0), T(0));
            if (c0 + i < nconn) { w[i] = cl.mel[c0 + i] * cx_exp(acc[i]); wtot += w[i]; }
        }
        if (GRAD) {
            for (int it = tid; it < NI; it += NT) {
                if (it < R2) {
                    int lay = it >= M, k = it - lay * M;
                    const int64_t ow = (lay ? o_wmu : o_wlam) + k;
                    C a = Ak[it], ap = Ak[R2 + it];
#pragma unroll
                    for (int i = 0; i < CB; i++) {
                        int c = c0 + i;
                        if (c < nconn) {
                            C wd = rscale(dcr[i * 2 * R2 + it], w[i]);
                            C wdp = rscale(dcr[i * 2 * R2 + R2 + it], w[i]);
                            a += wd; ap += wdp;
                            int nf = cl.nflip[c];
                            const uint8_t* s = cl.sites + c * 2 * MAXF;
                            // lambda rows hold S = sum w d^c v^c + w d'^c v'^c; mu rows hold D = ... - ...
                            for (int q = 0; q < (nf & 15); q++) { int j = s[q]; g[ow + (int64_t)M * j] += rscale(flip_delta<T>(hilb, get_bit(rb, j)), wd); }
                            for (int q = 0; q < (nf >> 4); q++) {
                                int j = s[MAXF + q];
                                T dv = flip_delta<T>(hilb, get_bit(cb, j));
                                g[ow + (int64_t)M * j] += rscale(lay ? -dv : dv, wdp);
                            }
                        }
                    }
                    Ak[it] = a; Ak[R2 + it] = ap;
                } else {
                    int a_ = it - R2;
                    C a = Ak[2 * R2 + a_];
#pragma unroll
                    for (int i = 0; i < CB; i++) {
                        int c = c0 + i;
                        if (c < nconn) {
                            C wd = w[i] * dcp[i * A + a_];
                            a += wd;
                            int nf = cl.nflip[c];
                            const uint8_t* s = cl.sites + c * 2 * MAXF;
                            // u_lam rows hold sum w dPi (v+v')^c, u_mu rows hold sum w dPi (v-v')^c
                            for (int q = 0; q < (nf & 15); q++) {
                                int j = s[q]; C x = rscale(flip_delta<T>(hilb, get_bit(rb, j)), wd);
                                g[o_ulam + a_ + (int64_t)A * j] += x; g[o_umu + a_ + (int64_t)A * j] += x;
                            }
                            for (int q = 0; q < (nf >> 4); q++) {
                                int j = s[MAXF + q]; C x = rscale(flip_delta<T>(hilb, get_bit(cb, j)), wd);
                                g[o_ulam + a_ + (int64_t)A * j] += x; g[o_umu + a_ + (int64_t)A * j] -= x;
                            }
                        }
                    }
                    Ak[2 * R2 + a_] = a;
                }
            }
            if (tid == 0) {
#pragma unroll
                for (int i = 0; i < CB; i++) {
                    int c = c0 + i;
                    if (c < nconn) {
                        int nf = cl.nflip[c];
                        const uint8_t* s = cl.sites + c * 2 * MAXF;
                        for (int q = 0; q < (nf & 15); q++) { int j = s[q]; vis[j] += rscale(flip_delta<T>(hilb, get_bit(rb, j)), w[i]); }
                        for (int q = 0; q < (nf >> 4); q++) { int j = s[MAXF + q]; vis[N + j] += rscale(flip_delta<T>(hilb, get_bit(cb, j)), w[i]); }
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) out_loc[b] = wtot;
    if (GRAD) {
        // finalise rows: add the base terms and apply the 1/2, i/2 prefactors
        for (int it = tid; it < NI; it += NT) {
            if (it < R2) {
                int lay = it >= M, k = it - lay * M;
                const int64_t ow = (lay ? o_wmu : o_wlam) + k;
                C a = Ak[it], ap = Ak[R2 + it];
                C hsum = lay ? (a - ap) : (a + ap);
                g[(lay ? o_hmu : o_hlam) + k] = lay ? C(-half * hsum.im, half * hsum.re) : rscale(half, hsum);
                for (int j = 0; j < N; j++) {
                    T x = digit_value<T>(hilb, get_bit(rb, j)), y = digit_value<T>(hilb, get_bit(cb, j));
                    C v = g[ow + (int64_t)M * j] + rscale(x, a) + rscale(lay ? -y : y, ap);
                    g[ow + (int64_t)M * j] = lay ? C(-half * v.im, half * v.re) : rscale(half, v);
                }
            } else {
                int a_ = it - R2;
                C a = Ak[2 * R2 + a_];
                g[o_dlam + a_] = a;
                for (int j = 0; j < N; j++) {
                    T x = digit_value<T>(hilb, get_bit(rb, j)), y = digit_value<T>(hilb, get_bit(cb, j));
                    C vl = g[o_ulam + a_ + (int64_t)A * j] + rscale(x + y, a);
                    C vm = g[o_umu + a_ + (int64_t)A * j] + rscale(x - y, a);
                    g[o_ulam + a_ + (int64_t)A * j] = rscale(half, vl);
                    g[o_umu + a_ + (int64_t)A * j] = C(-half * vm.im, half * vm.re);
                }
            }
        }
        for (int j = tid; j < N; j += NT) {
            T x = digit_value<T>(hilb, get_bit(rb, j)), y = digit_value<T>(hilb, get_bit(cb, j));
            C vl = rscale(x + y, wtot) + vis[j] + vis[N + j];
            C vm = rscale(x - y, wtot) + vis[j] - vis[N + j];
            g[o_blam + j] = rscale(half, vl);
            g[o_bmu + j] = C(-half * vm.im, half * vm.re);
        }
        __syncthreads();
        C* og = out_g + b * ld;
        for (int64_t i = tid; i < P; i += NT) og[i] = g[i];
    }
}

template <typename T>
size_t smem_common(int cap, int n_terms) { return conn_list_bytes<T>(cap, n_terms) + 32 * CB * sizeof(cx<T>); }

int round_threads(int items) {
    int nt = ((items + 31) / 32) * 32;
    if (nt < 64) nt = 64;
    if (nt > 512) nt = 512;
    return nt;
}

template <typename E, int ACT, bool DOUBLED>
int launch_local_rbm(nq_machine_t m, nq_operator_t op, const uint64_t* pr, const uint64_t* pc, int64_t B,
                     void* out_loc, void* out_g, int64_t ld) {
    typedef typename elem_traits<E>::real T;
    nq_ctx_t ctx = m->ctx;
    const bool grad = out_g != nullptr;
    int cap = (int)op->max_conn;
    size_t smem = smem_common<T>(cap, op->n_terms) + 2 * align16<T>((size_t)m->M * sizeof(E)) +
                  align16<T>((size_t)CB * m->M * sizeof(E));
    if (grad) smem += align16<T>((size_t)m->M * sizeof(cx<T>)) + align16<T>((size_t)2 * m->N * sizeof(cx<T>)) +
                      (size_t)m->P * sizeof(cx<T>);
    if (smem > ctx->smem_optin)
        return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "local estimator needs %zu B shared memory per CTA (limit %zu)", smem, ctx->smem_optin);
    int nt = round_threads(m->M);
    if (grad) {
        if (!DOUBLED) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "gradient estimator is defined for Liouvillians only");
        auto kern = local_rbm_kernel<E, ACT, DOUBLED, DOUBLED>;
        NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NQ_LAUNCH(ctx, kern, (unsigned)B, nt, smem, op_dev(op), (const E*)m->params, pr, pc, B, m->N, m->M, (int)m->hilb, cap, (cx<T>*)out_loc, (cx<T>*)out_g, ld);
    } else {
        auto kern = local_rbm_kernel<E, ACT, DOUBLED, false>;
        NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NQ_LAUNCH(ctx, kern, (unsigned)B, nt, smem, op_dev(op), (const E*)m->params, pr, pc, B, m->N, m->M, (int)m->hilb, cap, (cx<T>*)out_loc, (cx<T>*)nullptr, ld);
    }
    return NQ_OK;
}

template <typename T, int ACT>
int launch_local_ndm(nq_machine_t m, nq_operator_t op, const uint64_t* pr, const uint64_t* pc, int64_t B,
                     void* out_loc, void* out_g, int64_t ld) {
    nq_ctx_t ctx = m->ctx;
    const bool grad = out_g != nullptr;
    int cap = (int)op->max_conn;
    const int R2 = 2 * m->M, A = m->A;
    size_t smem = smem_common<T>(cap, op->n_terms) + 2 * align16<T>((size_t)2 * R2 * sizeof(T)) +
                  2 * align16<T>((size_t)A * sizeof(cx<T>));
    if (grad) smem += align16<T>((size_t)CB * 2 * R2 * sizeof(T)) + align16<T>((size_t)CB * A * sizeof(cx<T>)) +
                      align16<T>((size_t)2 * R2 * sizeof(T)) + align16<T>((size_t)(2 * R2 + A) * sizeof(cx<T>)) +
                      align16<T>((size_t)2 * m->N * sizeof(cx<T>)) + (size_t)m->P * sizeof(cx<T>);
    if (smem > ctx->smem_optin)
        return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "local estimator needs %zu B shared memory per CTA (limit %zu)", smem, ctx->smem_optin);
    int nt = round_threads(R2 + A);
    if (grad) {
        auto kern = local_ndm_kernel<T, ACT, true>;
        NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NQ_LAUNCH(ctx, kern, (unsigned)B, nt, smem, op_dev(op), (const T*)m->params, pr, pc, B, m->N, m->M, A, (int)m->hilb, cap, (cx<T>*)out_loc, (cx<T>*)out_g, ld);
    } else {
        auto kern = local_ndm_kernel<T, ACT, false>;
        NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NQ_LAUNCH(ctx, kern, (unsigned)B, nt, smem, op_dev(op), (const T*)m->params, pr, pc, B, m->N, m->M, A, (int)m->hilb, cap, (cx<T>*)out_loc, (cx<T>*)nullptr, ld);
    }
    return NQ_OK;
}

template <typename E>
int dispatch_local_rbm(nq_machine_t m, nq_operator_t op, const uint64_t* pr, const uint64_t* pc, int64_t B,
                       void* out_loc, void* out_g, int64_t ld) {
    if (m->kind == NQ_RBMSPLIT) return launch_local_rbm<E, NQ_SOFTPLUS, true>(m, op, pr, pc, B, out_loc, out_g, ld);
    if (m->act == NQ_SOFTPLUS) return launch_local_rbm<E, NQ_SOFTPLUS, false>(m, op, pr, pc, B, out_loc, out_g, ld);
    return launch_local_rbm<E, NQ_LOGCOSH, false>(m, op, pr, pc, B, out_loc, out_g, ld);
}

}  // namespace

int nq_local_device(nq_machine_t m, nq_operator_t op, const uint64_t* pr, const uint64_t* pc, int64_t B,
                    void* out_loc, void* out_g, int64_t ld) {
    nq_ctx_t ctx = m->ctx;
    if (op->ctx != ctx) return nq_fail(ctx, NQ_ERR_ARG, "machine and operator belong to different contexts");
    if (op->N != m->N) return nq_fail(ctx, NQ_ERR_SHAPE, "operator acts on %d sites, machine on %d", op->N, m->N);
    if ((op->space == NQ_SUPER) != m->doubled())
        return nq_fail(ctx, NQ_ERR_ARG, "ket operators pair with RBM, Liouvillians with RBMSplit/NDM");
    if (op->max_part_sites > MAXF) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "local terms on more than %d sites", MAXF);
    if (m->N > 256) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "estimator kernels index sites with 8 bits (N <= 256)");
    if (B == 0) return NQ_OK;
    if (m->kind == NQ_NDM) {
        if (m->dtype == NQ_F64)
            return m->act == NQ_SOFTPLUS ? launch_local_ndm<double, NQ_SOFTPLUS>(m, op, pr, pc, B, out_loc, out_g, ld)
                                         : launch_local_ndm<double, NQ_LOGCOSH>(m, op, pr, pc, B, out_loc, out_g, ld);
        return m->act == NQ_SOFTPLUS ? launch_local_ndm<float, NQ_SOFTPLUS>(m, op, pr, pc, B, out_loc, out_g, ld)
                                     : launch_local_ndm<float, NQ_LOGCOSH>(m, op, pr, pc, B, out_loc, out_g, ld);
    }
    switch (m->dtype) {
        case NQ_F32: return dispatch_local_rbm<float>(m, op, pr, pc, B, out_loc, out_g, ld);
        case NQ_F64: return dispatch_local_rbm<double>(m, op, pr, pc, B, out_loc, out_g, ld);
        case NQ_C64: return dispatch_local_rbm<cxf>(m, op, pr, pc, B, out_loc, out_g, ld);
        default: return dispatch_local_rbm<cxd>(m, op, pr, pc, B, out_loc, out_g, ld);
    }
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
template <typename X>
static int upload(nq_ctx_t ctx, X** dst, const X* src, size_t n) {
    *dst = nullptr;
    if (cudaMalloc((void**)dst, (n ? n : 1) * sizeof(X)) != cudaSuccess) { cudaGetLastError(); return nq_fail(ctx, NQ_ERR_ALLOC, "table allocation failed"); }
    if (n) NQ_CUDA(ctx, cudaMemcpy(*dst, src, n * sizeof(X), cudaMemcpyHostToDevice));
    return NQ_OK;
}

extern "C" int nq_operator_destroy(nq_operator_t op) {
    if (!op) return NQ_ERR_ARG;
    cudaSetDevice(op->ctx->device);
    cudaStreamSynchronize(op->ctx->stream);
    cudaFree(op->part_nsites); cudaFree(op->part_site_ptr); cudaFree(op->part_sites); cudaFree(op->part_row0);
    cudaFree(op->row_ptr); cudaFree(op->entry_mel); cudaFree(op->entry_flip); cudaFree(op->term_left); cudaFree(op->term_right);
    delete op;
    return NQ_OK;
}

extern "C" int nq_operator_create(nq_ctx_t ctx, nq_space space, int N, int n_parts, const int32_t* part_nsites,
                                  const int32_t* part_sites, const int64_t* row_ptr, const double* entry_mel,
                                  const uint32_t* entry_flip, int n_terms, const int32_t* term_left,
                                  const int32_t* term_right, nq_operator_t* out) {
    if (!ctx || !out) return NQ_ERR_ARG;
    *out = nullptr;
    if (N <= 0 || n_parts < 0 || n_terms < 0) return nq_fail(ctx, NQ_ERR_SHAPE, "negative table sizes");
    if (n_parts && (!part_nsites || !part_sites || !row_ptr || !entry_mel || !entry_flip)) return NQ_ERR_ARG;
    if (n_terms && !term_left) return NQ_ERR_ARG;
    if (space == NQ_SUPER && n_terms && !term_right) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int32_t> site_ptr(n_parts + 1, 0);
    std::vector<int64_t> row0(n_parts, 0);
    std::vector<int64_t> max_row(n_parts, 0);
    int64_t rows = 0;
    int maxk = 0;
    for (int p = 0; p < n_parts; p++) {
        int k = part_nsites[p];
        if (k < 0 || k > 16) return nq_fail(ctx, NQ_ERR_SHAPE, "part %d acts on %d sites", p, k);
        maxk = k > maxk ? k : maxk;
        site_ptr[p + 1] = site_ptr[p] + k;
        row0[p] = rows;
        rows += (int64_t)1 << k;
    }
    for (int i = 0; i < site_ptr[n_parts]; i++)
        if (part_sites[i] < 0 || part_sites[i] >= N) return nq_fail(ctx, NQ_ERR_SHAPE, "site index %d out of range", part_sites[i]);
    int64_t n_entries = n_parts ? row_ptr[rows] : 0;
    for (int p = 0; p < n_parts; p++)
        for (int64_t r = 0; r < ((int64_t)1 << part_nsites[p]); r++) {
            int64_t len = row_ptr[row0[p] + r + 1] - row_ptr[row0[p] + r];
            if (len < 0) return nq_fail(ctx, NQ_ERR_SHAPE, "row_ptr not monotone");
            max_row[p] = len > max_row[p] ? len : max_row[p];
        }
    int64_t max_conn = 0;
    for (int t = 0; t < n_terms; t++) {
        int L = term_left[t], R = space == NQ_SUPER ? term_right[t] : -1;
        if (L >= n_parts || R >= n_parts || (L < 0 && R < 0)) return nq_fail(ctx, NQ_ERR_SHAPE, "term %d references no valid part", t);
        max_conn += (L >= 0 ? max_row[L] : 1) * (R >= 0 ? max_row[R] : 1);
    }
    nq_operator_t op = new nq_operator_s();
    memset(op, 0, sizeof(*op));
    op->ctx = ctx; op->space = space; op->N = N; op->n_parts = n_parts; op->n_terms = n_terms;
    op->n_rows = rows; op->n_entries = n_entries; op->max_conn = max_conn > 0 ? max_conn : 1; op->max_part_sites = maxk;
    std::vector<int32_t> tr(n_terms, -1);
    if (space == NQ_SUPER) for (int t = 0; t < n_terms; t++) tr[t] = term_right[t];
    int s = NQ_OK;
    std::vector<int64_t> rp0(1, 0);
    if ((s = upload(ctx, &op->part_nsites, part_nsites, n_parts)) == NQ_OK &&
        (s = upload(ctx, &op->part_site_ptr, site_ptr.data(), n_parts + 1)) == NQ_OK &&
        (s = upload(ctx, &op->part_sites, part_sites, site_ptr[n_parts])) == NQ_OK &&
        (s = upload(ctx, &op->part_row0, row0.data(), n_parts)) == NQ_OK &&
        (s = upload(ctx, &op->row_ptr, n_parts ? row_ptr : rp0.data(), rows + 1)) == NQ_OK &&
        (s = upload(ctx, &op->entry_mel, entry_mel, 2 * n_entries)) == NQ_OK &&
        (s = upload(ctx, &op->entry_flip, entry_flip, n_entries)) == NQ_OK &&
        (s = upload(ctx, &op->term_left, term_left, n_terms)) == NQ_OK &&
        (s = upload(ctx, &op->term_right, tr.data(), n_terms)) == NQ_OK) {
        *out = op;
        return NQ_OK;
    }
    nq_operator_destroy(op);
    return s;
}

extern "C" int nq_operator_max_connections(nq_operator_t op, int64_t* out) {
    if (!op || !out) return NQ_ERR_ARG;
    *out = op->max_conn;
    return NQ_OK;
}

extern "C" int nq_connections(nq_operator_t op, nq_hilbert h, const void* srow, const void* scol, nq_dtype sdtype,
                              int64_t B, int64_t max_conn, int32_t* counts, double* mels, uint64_t* flips_row,
                              uint64_t* flips_col) {
    if (!op || !srow || !counts || !mels || !flips_row || B < 0 || max_conn <= 0) return NQ_ERR_ARG;
    nq_ctx_t ctx = op->ctx;
    if (op->space == NQ_SUPER && !scol) return nq_fail(ctx, NQ_ERR_ARG, "Liouvillian needs sigma and sigma'");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const int W64 = nq_words(op->N);
    size_t fbytes = (size_t)B * op->N * nq_dtype_size(sdtype), pbytes = (size_t)(B ? B : 1) * W64 * 8;
    const void* dr = st.in(SL_IN0, srow, fbytes);
    const void* dc = scol ? st.in(SL_IN1, scol, fbytes) : nullptr;
    if (st.status != NQ_OK) return st.status;
    uint64_t* pr = (uint64_t*)nq_scratch(ctx, SL_PROW, pbytes);
    uint64_t* pc = scol ? (uint64_t*)nq_scratch(ctx, SL_PCOL, pbytes) : nullptr;
    if (!pr || (scol && !pc)) return NQ_ERR_ALLOC;
    NQ_CHECK(nq_pack_device(ctx, h, op->N, B, dr, sdtype, pr));
    if (scol) NQ_CHECK(nq_pack_device(ctx, h, op->N, B, dc, sdtype, pc));
    int32_t* dcounts = (int32_t*)st.out(SL_OUT0, counts, (size_t)B * 4);
    double* dmels = (double*)st.out(SL_OUT1, mels, (size_t)B * max_conn * 16);
    uint64_t* dfr = (uint64_t*)st.out(SL_OUT2, flips_row, (size_t)B * max_conn * W64 * 8);
    uint64_t* dfc = flips_col ? (uint64_t*)st.out(SL_OUT3, flips_col, (size_t)B * max_conn * W64 * 8) : nullptr;
    if (st.status != NQ_OK) return st.status;
    if (B > 0) NQ_LAUNCH(ctx, connections_kernel, (unsigned)((B + 127) / 128), 128, 0, op_dev(op), pr, pc, B, W64, max_conn, dcounts, dmels, dfr, dfc);
    return st.finish();
}

static int local_common(nq_machine_t m, nq_operator_t op, const void* srow, const void* scol, nq_dtype sdtype,
                        int64_t B, void* out_loc, void* out_g, int64_t ld, bool packed) {
    if (!m || !op || !srow || !out_loc || B < 0) return NQ_ERR_ARG;
    nq_ctx_t ctx = m->ctx;
    if (m->doubled() != (scol != nullptr)) return nq_fail(ctx, NQ_ERR_ARG, "row/col configuration mismatch");
    if (out_g && ld < m->P) return nq_fail(ctx, NQ_ERR_SHAPE, "ld < P");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const uint64_t *pr, *pc;
    if (packed) {
        size_t pbytes = (size_t)B * nq_words(m->N) * 8;
        pr = (const uint64_t*)st.in(SL_PROW, srow, pbytes);
        pc = scol ? (const uint64_t*)st.in(SL_PCOL, scol, pbytes) : nullptr;
        if (st.status != NQ_OK) return st.status;
    } else {
        NQ_CHECK(nq_stage_pack(m, st, srow, scol, sdtype, B, &pr, &pc));
    }
    size_t cs = nq_dtype_size(nq_complex_of(m->dtype));
    void* dl = st.out(SL_OUT0, out_loc, (size_t)B * cs);
    void* dg = out_g ? st.out2d(SL_OUT1, out_g, (size_t)m->P * cs, (size_t)ld * cs, (size_t)B) : nullptr;
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(nq_local_device(m, op, pr, pc, B, dl, dg, ld));
    return st.finish();
}

extern "C" int nq_local_scalar(nq_machine_t m, nq_operator_t op, const void* srow, const void* scol, nq_dtype sdtype,
                               int64_t B, const void* /*logpsi: ratios are evaluated incrementally*/, void* out_loc) {
    return local_common(m, op, srow, scol, sdtype, B, out_loc, nullptr, 0, false);
}
extern "C" int nq_local_grad(nq_machine_t m, nq_operator_t op, const void* srow, const void* scol, nq_dtype sdtype,
                             int64_t B, const void* /*logpsi*/, void* out_loc, void* out_gloc, int64_t ld) {
    if (m && op && op->space != NQ_SUPER) return nq_fail(m->ctx, NQ_ERR_UNSUPPORTED, "gradient estimator is defined for Liouvillians only");
    return local_common(m, op, srow, scol, sdtype, B, out_loc, out_gloc, ld, false);
}
extern "C" int nq_local_scalar_packed(nq_machine_t m, nq_operator_t op, const uint64_t* prow, const uint64_t* pcol,
                                      int64_t B, void* out_loc) {
    return local_common(m, op, prow, pcol, NQ_F64, B, out_loc, nullptr, 0, true);
}
extern "C" int nq_local_grad_packed(nq_machine_t m, nq_operator_t op, const uint64_t* prow, const uint64_t* pcol,
                                    int64_t B, void* out_loc, void* out_gloc, int64_t ld) {
    if (m && op && op->space != NQ_SUPER) return nq_fail(m->ctx, NQ_ERR_UNSUPPORTED, "gradient estimator is defined for Liouvillians only");
    return local_common(m, op, prow, pcol, NQ_F64, B, out_loc, out_gloc, ld, true);
}
