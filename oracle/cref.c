/*
 * cref.c -- C restatement of the reference's CPU algorithm for the steady-state hot path
 * (NDM evaluation + gradient, Liouvillian local estimator with gradient).  TEST / BASELINE
 * INFRASTRUCTURE ONLY: it is compiled by oracle/build_cref.py into oracle/_ref/libcref.so and used by
 * bench.py's CPU legs and by tests; the product never links it.
 *
 * It does the work the Julia code does, in the same structure, not an optimised CPU algorithm:
 *   - every connected configuration with a non-zero matrix element (diagonal ones included) is
 *     evaluated with a FULL forward pass and a FULL gradient (outer products filling all P rows):
 *     src/IterativeInterface/Accumulators/AccumulatorObsGrad.jl:63-122, AccumulatorLogGradPsi.jl:109-120
 *   - the forward pass / gradient follow src/Networks/MixedDensityMatrix/NDMBatched.jl:177-280
 *     (six matrix-vector products, separate activation passes, utils/math.jl:23-77 outer products)
 *   - activations: src/Networks/activation.jl:5-29
 *   - connection enumeration: src/Operators/Operators/KLocalOperatorTensor.jl:129-157,
 *     KLocalLiouvillian.jl:46-52 (tables flattened by oracle/cref.py)
 * Samples are independent: an OpenMP loop over samples stands in for the reference's
 * threads / MPI ranks (src/Parallel).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

static inline double softplus_r(double x) { return log1p(exp(x)); }
static inline double dsoftplus_r(double x) { return 1.0 / (1.0 + exp(-x)); }
static inline cplx softplus_c(cplx x) { return clog(1.0 + cexp(x)); }
static inline cplx dsoftplus_c(cplx x) { return 1.0 / (1.0 + cexp(-x)); }
static inline double logcosh_r(double x) { return fabs(x) <= 12.0 ? log(cosh(x)) : fabs(x) - log(2.0); }
static inline cplx logcosh_c(cplx x) {
    double re = creal(x), im = cimag(x);
    return logcosh_r(re) + clog(cos(im) + I * tanh(re) * sin(im));
}

typedef struct {
    int N, M, A, act;
    const double *b_mu, *h_mu, *w_mu, *u_mu, *b_lam, *h_lam, *d_lam, *w_lam, *u_lam;
    int64_t P;
} ndm_t;

static void ndm_bind(ndm_t* n, const double* par, int N, int M, int A, int act) {
    n->N = N; n->M = M; n->A = A; n->act = act;
    const double* p = par;
    n->b_mu = p; p += N; n->h_mu = p; p += M; n->w_mu = p; p += (int64_t)M * N; n->u_mu = p; p += (int64_t)A * N;
    n->b_lam = p; p += N; n->h_lam = p; p += M; n->d_lam = p; p += A; n->w_lam = p; p += (int64_t)M * N;
    n->u_lam = p; p += (int64_t)A * N;
    n->P = p - par;
}

/* theta = h + W s  (W column-major [K, N]) */
static void gemv(double* th, const double* h, const double* W, const double* s, int K, int N) {
    for (int k = 0; k < K; k++) th[k] = h[k];
    for (int j = 0; j < N; j++) {
        double sj = s[j];
        const double* col = W + (int64_t)K * j;
        for (int k = 0; k < K; k++) th[k] += col[k] * sj;
    }
}

/* log rho and its gradient for ONE configuration; scratch: 6*M + 4*A doubles + 2*N */
static cplx ndm_logpsi_grad(const ndm_t* n, const double* sr, const double* sc, cplx* g, double* scratch) {
    const int N = n->N, M = n->M, A = n->A;
    double *tl = scratch, *tm = tl + M, *tlp = tm + M, *tmp = tlp + M, *ssum = tmp + M, *sdif = ssum + N;
    double *pr = sdif + N, *pi = pr + A;
    cplx* dpi = (cplx*)(pi + A);
    gemv(tl, n->h_lam, n->w_lam, sr, M, N);
    gemv(tm, n->h_mu, n->w_mu, sr, M, N);
    gemv(tlp, n->h_lam, n->w_lam, sc, M, N);
    gemv(tmp, n->h_mu, n->w_mu, sc, M, N);
    for (int j = 0; j < N; j++) { ssum[j] = sr[j] + sc[j]; sdif[j] = sr[j] - sc[j]; }
    for (int a = 0; a < A; a++) { pr[a] = 0.0; pi[a] = 0.0; }
    for (int j = 0; j < N; j++)
        for (int a = 0; a < A; a++) { pr[a] += n->u_lam[a + (int64_t)A * j] * ssum[j]; pi[a] += n->u_mu[a + (int64_t)A * j] * sdif[j]; }
    double sl = 0, sm = 0, slp = 0, smp = 0, bl = 0, bm = 0;
    cplx spi = 0;
    const int lc = n->act == 1;
    for (int k = 0; k < M; k++) {
        sl += lc ? logcosh_r(tl[k]) : softplus_r(tl[k]);
        sm += lc ? logcosh_r(tm[k]) : softplus_r(tm[k]);
        slp += lc ? logcosh_r(tlp[k]) : softplus_r(tlp[k]);
        smp += lc ? logcosh_r(tmp[k]) : softplus_r(tmp[k]);
    }
    for (int j = 0; j < N; j++) { bl += n->b_lam[j] * ssum[j]; bm += n->b_mu[j] * sdif[j]; }
    for (int a = 0; a < A; a++) {
        cplx x = 0.5 * pr[a] + n->d_lam[a] + 0.5 * I * pi[a];
        spi += lc ? logcosh_c(x) : softplus_c(x);
        dpi[a] = lc ? ctanh(x) : dsoftplus_c(x);
    }
    cplx out = 0.5 * (sl + slp + bl) + I * 0.5 * (sm - smp + bm) + spi;
    if (!g) return out;
    /* derivatives of the hidden layers (separate passes, like the reference) */
    for (int k = 0; k < M; k++) {
        tl[k] = lc ? tanh(tl[k]) : dsoftplus_r(tl[k]);
        tm[k] = lc ? tanh(tm[k]) : dsoftplus_r(tm[k]);
        tlp[k] = lc ? tanh(tlp[k]) : dsoftplus_r(tlp[k]);
        tmp[k] = lc ? tanh(tmp[k]) : dsoftplus_r(tmp[k]);
    }
    cplx* p = g;
    for (int j = 0; j < N; j++) *p++ = 0.5 * I * sdif[j];                                   /* b_mu  */
    for (int k = 0; k < M; k++) *p++ = 0.5 * I * (tm[k] - tmp[k]);                          /* h_mu  */
    for (int j = 0; j < N; j++) for (int k = 0; k < M; k++) *p++ = 0.5 * I * (tm[k] * sr[j] - tmp[k] * sc[j]);   /* w_mu */
    for (int j = 0; j < N; j++) for (int a = 0; a < A; a++) *p++ = 0.5 * I * dpi[a] * sdif[j];                  /* u_mu */
    for (int j = 0; j < N; j++) *p++ = 0.5 * ssum[j];                                       /* b_lam */
    for (int k = 0; k < M; k++) *p++ = 0.5 * (tl[k] + tlp[k]);                              /* h_lam */
    for (int a = 0; a < A; a++) *p++ = dpi[a];                                              /* d_lam */
    for (int j = 0; j < N; j++) for (int k = 0; k < M; k++) *p++ = 0.5 * (tl[k] * sr[j] + tlp[k] * sc[j]);       /* w_lam */
    for (int j = 0; j < N; j++) for (int a = 0; a < A; a++) *p++ = 0.5 * dpi[a] * ssum[j];                      /* u_lam */
    return out;
}

/* batched log rho + gradient: O [P, B] column-major */
void cref_ndm_logpsi_grad(const double* par, int N, int M, int A, int act, const double* sr, const double* sc,
                          int64_t B, double* out /* complex [B] */, double* O /* complex [P,B] or NULL */, int nthreads) {
    ndm_t n;
    ndm_bind(&n, par, N, M, A, act);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double* scratch = (double*)malloc(sizeof(double) * (6 * M + 4 * A + 2 * N + 16));
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; b++) {
            cplx v = ndm_logpsi_grad(&n, sr + b * N, sc + b * N, O ? (cplx*)O + b * n.P : NULL, scratch);
            out[2 * b] = creal(v); out[2 * b + 1] = cimag(v);
        }
        free(scratch);
    }
}

typedef struct {
    int n_parts, n_terms;
    const int32_t *part_nsites, *part_site_ptr, *part_sites;
    const int64_t *part_row0, *row_ptr;
    const double* mel; /* complex */
    const uint32_t* flip;
    const int32_t *term_left, *term_right;
} optab_t;

static int local_row(const optab_t* t, int p, const double* s, int fock) {
    int k = t->part_nsites[p], r = 0;
    const int32_t* st = t->part_sites + t->part_site_ptr[p];
    for (int i = 0; i < k; i++) {
        int d = fock ? (int)s[st[i]] : (int)((s[st[i]] + 1.0) / 2.0);
        r |= d << i;
    }
    return r;
}
static void apply_flips(const optab_t* t, int p, uint32_t f, double* s, int fock) {
    const int32_t* st = t->part_sites + t->part_site_ptr[p];
    for (int i = 0; f >> i; i++)
        if ((f >> i) & 1u) s[st[i]] = fock ? 1.0 - s[st[i]] : -s[st[i]];
}

/* L_loc [B] and grad L_loc [P,B] (complex), AccumulatorObsGrad semantics */
void cref_local_grad(const double* par, int N, int M, int A, int act, int fock,
                     int n_parts, const int32_t* part_nsites, const int32_t* part_site_ptr, const int32_t* part_sites,
                     const int64_t* part_row0, const int64_t* row_ptr, const double* mel, const uint32_t* flip,
                     int n_terms, const int32_t* term_left, const int32_t* term_right,
                     const double* sr, const double* sc, int64_t B, double* out_loc, double* out_g, int nthreads) {
    ndm_t n;
    ndm_bind(&n, par, N, M, A, act);
    optab_t t = {n_parts, n_terms, part_nsites, part_site_ptr, part_sites, part_row0, row_ptr, mel, flip, term_left, term_right};
    const int64_t P = n.P;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double* scratch = (double*)malloc(sizeof(double) * (6 * M + 4 * A + 2 * N + 16));
        double* er = (double*)malloc(sizeof(double) * 2 * N);
        double* ec = er + N;
        cplx* g = (cplx*)malloc(sizeof(cplx) * P);
#pragma omp for schedule(dynamic, 8)
        for (int64_t b = 0; b < B; b++) {
            const double *r0 = sr + b * N, *c0 = sc + b * N;
            cplx lp0 = ndm_logpsi_grad(&n, r0, c0, NULL, scratch);
            cplx acc = 0;
            cplx* G = out_g ? (cplx*)out_g + b * P : NULL;
            if (G) memset(G, 0, sizeof(cplx) * P);
            for (int tt = 0; tt < n_terms; tt++) {
                int L = term_left[tt], R = term_right[tt];
                int64_t l0 = 0, l1 = 1, q0 = 0, q1 = 1;
                if (L >= 0) { int64_t row = part_row0[L] + local_row(&t, L, r0, fock); l0 = row_ptr[row]; l1 = row_ptr[row + 1]; }
                if (R >= 0) { int64_t row = part_row0[R] + local_row(&t, R, c0, fock); q0 = row_ptr[row]; q1 = row_ptr[row + 1]; }
                for (int64_t el = l0; el < l1; el++)
                    for (int64_t eq = q0; eq < q1; eq++) {
                        cplx m = 1.0;
                        if (L >= 0) m *= mel[2 * el] + I * mel[2 * el + 1];
                        if (R >= 0) m *= mel[2 * eq] + I * mel[2 * eq + 1];
                        if (m == 0.0) continue;
                        memcpy(er, r0, sizeof(double) * N);
                        memcpy(ec, c0, sizeof(double) * N);
                        if (L >= 0) apply_flips(&t, L, flip[el], er, fock);
                        if (R >= 0) apply_flips(&t, R, flip[eq], ec, fock);
                        cplx lp = ndm_logpsi_grad(&n, er, ec, G ? g : NULL, scratch);
                        cplx w = m * cexp(lp - lp0);
                        acc += w;
                        if (G) for (int64_t i = 0; i < P; i++) G[i] += w * g[i];
                    }
            }
            out_loc[2 * b] = creal(acc); out_loc[2 * b + 1] = cimag(acc);
        }
        free(scratch); free(er); free(g);
    }
}

int cref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
