/* cabi_smoke.c -- a plain-C host of libnqcuda's ABI (include/nqcuda.h), HOST buffers only: what a ccall / cgo / JNI
 * binding does.  create -> set_params -> logpsi_grad -> local_scalar -> center -> force -> sr_setup -> solve -> update,
 * each step checked inside C against quantities recomputed through other entry points of the same ABI:
 *   O        vs central differences of nq_logpsi in the parameters,
 *   E_loc    vs sum_c mel_c psi(eta_c)/psi(sigma) with psi(eta_c) from nq_logpsi on explicitly flipped configurations,
 *   <O>, F   vs direct sums in C,   S vs the definition,   dw vs the residual (S + eps I) dw - F.
 * Model: TFIM 1D (h = 1, J = 1, periodic), N = 6, real RBM alpha = 2 logcosh (examples/ising1d.jl scaled down).
 * build: gcc -O2 -std=c11 tests/cabi_smoke.c -Iinclude -L neuralquantum.jl_b200 -lnqcuda -lm -o tests/cabi_smoke */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "nqcuda.h"

#define N 6
#define M 12
#define P (N + M + M * N)
#define B 64
#define CK(call) do { int s_ = (call); if (s_ != NQ_OK) { fprintf(stderr, "%s -> %d (%s): %s\n", #call, s_, \
    nq_status_string(s_), ctx ? nq_last_error(ctx) : ""); return 1; } } while (0)

static unsigned long long rng_state = 88172645463325252ull;
static double urand(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
                            return (double)(rng_state >> 11) / 9007199254740992.0; }

int main(void) {
    nq_ctx_t ctx = NULL; nq_machine_t m = NULL; nq_operator_t op = NULL;
    CK(nq_ctx_create(0, NULL, &ctx));
    CK(nq_machine_create(ctx, NQ_RBM, NQ_SPIN, N, M, 0, NQ_LOGCOSH, NQ_F64, &m));
    int64_t np = 0; CK(nq_machine_nparams(m, &np));
    if (np != P) { fprintf(stderr, "P = %lld, expected %d\n", (long long)np, P); return 1; }
    static double w[P], sigma[N * B], logpsi[B], O[P * B];
    for (int i = 0; i < P; i++) w[i] = 0.2 * (urand() - 0.5);
    for (int i = 0; i < N * B; i++) sigma[i] = urand() < 0.5 ? -1.0 : 1.0;
    CK(nq_machine_set_params(m, w, P));
    CK(nq_logpsi_grad(m, sigma, NULL, NQ_F64, B, logpsi, O, P));

    /* O vs central differences on a few parameters */
    const int probe[4] = {0, N + 3, N + M + 7, P - 1};
    static double lp_p[B], lp_m[B];
    for (int q = 0; q < 4; q++) {
        const int k = probe[q]; const double hstep = 1e-5, w0 = w[k];
        w[k] = w0 + hstep; CK(nq_machine_set_params(m, w, P)); CK(nq_logpsi(m, sigma, NULL, NQ_F64, B, lp_p));
        w[k] = w0 - hstep; CK(nq_machine_set_params(m, w, P)); CK(nq_logpsi(m, sigma, NULL, NQ_F64, B, lp_m));
        w[k] = w0;
        for (int b = 0; b < B; b++) {
            const double fd = (lp_p[b] - lp_m[b]) / (2 * hstep);
            if (fabs(fd - O[k + P * b]) > 1e-8) { fprintf(stderr, "O[%d,%d] = %g, finite difference %g\n", k, b, O[k + P * b], fd); return 1; }
        }
    }
    CK(nq_machine_set_params(m, w, P));

    /* operator tables: per site a 1-site part (-h sx) and a 2-site part (J sz sz); KLocalOperator.jl:54-112 */
    enum { NPARTS = 2 * N };
    int32_t nsites[NPARTS], sites[3 * N], left[NPARTS], right[NPARTS];
    int64_t rowptr[2 * N + 4 * N + 1];
    double mel[2 * (4 * N + 4 * N)]; uint32_t flip[4 * N + 4 * N];
    int ns = 0, ne = 0, nr = 0; rowptr[0] = 0;
    for (int i = 0; i < N; i++) {
        nsites[2 * i] = 1; sites[ns++] = i;                       /* rows: local index r = digit (0: value -1) */
        for (int r = 0; r < 2; r++) {
            mel[2 * ne] = 0.0; mel[2 * ne + 1] = 0.0; flip[ne++] = 0;       /* diagonal entry first */
            mel[2 * ne] = -1.0; mel[2 * ne + 1] = 0.0; flip[ne++] = 1;      /* -h sigma^x */
            rowptr[++nr] = ne;
        }
        nsites[2 * i + 1] = 2; sites[ns++] = i < (i + 1) % N ? i : (i + 1) % N; sites[ns++] = i < (i + 1) % N ? (i + 1) % N : i;
        for (int r = 0; r < 4; r++) {                              /* sz = diag(+1, -1) on local index 1, 2 */
            const double z0 = (r & 1) ? -1.0 : 1.0, z1 = (r & 2) ? -1.0 : 1.0;
            mel[2 * ne] = z0 * z1; mel[2 * ne + 1] = 0.0; flip[ne++] = 0;
            rowptr[++nr] = ne;
        }
    }
    for (int t = 0; t < NPARTS; t++) { left[t] = t; right[t] = -1; }
    CK(nq_operator_create(ctx, NQ_KET, N, NPARTS, nsites, sites, rowptr, mel, flip, NPARTS, left, right, &op));
    static double Eloc[2 * B];
    CK(nq_local_scalar(m, op, sigma, NULL, NQ_F64, B, NULL, Eloc));
    /* E_loc by hand: flipped configurations through nq_logpsi */
    static double flipped[N * B], lpf[B];
    static double Eref[B];
    for (int b = 0; b < B; b++) { Eref[b] = 0; for (int i = 0; i < N; i++) Eref[b] += sigma[i + N * b] * sigma[(i + 1) % N + N * b]; }
    for (int i = 0; i < N; i++) {
        memcpy(flipped, sigma, sizeof(sigma));
        for (int b = 0; b < B; b++) flipped[i + N * b] = -flipped[i + N * b];
        CK(nq_logpsi(m, flipped, NULL, NQ_F64, B, lpf));
        for (int b = 0; b < B; b++) Eref[b] -= exp(lpf[b] - logpsi[b]);
    }
    for (int b = 0; b < B; b++)
        if (fabs(Eloc[2 * b] - Eref[b]) > 1e-10 * (1 + fabs(Eref[b])) || fabs(Eloc[2 * b + 1]) > 1e-12) {
            fprintf(stderr, "E_loc[%d] = %.15g%+.3gi, by hand %.15g\n", b, Eloc[2 * b], Eloc[2 * b + 1], Eref[b]); return 1; }

    /* centre, force, S, solve -- all on host arrays */
    static double avg[P], Oc[P * B], gradC[2 * P], S[P * P], F[P], dw[P];
    memcpy(Oc, O, sizeof(O));
    CK(nq_center(ctx, Oc, P, P, B, NQ_F64, avg));
    for (int k = 0; k < P; k++) {
        double a = 0; for (int b = 0; b < B; b++) a += O[k + P * b]; a /= B;
        if (fabs(a - avg[k]) > 1e-13) { fprintf(stderr, "<O>[%d]\n", k); return 1; }
        for (int b = 0; b < B; b++) if (fabs(Oc[k + P * b] - (O[k + P * b] - a)) > 1e-13) { fprintf(stderr, "Oc[%d,%d]\n", k, b); return 1; }
    }
    CK(nq_force_ket(ctx, Oc, P, P, B, NQ_F64, Eloc, gradC));
    CK(nq_sr_setup(ctx, Oc, P, P, B, B, NQ_F64, gradC, 1, S, F));
    for (int k = 0; k < P; k += 7) {
        double f = 0; for (int b = 0; b < B; b++) f += Eloc[2 * b] * Oc[k + P * b]; f /= B;
        if (fabs(f - F[k]) > 1e-12 * (1 + fabs(f))) { fprintf(stderr, "F[%d] = %g vs %g\n", k, F[k], f); return 1; }
        for (int l = 0; l < P; l += 5) {
            double sv = 0; for (int b = 0; b < B; b++) sv += Oc[k + P * b] * Oc[l + P * b]; sv /= B;
            if (fabs(sv - S[k + P * l]) > 1e-12 * (1 + fabs(sv))) { fprintf(stderr, "S[%d,%d]\n", k, l); return 1; }
        }
    }
    static double S0[P * P];
    memcpy(S0, S, sizeof(S));
    const double eps = 0.01; int64_t its = 0;
    CK(nq_sr_solve(ctx, S, F, P, NQ_F64, eps, NQ_SOLVE_CHOLESKY, 0.0, 0, dw, &its));
    double rmax = 0, fmax = 0;
    for (int k = 0; k < P; k++) {
        double r = eps * dw[k] - F[k]; for (int l = 0; l < P; l++) r += S0[k + P * l] * dw[l];
        if (fabs(r) > rmax) rmax = fabs(r);
        if (fabs(F[k]) > fmax) fmax = fabs(F[k]);
    }
    if (rmax > 1e-10 * fmax) { fprintf(stderr, "Cholesky residual %g (|F| %g)\n", rmax, fmax); return 1; }
    memcpy(S, S0, sizeof(S));
    static double dw2[P];
    CK(nq_sr_solve(ctx, S, F, P, NQ_F64, eps, NQ_SOLVE_QLP, 1e-12, 0, dw2, &its));
    for (int k = 0; k < P; k++) if (fabs(dw2[k] - dw[k]) > 1e-7 * (1 + fabs(dw[k]))) { fprintf(stderr, "QLP dw[%d]\n", k); return 1; }
    /* update: w <- w - eta dw */
    static double w2[P];
    CK(nq_update(m, dw, 0.1));
    CK(nq_machine_get_params(m, w2, P));
    for (int k = 0; k < P; k++) if (fabs(w2[k] - (w[k] - 0.1 * dw[k])) > 1e-14) { fprintf(stderr, "update[%d]\n", k); return 1; }
    uint64_t launches = 0; CK(nq_ctx_launch_count(ctx, &launches));
    printf("CABI_SMOKE_OK P=%d B=%d qlp_iters=%lld launches=%llu\n", P, B, (long long)its, (unsigned long long)launches);
    nq_operator_destroy(op); nq_machine_destroy(m); nq_ctx_destroy(ctx);
    return 0;
}
