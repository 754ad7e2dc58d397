// nq_sampler.cu -- K4: Metropolis-Hastings local-flip sampler, one Markov chain per warp.
//
// The reference re-evaluates the whole machine for every proposal (Samplers/Metropolis.jl:141-146,
// O(M N) per pass) and loops serially over chains to propose (MCMCRules/LocalRule.jl:19-28).  Here a
// warp owns a chain for the whole burn-in + sampling run inside ONE launch: the pre-activations
// theta live in shared memory, a single flip changes theta by +-column of W (O(M) per pass), the
// lanes own hidden units, the log-probability ratio is a warp-shuffle reduction and randomness is
// Philox4x32-10 keyed by (seed, global chain id) so the chains do not depend on the GPU count.
// Replay mode takes the proposal sites and uniforms from the caller instead: accept/reject
// decisions are then comparable bit for bit with the oracle.
//
// Accept rule (Metropolis.jl:148-154): accept iff  u - exp(logp' - logp) < 0,  logp = 2 Re log psi,
// evaluated in the machine's real precision.
#include "nq_internal.cuh"
#include "nq_opdev.cuh"

struct nq_sampler_s {
    nq_ctx_t ctx;
    nq_machine_t m;
    int64_t B;
    int passes;
    uint64_t seed;
    int64_t chain_offset;
    uint64_t pass_base;       // proposals already drawn per chain (Philox counter)
    int diag;                 // 1: chain over the diagonal rho(sigma, sigma) (nq_sampler_set_mode)
    uint64_t* prow;           // device [B][W64]
    uint64_t* pcol;
    unsigned long long* accepted;   // device counter
    int64_t passes_done;
    // transition rule (nq_sampler_set_rule): couplings of Exchange / Nagy, operator of the OperatorRule
    int rule;
    int n_coup;
    int32_t* coup;            // device [n_coup][2], 0-based sites
    nq_operator_t rule_op;
};

namespace {

constexpr int MAXW = 4;   // N <= 256
// The tracked f' values follow accepted moves exactly (in exact arithmetic); they are rebuilt from the configuration at the
// start of every launch and then every REFRESH_STEPS stored steps, which bounds the accumulated rounding (~1e-16 per
// accepted move) without paying a full pass over the weights per step: on cfg3 that pass moved as many bytes as all the
// proposals of the step.
constexpr int REFRESH_STEPS = 16;

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t ka, uint32_t kb) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
        uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ ka, n1 = lo1, n2 = hi0 ^ c[3] ^ kb, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __device__ __forceinline__ void gen(uint32_t (&c)[4]) {
        uint32_t ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            round(c, ka, kb);
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
    }
};

template <typename T> __device__ __forceinline__ T uniform01(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ float uniform01<float>(uint32_t a, uint32_t) { return (float)(a >> 8) * 5.9604644775390625e-8f; }
template <> __device__ __forceinline__ double uniform01<double>(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__global__ void randomize_kernel(uint64_t* __restrict__ prow, uint64_t* __restrict__ pcol, int64_t B, int N, int W64,
                                 uint64_t seed, int64_t chain_offset, uint64_t epoch) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= B * W64) return;
    int64_t b = i / W64;
    int w = (int)(i % W64);
    uint64_t gid = (uint64_t)(chain_offset + b);
    Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(epoch * 8 + w), 0x52414e44u /* "RAND" domain */};
    ph.gen(c);
    int nbits = N - 64 * w; nbits = nbits > 64 ? 64 : nbits;
    uint64_t mask = nbits == 64 ? ~0ull : ((1ull << nbits) - 1);
    prow[i] = (((uint64_t)c[1] << 32) | c[0]) & mask;
    if (pcol) pcol[i] = (((uint64_t)c[3] << 32) | c[2]) & mask;
}

// shared memory per warp: items * (theta, f, theta_tentative, f_tentative)
struct RunArgs {
    int64_t B;
    int N, M, A, hilb, passes, burn, L, replay, diag;
    uint64_t seed, pass_base;
    int64_t chain_offset;
    const int32_t* sites;      // replay: [passes][B], 1-based
    const void* uniforms;      // replay: [passes][B]
    uint8_t* accept_out;       // replay: [passes][B]
    uint64_t* out_prow;        // [L][B][W64]
    uint64_t* out_pcol;
    unsigned long long* accepted;
    int rule, n_coup;
    const int32_t* coup;
    const int32_t* draws;      // replay of a rule: [passes][B][4]
};

template <typename T>
__device__ __forceinline__ void draw(const RunArgs& a, int64_t chain, uint64_t pass_idx, int pass_in_call, int nsites,
                                     int& site, T& u) {
    if (a.replay) {
        site = a.sites[(int64_t)pass_in_call * a.B + chain] - 1;
        u = ((const T*)a.uniforms)[(int64_t)pass_in_call * a.B + chain];
    } else {
        uint64_t gid = (uint64_t)(a.chain_offset + chain);
        Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
        uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)pass_idx, (uint32_t)(pass_idx >> 32)};
        ph.gen(c);
        site = (int)(((uint64_t)c[0] * (uint64_t)nsites) >> 32);
        u = uniform01<T>(c[1], c[2]);
    }
}

// Proposal ps of a stored step: the 32 lanes generate the (site, uniform) pairs of 32 consecutive proposals at
// once (lane l -> proposal ps0 + l; same Philox counters as one call per proposal) and every proposal reads
// its pair with shuffles.  Replay mode reads the caller's arrays.
template <typename T>
struct DrawBatch {
    int site_l; T u_l;
    __device__ __forceinline__ void fill(const RunArgs& a, int64_t chain, int pic0, int nleft, int nsites, int lane) {
        site_l = 0; u_l = T(0);
        if (lane < nleft) draw<T>(a, chain, a.pass_base + (uint64_t)(pic0 + lane), pic0 + lane, nsites, site_l, u_l);
    }
    __device__ __forceinline__ void get(int idx, int& site, T& u) const {
        site = __shfl_sync(0xffffffffu, site_l, idx);
        u = __shfl_sync(0xffffffffu, u_l, idx);
    }
};

// ---- RBM / RBMSplit -------------------------------------------------------------------
// log cosh(theta) = softplus(2 theta) - theta - ln 2, so both activations run ONE recurrence on q_k = sigmoid(g theta_k)
// (g = 2 logcosh, 1 softplus; q = (1 + tanh theta)/2 or f'(theta), no extra transcendental):
//   psi(eta)/psi(sigma) = e^{dv (a_j - [logcosh] sum_k W_kj)} prod_k [1 + q_k T_kj],   T = exp(+-c g W_kj) - 1
//   accepted move:  q_k <- q_k (1 + T_kj) / (1 + q_k T_kj)
// A proposal is one 16-byte table load and one complex multiply-add per hidden unit.  The first version kept the tanh
// values in shared memory and read two table entries per unit (exp(+-c W)); it was bound by L1 / shared-memory bandwidth
// (cfg3: 61 M proposals x 144 units x 64 B = 21 TB/s).  Here q, T and the factors of the current proposal live in
// REGISTERS (KU = ceil(M / 32) units per lane), so a proposal moves 16 B per unit and an accepted one nothing more.
// theta is rebuilt from the configuration once per stored sample from site values staged per warp.
__device__ __forceinline__ double fast_rcp(double x) {           // MUFU.RCP64H seed + cubic step, ~1 ulp, no slow path
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    return fma(r, e, r);
}
__device__ __forceinline__ float fast_rcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float q_div(float a, float b) { return a * fast_rcp(b); }
__device__ __forceinline__ double q_div(double a, double b) { return a * fast_rcp(b); }
template <typename T> __device__ __forceinline__ cx<T> q_div(cx<T> a, cx<T> b) {
    const T inv = fast_rcp(b.re * b.re + b.im * b.im);
    return cx<T>((a.re * b.re + a.im * b.im) * inv, (a.im * b.re - a.re * b.im) * inv);
}

template <typename E, int ACT, bool DOUBLED, int KU>
__global__ void __launch_bounds__(256) sampler_rbm_reg_kernel(const E* __restrict__ par, const E* __restrict__ qtab,
                                                               uint64_t* __restrict__ st_row, uint64_t* __restrict__ st_col, RunArgs a) {
    typedef typename elem_traits<E>::real T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)wpb + warp;
    const int N = a.N, M = a.M, W64 = (N + 63) >> 6;
    const int64_t off_b = DOUBLED ? 2 * N : N;
    const int64_t MN = (int64_t)M * N;
    constexpr int nmat = DOUBLED ? 2 : 1;
    const E* __restrict__ wsum = qtab + 2 * nmat * MN;
    // eb[(sign * nmat + mat) * N + j] = |exp(dv (a_j - [logcosh] sum_k W_kj))|^2 for the two flips of site j (CTA-wide)
    T* eb = (T*)smem_raw;
    for (int i = threadIdx.x; i < 2 * nmat * N; i += blockDim.x) {
        const int sgn = i / (nmat * N), rem = i - sgn * nmat * N, mt = rem / N, jj = rem - mt * N;
        const T dvv = a.hilb == NQ_SPIN ? (sgn ? T(-2) : T(2)) : (sgn ? T(-1) : T(1));
        E lin = par[(mt ? N : 0) + jj];
        if (ACT == NQ_LOGCOSH) lin = lin - wsum[mt * N + jj];
        eb[i] = (T)abs2_d(e_exp(rscale(dvv, lin)));
    }
    __syncthreads();
    if (chain >= a.B) return;
    const int nsv = DOUBLED ? 2 * N : N;
    T* vsw = (T*)(smem_raw + ((size_t)4 * N * sizeof(T) + 15) / 16 * 16) + (size_t)warp * ((nsv + 1) & ~1);
    const E* __restrict__ Wr = par + off_b + M;
    uint64_t rb[MAXW], cb[MAXW];
#pragma unroll
    for (int w = 0; w < MAXW; w++) {
        rb[w] = w < W64 ? st_row[chain * W64 + w] : 0ull;
        cb[w] = (DOUBLED && w < W64) ? st_col[chain * W64 + w] : 0ull;
    }
    const int nsites = DOUBLED ? 2 * N : N;
    unsigned nacc = 0;
    const int nsteps = a.burn + a.L;
    E q[KU];
    for (int step = 0; step < nsteps; step++) {
        if (step % REFRESH_STEPS == 0) {
        for (int i = lane; i < nsv; i += 32) vsw[i] = digit_value<T>(a.hilb, get_bit(i < N ? rb : cb, i < N ? i : i - N));
        __syncwarp();
        {   // Wr and Wc are contiguous: column i of the [M, nsv] matrix belongs to vsw[i]
            E t[KU];
#pragma unroll
            for (int u = 0; u < KU; u++) t[u] = lane + 32 * u < M ? par[off_b + lane + 32 * u] : make_zero<E>();
            for (int i = 0; i < nsv; i++) {
                const T v = vsw[i];
                const E* __restrict__ wc = Wr + (int64_t)M * i + lane;
#pragma unroll
                for (int u = 0; u < KU; u++) if (lane + 32 * u < M) t[u] += rscale(v, wc[32 * u]);
            }
#pragma unroll
            for (int u = 0; u < KU; u++) {
                E f, d;
                act_eval<ACT>(t[u], f, d);
                q[u] = ACT == NQ_LOGCOSH ? rscale(T(0.5), e_one<E>() + d) : d;
            }
        }
        __syncwarp();
        }
        DrawBatch<T> db;
#pragma unroll 1
        for (int ps = 0; ps < a.passes; ps++) {
            int pic = step * a.passes + ps;
            if ((ps & 31) == 0) db.fill(a, chain, pic, a.passes - ps, nsites, lane);
            int site; T u01;
            db.get(ps & 31, site, u01);
            const bool col = DOUBLED && site >= N;
            const int j = col ? site - N : site;
            const T dv = flip_delta<T>(a.hilb, get_bit(col ? cb : rb, j));
            const int sg = dv > T(0) ? 0 : 1, mat = col ? 1 : 0;
            const E* __restrict__ tp = qtab + (int64_t)(sg * nmat + mat) * MN + (int64_t)M * j + lane;
            E Tk[KU], fac[KU];
            double prod = 1.0;
#pragma unroll
            for (int u = 0; u < KU; u++) {
                Tk[u] = lane + 32 * u < M ? tp[32 * u] : make_zero<E>();
                fac[u] = e_one<E>() + e_mul(q[u], Tk[u]);
                prod *= abs2_d(fac[u]);
            }
            prod = warp_prod(prod);
            const double pr = prod * (double)eb[(sg * nmat + mat) * N + j];
            const bool acc = (u01 - (T)pr) < T(0);
            if (acc) {
#pragma unroll
                for (int u = 0; u < KU; u++) q[u] = q_div(q[u] + e_mul(q[u], Tk[u]), fac[u]);
                if (col) cb[j >> 6] ^= 1ull << (j & 63); else rb[j >> 6] ^= 1ull << (j & 63);
                nacc++;
            }
            if (a.replay && a.accept_out && lane == 0) a.accept_out[(int64_t)pic * a.B + chain] = acc ? 1 : 0;
        }
        if (step >= a.burn && a.out_prow && lane == 0) {
            int64_t o = ((int64_t)(step - a.burn) * a.B + chain) * W64;
            for (int w = 0; w < W64; w++) { a.out_prow[o + w] = rb[w]; if (DOUBLED && a.out_pcol) a.out_pcol[o + w] = cb[w]; }
        }
        __syncwarp();
    }
    if (lane == 0) {
        for (int w = 0; w < W64; w++) { st_row[chain * W64 + w] = rb[w]; if (DOUBLED) st_col[chain * W64 + w] = cb[w]; }
        if (nacc) atomicAdd(a.accepted, (unsigned long long)nacc);
    }
}

// ---- NDM ------------------------------------------------------------------------------
// p(sigma,sigma') = |rho|^2: the mu hidden layer only enters the phase, so the chain tracks the lambda layer
// (on sigma and sigma') and the ancilla layer:  |rho(eta)/rho(sigma)|^2 = e^{b_lam_j dv} prod_lambda |prod_Pi|^2.
// per warp: spi[A], tspi[A] (complex); sl[2M] (side*M + k), tsl[M]
template <typename T, int ACT>
__global__ void sampler_ndm_kernel(const T* __restrict__ par, const T* __restrict__ tabr, const cx<T>* __restrict__ tabc,
                                   uint64_t* __restrict__ st_row, uint64_t* __restrict__ st_col, RunArgs a) {
    typedef cx<T> C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)wpb + warp;
    const int N = a.N, M = a.M, A = a.A, W64 = (N + 63) >> 6;
    const int64_t MN = (int64_t)M * N, AN = (int64_t)A * N;
    const int64_t o_wmu = N + M, o_umu = o_wmu + MN, o_blam = o_umu + AN,
                  o_hlam = o_blam + N, o_dlam = o_hlam + M, o_wlam = o_dlam + A, o_ulam = o_wlam + MN;
    // eb[sign * N + j] = exp(dv b_lam_j) for the two possible flips of site j (CTA-wide table)
    T* eb = (T*)smem_raw;
    for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
        const int sgn = i / N, jj = i - sgn * N;
        const T dvv = a.hilb == NQ_SPIN ? (sgn ? T(-2) : T(2)) : (sgn ? T(-1) : T(1));
        eb[i] = (T)exp((double)(dvv * par[o_blam + jj]));
    }
    __syncthreads();
    if (chain >= a.B) return;
    const size_t per_warp = (size_t)(4 * M + 2 * N) * sizeof(T) + (size_t)(3 * A) * sizeof(C);
    unsigned char* base = smem_raw + ((size_t)2 * N * sizeof(T) + 15) / 16 * 16 + (size_t)warp * ((per_warp + 15) / 16 * 16);
    C* spi = (C*)base;            // [A]
    C* tspi = spi + A;            // [2A]: numerators, then factors
    T* sl = (T*)(tspi + 2 * A);   // [2M]
    T* tsl = sl + 2 * M;          // [2M]: numerators, then factors
    T* vsw = tsl + 2 * M;         // [2N]: site values of (sigma, sigma') for the refresh of theta
    const T half = T(0.5);
    uint64_t rb[MAXW], cb[MAXW];
#pragma unroll
    for (int w = 0; w < MAXW; w++) {
        rb[w] = w < W64 ? st_row[chain * W64 + w] : 0ull;
        cb[w] = w < W64 ? st_col[chain * W64 + w] : 0ull;
    }
    unsigned nacc = 0;
    const int nsteps = a.burn + a.L;
    for (int step = 0; step < nsteps; step++) {
        if (step % REFRESH_STEPS == 0) {
        for (int i = lane; i < 2 * N; i += 32) vsw[i] = digit_value<T>(a.hilb, get_bit(i < N ? rb : cb, i < N ? i : i - N));
        __syncwarp();
        for (int k = lane; k < M; k += 32) {
            const T* __restrict__ w = par + o_wlam + k;
            T t = par[o_hlam + k], tp = t;
            for (int j = 0; j < N; j++) {
                T wv = w[(int64_t)M * j];
                t += wv * vsw[j];
                tp += wv * vsw[N + j];
            }
            T f, d, fp, dp;
            act_eval<ACT>(t, f, d);
            act_eval<ACT>(tp, fp, dp);
            sl[k] = d; sl[M + k] = dp;
        }
        for (int q = lane; q < A; q += 32) {
            T pr = par[o_dlam + q], pim = T(0);
            for (int j = 0; j < N; j++) {
                T x = vsw[j], y = vsw[N + j];
                pr += half * par[o_ulam + q + (int64_t)A * j] * (x + y);
                pim += half * par[o_umu + q + (int64_t)A * j] * (x - y);
            }
            C f, d;
            act_eval<ACT>(C(pr, pim), f, d);
            spi[q] = d;
        }
        __syncwarp();
        }
        DrawBatch<T> db;
#pragma unroll 1
        for (int ps = 0; ps < a.passes; ps++) {
            int pic = step * a.passes + ps;
            if ((ps & 31) == 0) db.fill(a, chain, pic, a.passes - ps, a.diag ? N : 2 * N, lane);
            int site; T u;
            db.get(ps & 31, site, u);
            // diagonal mode: sigma' = sigma, a proposal flips site j of both (p ~ rho(sigma, sigma), which is real)
            const bool col = !a.diag && site >= N;
            const int j = col ? site - N : site;
            const T dv = flip_delta<T>(a.hilb, get_bit(col ? cb : rb, j));
            const int sg = dv > T(0) ? 0 : 1;
            const int so = col ? M : 0;
            const T* __restrict__ tp = tabr + (int64_t)sg * 2 * MN + (int64_t)M * j;          // lambda layer = lay 0
            const T* __restrict__ tm = tabr + (int64_t)(1 - sg) * 2 * MN + (int64_t)M * j;
            const C* __restrict__ cp = tabc + (int64_t)sg * AN + (int64_t)A * j;
            const C* __restrict__ cm = tabc + (int64_t)(1 - sg) * AN + (int64_t)A * j;
            // |rho(eta)/rho(sigma)|^2 = e^{b_lam dv} prod_k fac_k prod_q |fac_q|^2: ONE real product; the new
            // f' values (num / den) are only formed when the move is accepted
            double pl = 1.0;
            for (int k = lane; k < M; k += 32) {
                T Ep = tp[k], Em = ACT == NQ_LOGCOSH ? tm[k] : T(1);
                T fac, num;
                ratio_parts<ACT>(sl[so + k], Ep, Em, fac, num);
                pl *= (double)fac;
                tsl[k] = num; tsl[M + k] = fac;
            }
            for (int q = lane; q < A; q += 32) {
                C Ep = cp[q], Em = ACT == NQ_LOGCOSH ? cm[q] : C(T(1), T(0));
                if (col) { Ep.im = -Ep.im; Em.im = -Em.im; }
                if (a.diag) {              // row and column factors multiply: E conj(E) = |E|^2, the ancilla stays real
                    Ep = C(Ep.re * Ep.re + Ep.im * Ep.im, T(0));
                    Em = C(Em.re * Em.re + Em.im * Em.im, T(0));
                }
                C fac, num;
                ratio_parts<ACT>(spi[q], Ep, Em, fac, num);
                pl *= a.diag ? (double)fac.re : abs2_d(fac);
                tspi[q] = num; tspi[A + q] = fac;
            }
            pl = warp_prod(pl);
            const double pr = pl * (double)eb[sg * N + j];
            const bool acc = (u - (T)pr) < T(0);
            if (acc) {
                for (int k = lane; k < M; k += 32) {
                    const T sn = ratio_finish<ACT>(tsl[M + k], tsl[k]);
                    sl[so + k] = sn;
                    if (a.diag) sl[M + k] = sn;
                }
                for (int q = lane; q < A; q += 32) spi[q] = ratio_finish<ACT>(tspi[A + q], tspi[q]);
                if (col || a.diag) cb[j >> 6] ^= 1ull << (j & 63);
                if (!col) rb[j >> 6] ^= 1ull << (j & 63);
                nacc++;
            }
            __syncwarp();
            if (a.replay && a.accept_out && lane == 0) a.accept_out[(int64_t)pic * a.B + chain] = acc ? 1 : 0;
        }
        if (step >= a.burn && a.out_prow && lane == 0) {
            int64_t o = ((int64_t)(step - a.burn) * a.B + chain) * W64;
            for (int w = 0; w < W64; w++) { a.out_prow[o + w] = rb[w]; if (a.out_pcol) a.out_pcol[o + w] = cb[w]; }
        }
    }
    if (lane == 0) {
        for (int w = 0; w < W64; w++) { st_row[chain * W64 + w] = rb[w]; st_col[chain * W64 + w] = cb[w]; }
        if (nacc) atomicAdd(a.accepted, (unsigned long long)nacc);
    }
}


// ---- transition rules other than LocalRule ------------------------------------------------
// ExchangeRule (MCMCRules/ExchangeRule.jl:36-68, ket states), NagyRule (Nagy.jl:40-115, doubled states) and
// OperatorRule (OperatorRule.jl:27-59: a uniformly drawn connection of the operator, log_prob_bias =
// log(n_forward / n_back)).  Proposals flip several sites, so this kernel tracks the pre-activations theta_k of
// every unit and re-evaluates f on theta_k + sum_flips dv W_kj (O(units) transcendentals per proposal, in double
// precision whatever the machine's type; the accept test is rounded to the machine's real type like the reference).
// One warp per chain, all lanes build the proposal redundantly (integer work on the same addresses).
//   log p = lin + sum_k c_k Re f(theta_k):   RBM/RBMSplit c = 2, lin = 2 Re(a.v);   NDM: lambda units of sigma and
//   sigma' (c = 1, real), ancillas (c = 2, complex), lin = b_lam.(v + v').
constexpr int MAXFLIP = 16;

__device__ __forceinline__ cxd as_cxd(float a) { return cxd((double)a, 0.0); }
__device__ __forceinline__ cxd as_cxd(double a) { return cxd(a, 0.0); }
template <typename T> __device__ __forceinline__ cxd as_cxd(cx<T> a) { return cxd((double)a.re, (double)a.im); }

template <typename E, int KIND>
struct Units {
    const E* par; int N, M, A;
    __device__ __forceinline__ int count() const { return KIND == NQ_NDM ? 2 * M + A : M; }
    // weight of a unit in log p and whether its pre-activation is complex
    __device__ __forceinline__ double weight(int k) const { return (KIND == NQ_NDM && k < 2 * M) ? 1.0 : 2.0; }
    __device__ __forceinline__ cxd bias(int k) const {
        if (KIND == NQ_RBM) return as_cxd(par[N + k]);
        if (KIND == NQ_RBMSPLIT) return as_cxd(par[2 * N + k]);
        const int64_t MN = (int64_t)M * N, AN = (int64_t)A * N;
        const int64_t o_hlam = (N + M + MN + AN) + N, o_dlam = o_hlam + M;
        if (k < 2 * M) return as_cxd(par[o_hlam + (k < M ? k : k - M)]);
        return as_cxd(par[o_dlam + (k - 2 * M)]);
    }
    // d theta_k / d value of site j on side `side`
    __device__ __forceinline__ cxd slope(int k, int side, int j) const {
        if (KIND == NQ_RBM) return as_cxd(par[N + M + k + (int64_t)M * j]);
        if (KIND == NQ_RBMSPLIT) return as_cxd(par[2 * N + M + (side ? (int64_t)M * N : 0) + k + (int64_t)M * j]);
        const int64_t MN = (int64_t)M * N, AN = (int64_t)A * N;
        const int64_t o_umu = N + M + MN, o_wlam = o_umu + AN + N + M + A, o_ulam = o_wlam + MN;
        if (k < M) return side == 0 ? as_cxd(par[o_wlam + k + (int64_t)M * j]) : cxd(0.0, 0.0);
        if (k < 2 * M) return side == 1 ? as_cxd(par[o_wlam + (k - M) + (int64_t)M * j]) : cxd(0.0, 0.0);
        const int q = k - 2 * M;
        const double ul = 0.5 * (double)real_part(par[o_ulam + q + (int64_t)A * j]), um = 0.5 * (double)real_part(par[o_umu + q + (int64_t)A * j]);
        return cxd(ul, side == 0 ? um : -um);
    }
    // d lin / d value of site j on side `side`
    __device__ __forceinline__ double lin_slope(int side, int j) const {
        if (KIND == NQ_RBM) return 2.0 * (double)real_part(par[j]);
        if (KIND == NQ_RBMSPLIT) return 2.0 * (double)real_part(par[(side ? N : 0) + j]);
        const int64_t o_blam = N + M + (int64_t)M * N + (int64_t)A * N;
        return (double)real_part(par[o_blam + j]);
    }
};

template <int ACT>
__device__ __forceinline__ double unit_value(cxd th, bool is_real) {
    if (is_real) { double f, d; act_eval<ACT>(th.re, f, d); return f; }
    cxd f, d;
    act_eval<ACT>(th, f, d);
    return f.re;
}

struct Proposal { int n; int idx[MAXFLIP]; double bias; };   // idx = side * N + site

__device__ __forceinline__ void toggle(uint64_t* m, int j) { m[j >> 6] ^= 1ull << (j & 63); }

// number of connections of (rb, cb) in reference order (zero matrix elements included: length(row_valdiff!(...)))
__device__ int op_count(const OpDev& op, const uint64_t* rb, const uint64_t* cb) {
    int n = 0;
    for (int t = 0; t < op.n_terms; t++) visit_term(op, t, rb, cb, [&](double, double, int, uint32_t, int, uint32_t) { n++; });
    return n;
}
// flip masks of connection number `pick`
__device__ void op_pick(const OpDev& op, const uint64_t* rb, const uint64_t* cb, int pick, uint64_t* fr, uint64_t* fc) {
    int n = 0;
    for (int t = 0; t < op.n_terms; t++)
        visit_term(op, t, rb, cb, [&](double, double, int L, uint32_t fl, int R, uint32_t frr) {
            if (n == pick) {
                if (L >= 0) {
                    const int32_t* s = op.part_sites + op.part_site_ptr[L];
                    for (int i = 0; fl >> i; i++) if ((fl >> i) & 1u) toggle(fr, s[i]);
                }
                if (R >= 0) {
                    const int32_t* s = op.part_sites + op.part_site_ptr[R];
                    for (int i = 0; frr >> i; i++) if ((frr >> i) & 1u) toggle(fc, s[i]);
                }
            }
            n++;
        });
}

// d[0..3]: production = raw 32-bit randoms; replay = the reference's integer draws (1-based), see nq_sampler_replay_rule
template <bool DOUBLED>
__device__ void make_proposal(const RunArgs& a, const OpDev& op, bool replay, const uint32_t (&d)[4], const uint64_t* rb,
                              const uint64_t* cb, Proposal& p) {
    const int N = a.N, W64 = (N + 63) >> 6;
    uint64_t fr[MAXW], fc[MAXW];
#pragma unroll
    for (int w = 0; w < MAXW; w++) { fr[w] = 0; fc[w] = 0; }
    p.bias = 0.0;
    auto pick = [&](uint32_t raw, int range) { return replay ? (int)raw - 1 : (int)(((uint64_t)raw * (uint64_t)range) >> 32); };
    if (a.rule == NQ_RULE_EXCHANGE) {
        const int c = pick(d[0], a.n_coup);
        const int i = a.coup[2 * c], j = a.coup[2 * c + 1];
        if (get_bit(rb, i) != get_bit(rb, j)) { toggle(fr, i); toggle(fr, j); }
    } else if (a.rule == NQ_RULE_NAGY) {
        const int move = pick(d[0], 8) + 1;
        const int s1 = pick(d[1], N);
        if (move <= 4) {                               // hopping in sigma (1, 2) or sigma' (3, 4)
            const int s2 = a.coup[2 * s1 + pick(d[2], 2)];       // rand(adjacency_list[s1]): one element of the s1-th couple
            uint64_t* f = move <= 2 ? fr : fc;
            toggle(f, s1); toggle(f, s2);
        } else if (move == 5) {
            toggle(fr, s1);
        } else if (move == 6) {
            toggle(fc, s1);
        } else if (move == 7) {                        // dissipator: an empty site is excited with probability 1/10
            const bool r_empty = a.hilb == NQ_FOCK && get_bit(rb, s1) == 0;
            const bool c_empty = a.hilb == NQ_FOCK && get_bit(cb, s1) == 0;
            if (!r_empty || pick(d[2], 10) == 0) toggle(fr, s1);
            if (!c_empty || pick(d[3], 10) == 0) toggle(fc, s1);
        } else {                                       // jumper
            toggle(fr, s1);
            toggle(fc, pick(d[2], N));
        }
    } else {                                           // NQ_RULE_OPERATOR
        const int nf = op_count(op, rb, DOUBLED ? cb : nullptr);
        const int k = replay ? (int)(((uint64_t)d[0] * (uint64_t)nf) >> 32) : (int)(((uint64_t)d[0] * (uint64_t)nf) >> 32);
        op_pick(op, rb, DOUBLED ? cb : nullptr, k, fr, fc);
        uint64_t nr[MAXW], nc[MAXW];
#pragma unroll
        for (int w = 0; w < MAXW; w++) { nr[w] = rb[w] ^ fr[w]; nc[w] = DOUBLED ? cb[w] ^ fc[w] : 0ull; }
        const int nb = op_count(op, nr, DOUBLED ? nc : nullptr);
        p.bias = log((double)nf / (double)nb);
    }
    p.n = 0;
    for (int w = 0; w < W64; w++) {
        uint64_t m = fr[w];
        while (m && p.n < MAXFLIP) { int b = __ffsll((long long)m) - 1; m &= m - 1; p.idx[p.n++] = w * 64 + b; }
        m = DOUBLED ? fc[w] : 0ull;
        while (m && p.n < MAXFLIP) { int b = __ffsll((long long)m) - 1; m &= m - 1; p.idx[p.n++] = N + w * 64 + b; }
    }
}

template <typename E, int KIND, int ACT>
__global__ void sampler_rule_kernel(const E* __restrict__ par, OpDev op, uint64_t* __restrict__ st_row,
                                    uint64_t* __restrict__ st_col, RunArgs a) {
    typedef typename elem_traits<E>::real T;
    constexpr bool DOUBLED = KIND != NQ_RBM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)wpb + warp;
    if (chain >= a.B) return;
    const int N = a.N, W64 = (N + 63) >> 6;
    Units<E, KIND> un{par, N, a.M, a.A};
    const int U = un.count();
    // per warp: theta[2][U] (current / tentative), fval[2][U]
    cxd* th = (cxd*)smem_raw + (size_t)warp * 3 * U;
    double* fv = (double*)(th + 2 * U);
    int cur = 0;
    uint64_t rb[MAXW], cb[MAXW];
#pragma unroll
    for (int w = 0; w < MAXW; w++) {
        rb[w] = w < W64 ? st_row[chain * W64 + w] : 0ull;
        cb[w] = (DOUBLED && w < W64) ? st_col[chain * W64 + w] : 0ull;
    }
    unsigned nacc = 0;
    const int nsteps = a.burn + a.L;
    const uint64_t gid = (uint64_t)(a.chain_offset + chain);
    for (int step = 0; step < nsteps; step++) {
        for (int k = lane; k < U; k += 32) {
            cxd t = un.bias(k);
            for (int j = 0; j < N; j++) {
                t += rscale((double)digit_value<T>(a.hilb, get_bit(rb, j)), un.slope(k, 0, j));
                if (DOUBLED) t += rscale((double)digit_value<T>(a.hilb, get_bit(cb, j)), un.slope(k, 1, j));
            }
            th[cur * U + k] = t;
            fv[cur * U + k] = un.weight(k) * unit_value<ACT>(t, KIND == NQ_NDM && k < 2 * a.M);
        }
        __syncwarp();
#pragma unroll 1
        for (int ps = 0; ps < a.passes; ps++) {
            const int pic = step * a.passes + ps;
            uint32_t d[4];
            T u;
            if (a.replay) {
                const int32_t* dr = a.draws + ((int64_t)pic * a.B + chain) * 4;
                d[0] = (uint32_t)dr[0]; d[1] = (uint32_t)dr[1]; d[2] = (uint32_t)dr[2]; d[3] = (uint32_t)dr[3];
                u = ((const T*)a.uniforms)[(int64_t)pic * a.B + chain];
            } else {
                const uint64_t pidx = a.pass_base + (uint64_t)pic;
                Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
                uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)pidx, (uint32_t)(pidx >> 32)};
                ph.gen(c);
                u = uniform01<T>(c[1], c[2]);
                Philox pr{(uint32_t)a.seed ^ 0x52554c45u /* "RULE" domain */, (uint32_t)(a.seed >> 32)};
                uint32_t e[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)pidx, (uint32_t)(pidx >> 32)};
                pr.gen(e);
                d[0] = e[0]; d[1] = e[1]; d[2] = e[2]; d[3] = e[3];
            }
            Proposal p;
            make_proposal<DOUBLED>(a, op, a.replay != 0, d, rb, cb, p);
            // change of the site values and of the linear term
            double dvs[MAXFLIP];
            double dlp = 0.0;
            for (int i = 0; i < p.n; i++) {
                const int side = p.idx[i] >= N ? 1 : 0, j = p.idx[i] - side * N;
                dvs[i] = (double)flip_delta<T>(a.hilb, get_bit(side ? cb : rb, j));
                dlp += dvs[i] * un.lin_slope(side, j);
            }
            const int nxt = cur ^ 1;
            double acc_sum = 0.0;
            if (p.n > 0) {
                for (int k = lane; k < U; k += 32) {
                    cxd t = th[cur * U + k];
                    for (int i = 0; i < p.n; i++) {
                        const int side = p.idx[i] >= N ? 1 : 0, j = p.idx[i] - side * N;
                        t += rscale(dvs[i], un.slope(k, side, j));
                    }
                    const double f = un.weight(k) * unit_value<ACT>(t, KIND == NQ_NDM && k < 2 * a.M);
                    acc_sum += f - fv[cur * U + k];
                    th[nxt * U + k] = t; fv[nxt * U + k] = f;
                }
                acc_sum = warp_sum(acc_sum);
            }
            const double pr_ratio = exp(acc_sum + dlp + p.bias);
            const bool acc = (u - (T)pr_ratio) < T(0);
            if (acc) {
                if (p.n > 0) cur = nxt;
                for (int i = 0; i < p.n; i++) {
                    const int side = p.idx[i] >= N ? 1 : 0, j = p.idx[i] - side * N;
                    toggle(side ? cb : rb, j);
                }
                nacc++;
            }
            __syncwarp();
            if (a.replay && a.accept_out && lane == 0) a.accept_out[(int64_t)pic * a.B + chain] = acc ? 1 : 0;
        }
        if (step >= a.burn && a.out_prow && lane == 0) {
            int64_t o = ((int64_t)(step - a.burn) * a.B + chain) * W64;
            for (int w = 0; w < W64; w++) { a.out_prow[o + w] = rb[w]; if (DOUBLED && a.out_pcol) a.out_pcol[o + w] = cb[w]; }
        }
    }
    if (lane == 0) {
        for (int w = 0; w < W64; w++) { st_row[chain * W64 + w] = rb[w]; if (DOUBLED) st_col[chain * W64 + w] = cb[w]; }
        if (nacc) atomicAdd(a.accepted, (unsigned long long)nacc);
    }
}

template <typename E, int KIND, int ACT>
int launch_sampler_rule(nq_sampler_t s, const RunArgs& a) {
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    const int U = KIND == NQ_NDM ? 2 * m->M + m->A : m->M;
    const size_t per_warp = (size_t)3 * U * sizeof(cxd);
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > 96 * 1024) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    if (smem > ctx->smem_optin) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "sampler needs %zu B shared memory per chain", per_warp);
    OpDev op;
    memset(&op, 0, sizeof op);
    if (s->rule == NQ_RULE_OPERATOR) op = op_dev(s->rule_op);
    auto kern = sampler_rule_kernel<E, KIND, ACT>;
    NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)((s->B + wpb - 1) / wpb);
    NQ_LAUNCH(ctx, kern, grid, wpb * 32, smem, (const E*)m->params, op, s->prow, s->pcol, a);
    return NQ_OK;
}

template <typename E>
int dispatch_sampler_rule(nq_sampler_t s, const RunArgs& a) {
    nq_machine_t m = s->m;
    if (m->kind == NQ_RBMSPLIT) return launch_sampler_rule<E, NQ_RBMSPLIT, NQ_SOFTPLUS>(s, a);
    if (m->kind == NQ_RBM)
        return m->act == NQ_SOFTPLUS ? launch_sampler_rule<E, NQ_RBM, NQ_SOFTPLUS>(s, a) : launch_sampler_rule<E, NQ_RBM, NQ_LOGCOSH>(s, a);
    return nq_fail(m->ctx, NQ_ERR_ARG, "machine kind / dtype mismatch");
}

int run_sampler_rule(nq_sampler_t s, const RunArgs& a) {
    nq_machine_t m = s->m;
    if (m->kind == NQ_NDM) {
        if (m->dtype == NQ_F64)
            return m->act == NQ_SOFTPLUS ? launch_sampler_rule<double, NQ_NDM, NQ_SOFTPLUS>(s, a) : launch_sampler_rule<double, NQ_NDM, NQ_LOGCOSH>(s, a);
        return m->act == NQ_SOFTPLUS ? launch_sampler_rule<float, NQ_NDM, NQ_SOFTPLUS>(s, a) : launch_sampler_rule<float, NQ_NDM, NQ_LOGCOSH>(s, a);
    }
    switch (m->dtype) {
        case NQ_F32: return dispatch_sampler_rule<float>(s, a);
        case NQ_F64: return dispatch_sampler_rule<double>(s, a);
        case NQ_C64: return dispatch_sampler_rule<cxf>(s, a);
        default: return dispatch_sampler_rule<cxd>(s, a);
    }
}

// NDM with the tracked f' values in REGISTERS (N <= 64 sites: the configuration is two 64-bit registers; M <= 32 KM
// hidden units and A <= 32 KA ancillas): lane l owns lambda units l + 32 u of BOTH sides and ancillas l + 32 u.  Same
// recurrence as sampler_ndm_kernel; the factors and numerators of the current proposal stay in registers, so an accepted
// move re-reads nothing and the reciprocals skip the slow path (fast_rcp).
template <typename T, int ACT, int KM, int KA>
__global__ void __launch_bounds__(256) sampler_ndm_reg_kernel(const T* __restrict__ par, const T* __restrict__ tabr,
                                                               const cx<T>* __restrict__ tabc, uint64_t* __restrict__ st_row,
                                                               uint64_t* __restrict__ st_col, RunArgs a) {
    typedef cx<T> C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)wpb + warp;
    const int N = a.N, M = a.M, A = a.A;
    const int64_t MN = (int64_t)M * N, AN = (int64_t)A * N;
    const int64_t o_wmu = N + M, o_umu = o_wmu + MN, o_blam = o_umu + AN,
                  o_hlam = o_blam + N, o_dlam = o_hlam + M, o_wlam = o_dlam + A, o_ulam = o_wlam + MN;
    T* eb = (T*)smem_raw;                                       // eb[sign * N + j] = exp(dv b_lam_j)
    for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
        const int sgn = i / N, jj = i - sgn * N;
        const T dvv = a.hilb == NQ_SPIN ? (sgn ? T(-2) : T(2)) : (sgn ? T(-1) : T(1));
        eb[i] = (T)exp((double)(dvv * par[o_blam + jj]));
    }
    __syncthreads();
    if (chain >= a.B) return;
    T* vsw = (T*)(smem_raw + ((size_t)2 * N * sizeof(T) + 15) / 16 * 16) + (size_t)warp * 2 * N;
    const T half = T(0.5);
    uint64_t rb = st_row[chain], cb = st_col[chain];
    unsigned nacc = 0;
    const int nsteps = a.burn + a.L;
    T slr[KM], slc[KM];
    C spi[KA];
    for (int step = 0; step < nsteps; step++) {
        if (step % REFRESH_STEPS == 0) {
            for (int i = lane; i < 2 * N; i += 32) vsw[i] = digit_value<T>(a.hilb, (int)(((i < N ? rb : cb) >> (i < N ? i : i - N)) & 1ull));
            __syncwarp();
#pragma unroll
            for (int u = 0; u < KM; u++) {
                const int k = lane + 32 * u;
                T t = T(0), tp = T(0);
                if (k < M) {
                    const T* __restrict__ w = par + o_wlam + k;
                    t = tp = par[o_hlam + k];
                    for (int j = 0; j < N; j++) {
                        const T wv = w[(int64_t)M * j];
                        t += wv * vsw[j];
                        tp += wv * vsw[N + j];
                    }
                }
                T f, d, fp, dp;
                act_eval<ACT>(t, f, d);
                act_eval<ACT>(tp, fp, dp);
                slr[u] = d; slc[u] = dp;
            }
#pragma unroll
            for (int u = 0; u < KA; u++) {
                const int q = lane + 32 * u;
                T pr = T(0), pim = T(0);
                if (q < A) {
                    pr = par[o_dlam + q];
                    for (int j = 0; j < N; j++) {
                        const T x = vsw[j], y = vsw[N + j];
                        pr += half * par[o_ulam + q + (int64_t)A * j] * (x + y);
                        pim += half * par[o_umu + q + (int64_t)A * j] * (x - y);
                    }
                }
                C f, d;
                act_eval<ACT>(C(pr, pim), f, d);
                spi[u] = d;
            }
            __syncwarp();
        }
        DrawBatch<T> db;
#pragma unroll 1
        for (int ps = 0; ps < a.passes; ps++) {
            const int pic = step * a.passes + ps;
            if ((ps & 31) == 0) db.fill(a, chain, pic, a.passes - ps, a.diag ? N : 2 * N, lane);
            int site; T u01;
            db.get(ps & 31, site, u01);
            const bool col = !a.diag && site >= N;
            const int j = col ? site - N : site;
            const int bit = (int)(((col ? cb : rb) >> j) & 1ull);
            const int sg = bit;                                     // digit 0 -> the value increases (sign 0), 1 -> decreases
            const T* __restrict__ tp = tabr + (int64_t)sg * 2 * MN + (int64_t)M * j + lane;          // lambda layer = lay 0
            const T* __restrict__ tm = tabr + (int64_t)(1 - sg) * 2 * MN + (int64_t)M * j + lane;
            const C* __restrict__ cp = tabc + (int64_t)sg * AN + (int64_t)A * j + lane;
            const C* __restrict__ cm = tabc + (int64_t)(1 - sg) * AN + (int64_t)A * j + lane;
            T facl[KM], numl[KM];
            C facp[KA], nump[KA];
            double pl = 1.0;
#pragma unroll
            for (int u = 0; u < KM; u++) {
                const bool on = lane + 32 * u < M;
                const T Ep = on ? tp[32 * u] : T(1), Em = (ACT == NQ_LOGCOSH && on) ? tm[32 * u] : T(1);
                ratio_parts<ACT>(col ? slc[u] : slr[u], Ep, Em, facl[u], numl[u]);
                pl *= (double)facl[u];
            }
#pragma unroll
            for (int u = 0; u < KA; u++) {
                const bool on = lane + 32 * u < A;
                C Ep = on ? cp[32 * u] : C(T(1), T(0)), Em = (ACT == NQ_LOGCOSH && on) ? cm[32 * u] : C(T(1), T(0));
                if (col) { Ep.im = -Ep.im; Em.im = -Em.im; }
                if (a.diag) {
                    Ep = C(Ep.re * Ep.re + Ep.im * Ep.im, T(0));
                    Em = C(Em.re * Em.re + Em.im * Em.im, T(0));
                }
                ratio_parts<ACT>(spi[u], Ep, Em, facp[u], nump[u]);
                pl *= a.diag ? (double)facp[u].re : abs2_d(facp[u]);
            }
            pl = warp_prod(pl);
            const double pr = pl * (double)eb[sg * N + j];
            const bool acc = (u01 - (T)pr) < T(0);
            if (acc) {
#pragma unroll
                for (int u = 0; u < KM; u++) {
                    const T sn = q_div(numl[u], ACT == NQ_SOFTPLUS ? facl[u] : T(2) * facl[u]);
                    if (col) slc[u] = sn; else slr[u] = sn;
                    if (a.diag) slc[u] = sn;
                }
#pragma unroll
                for (int u = 0; u < KA; u++) spi[u] = q_div(nump[u], ACT == NQ_SOFTPLUS ? facp[u] : rscale(T(2), facp[u]));
                if (col || a.diag) cb ^= 1ull << j;
                if (!col) rb ^= 1ull << j;
                nacc++;
            }
            if (a.replay && a.accept_out && lane == 0) a.accept_out[(int64_t)pic * a.B + chain] = acc ? 1 : 0;
        }
        if (step >= a.burn && a.out_prow && lane == 0) {
            const int64_t o = (int64_t)(step - a.burn) * a.B + chain;
            a.out_prow[o] = rb;
            if (a.out_pcol) a.out_pcol[o] = cb;
        }
    }
    if (lane == 0) {
        st_row[chain] = rb; st_col[chain] = cb;
        if (nacc) atomicAdd(a.accepted, (unsigned long long)nacc);
    }
}

template <typename T, int ACT, int KM, int KA>
int launch_sampler_ndm_reg(nq_sampler_t s, const RunArgs& a) {
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    const int wpb = 8;
    const size_t smem = ((size_t)2 * m->N * sizeof(T) + 15) / 16 * 16 + (size_t)wpb * 2 * m->N * sizeof(T);
    const cx<T>* tabc = (const cx<T>*)((const char*)m->etab + ((size_t)4 * m->M * m->N * sizeof(T) + 15) / 16 * 16);
    auto kern = sampler_ndm_reg_kernel<T, ACT, KM, KA>;
    unsigned grid = (unsigned)((s->B + wpb - 1) / wpb);
    NQ_LAUNCH(ctx, kern, grid, wpb * 32, smem, (const T*)m->params, (const T*)m->etab, tabc, s->prow, s->pcol, a);
    return NQ_OK;
}

// M, A <= 64 and N <= 64: the register kernel; returns false when the shape needs the shared-memory kernel
template <typename T, int ACT>
bool try_sampler_ndm_reg(nq_sampler_t s, const RunArgs& a, int* status) {
    nq_machine_t m = s->m;
    if (m->N > 64 || m->M > 64 || m->A > 64) return false;
    const int km = m->M <= 32 ? 1 : 2, ka = m->A <= 32 ? 1 : 2;
    if (km == 1 && ka == 1) *status = launch_sampler_ndm_reg<T, ACT, 1, 1>(s, a);
    else if (km == 1) *status = launch_sampler_ndm_reg<T, ACT, 1, 2>(s, a);
    else if (ka == 1) *status = launch_sampler_ndm_reg<T, ACT, 2, 1>(s, a);
    else *status = launch_sampler_ndm_reg<T, ACT, 2, 2>(s, a);
    return true;
}

template <typename E, int ACT, bool DOUBLED, int KU>
int launch_sampler_rbm_ku(nq_sampler_t s, const RunArgs& a) {
    typedef typename elem_traits<E>::real T;
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    const int nsv = (DOUBLED ? 2 : 1) * m->N;
    const int wpb = 8;
    const size_t smem = ((size_t)4 * m->N * sizeof(T) + 15) / 16 * 16 + (size_t)wpb * ((nsv + 1) & ~1) * sizeof(T);
    if (smem > ctx->smem_optin) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "sampler needs %zu B shared memory", smem);
    auto kern = sampler_rbm_reg_kernel<E, ACT, DOUBLED, KU>;
    NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)((s->B + wpb - 1) / wpb);
    NQ_LAUNCH(ctx, kern, grid, wpb * 32, smem, (const E*)m->params, (const E*)nq_machine_qtab(m), s->prow, s->pcol, a);
    return NQ_OK;
}

template <typename E, int ACT, bool DOUBLED>
int launch_sampler_rbm(nq_sampler_t s, const RunArgs& a) {
    const int M = s->m->M;
    if (M <= 32) return launch_sampler_rbm_ku<E, ACT, DOUBLED, 1>(s, a);
    if (M <= 64) return launch_sampler_rbm_ku<E, ACT, DOUBLED, 2>(s, a);
    if (M <= 96) return launch_sampler_rbm_ku<E, ACT, DOUBLED, 3>(s, a);
    if (M <= 128) return launch_sampler_rbm_ku<E, ACT, DOUBLED, 4>(s, a);
    if (M <= 160) return launch_sampler_rbm_ku<E, ACT, DOUBLED, 5>(s, a);
    if (M <= 256) return launch_sampler_rbm_ku<E, ACT, DOUBLED, 8>(s, a);
    if (M <= 512) return launch_sampler_rbm_ku<E, ACT, DOUBLED, 16>(s, a);
    if (M <= 1024) return launch_sampler_rbm_ku<E, ACT, DOUBLED, 32>(s, a);
    return nq_fail(s->m->ctx, NQ_ERR_UNSUPPORTED, "sampler supports M <= 1024 hidden units (M = %d)", M);
}

template <typename T, int ACT>
int launch_sampler_ndm(nq_sampler_t s, const RunArgs& a) {
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    static const bool smem_only = [] { const char* e = getenv("NQ_SAMPLER_NDM"); return e && !strcmp(e, "smem"); }();
    int st = NQ_OK;
    if (!smem_only && try_sampler_ndm_reg<T, ACT>(s, a, &st)) return st;
    size_t per_warp = ((size_t)(4 * m->M + 2 * m->N) * sizeof(T) + (size_t)3 * m->A * sizeof(cx<T>) + 15) / 16 * 16;
    const size_t tab_bytes = ((size_t)2 * m->N * sizeof(T) + 15) / 16 * 16;
    const cx<T>* tabc = (const cx<T>*)((const char*)m->etab + ((size_t)4 * m->M * m->N * sizeof(T) + 15) / 16 * 16);
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > 96 * 1024) wpb >>= 1;
    size_t smem = tab_bytes + per_warp * wpb;
    if (smem > ctx->smem_optin) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "sampler needs %zu B shared memory per chain", per_warp);
    auto kern = sampler_ndm_kernel<T, ACT>;
    NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)((s->B + wpb - 1) / wpb);
    NQ_LAUNCH(ctx, kern, grid, wpb * 32, smem, (const T*)m->params, (const T*)m->etab, tabc, s->prow, s->pcol, a);
    return NQ_OK;
}

template <typename E>
int dispatch_sampler_rbm(nq_sampler_t s, const RunArgs& a) {
    nq_machine_t m = s->m;
    if (m->kind == NQ_RBMSPLIT) return launch_sampler_rbm<E, NQ_SOFTPLUS, true>(s, a);
    if (m->act == NQ_SOFTPLUS) return launch_sampler_rbm<E, NQ_SOFTPLUS, false>(s, a);
    return launch_sampler_rbm<E, NQ_LOGCOSH, false>(s, a);
}

int run_sampler(nq_sampler_t s, const RunArgs& a) {
    nq_machine_t m = s->m;
    if (m->N > 64 * MAXW) return nq_fail(m->ctx, NQ_ERR_UNSUPPORTED, "sampler supports N <= %d", 64 * MAXW);
    if (s->rule != NQ_RULE_LOCAL) {
        if (s->diag) return nq_fail(m->ctx, NQ_ERR_UNSUPPORTED, "the diagonal chain uses LocalRule");
        return run_sampler_rule(s, a);
    }
    NQ_CHECK(nq_machine_ensure_tables(m));
    if (m->kind == NQ_NDM) {
        if (m->dtype == NQ_F64)
            return m->act == NQ_SOFTPLUS ? launch_sampler_ndm<double, NQ_SOFTPLUS>(s, a) : launch_sampler_ndm<double, NQ_LOGCOSH>(s, a);
        return m->act == NQ_SOFTPLUS ? launch_sampler_ndm<float, NQ_SOFTPLUS>(s, a) : launch_sampler_ndm<float, NQ_LOGCOSH>(s, a);
    }
    switch (m->dtype) {
        case NQ_F32: return dispatch_sampler_rbm<float>(s, a);
        case NQ_F64: return dispatch_sampler_rbm<double>(s, a);
        case NQ_C64: return dispatch_sampler_rbm<cxf>(s, a);
        default: return dispatch_sampler_rbm<cxd>(s, a);
    }
}

RunArgs base_args(nq_sampler_t s) {
    RunArgs a;
    memset(&a, 0, sizeof a);
    nq_machine_t m = s->m;
    a.B = s->B; a.N = m->N; a.M = m->M; a.A = m->A; a.hilb = (int)m->hilb; a.passes = s->passes;
    a.seed = s->seed; a.pass_base = s->pass_base; a.chain_offset = s->chain_offset; a.accepted = s->accepted;
    a.diag = s->diag;
    a.rule = s->rule; a.n_coup = s->n_coup; a.coup = s->coup;
    return a;
}

}  // namespace

extern "C" int nq_sampler_create(nq_machine_t m, int64_t B, int passes, uint64_t seed, int64_t chain_offset,
                                 nq_sampler_t* out) {
    if (!m || !out) return NQ_ERR_ARG;
    *out = nullptr;
    nq_ctx_t ctx = m->ctx;
    if (B <= 0 || passes <= 0) return nq_fail(ctx, NQ_ERR_ARG, "B and passes must be positive");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    nq_sampler_t s = new nq_sampler_s();
    s->ctx = ctx; s->m = m; s->B = B;
    s->passes = (passes % 2 == 0) ? passes + 1 : passes;   // Metropolis.jl:30-38
    s->seed = seed; s->chain_offset = chain_offset; s->pass_base = 0; s->passes_done = 0; s->diag = 0;
    s->prow = s->pcol = nullptr; s->accepted = nullptr;
    s->rule = NQ_RULE_LOCAL; s->n_coup = 0; s->coup = nullptr; s->rule_op = nullptr;
    size_t pbytes = (size_t)B * nq_words(m->N) * 8;
    bool ok = cudaMalloc((void**)&s->prow, pbytes) == cudaSuccess &&
              (!m->doubled() || cudaMalloc((void**)&s->pcol, pbytes) == cudaSuccess) &&
              cudaMalloc((void**)&s->accepted, 8) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        cudaFree(s->prow); cudaFree(s->pcol); cudaFree(s->accepted);
        delete s;
        return nq_fail(ctx, NQ_ERR_ALLOC, "sampler allocation failed");
    }
    cudaMemsetAsync(s->prow, 0, pbytes, ctx->stream);
    if (s->pcol) cudaMemsetAsync(s->pcol, 0, pbytes, ctx->stream);
    cudaMemsetAsync(s->accepted, 0, 8, ctx->stream);
    *out = s;
    return NQ_OK;
}

extern "C" int nq_sampler_destroy(nq_sampler_t s) {
    if (!s) return NQ_ERR_ARG;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    cudaFree(s->prow); cudaFree(s->pcol); cudaFree(s->accepted); cudaFree(s->coup);
    delete s;
    return NQ_OK;
}

extern "C" int nq_sampler_set_rule(nq_sampler_t s, nq_rule rule, int n_couplings, const int32_t* couplings, nq_operator_t op) {
    if (!s) return NQ_ERR_ARG;
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (rule == NQ_RULE_EXCHANGE) {
        if (m->doubled()) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "ExchangeRule is defined for ket states (ExchangeRule.jl:46-48)");
        if (n_couplings <= 0 || !couplings) return nq_fail(ctx, NQ_ERR_ARG, "ExchangeRule needs at least one coupling");
    } else if (rule == NQ_RULE_NAGY) {
        if (!m->doubled()) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "NagyRule is defined for doubled states (Nagy.jl:40)");
        // adjacency_list[site] indexes the list of couples by SITE number (Nagy.jl:51): fewer couples than sites is a
        // BoundsError in the reference
        if (n_couplings < m->N || !couplings) return nq_fail(ctx, NQ_ERR_ARG, "NagyRule indexes its couplings by site: %d couplings < %d sites", n_couplings, m->N);
    } else if (rule == NQ_RULE_OPERATOR) {
        if (!op) return nq_fail(ctx, NQ_ERR_ARG, "OperatorRule needs an operator");
        if (op->N != m->N || (op->space == NQ_SUPER) != m->doubled())
            return nq_fail(ctx, NQ_ERR_SHAPE, "OperatorRule: the operator does not act on the machine's space");
        if (op->max_part_sites > MAXFLIP / 2) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "OperatorRule: parts of more than %d sites", MAXFLIP / 2);
    } else if (rule != NQ_RULE_LOCAL) {
        return nq_fail(ctx, NQ_ERR_ARG, "unknown rule %d", (int)rule);
    }
    if (rule == NQ_RULE_EXCHANGE || rule == NQ_RULE_NAGY) {
        std::vector<int32_t> h(2 * (size_t)n_couplings);
        NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (nq_is_device_ptr(couplings)) NQ_CUDA(ctx, cudaMemcpy(h.data(), couplings, h.size() * 4, cudaMemcpyDeviceToHost));
        else memcpy(h.data(), couplings, h.size() * 4);
        for (int32_t v : h) if (v < 0 || v >= m->N) return nq_fail(ctx, NQ_ERR_ARG, "coupling site %d outside 0..%d", (int)v, m->N - 1);
        cudaFree(s->coup); s->coup = nullptr;
        NQ_CUDA(ctx, cudaMalloc((void**)&s->coup, h.size() * 4));
        NQ_CUDA(ctx, cudaMemcpy(s->coup, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        s->n_coup = n_couplings;
    }
    s->rule = (int)rule;
    s->rule_op = rule == NQ_RULE_OPERATOR ? op : nullptr;
    return NQ_OK;
}

extern "C" int nq_sampler_replay_rule(nq_sampler_t s, const int32_t* draws, const void* uniforms, uint8_t* accept_out) {
    if (!s || !draws || !uniforms) return NQ_ERR_ARG;
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    if (s->rule == NQ_RULE_LOCAL) return nq_fail(ctx, NQ_ERR_ARG, "LocalRule replays through nq_sampler_replay");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    size_t n = (size_t)s->passes * s->B;
    RunArgs a = base_args(s);
    a.replay = 1; a.burn = 0; a.L = 1;
    a.draws = (const int32_t*)st.in(SL_IN0, draws, n * 16);
    a.uniforms = st.in(SL_IN1, uniforms, n * nq_dtype_size(nq_real_of(m->dtype)));
    a.accept_out = accept_out ? (uint8_t*)st.out(SL_OUT0, accept_out, n) : nullptr;
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(run_sampler(s, a));
    s->passes_done += (int64_t)n;
    return st.finish();
}

extern "C" int nq_sampler_set_state(nq_sampler_t s, const void* srow, const void* scol, nq_dtype sdtype) {
    if (!s || !srow) return NQ_ERR_ARG;
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    if (s->diag) scol = srow;             // diagonal chain: sigma' = sigma
    if (m->doubled() != (scol != nullptr)) return nq_fail(ctx, NQ_ERR_ARG, "row/col configuration mismatch");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    size_t fbytes = (size_t)s->B * m->N * nq_dtype_size(sdtype);
    const void* dr = st.in(SL_IN0, srow, fbytes);
    const void* dc = scol ? st.in(SL_IN1, scol, fbytes) : nullptr;
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(nq_pack_device(ctx, m->hilb, m->N, s->B, dr, sdtype, s->prow));
    if (scol) NQ_CHECK(nq_pack_device(ctx, m->hilb, m->N, s->B, dc, sdtype, s->pcol));
    NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NQ_OK;
}

extern "C" int nq_sampler_get_state(nq_sampler_t s, void* srow, void* scol, nq_dtype sdtype) {
    if (!s || !srow) return NQ_ERR_ARG;
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    if (m->doubled() != (scol != nullptr)) return nq_fail(ctx, NQ_ERR_ARG, "row/col configuration mismatch");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    size_t fbytes = (size_t)s->B * m->N * nq_dtype_size(sdtype);
    void* dr = st.out(SL_OUT0, srow, fbytes);
    void* dc = scol ? st.out(SL_OUT1, scol, fbytes) : nullptr;
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(nq_unpack_device(ctx, m->hilb, m->N, s->B, s->prow, dr, sdtype));
    if (scol) NQ_CHECK(nq_unpack_device(ctx, m->hilb, m->N, s->B, s->pcol, dc, sdtype));
    return st.finish();
}

extern "C" int nq_sampler_randomize(nq_sampler_t s) {
    if (!s) return NQ_ERR_ARG;
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    int W64 = nq_words(m->N);
    int64_t n = s->B * W64;
    NQ_LAUNCH(ctx, randomize_kernel, (unsigned)((n + 255) / 256), 256, 0, s->prow, s->pcol, s->B, m->N, W64, s->seed,
              s->chain_offset, s->pass_base);
    if (s->diag) NQ_CUDA(ctx, cudaMemcpyAsync(s->pcol, s->prow, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return NQ_OK;
}

extern "C" int nq_sampler_set_mode(nq_sampler_t s, int diagonal) {
    if (!s) return NQ_ERR_ARG;
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    if (diagonal && m->kind != NQ_NDM)
        return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "the diagonal chain needs a density matrix with a real positive diagonal (NDM)");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    s->diag = diagonal ? 1 : 0;
    if (s->diag)
        NQ_CUDA(ctx, cudaMemcpyAsync(s->pcol, s->prow, (size_t)s->B * nq_words(m->N) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return NQ_OK;
}

extern "C" int nq_sampler_replay(nq_sampler_t s, const int32_t* sites, const void* uniforms, uint8_t* accept_out) {
    if (!s || !sites || !uniforms) return NQ_ERR_ARG;
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    size_t n = (size_t)s->passes * s->B;
    RunArgs a = base_args(s);
    a.replay = 1; a.burn = 0; a.L = 1;
    a.sites = (const int32_t*)st.in(SL_IN0, sites, n * 4);
    a.uniforms = st.in(SL_IN1, uniforms, n * nq_dtype_size(nq_real_of(m->dtype)));
    a.accept_out = accept_out ? (uint8_t*)st.out(SL_OUT0, accept_out, n) : nullptr;
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(run_sampler(s, a));
    s->passes_done += (int64_t)n;
    return st.finish();
}

extern "C" int nq_sampler_sample(nq_sampler_t s, int burn, int L, uint64_t* prow, uint64_t* pcol, void* srow,
                                 void* scol, nq_dtype sdtype) {
    if (!s || burn < 0 || L < 0) return NQ_ERR_ARG;
    nq_machine_t m = s->m;
    nq_ctx_t ctx = m->ctx;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (burn + L == 0) return NQ_OK;
    NqStage st(ctx);
    const int W64 = nq_words(m->N);
    size_t pbytes = (size_t)(L ? L : 1) * s->B * W64 * 8;
    size_t fbytes = (size_t)L * s->B * m->N * nq_dtype_size(sdtype);
    RunArgs a = base_args(s);
    a.replay = 0; a.burn = burn; a.L = L;
    const bool want = L > 0 && (prow || srow);
    if (want) {
        a.out_prow = prow ? (uint64_t*)st.out(SL_OUT0, prow, pbytes) : (uint64_t*)nq_scratch(ctx, SL_W0, pbytes);
        if (m->doubled()) a.out_pcol = pcol ? (uint64_t*)st.out(SL_OUT1, pcol, pbytes) : (uint64_t*)nq_scratch(ctx, SL_W1, pbytes);
        if (!a.out_prow || (m->doubled() && !a.out_pcol)) return NQ_ERR_ALLOC;
    }
    void* dsr = (L > 0 && srow) ? st.out(SL_OUT2, srow, fbytes) : nullptr;
    void* dsc = (L > 0 && scol && m->doubled()) ? st.out(SL_OUT3, scol, fbytes) : nullptr;
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(run_sampler(s, a));
    s->pass_base += (uint64_t)(burn + L) * s->passes;
    s->passes_done += (int64_t)(burn + L) * s->passes * s->B;
    if (dsr) NQ_CHECK(nq_unpack_device(ctx, m->hilb, m->N, (int64_t)L * s->B, a.out_prow, dsr, sdtype));
    if (dsc) NQ_CHECK(nq_unpack_device(ctx, m->hilb, m->N, (int64_t)L * s->B, a.out_pcol, dsc, sdtype));
    return st.finish();
}

extern "C" int nq_sampler_counters(nq_sampler_t s, int64_t* passes_done, int64_t* passes_accepted) {
    if (!s) return NQ_ERR_ARG;
    nq_ctx_t ctx = s->m->ctx;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long acc = 0;
    NQ_CUDA(ctx, cudaMemcpyAsync(&acc, s->accepted, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (passes_done) *passes_done = s->passes_done;
    if (passes_accepted) *passes_accepted = (int64_t)acc;
    return NQ_OK;
}
