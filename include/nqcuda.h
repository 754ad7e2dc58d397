/*
 * nqcuda.h -- C ABI of libnqcuda, the B200 (sm_100a) implementation of NeuralQuantum.jl's
 * variational-Monte-Carlo hot path.
 *
 * Every entry point replaces one interface of the reference (cited "ref:" with file:line relative
 * to the NeuralQuantum.jl v0.2.0 checkout) and is what a Julia `ccall` shim binds.
 *
 * Conventions
 *   - All functions return an int status: NQ_OK (0) or a negative NQ_ERR_* code.  No exception or
 *     exit crosses the boundary; nq_last_error(ctx) returns the message of the last failure.
 *   - Arrays are dense, COLUMN-MAJOR and 0-offset exactly as Julia lays them out:
 *       states sigma [N, B]   (site fastest), values -1/+1 (HomogeneousSpin) or 0/1 (HomogeneousFock)
 *       log psi      [B]
 *       O = grad     [P, B]   (parameter fastest; leading dimension ldO >= P), functor order:
 *                     RBM a,b,W | RBMSplit ar,ac,b,Wr,Wc | NDM b_mu,h_mu,w_mu,u_mu,b_lam,h_lam,d_lam,w_lam,u_lam
 *                     with matrices W[M,N] flattened column-major (k + M*j)
 *       S            [P, P]
 *     Complex numbers are interleaved (re, im).
 *   - Every data pointer may be a HOST pointer or a DEVICE pointer (detected with
 *     cudaPointerGetAttributes).  Host buffers are staged through device scratch inside the call
 *     and the call returns after the results are back on the host.  With device pointers the call
 *     only enqueues work on the context's stream.
 *   - Caller owns all buffers it passes; the library never keeps a caller pointer after return.
 *     Device state (parameters, tables, chains, workspaces) lives behind opaque handles.
 *   - Handles are not thread-safe; distinct contexts may be used from distinct threads.
 *   - There is NO CPU fallback: without a CUDA device nq_ctx_create fails with NQ_ERR_CUDA.
 */
#ifndef NQCUDA_H
#define NQCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NQ_VERSION 100

/* ---- status codes (ref: Julia exceptions; SRDirect.jl:79 PosDefException; SRIterative.jl:133-150) */
enum {
    NQ_OK = 0,
    NQ_ERR_ARG = -1,            /* bad argument / null handle                      */
    NQ_ERR_SHAPE = -2,          /* inconsistent sizes                              */
    NQ_ERR_CUDA = -3,           /* CUDA runtime failure (message in nq_last_error) */
    NQ_ERR_NCCL = -4,           /* NCCL failure or communicator not initialised    */
    NQ_ERR_NOT_POSDEF = -5,     /* Cholesky met a non-positive pivot (info = index)*/
    NQ_ERR_NOT_CONVERGED = -6,  /* CG hit maxiter; dw is still returned            */
    NQ_ERR_UNSUPPORTED = -7,    /* valid request outside the built path            */
    NQ_ERR_ALLOC = -8
};

/* ---- enumerations */
typedef enum { NQ_RBM = 0, NQ_RBMSPLIT = 1, NQ_NDM = 2 } nq_machine_kind;
typedef enum { NQ_SOFTPLUS = 0, NQ_LOGCOSH = 1 } nq_activation;          /* ref: Networks/activation.jl:5-29 */
typedef enum { NQ_F32 = 0, NQ_F64 = 1, NQ_C64 = 2, NQ_C128 = 3 } nq_dtype;
typedef enum { NQ_SPIN = 0, NQ_FOCK = 1 } nq_hilbert;                    /* values -1/+1 | 0/1, local dim 2 */
typedef enum { NQ_KET = 0, NQ_SUPER = 1 } nq_space;                      /* H on sigma | Liouvillian on (sigma,sigma') */
/* ref: SR/SR.jl sr_cholesky | sr_cg | sr_minres | sr_qlp.  NQ_SOLVE_QLP_WARM = MINRES-QLP started from the dw passed in
 * (the warm-started restarts of SRIterative.jl:133-150). */
typedef enum { NQ_SOLVE_CHOLESKY = 0, NQ_SOLVE_CG = 1, NQ_SOLVE_MINRES = 2, NQ_SOLVE_QLP = 3, NQ_SOLVE_QLP_WARM = 4 } nq_solver;

/* ref: Samplers/MCMCRules/{LocalRule,ExchangeRule,Nagy,OperatorRule}.jl */
typedef enum { NQ_RULE_LOCAL = 0, NQ_RULE_EXCHANGE = 1, NQ_RULE_NAGY = 2, NQ_RULE_OPERATOR = 3 } nq_rule;

typedef struct nq_ctx_s* nq_ctx_t;
typedef struct nq_machine_s* nq_machine_t;
typedef struct nq_operator_s* nq_operator_t;
typedef struct nq_sampler_s* nq_sampler_t;
typedef struct nq_symm_s* nq_symm_t;

/* ======================================================================================
 * Context.  One context = one device + one stream.  ref: one sampler object per task/rank
 * (Parallel/Threads/threads_wrappers.jl:14-31, Parallel/MPI/mpi.jl).
 * `stream` is a cudaStream_t used as given for every launch and copy (NULL = the CUDA default stream).
 * ==================================================================================== */
int nq_version(void);
const char* nq_status_string(int status);
int nq_ctx_create(int device, void* stream, nq_ctx_t* out);
int nq_ctx_destroy(nq_ctx_t ctx);
const char* nq_last_error(nq_ctx_t ctx);
int nq_ctx_sync(nq_ctx_t ctx);
/* number of kernels this library launched on the context so far (bench.py "gpu_launches") */
int nq_ctx_launch_count(nq_ctx_t ctx, uint64_t* out);
/* `info` of the last NQ_ERR_NOT_POSDEF (failing pivot, 0-based) */
int nq_ctx_last_info(nq_ctx_t ctx, int64_t* out);

/* ======================================================================================
 * Configurations: reference float arrays <-> bit-packed device words.
 * Packed layout: uint64 words[B][W64], W64 = ceil(N/64); bit (j%64) of word j/64 = local digit of
 * site j (0-based): digit = (v+1)/2 for NQ_SPIN, v for NQ_FOCK.
 * ref: States/States.jl:12-32 (dense [N,B,L] float arrays), Hilbert/HomogeneousSpin.jl:156-179.
 * `sdtype` is the element type of the float arrays (NQ_F32 or NQ_F64).
 * ==================================================================================== */
int nq_states_words(int N);
int nq_pack_states(nq_ctx_t ctx, nq_hilbert h, int N, int64_t B, const void* sigma, nq_dtype sdtype,
                   uint64_t* packed);
int nq_unpack_states(nq_ctx_t ctx, nq_hilbert h, int N, int64_t B, const uint64_t* packed,
                     void* sigma, nq_dtype sdtype);

/* ======================================================================================
 * Machines.  ref: machine plugin interface, base_batched_networks.jl:20-104
 *   cache(net,batch_sz) / logpsi!(out,net,cache,sigma...) / logpsi_and_grad!(grad,out,net,cache,sigma...)
 * RBM:      Networks/ClosedSystems/RBM.jl:50-51, RBMBatched.jl:37-91
 * RBMSplit: Networks/MixedDensityMatrix/RBMSplit.jl:49-50, RBMSplitBatched.jl:35-101
 * NDM:      Networks/MixedDensityMatrix/NDM.jl:76-96, NDMBatched.jl:94-280
 * dtype = parameter type: NQ_F32/NQ_F64 (real weights) or NQ_C64/NQ_C128 (complex weights);
 * NDM takes real parameters only and outputs complex.  out_type: RBM/RBMSplit = dtype, NDM = complex.
 * M = hidden units, A = ancillas (NDM only, else 0).  RBMSplit ignores `act` (softplus hard-wired).
 * ==================================================================================== */
int nq_machine_create(nq_ctx_t ctx, nq_machine_kind kind, nq_hilbert h, int N, int M, int A,
                      nq_activation act, nq_dtype dtype, nq_machine_t* out);
int nq_machine_destroy(nq_machine_t m);
int nq_machine_nparams(nq_machine_t m, int64_t* P);
int nq_machine_out_dtype(nq_machine_t m, nq_dtype* out);   /* ref: out_type(net) */
/* flat parameter vector in functor order, P elements of the machine dtype. ref: functor.jl:41, utils/loading.jl:8-14 */
int nq_machine_set_params(nq_machine_t m, const void* params, int64_t P);
int nq_machine_get_params(nq_machine_t m, void* params, int64_t P);

/* logpsi!(out, net, cache, sigma[, sigma'])  -- `scol` must be NULL for RBM. out: [B] of out_type.
 * ref: base_batched_networks.jl:35-40, RBMBatched.jl:37-56, RBMSplitBatched.jl:35-62, NDMBatched.jl:94-175 */
int nq_logpsi(nq_machine_t m, const void* srow, const void* scol, nq_dtype sdtype, int64_t B, void* out);
/* log_prob_psi!: 2*Re(log psi).  ref: base_batched_networks.jl:255-259.  out: [B] real */
int nq_log_prob(nq_machine_t m, const void* srow, const void* scol, nq_dtype sdtype, int64_t B, void* out);
/* logpsi_and_grad!: out [B], O [P,B] with leading dimension ldO (elements of out_type) */
int nq_logpsi_grad(nq_machine_t m, const void* srow, const void* scol, nq_dtype sdtype, int64_t B,
                   void* out, void* O, int64_t ldO);
/* same on bit-packed configurations already on the device (no conversion pass) */
int nq_logpsi_packed(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B, void* out);
int nq_logpsi_grad_packed(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B,
                          void* out, void* O, int64_t ldO);

/* ======================================================================================
 * Operators: per-local-row connection tables.
 * ref: Operators/Operators/KLocalOperator.jl:54-112 (tables), :183-199 (enumeration),
 *      KLocalOperatorSum.jl:65-72, KLocalOperatorTensor.jl:129-157, KLocalLiouvillian.jl:46-52.
 * A "part" is one KLocalOperator: k sites (0-based, local digit i <-> part_sites[i]) and 2^k rows;
 * row r holds entries [row_ptr[r], row_ptr[r+1]) = (mel, local flip mask: bit i set <=> site i changes).
 * Rows of all parts are concatenated in part order (row_ptr has sum_p 2^k_p + 1 entries).  Entry 0 of
 * every row is the diagonal one (flip mask 0).
 * A "term" is visited in order and references a left part (acts on sigma; row index from sigma) and/or a
 * right part (acts on sigma'), -1 for identity; with both, entries are enumerated left-major and
 * mel = mel_l*mel_r (KLocalOperatorTensor.jl:146-154).  For NQ_KET only term_left is used.
 * ==================================================================================== */
int nq_operator_create(nq_ctx_t ctx, nq_space space, int N,
                       int n_parts, const int32_t* part_nsites, const int32_t* part_sites,
                       const int64_t* row_ptr, const double* entry_mel /* complex128 */,
                       const uint32_t* entry_flip,
                       int n_terms, const int32_t* term_left, const int32_t* term_right,
                       nq_operator_t* out);
int nq_operator_destroy(nq_operator_t op);
/* upper bound on the number of connections of one configuration */
int nq_operator_max_connections(nq_operator_t op, int64_t* out);
/* row_valdiff! over a batch (integer-parity path): for sample b, counts[b] connections in reference
 * order, zero matrix elements included; mels [max_conn,B] complex128; flips_row/flips_col
 * [W64, max_conn, B] uint64 global flip masks (flips_col may be NULL for NQ_KET).
 * ref: Operators/BaseOperators.jl:25-36, KLocalOperator.jl:151-159 */
int nq_connections(nq_operator_t op, nq_hilbert h, const void* srow, const void* scol, nq_dtype sdtype,
                   int64_t B, int64_t max_conn, int32_t* counts, double* mels,
                   uint64_t* flips_row, uint64_t* flips_col);

/* E_loc(sigma) = sum_c mel_c psi(eta_c)/psi(sigma).  out_loc [B] complex (C64 if machine is 32-bit, else C128).
 * logpsi may be NULL (recomputed).  Works for NQ_KET on RBM and for NQ_SUPER on RBMSplit/NDM.
 * ref: Accumulators/AccumulatorObsScalar.jl:52-137, AccumulatorLogPsi.jl:56-126, BatchedValSampler.jl:70-95 */
int nq_local_scalar(nq_machine_t m, nq_operator_t op, const void* srow, const void* scol, nq_dtype sdtype,
                    int64_t B, const void* logpsi, void* out_loc);
/* L_loc and grad L_loc = sum_c mel_c r_c grad log rho(eta_c) over ALL non-zero connections.
 * out_gloc [P,B] complex, leading dimension ld; may be NULL (value only).
 * ref: Accumulators/AccumulatorObsGrad.jl:39-127, AccumulatorLogGradPsi.jl:57-120, BatchedGradSampler.jl:87-97 */
int nq_local_grad(nq_machine_t m, nq_operator_t op, const void* srow, const void* scol, nq_dtype sdtype,
                  int64_t B, const void* logpsi, void* out_loc, void* out_gloc, int64_t ld);
/* fused iteration step on device-resident packed configurations: log psi [B], O [P,B] (ldO), the local
 * estimator [B] and, for Liouvillians, its gradient [P,B] (ld; may be NULL) in one pass over the batch.
 * ref: BatchedGradSampler.jl:83-97, BatchedValSampler.jl:122-125 */
int nq_logpsi_grad_local_packed(nq_machine_t m, nq_operator_t op, const uint64_t* prow, const uint64_t* pcol,
                                int64_t B, void* out_logpsi, void* O, int64_t ldO, void* out_loc, void* out_gloc,
                                int64_t ld);
/* the same step for configurations held by the HOST (srow / scol [N,B] float arrays, pinned memory makes the copies
 * asynchronous): copy + packing of piece c+1 run on a side stream of the context while the fused kernel works on piece c
 * (pieces of 2 rounds, 7 rounds, the rest of the persistent kernel); prow / pcol [W64,B] (device) receive the packed words,
 * all outputs are device buffers as above.  Results are bit-identical to nq_pack_states + nq_logpsi_grad_local_packed.
 * The call returns once everything is enqueued: with pinned arrays the last copies may still be in flight, so srow / scol
 * must not be modified before the context's stream has been synchronised (nq_ctx_sync; the stream waits for every copy).
 * ref: BatchedGradSampler.jl:83-97 (sample, then evaluate the batch), Samplers/Metropolis.jl:124-167 (host float states) */
int nq_logpsi_grad_local_host(nq_machine_t m, nq_operator_t op, const void* srow, const void* scol, nq_dtype sdtype,
                              int64_t B, uint64_t* prow, uint64_t* pcol, void* out_logpsi, void* O, int64_t ldO,
                              void* out_loc, void* out_gloc, int64_t ld);
int nq_local_scalar_packed(nq_machine_t m, nq_operator_t op, const uint64_t* prow, const uint64_t* pcol,
                           int64_t B, void* out_loc);
int nq_local_grad_packed(nq_machine_t m, nq_operator_t op, const uint64_t* prow, const uint64_t* pcol,
                         int64_t B, void* out_loc, void* out_gloc, int64_t ld);

/* ======================================================================================
 * Metropolis-Hastings sampler, LocalRule (single-site flip).
 * ref: Samplers/Metropolis.jl:4-40 (ctor), :101-115 (init_sampler!), :124-167 (samplenext!),
 *      Samplers/MCMCRules/LocalRule.jl:19-28, Hilbert/DoubledHilbert.jl:19-27.
 * B chains; one stored sample = `passes` proposals (even `passes` is bumped to odd like the reference).
 * Production mode draws (site, uniform) from Philox4x32-10 keyed by (seed, global chain id =
 * chain_offset + chain), so results do not depend on how chains are sharded over GPUs.
 * ==================================================================================== */
int nq_sampler_create(nq_machine_t m, int64_t B, int passes, uint64_t seed, int64_t chain_offset,
                      nq_sampler_t* out);
int nq_sampler_destroy(nq_sampler_t s);
int nq_sampler_set_state(nq_sampler_t s, const void* srow, const void* scol, nq_dtype sdtype);
int nq_sampler_get_state(nq_sampler_t s, void* srow, void* scol, nq_dtype sdtype);
/* rand!(rng, sigma, hilb): uniformly random configurations (Metropolis.jl:106) */
int nq_sampler_randomize(nq_sampler_t s);
/* Density-matrix machines (NDM): diagonal != 0 turns the chain into one over the diagonal rho(sigma, sigma) --
 * sigma' follows sigma, proposals are drawn in 1..N and flip the site on both sides, p(sigma) ~ rho(sigma, sigma).
 * This is the chain of the observables sampler (ref: IterativeInterface/Samplers/Obs/BatchedObsDMSampler.jl:59-105,
 * base_batched_networks.jl:244-253).  The reference takes abs.(log rho) as the log-probability there (SURVEY quirk
 * Q4); the library uses Re log rho(sigma, sigma), which is what the estimator needs.  Default 0 (joint chain). */
int nq_sampler_set_mode(nq_sampler_t s, int diagonal);
/* one samplenext! with supplied randomness: sites [passes,B] int32 1-based in 1..N (1..2N doubled),
 * uniforms [passes,B] of the machine's real type; accept_out [passes,B] uint8 (1 = accepted), may be NULL */
int nq_sampler_replay(nq_sampler_t s, const int32_t* sites, const void* uniforms, uint8_t* accept_out);
/* `burn` discarded + `L` stored samples per chain.  Outputs (each may be NULL):
 * packed [L][B][W64] device/host words, or float arrays [N,B,L] of sdtype. */
int nq_sampler_sample(nq_sampler_t s, int burn, int L, uint64_t* prow, uint64_t* pcol,
                      void* srow, void* scol, nq_dtype sdtype);
int nq_sampler_counters(nq_sampler_t s, int64_t* passes_done, int64_t* passes_accepted);
/* Transition rule of the chain (default NQ_RULE_LOCAL).
 *   NQ_RULE_EXCHANGE (ket machines): a random couple (i, j) of `couplings` [n_couplings][2] (int32, 0-based) is swapped.
 *                    ref: MCMCRules/ExchangeRule.jl:14-68 (couples = the 2-site couplings of an operator)
 *   NQ_RULE_NAGY (density-matrix machines): one of 8 moves -- hopping in sigma / sigma' (site s and one element of the
 *                    s-th couple, the reference indexes its couple list by SITE), single flips, the dissipator move
 *                    (an empty Fock site is excited with probability 1/10), the jumper.  ref: MCMCRules/Nagy.jl:40-115
 *   NQ_RULE_OPERATOR: a uniformly drawn connection of `op` (diagonal and zero elements included, reference order) is
 *                    applied; log_prob_bias = log(n_forward / n_back) enters the accept test.
 *                    ref: MCMCRules/OperatorRule.jl:27-59, Metropolis.jl:148
 * These rules flip several sites per proposal: the chain re-evaluates the activations of the moved pre-activations. */
int nq_sampler_set_rule(nq_sampler_t s, nq_rule rule, int n_couplings, const int32_t* couplings, nq_operator_t op);
/* one samplenext! of a non-local rule with supplied randomness: draws [passes,B,4] int32 = the integers the reference
 * draws per proposal (1-based): Exchange (couple, -, -, -); Nagy (move 1..8, site 1..N, aux, aux2) with aux = element
 * 1..2 of the couple (moves 1-4), rand(1:10) of the row (move 7; aux2 = that of the column), column site (move 8);
 * Operator (r, -, -, -): r is the raw 32-bit random, the connection index is floor(r n_forward / 2^32).
 * uniforms / accept_out as in nq_sampler_replay. */
int nq_sampler_replay_rule(nq_sampler_t s, const int32_t* draws, const void* uniforms, uint8_t* accept_out);

/* ======================================================================================
 * Symmetrised machines (NDMSymm): a bare machine whose parameters are tied by site permutations.
 * The handle owns the Ps symmetric parameters (machine dtype, real) and two maps:
 *   parameters:  bare[q] = w[src[q]]   after the ranges avg_ranges[i] = [a, b) of w were replaced by their mean
 *                (set_bare_params!, NDMSymm.jl:79-128: the local biases b_mu, b_lam are averaged)
 *   gradients:   Osymm[p, s] = scale[p] * sum_{e in [ptr[p], ptr[p+1])} Obare[idx[e], s]
 *                (symmetrize_grad_NDM_batched!, NDMSymmBatched.jl:22-36 with the 0/1 matrices of NDMSymm.jl:130-181)
 * All index arrays are HOST arrays, 0-based.  The bare machine must outlive the handle.
 * ==================================================================================== */
int nq_symm_create(nq_machine_t bare, int64_t Ps, const int64_t* ptr, const int32_t* idx, const double* scale,
                   const int32_t* src, int n_avg, const int64_t* avg_ranges, nq_symm_t* out);
int nq_symm_destroy(nq_symm_t g);
int nq_symm_nparams(nq_symm_t g, int64_t* Ps);
int nq_symm_set_params(nq_symm_t g, const void* w, int64_t Ps);    /* stores w, averages, expands into the bare machine */
int nq_symm_get_params(nq_symm_t g, void* w, int64_t Ps);
/* update!(opt, cnet::NDMSymm, dw): w <- w - eta dw, then set_bare_params! (NDMSymm.jl:27-30) */
int nq_symm_update(nq_symm_t g, const void* dw, double eta);
/* rows of the bare gradient [Pb, Ns] (ldb) -> rows of the symmetrised gradient [Ps, Ns] (lds); dtype = element type */
int nq_symm_gradient(nq_symm_t g, const void* Obare, int64_t ldb, int64_t Ns, nq_dtype dtype, void* Osymm, int64_t lds);

/* ======================================================================================
 * Full-space tools (the reference's validation path; indexable spaces: at most 2^30 table entries).
 * Basis number i (1-based) <-> digits of i-1, site 1 least significant (Hilbert/HomogeneousSpin.jl:156-179);
 * density matrices: super index = (col-1) D + row, D = 2^N.
 * ==================================================================================== */
int nq_fullspace_size(nq_machine_t m, int64_t* size);          /* 2^N (RBM) or 4^N (RBMSplit, NDM) */
/* ket(net, hilb, norm) / densitymatrix(net, hilb, norm): out[i] = exp(log psi(state i)), [2^N] or column-major [D,D]
 * (rho[row,col]) of out_type; norm != 0 divides the ket by its 2-norm and the density matrix by its trace.
 * ref: utils/densitymatrix.jl:9-62 */
int nq_fullspace_state(nq_machine_t m, int norm, void* out);
/* ExactSampler init_sampler!: cdf[i] = sum_{k<=i} p_k / sum_k p_k with p_k = exp(log_prob_psi(k)) (evaluated as
 * exp(lp - max lp)), [size] doubles.  ref: Samplers/Exact.jl:135-162 */
int nq_exact_table(nq_machine_t m, double* cdf);
/* ExactSampler samplenext! for L slots x B chains: basis number = searchsortedfirst(cdf, r), written as packed words
 * [L][B] (pcol NULL for RBM) and optionally as 1-based numbers in `indices` [L][B].  r comes from `uniforms` [L][B]
 * doubles when given (replay), else from Philox4x32-10 keyed by (seed, chain_offset + chain) at counter
 * draw_base + slot -- independent of how chains are sharded.  ref: Samplers/Exact.jl:166-181 */
int nq_exact_sample(nq_machine_t m, const double* cdf, uint64_t seed, int64_t chain_offset, uint64_t draw_base,
                    int64_t B, int64_t L, const double* uniforms, uint64_t* prow, uint64_t* pcol, int64_t* indices);

/* ======================================================================================
 * Stochastic reconfiguration.
 * dtype below = element type of O (NQ_F32/F64 for real-weight RBM/RBMSplit; NQ_C64/C128 otherwise).
 * ==================================================================================== */
/* <O> and O <- O - <O> in place.  ref: BaseIterativeSampler.jl:19-26.  avg [P] */
int nq_center(nq_ctx_t ctx, void* O, int64_t ldO, int64_t P, int64_t Ns, nq_dtype dtype, void* avg);
/* The same with the subtraction DEFERRED when the next consumer can do it on the fly: avg is returned, and if
 * *deferred = 1 the matrix is left uncentred -- the FP64 S assembly of nq_sr_setup (integer tensor-core path) then subtracts
 * the means while it slices the rows, which saves the read-modify-write pass over O; any other nq_sr_setup path centres in
 * place first.  Device-resident FP64 O with >= 4096 samples only (otherwise it is nq_center and *deferred = 0).
 * nq_center_finish applies a pending subtraction (no-op if none is pending for this matrix): call it before reading O as
 * the centred matrix (nq_force_ket, matrix-free solves).  ref: BaseIterativeSampler.jl:19-26 */
int nq_center_lazy(nq_ctx_t ctx, void* O, int64_t ldO, int64_t P, int64_t Ns, nq_dtype dtype, void* avg, int* deferred);
int nq_center_finish(nq_ctx_t ctx, void* O, int64_t ldO, int64_t P, int64_t Ns, nq_dtype dtype);
/* gradC = E_loc * Oc' / Ns  (F_k = <E_loc conj(Oc_k)>).  ref: BatchedValSampler.jl:97-115.
 * Eloc [Ns] complex of matching precision; gradC [P] complex */
int nq_force_ket(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, nq_dtype dtype,
                 const void* Eloc, void* gradC);
/* gradC = conj(L_loc * gradL' / Ns - <|L_loc|^2> avg').  ref: BatchedGradSampler.jl:99-118.
 * Returns the cost <|L_loc|^2> in *cost (host double). */
int nq_force_liouvillian(nq_ctx_t ctx, const void* Lloc, const void* gLloc, int64_t ld, int64_t P,
                         int64_t Ns, nq_dtype dtype, const void* avg, void* gradC, double* cost);
/* S and F from the centred O.  Implementation (all hand-written, chosen by size and precision):
 *   FP64, >= 4096 samples: Ozaki scheme on the integer tensor cores (tcgen05 kind::i8; rows scaled by their exponent, 7
 *     signed 7-bit digits, 34 exact products, FP64 recombination) -- assumes rows without extreme outliers (the digits
 *     keep 48 bits below the largest entry of a row; gradients of the machines here are bounded); NQ_SR_FP64=dmma in the
 *     environment forces the FP64 tensor instruction (DMMA), which is also used below 4096 samples and as the fallback;
 *   FP32 mode: 3xTF32 on tcgen05 fed by TMA (NQ_SR_TF32=stage: the staging kernel), split-K partials summed in FP64.
 * real_params != 0: S = Re(Oc Oc^H)/Ns (real [P,P]), F = Re(gradC);
 * else S = conj(Oc Oc^H)/Ns (complex), F = gradC.  Ns_total = global sample count used to
 * normalise (== Ns on one GPU; under sharding the partial S is all-reduced by the caller, quirk Q5).
 * ref: SR/SRDirect.jl:26-49, SRIterative.jl:45-62 */
int nq_sr_setup(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, int64_t Ns_total,
                nq_dtype dtype, const void* gradC, int real_params, void* S, void* F);
/* One-shot structural hint for the NEXT nq_sr_setup on this context: row_planes [P] host bytes, bit 0 = the real part of
 * the row may be non-zero, bit 1 = the imaginary part may be.  For the real-parameter NDM the gradient rows of the lambda
 * biases / weights are purely real and those of the mu biases / weights purely imaginary (NDMBatched.jl:262-277); with
 * the hint the assembly skips those operand planes without scanning O for them (it scans when no hint is given).  A hint
 * that clears a plane which is not identically zero gives a wrong S.  NULL clears the hint. */
int nq_sr_hint_row_planes(nq_ctx_t ctx, const uint8_t* row_planes, int64_t P);
/* (S + eps I) dw = F.  sdtype = element type of S/F/dw (real or complex).  S is overwritten.
 * CHOLESKY: NQ_ERR_NOT_POSDEF if a pivot <= 0.  CG: IterativeSolvers-0.8.1 semantics, x0 = 0,
 * stop ||r|| <= tol ||F||, maxiter (<=0: 10 P); NQ_ERR_NOT_CONVERGED when exhausted.
 * ref: SRDirect.jl:51-90, SRIterative.jl:71-153 */
int nq_sr_solve(nq_ctx_t ctx, void* S, const void* F, int64_t P, nq_dtype sdtype, double eps,
                nq_solver algo, double tol, int64_t maxiter, void* dw, int64_t* iters);
/* matrix-free CG: v -> eps v + conj(Oc)(conj(Oc)^H v)/Ns_total (complex nets) or
 * (Or Or^T + Oi Oi^T) v/Ns_total (real_params).  With a communicator the partial products are
 * all-reduced every iteration.  ref: SR/SR_notfull.jl:47-148, SRIterative.jl:117-132 */
int nq_sr_solve_matfree(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, int64_t Ns_total,
                        nq_dtype dtype, const void* F, int real_params, double eps, double tol,
                        int64_t maxiter, void* dw, int64_t* iters);
/* The same with the solver chosen by the caller: NQ_SOLVE_CG, NQ_SOLVE_MINRES (Paige-Saunders MINRES on the
 * Hermitian operator, x0 = 0, stop when the recurrence residual <= tol ||F||) or NQ_SOLVE_QLP / NQ_SOLVE_QLP_WARM
 * (MINRES-QLP with the reference's stopping rules: min(relres, relAres) <= tol, xnorm and Acond limits 1e7; the exit
 * flag 1..9 is left in nq_ctx_last_info; NQ_ERR_NOT_CONVERGED only for flag 8, the iteration limit).  All of them are also
 * accepted by nq_sr_solve on an explicit S.
 * ref: SRIterative.jl:92-125, External/IterativeSolvers/minresqlp.jl (quirks Q19-Q22 in oracle/minresqlp.py). */
int nq_sr_solve_matfree_algo(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, int64_t Ns_total,
                             nq_dtype dtype, const void* F, int real_params, double eps, nq_solver algo, double tol,
                             int64_t maxiter, void* dw, int64_t* iters);
/* Streaming S assembly for batches whose gradient matrix does not fit in memory (BASELINE cfg5): rows are produced chunk
 * by chunk into one reused device buffer (nq_logpsi_grad_packed) and consumed at once.  O: UNCENTRED rows [P, Nc] of the
 * chunk, OVERWRITTEN (shifted by a provisional mean, the mean of the first chunk); Sacc [P,P] (element type of S, see
 * nq_sr_setup) accumulates the Gram matrix of the shifted rows / Ns_total; state [2 P] complex128 holds the running sums
 * and the shift; first != 0 initialises both.  nq_sr_finish applies the rank-one correction (and all-reduces the partial
 * sums under sharding): Sacc is then the S of nq_center + nq_sr_setup on the whole batch and state[0..P) = <O>.
 * Device buffers only.
 * ref: BaseIterativeSampler.jl:19-26 + SRDirect.jl:26-49 (the reference needs O of the whole batch in memory) */
int nq_sr_accumulate(nq_ctx_t ctx, void* O, int64_t ldO, int64_t P, int64_t Nc, int64_t Ns_total, nq_dtype dtype,
                     int real_params, void* Sacc, void* state, int first);
int nq_sr_finish(nq_ctx_t ctx, void* Sacc, void* state, int64_t P, int64_t Ns_total, nq_dtype dtype, int real_params);
/* Nesterov(lr, mu): d = mu^2 v - (1 + mu) lr dw; v <- mu v - lr dw; delta = -d (apply with nq_update(m, delta, 1)).
 * velocity, dw, delta: device vectors of n elements of `dtype`.  ref: Optimisers/rules.jl:36-55 */
int nq_nesterov(nq_ctx_t ctx, void* velocity, const void* dw, int64_t n, nq_dtype dtype, double lr, double mu, void* delta);
/* w <- w - eta dw on the machine's device parameters.  dw has the machine dtype (real for NDM).
 * ref: Optimisers/rules.jl:11-17, apply.jl:25-73 */
int nq_update(nq_machine_t m, const void* dw, double eta);
/* sr_multiplicative regulariser: S <- S + lambda Diagonal(diag(S)), in place on the device or host S (then solve with
 * eps = 0).  ref: SRDirect.jl:66-72, SRIterative.jl:84-90 (lambda = max(lambda0 b^iter, lambda_min)). */
int nq_sr_scale_diagonal(nq_ctx_t ctx, void* S, int64_t P, nq_dtype sdtype, double lambda);
/* stat_analysis over [B chains, L] values (column-major [B,L]); out = {mean_re, mean_im, error,
 * variance, tau, R} host doubles.  vdtype: NQ_F32/F64/C64/C128.  ref: utils/stats.jl:26-50.
 * With a communicator the statistics are those of the UNION of the ranks' chains (the chain moments are all-reduced;
 * the reference reduces only the mean, utils/stats.jl:52-77, quirk Q15): every rank gets the same six numbers. */
int nq_stat_analysis(nq_ctx_t ctx, const void* vals, int64_t B, int64_t L, nq_dtype vdtype, double out[6]);

/* out[i] = |vals[i]|^2 (real of the same precision).  ref: BatchedGradSampler.jl:99 (abs2.(local_vals)) */
int nq_abs2(nq_ctx_t ctx, const void* vals, int64_t n, nq_dtype vdtype, void* out);

/* ======================================================================================
 * Parallel backend: chains are sharded over ranks (one GPU each); partial sums are reduced with
 * NCCL all-reduce over NVLink.  ref: Parallel/not_parallel.jl:1-19, Parallel/MPI/mpi.jl:21-74
 * (workers_sum!, workers_mean!, num_workers, worker_local_seed).
 * ==================================================================================== */
#define NQ_UNIQUE_ID_BYTES 128
int nq_comm_unique_id(uint8_t id[NQ_UNIQUE_ID_BYTES]);
int nq_comm_init(nq_ctx_t ctx, int nranks, int rank, const uint8_t id[NQ_UNIQUE_ID_BYTES]);
int nq_comm_destroy(nq_ctx_t ctx);
int nq_comm_size(nq_ctx_t ctx, int* nranks, int* rank);   /* 1, 0 without a communicator */
int nq_allreduce_sum(nq_ctx_t ctx, void* buf, int64_t n, nq_dtype dtype);   /* device buffer, in place */
int nq_allreduce_mean(nq_ctx_t ctx, void* buf, int64_t n, nq_dtype dtype);
/* Global sample count of the iteration when ranks own DIFFERENT numbers of samples (chains do not divide by the number
 * of ranks): nq_center / nq_force_* normalise by it.  0 (default) = Ns * nranks.  The reference shards the chain
 * length with ceil (Metropolis.jl:75) and divides by the worker count (mpi.jl:21-34). */
int nq_comm_set_global_samples(nq_ctx_t ctx, int64_t ns_total);

#ifdef __cplusplus
}
#endif
#endif /* NQCUDA_H */
