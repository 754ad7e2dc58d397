// nq_internal.cuh -- handle layouts and cross-file internal entry points of libnqcuda.
#pragma once
#include "nq_common.cuh"

// scratch slot ids (one live buffer per slot per call)
enum {
    SL_IN0 = 0, SL_IN1, SL_IN2, SL_IN3, SL_IN4,
    SL_OUT0, SL_OUT1, SL_OUT2, SL_OUT3,
    SL_PROW, SL_PCOL, SL_LOGPSI, SL_W0, SL_W1, SL_W2, SL_W3, SL_W4, SL_W5,
    SL_HOSTO, SL_HOSTG          // device copies of host-resident O / grad L_loc matrices (SR entry points)
};

struct nq_machine_s {
    nq_ctx_t ctx;
    nq_machine_kind kind;
    nq_hilbert hilb;
    int N, M, A;
    nq_activation act;
    nq_dtype dtype;       // parameter dtype
    nq_dtype out_dtype;   // log psi / gradient dtype
    int64_t P;
    void* params;         // device, P elements of dtype
    // single-flip ratio tables exp(+-c W) (c = |change of a site value|), rebuilt lazily after every
    // parameter change; see nq_machine_ensure_tables (nq_machines.cu)
    void* etab;
    bool etab_valid;
    bool doubled() const { return kind != NQ_RBM; }
};

struct nq_operator_s {
    nq_ctx_t ctx;
    nq_space space;
    int N;
    int n_parts, n_terms;
    int64_t n_rows, n_entries;
    int64_t max_conn;
    int max_part_sites;
    int site_local;         // every connection flips at most one site index
    // device tables
    int32_t* part_nsites;   // [n_parts]
    int32_t* part_site_ptr; // [n_parts+1] offsets into part_sites
    int32_t* part_sites;    // 0-based
    int64_t* part_row0;     // [n_parts] first row of the part in row_ptr
    int64_t* row_ptr;       // [n_rows+1]
    double* entry_mel;      // complex128 interleaved [n_entries]
    uint32_t* entry_flip;   // [n_entries]
    int32_t* term_left;     // [n_terms]
    int32_t* term_right;    // [n_terms]
    // flat records (N <= 64): [n_recs][8] = rmask, rval, cmask, cval, rflip, cflip, mel_re, mel_im (doubles as bits)
    uint64_t* recs;
    int n_recs;
    // site tables (site-local operators, N <= 64): the summed matrix element of every (site, pattern) slot and of the
    // diagonal as lookup tables over the few configuration bits it depends on; see build_site_luts (nq_operator.cu)
    int32_t* lut_units;     // [n_lut_units][2] group range of a unit; units 0..3N-1 = slots 3 j + pattern, then diagonal units
    uint64_t* lut_groups;   // [n_groups][2] = (8 bit selectors: side << 6 | site, 0xFF unused ; table offset)
    double* lut_tab;        // complex128 entries
    int n_lut_units;
};

// device-pointer internals shared between translation units
int nq_pack_device(nq_ctx_t ctx, nq_hilbert h, int N, int64_t B, const void* dsigma, nq_dtype sdtype, uint64_t* dpacked);
int nq_unpack_device(nq_ctx_t ctx, nq_hilbert h, int N, int64_t B, const uint64_t* dpacked, void* dsigma, nq_dtype sdtype);
int nq_machine_ensure_tables(nq_machine_t m);
// RBM / RBMSplit: tables of the register-resident sampler, [sign][mat][k + M j] = exp(+-c g W) - 1, then wsum[mat][j]
const void* nq_machine_qtab(nq_machine_t m);
int nq_machine_eval_device(nq_machine_t m, const uint64_t* prow, const uint64_t* pcol, int64_t B,
                           void* out, void* O, int64_t ldO);
struct NqStage;
int nq_stage_pack(nq_machine_t m, NqStage& st, const void* srow, const void* scol, nq_dtype sdtype, int64_t B,
                  const uint64_t** prow, const uint64_t** pcol);
// nq_syrk_tf32.cu: FP32-mode S assembly on tcgen05 (3xTF32); S written as float / interleaved complex float
int nq_syrk_tf32_device(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, int64_t Ns_total, bool o_complex,
                        bool out_complex, void* dS);
// nq_syrk_ozaki.cu: FP64 S assembly (real part) on the integer tensor cores; *used = false -> run the DMMA kernel instead
int nq_syrk_ozaki_device(nq_ctx_t ctx, const double* Xr, int64_t ldr, int64_t P, int64_t Ns, int NC, int ntile, int nsplit, double* W,
                         double* Wim, const unsigned long long* known_rowmax, const double* shift /* complex128 [P] or NULL */,
                         bool* used);
