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
#include "nq_opdev.cuh"

namespace {

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

#include "nq_estimators.inc"
#include "nq_estimators_ndm3.inc"
#include "nq_estimators_ndm5.inc"

}  // namespace

// out_logpsi / O may be null; when given (fused eval+grad+estimator) the machines without a fused kernel
// run the eval+grad kernel first.
int nq_local_device(nq_machine_t m, nq_operator_t op, const uint64_t* pr, const uint64_t* pc, int64_t B,
                    void* out_logpsi, void* O, int64_t ldO, void* out_loc, void* out_g, int64_t ld) {
    nq_ctx_t ctx = m->ctx;
    if (op->ctx != ctx) return nq_fail(ctx, NQ_ERR_ARG, "machine and operator belong to different contexts");
    if (op->N != m->N) return nq_fail(ctx, NQ_ERR_SHAPE, "operator acts on %d sites, machine on %d", op->N, m->N);
    if ((op->space == NQ_SUPER) != m->doubled())
        return nq_fail(ctx, NQ_ERR_ARG, "ket operators pair with RBM, Liouvillians with RBMSplit/NDM");
    if (op->max_part_sites > MAXF) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "local terms on more than %d sites", MAXF);
    if (m->N > 256) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "estimator kernels index sites with 8 bits (N <= 256)");
    if (B == 0) return NQ_OK;
    NQ_CHECK(nq_machine_ensure_tables(m));
    if (m->kind == NQ_NDM) {
        bool used = false;
        // v5 (warp per configuration) where it applies, else v3t / the list kernel; NQ_NDM_KERNEL=v3 forces the latter
        const char* envk = getenv("NQ_NDM_KERNEL");
        if (!envk || !strcmp(envk, "v5")) {
            if (m->dtype == NQ_F64) {
                NQ_CHECK((m->act == NQ_SOFTPLUS ? launch_local_ndm5<double, NQ_SOFTPLUS>(m, op, pr, pc, B, out_logpsi, O, ldO, out_loc, out_g, ld, &used)
                                                 : launch_local_ndm5<double, NQ_LOGCOSH>(m, op, pr, pc, B, out_logpsi, O, ldO, out_loc, out_g, ld, &used)));
            } else {
                NQ_CHECK((m->act == NQ_SOFTPLUS ? launch_local_ndm5<float, NQ_SOFTPLUS>(m, op, pr, pc, B, out_logpsi, O, ldO, out_loc, out_g, ld, &used)
                                                 : launch_local_ndm5<float, NQ_LOGCOSH>(m, op, pr, pc, B, out_logpsi, O, ldO, out_loc, out_g, ld, &used)));
            }
            if (used) return NQ_OK;
        }
        if (m->dtype == NQ_F64) {
            NQ_CHECK((m->act == NQ_SOFTPLUS ? launch_local_ndm3<double, NQ_SOFTPLUS>(m, op, pr, pc, B, out_logpsi, O, ldO, out_loc, out_g, ld, &used)
                                             : launch_local_ndm3<double, NQ_LOGCOSH>(m, op, pr, pc, B, out_logpsi, O, ldO, out_loc, out_g, ld, &used)));
        } else {
            NQ_CHECK((m->act == NQ_SOFTPLUS ? launch_local_ndm3<float, NQ_SOFTPLUS>(m, op, pr, pc, B, out_logpsi, O, ldO, out_loc, out_g, ld, &used)
                                             : launch_local_ndm3<float, NQ_LOGCOSH>(m, op, pr, pc, B, out_logpsi, O, ldO, out_loc, out_g, ld, &used)));
        }
        if (used) return NQ_OK;
    }
    if (out_logpsi) NQ_CHECK(nq_machine_eval_device(m, pr, pc, B, out_logpsi, O, ldO));
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


// Site tables for site-local operators (N <= 64).  Every flat record contributes its matrix element to one slot:
// the diagonal, or (site j, pattern p) with p = 0 flip on sigma, 1 flip on sigma', 2 flip on both.  The records of a
// slot are cut into groups whose conditions read at most LUT_BITS configuration bits; a group becomes a table
// indexed by those bits holding the sum (in record = reference order) of the matching matrix elements.  A
// configuration then costs one bit gather and one 16-byte load per group instead of a pass over all records.
// Diagonal groups are independent units (summed across lanes), the groups of an off-diagonal slot form one unit.
static constexpr int LUT_BITS = 8;
static int build_site_luts(nq_ctx_t ctx, nq_operator_t op, const std::vector<uint64_t>& recs, int N) {
    struct Rec { uint64_t rm, rv, cm, cv; double mr, mi; };
    struct Group { uint64_t rmask = 0, cmask = 0; std::vector<Rec> recs; };
    const int n_slots = 3 * N + 1, diag = 3 * N;
    std::vector<std::vector<Group>> slots(n_slots);
    auto popc = [](uint64_t x) { return __builtin_popcountll(x); };
    for (size_t r = 0; r + 8 <= recs.size(); r += 8) {
        const uint64_t rf = recs[r + 4], cf = recs[r + 5];
        Rec rec{recs[r], recs[r + 1], recs[r + 2], recs[r + 3], 0.0, 0.0};
        memcpy(&rec.mr, &recs[r + 6], 8); memcpy(&rec.mi, &recs[r + 7], 8);
        int slot = diag;
        if (rf | cf) {
            if (popc(rf) > 1 || popc(cf) > 1 || (rf && cf && rf != cf)) return NQ_OK;     // not site-local: no tables
            slot = 3 * __builtin_ctzll(rf | cf) + ((rf ? 1 : 0) + (cf ? 2 : 0) - 1);
        }
        if (popc(rec.rm) + popc(rec.cm) > LUT_BITS) return NQ_OK;
        auto& gs = slots[slot];
        Group* g = nullptr;
        for (auto& cand : gs)          // first group whose support stays within LUT_BITS bits
            if (popc(cand.rmask | rec.rm) + popc(cand.cmask | rec.cm) <= LUT_BITS) { g = &cand; break; }
        if (!g) { gs.emplace_back(); g = &gs.back(); }
        g->rmask |= rec.rm; g->cmask |= rec.cm;
        g->recs.push_back(rec);
    }
    std::vector<int32_t> units;
    std::vector<uint64_t> groups;
    std::vector<double> tab;
    auto emit_group = [&](const Group& g) {
        int pos[LUT_BITS], nb = 0;
        for (int side = 0; side < 2; side++) {
            const uint64_t mk = side ? g.cmask : g.rmask;
            for (int b = 0; b < 64; b++) if ((mk >> b) & 1ull) pos[nb++] = (side << 6) | b;
        }
        uint64_t sel = 0;
        for (int i = 0; i < LUT_BITS; i++) sel |= (uint64_t)(i < nb ? pos[i] : 0xFF) << (8 * i);
        groups.push_back(sel);
        groups.push_back((uint64_t)(tab.size() / 2));
        for (int idx = 0; idx < (1 << nb); idx++) {
            uint64_t rb = 0, cb = 0;
            for (int i = 0; i < nb; i++) if ((idx >> i) & 1) ((pos[i] & 64) ? cb : rb) |= 1ull << (pos[i] & 63);
            double sr = 0.0, si = 0.0;
            for (const Rec& rec : g.recs)
                if ((rb & rec.rm) == rec.rv && (cb & rec.cm) == rec.cv) { sr += rec.mr; si += rec.mi; }
            tab.push_back(sr); tab.push_back(si);
        }
    };
    for (int slot = 0; slot < diag; slot++) {
        units.push_back((int32_t)(groups.size() / 2));
        for (const Group& g : slots[slot]) emit_group(g);
        units.push_back((int32_t)(groups.size() / 2));
    }
    for (const Group& g : slots[diag]) {
        units.push_back((int32_t)(groups.size() / 2));
        emit_group(g);
        units.push_back((int32_t)(groups.size() / 2));
    }
    if (tab.size() > ((size_t)1 << 22)) return NQ_OK;
    int s;
    if ((s = upload(ctx, &op->lut_units, units.data(), units.size())) != NQ_OK) return s;
    if ((s = upload(ctx, &op->lut_groups, groups.data(), groups.size())) != NQ_OK) return s;
    if ((s = upload(ctx, &op->lut_tab, tab.data(), tab.size())) != NQ_OK) return s;
    op->n_lut_units = (int)(units.size() / 2);
    return NQ_OK;
}

extern "C" int nq_operator_destroy(nq_operator_t op) {
    if (!op) return NQ_ERR_ARG;
    cudaSetDevice(op->ctx->device);
    cudaStreamSynchronize(op->ctx->stream);
    cudaFree(op->part_nsites); cudaFree(op->part_site_ptr); cudaFree(op->part_sites); cudaFree(op->part_row0);
    cudaFree(op->row_ptr); cudaFree(op->entry_mel); cudaFree(op->entry_flip); cudaFree(op->term_left); cudaFree(op->term_right);
    cudaFree(op->recs); cudaFree(op->lut_units); cudaFree(op->lut_groups); cudaFree(op->lut_tab);
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
    // site_local: every connection flips at most ONE site index (the same on sigma and sigma'), so the
    // estimator kernels can deal connections to warps by site without write conflicts
    int site_local = 1;
    {
        std::vector<uint32_t> part_mask(n_parts, 0u);
        for (int p = 0; p < n_parts; p++)
            for (int64_t e = row_ptr[row0[p]]; e < row_ptr[row0[p] + ((int64_t)1 << part_nsites[p])]; e++) part_mask[p] |= entry_flip[e];
        for (int t = 0; t < n_terms && site_local; t++) {
            int parts[2] = {term_left[t], space == NQ_SUPER ? term_right[t] : -1};
            int site = -1;
            for (int q = 0; q < 2; q++) {
                int p = parts[q];
                if (p < 0) continue;
                for (int i = 0; i < part_nsites[p]; i++)
                    if ((part_mask[p] >> i) & 1u) {
                        int sidx = part_sites[site_ptr[p] + i];
                        if (site >= 0 && site != sidx) site_local = 0;
                        site = sidx;
                    }
            }
        }
    }
    nq_operator_t op = new nq_operator_s();
    memset(op, 0, sizeof(*op));
    op->site_local = site_local;
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
        // Flat records for N <= 64: one per (term, local row[, local row'], entry[, entry']) with non-zero matrix
        // element, in reference order: "if (sigma & rmask) == rval and (sigma' & cmask) == cval then this
        // connection exists, flips rflip / cflip and weighs mel".  The estimator kernels test all records
        // in parallel instead of chasing term -> part -> row -> entry pointers per configuration.
        std::vector<uint64_t> recs;
        bool flat = N <= 64;
        auto bits_of = [](double v) { uint64_t u; memcpy(&u, &v, 8); return u; };
        for (int t = 0; t < n_terms && flat; t++) {
            const int L = term_left[t], R = tr[t];
            const int kl = L >= 0 ? part_nsites[L] : 0, kr = R >= 0 ? part_nsites[R] : 0;
            const int32_t* sl = L >= 0 ? part_sites + site_ptr[L] : nullptr;
            const int32_t* sr = R >= 0 ? part_sites + site_ptr[R] : nullptr;
            auto site_mask = [](const int32_t* ps, int k, uint32_t local) {
                uint64_t m = 0;
                for (int i = 0; i < k; i++) if ((local >> i) & 1u) m |= (uint64_t)1 << ps[i];
                return m;
            };
            for (int64_t rl = 0; rl < ((int64_t)1 << kl) && flat; rl++)
                for (int64_t rr = 0; rr < ((int64_t)1 << kr) && flat; rr++) {
                    const int64_t l0 = L >= 0 ? row_ptr[row0[L] + rl] : 0, l1 = L >= 0 ? row_ptr[row0[L] + rl + 1] : 1;
                    const int64_t r0 = R >= 0 ? row_ptr[row0[R] + rr] : 0, r1 = R >= 0 ? row_ptr[row0[R] + rr + 1] : 1;
                    for (int64_t el = l0; el < l1; el++)
                        for (int64_t er = r0; er < r1; er++) {
                            double mr, mi;
                            if (L >= 0 && R >= 0) {
                                const double ar = entry_mel[2 * el], ai = entry_mel[2 * el + 1], br = entry_mel[2 * er], bi = entry_mel[2 * er + 1];
                                volatile double p0 = ar * br, p1 = ai * bi, p2 = ar * bi, p3 = ai * br;   // no contraction
                                mr = p0 - p1; mi = p2 + p3;
                            } else {
                                const int64_t e = L >= 0 ? el : er;
                                mr = entry_mel[2 * e]; mi = entry_mel[2 * e + 1];
                            }
                            if (mr == 0.0 && mi == 0.0) continue;
                            const uint64_t rec[8] = {site_mask(sl, kl, (1u << kl) - 1u), site_mask(sl, kl, (uint32_t)rl),
                                                     site_mask(sr, kr, (1u << kr) - 1u), site_mask(sr, kr, (uint32_t)rr),
                                                     L >= 0 ? site_mask(sl, kl, entry_flip[el]) : 0, R >= 0 ? site_mask(sr, kr, entry_flip[er]) : 0,
                                                     bits_of(mr), bits_of(mi)};
                            recs.insert(recs.end(), rec, rec + 8);
                            if (recs.size() > (size_t)8 << 16) flat = false;
                        }
                }
        }
        if (flat && !recs.empty()) {
            if ((s = upload(ctx, &op->recs, recs.data(), (int64_t)recs.size())) != NQ_OK) { nq_operator_destroy(op); return s; }
            op->n_recs = (int)(recs.size() / 8);
            if (site_local && (s = build_site_luts(ctx, op, recs, N)) != NQ_OK) { nq_operator_destroy(op); return s; }
        }
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
    NQ_CHECK(nq_local_device(m, op, pr, pc, B, nullptr, nullptr, 0, dl, dg, ld));
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

// fused iteration entry: log psi, O, local estimator (and its gradient for Liouvillians) of device-resident
// packed configurations in one pass.  ref: BatchedGradSampler.jl:83-97 / BatchedValSampler.jl:122-125
extern "C" int nq_logpsi_grad_local_packed(nq_machine_t m, nq_operator_t op, const uint64_t* prow, const uint64_t* pcol,
                                           int64_t B, void* out_logpsi, void* O, int64_t ldO, void* out_loc,
                                           void* out_gloc, int64_t ld) {
    if (!m || !op || !prow || !out_logpsi || !O || !out_loc || B < 0) return NQ_ERR_ARG;
    nq_ctx_t ctx = m->ctx;
    if (m->doubled() != (pcol != nullptr)) return nq_fail(ctx, NQ_ERR_ARG, "row/col configuration mismatch");
    if (ldO < m->P || (out_gloc && ld < m->P)) return nq_fail(ctx, NQ_ERR_SHAPE, "leading dimension < P");
    if (out_gloc && op->space != NQ_SUPER) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "gradient estimator is defined for Liouvillians only");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!nq_is_device_ptr(prow) || !nq_is_device_ptr(O) || !nq_is_device_ptr(out_logpsi) || !nq_is_device_ptr(out_loc) ||
        (out_gloc && !nq_is_device_ptr(out_gloc)))
        return nq_fail(ctx, NQ_ERR_ARG, "the fused entry point works on device-resident buffers");
    ctx->shift_pending = false; ctx->rowmax_ptr = nullptr;          // new rows: records of a centring pass are stale
    return nq_local_device(m, op, prow, pcol, B, out_logpsi, O, ldO, out_loc, out_gloc, ld);
}

// The same step for configurations held by the HOST as float arrays (what the reference's sampler hands over), software
// pipelined inside the library: the batch is cut into pieces that grow in units of one ROUND of the persistent kernel (one
// configuration per resident warp: 2 rounds, 7 rounds, the rest); the float arrays of piece c+1 are copied and packed on a
// side stream while the fused kernel works on piece c, so only the copy of the first piece is exposed (a piece copies ~3.6x
// faster than it computes) and whole rounds keep the total number of rounds.  prow / pcol receive the packed words.
// ref: BatchedGradSampler.jl:83-97 (the reference copies nothing: its arrays live where they are computed)
extern "C" int nq_logpsi_grad_local_host(nq_machine_t m, nq_operator_t op, const void* srow, const void* scol, nq_dtype sdtype,
                                         int64_t B, uint64_t* prow, uint64_t* pcol, void* out_logpsi, void* O, int64_t ldO,
                                         void* out_loc, void* out_gloc, int64_t ld) {
    if (!m || !op || !srow || !prow || !out_logpsi || !O || !out_loc || B < 0) return NQ_ERR_ARG;
    nq_ctx_t ctx = m->ctx;
    if (m->doubled() != (scol != nullptr) || m->doubled() != (pcol != nullptr)) return nq_fail(ctx, NQ_ERR_ARG, "row/col configuration mismatch");
    if (ldO < m->P || (out_gloc && ld < m->P)) return nq_fail(ctx, NQ_ERR_SHAPE, "leading dimension < P");
    if (out_gloc && op->space != NQ_SUPER) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "gradient estimator is defined for Liouvillians only");
    if (sdtype != NQ_F32 && sdtype != NQ_F64) return nq_fail(ctx, NQ_ERR_ARG, "state arrays must be NQ_F32 or NQ_F64");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (nq_is_device_ptr(srow) || (scol && nq_is_device_ptr(scol)))
        return nq_fail(ctx, NQ_ERR_ARG, "nq_logpsi_grad_local_host takes HOST configurations (device arrays: nq_pack_states + the packed entry)");
    if (!nq_is_device_ptr(prow) || (pcol && !nq_is_device_ptr(pcol)) || !nq_is_device_ptr(O) || !nq_is_device_ptr(out_logpsi) ||
        !nq_is_device_ptr(out_loc) || (out_gloc && !nq_is_device_ptr(out_gloc)))
        return nq_fail(ctx, NQ_ERR_ARG, "the outputs of the fused entry point are device-resident buffers");
    if (B == 0) return NQ_OK;
    const int N = m->N, W = nq_words(N);
    const size_t fs = nq_dtype_size(sdtype), es = nq_dtype_size(m->out_dtype), cs = nq_dtype_size(nq_complex_of(m->dtype));
    char* stage_r = (char*)nq_scratch(ctx, SL_IN0, (size_t)B * N * fs);
    char* stage_c = scol ? (char*)nq_scratch(ctx, SL_IN1, (size_t)B * N * fs) : nullptr;
    if (!stage_r || (scol && !stage_c)) return NQ_ERR_ALLOC;
    if (!ctx->side_stream) {
        NQ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
        for (auto& e : ctx->side_ev) NQ_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const int64_t rnd = (int64_t)ctx->num_sms * 16;
    int64_t bounds[4] = {0, B, B, B};
    int npiece = 1;
    if (B >= 16 * rnd) { bounds[1] = 2 * rnd; bounds[2] = 9 * rnd; bounds[3] = B; npiece = 3; }
    else if (B >= 6 * rnd) { bounds[1] = 2 * rnd; bounds[2] = B; npiece = 2; }
    ctx->shift_pending = false; ctx->rowmax_ptr = nullptr;          // new rows: records of a centring pass are stale
    cudaStream_t main_stream = ctx->stream, side = ctx->side_stream;
    // the staging and packed buffers may still be read by earlier work on the main stream
    NQ_CUDA(ctx, cudaEventRecord(ctx->side_ev[3], main_stream));
    NQ_CUDA(ctx, cudaStreamWaitEvent(side, ctx->side_ev[3], 0));
    for (int c = 0; c < npiece; c++) {
        const int64_t c0 = bounds[c], n = bounds[c + 1] - bounds[c];
        const size_t foff = (size_t)c0 * N * fs;
        NQ_CUDA(ctx, cudaMemcpyAsync(stage_r + foff, (const char*)srow + foff, (size_t)n * N * fs, cudaMemcpyHostToDevice, side));
        if (scol) NQ_CUDA(ctx, cudaMemcpyAsync(stage_c + foff, (const char*)scol + foff, (size_t)n * N * fs, cudaMemcpyHostToDevice, side));
        ctx->stream = side;                                         // the packing kernels follow their copies
        int rc = nq_pack_device(ctx, m->hilb, N, n, stage_r + foff, sdtype, prow + c0 * W);
        if (rc == NQ_OK && scol) rc = nq_pack_device(ctx, m->hilb, N, n, stage_c + foff, sdtype, pcol + c0 * W);
        ctx->stream = main_stream;
        if (rc != NQ_OK) return rc;
        NQ_CUDA(ctx, cudaEventRecord(ctx->side_ev[c], side));
        NQ_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->side_ev[c], 0));
        NQ_CHECK(nq_local_device(m, op, prow + c0 * W, pcol ? pcol + c0 * W : nullptr, n, (char*)out_logpsi + (size_t)c0 * es,
                                 (char*)O + (size_t)c0 * ldO * es, ldO, (char*)out_loc + (size_t)c0 * cs,
                                 out_gloc ? (char*)out_gloc + (size_t)c0 * ld * cs : nullptr, ld));
    }
    return NQ_OK;
}

#ifdef NQ_PROFILE_PHASES
extern "C" int nq_debug_phase_cycles(unsigned long long out[8], int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, nq_phase_cycles, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(nq_phase_cycles, z, sizeof z); }
    return 0;
}
#endif
