// nq_opdev.cuh -- device view of an operator's connection tables and the in-order visit of the connections of one
// term (shared by the estimator kernels of nq_operator.cu and the OperatorRule sampler of nq_sampler.cu).
// ref: Operators/Operators/KLocalOperator.jl:183-199 (row_valdiff!), KLocalOperatorTensor.jl:129-157.
#pragma once
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
    const uint64_t* recs;   // flat (condition, flips, mel) records, see nq_operator_create
    int n_recs;
    const int2* lut_units;  // site tables, see build_site_luts
    const ulonglong2* lut_groups;
    const double2* lut_tab;
    int n_lut_units;
};

OpDev op_dev(nq_operator_t op) {
    OpDev d;
    d.n_terms = op->n_terms; d.super = op->space == NQ_SUPER;
    d.part_nsites = op->part_nsites; d.part_site_ptr = op->part_site_ptr; d.part_sites = op->part_sites;
    d.part_row0 = op->part_row0; d.row_ptr = op->row_ptr; d.entry_mel = op->entry_mel;
    d.entry_flip = op->entry_flip; d.term_left = op->term_left; d.term_right = op->term_right;
    d.recs = op->recs; d.n_recs = op->n_recs;
    d.lut_units = (const int2*)op->lut_units; d.lut_groups = (const ulonglong2*)op->lut_groups;
    d.lut_tab = (const double2*)op->lut_tab; d.n_lut_units = op->n_lut_units;
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

}  // namespace
