// nq_fullspace.cu -- the reference's validation tools on the device: full-space reconstruction of the
// state (`ket`, `densitymatrix`) and the ExactSampler (probability table of ALL basis states, inverse-CDF
// draws).  Indexable spaces only: 2^N (ket) or 4^N (density matrix) entries.
//
// ref: utils/densitymatrix.jl:9-62 (densitymatrix, ket), Samplers/Exact.jl:135-181 (init_sampler!,
//      samplenext!), Hilbert/HomogeneousSpin.jl:156-179 (set_index!: basis number i <-> digits of i-1,
//      site 1 least significant), Hilbert/SuperOpSpace (super index = (col-1) D + row).
//
// Basis numbers are generated on the fly (iota kernel -> packed words, one word per configuration since
// N <= 31 here) and go through the machine kernels of nq_machines.cu in chunks; the table is a
// deterministic three-phase scan in double precision (block scans -> scan of the block sums -> offsets), followed by
// an exact running maximum so that the table is sorted whatever the rounding.
#include "nq_internal.cuh"

namespace {

constexpr int64_t FS_CHUNK = 1 << 20;   // configurations per machine launch
constexpr int FS_BLK = 1024;            // elements per scan block

__global__ void iota_kernel(uint64_t* __restrict__ prow, uint64_t* __restrict__ pcol, int64_t i0, int64_t n, int N) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t idx = (uint64_t)(i0 + i);
    if (pcol) {
        prow[i] = idx & ((1ull << N) - 1);
        pcol[i] = idx >> N;
    } else {
        prow[i] = idx;
    }
}

__device__ __forceinline__ cxd to_cxd(float a) { return cxd((double)a, 0.0); }
__device__ __forceinline__ cxd to_cxd(double a) { return cxd(a, 0.0); }
template <typename T> __device__ __forceinline__ cxd to_cxd(cx<T> a) { return cxd((double)a.re, (double)a.im); }
template <typename E> __device__ __forceinline__ E from_cxd(cxd z);
template <> __device__ __forceinline__ float from_cxd<float>(cxd z) { return (float)z.re; }
template <> __device__ __forceinline__ double from_cxd<double>(cxd z) { return z.re; }
template <> __device__ __forceinline__ cxf from_cxd<cxf>(cxd z) { return cxf((float)z.re, (float)z.im); }
template <> __device__ __forceinline__ cxd from_cxd<cxd>(cxd z) { return z; }

template <typename E> __device__ __forceinline__ double lp_of(E v) { return 2.0 * (double)v; }
template <> __device__ __forceinline__ double lp_of<cxf>(cxf v) { return 2.0 * (double)v.re; }
template <> __device__ __forceinline__ double lp_of<cxd>(cxd v) { return 2.0 * v.re; }

__device__ __forceinline__ double block_reduce(double v, bool is_max, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        double o = __shfl_xor_sync(0xffffffffu, v, m);
        v = is_max ? fmax(v, o) : v + o;
    }
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double r = lane < nw ? sh[lane] : (is_max ? -INFINITY : 0.0);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            double o = __shfl_xor_sync(0xffffffffu, r, m);
            r = is_max ? fmax(r, o) : r + o;
        }
        if (lane == 0) sh[0] = r;
    }
    __syncthreads();
    double r = sh[0];
    __syncthreads();
    return r;
}

// lp[i0 + i] = 2 Re log psi_i (double) and the block maxima
template <typename E>
__global__ void logprob_kernel(const E* __restrict__ lpsi, double* __restrict__ lp, int64_t n, double* __restrict__ bmax) {
    __shared__ double sh[32];
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double v = -INFINITY;
    if (i < n) { v = lp_of(lpsi[i]); lp[i] = v; }
    double mx = block_reduce(v, true, sh);
    if (threadIdx.x == 0) bmax[blockIdx.x] = mx;
}

// single block: out[0] = max / sum of v[0..n)
__global__ void final_reduce_kernel(const double* __restrict__ v, int64_t n, int is_max, double* __restrict__ out) {
    __shared__ double sh[32];
    double a = is_max ? -INFINITY : 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a = is_max ? fmax(a, v[i]) : a + v[i];
    double r = block_reduce(a, is_max != 0, sh);
    if (threadIdx.x == 0) out[0] = r;
}

// p = exp(lp - max), scanned inside blocks of FS_BLK in place (inclusive); the block sum is the LAST scanned element, so
// that the table stays monotone across block boundaries whatever the rounding
__global__ void prob_scan_kernel(double* __restrict__ p, int64_t n, const double* __restrict__ mx, double* __restrict__ bsum) {
    __shared__ double sh[FS_BLK];
    int64_t i = blockIdx.x * (int64_t)FS_BLK + threadIdx.x;
    sh[threadIdx.x] = i < n ? exp(p[i] - mx[0]) : 0.0;
    __syncthreads();
    for (int d = 1; d < FS_BLK; d <<= 1) {
        double t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0.0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) p[i] = sh[threadIdx.x];
    if (threadIdx.x == FS_BLK - 1) bsum[blockIdx.x] = sh[FS_BLK - 1];
}

// exclusive scan of the block sums in place, sequential (at most 2^20 block sums, usually a few): boff[b+1] is exactly
// fl(boff[b] + S_b), which is what keeps cdf monotone; total -> tot[0]
__global__ void scan_sums_kernel(double* __restrict__ bsum, int64_t nb, double* __restrict__ tot) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double carry = 0.0;
    for (int64_t b = 0; b < nb; b++) {
        double v = bsum[b];
        bsum[b] = carry;
        carry += v;
    }
    tot[0] = carry;
}

// cdf = min((block offset + in-block scan) / total, 1), then an in-block running maximum: a parallel prefix sum associates
// differently for neighbouring elements, so two entries that differ by less than an ulp of the running total can come out
// in the wrong order; max is exact under any association, so the three-phase running maximum below gives the same table
// as a sequential pass and searchsortedfirst sees a sorted array.
__global__ void scan_finish_kernel(double* __restrict__ p, int64_t n, const double* __restrict__ boff,
                                   const double* __restrict__ tot, double* __restrict__ bmx) {
    __shared__ double sh[FS_BLK];
    int64_t i = blockIdx.x * (int64_t)FS_BLK + threadIdx.x;
    sh[threadIdx.x] = i < n ? fmin((boff[blockIdx.x] + p[i]) / tot[0], 1.0) : 0.0;
    __syncthreads();
    for (int d = 1; d < FS_BLK; d <<= 1) {
        double t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0.0;
        __syncthreads();
        sh[threadIdx.x] = fmax(sh[threadIdx.x], t);
        __syncthreads();
    }
    if (i < n) p[i] = sh[threadIdx.x];
    if (threadIdx.x == FS_BLK - 1) bmx[blockIdx.x] = sh[FS_BLK - 1];
}
__global__ void maxscan_sums_kernel(double* __restrict__ bmx, int64_t nb) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double carry = 0.0;
    for (int64_t b = 0; b < nb; b++) {
        double v = bmx[b];
        bmx[b] = carry;
        carry = fmax(carry, v);
    }
}
__global__ void apply_max_kernel(double* __restrict__ p, int64_t n, const double* __restrict__ bmx) {
    int64_t i = blockIdx.x * (int64_t)FS_BLK + threadIdx.x;
    if (i < n) p[i] = fmax(p[i], bmx[blockIdx.x]);
}

struct PhiloxFS {
    uint32_t k0, k1;
    __device__ __forceinline__ void gen(uint32_t (&c)[4]) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
            uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
            uint32_t n0 = hi1 ^ c[1] ^ ka, n2 = hi0 ^ c[3] ^ kb;
            c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
    }
};

// one draw per (slot, chain): r uniform in [0,1) -> searchsortedfirst(cdf, r) (first index with cdf >= r),
// clamped to the table; the basis number becomes the packed configuration
__global__ void exact_draw_kernel(const double* __restrict__ cdf, int64_t size, int N, int doubled, uint64_t seed,
                                  int64_t chain_offset, uint64_t draw_base, int64_t B, int64_t L,
                                  const double* __restrict__ uniforms, uint64_t* __restrict__ prow,
                                  uint64_t* __restrict__ pcol, int64_t* __restrict__ indices) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= B * L) return;
    int64_t slot = t / B, chain = t - slot * B;
    double r;
    if (uniforms) {
        r = uniforms[t];
    } else {
        uint64_t gid = (uint64_t)(chain_offset + chain), ctr = draw_base + (uint64_t)slot;
        PhiloxFS ph{(uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x45584143u /* "EXAC" domain */};
        uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32)};
        ph.gen(c);
        r = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6)) * (1.0 / 9007199254740992.0);
    }
    int64_t lo = 0, hi = size;                  // first index with cdf[idx] >= r
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] < r) lo = mid + 1; else hi = mid;
    }
    if (lo > size - 1) lo = size - 1;
    if (doubled) {
        prow[t] = (uint64_t)lo & ((1ull << N) - 1);
        pcol[t] = (uint64_t)lo >> N;
    } else {
        prow[t] = (uint64_t)lo;
    }
    if (indices) indices[t] = lo + 1;
}

// out[i] = exp(log psi_i); block partial sums of |psi|^2 (ket) or of the diagonal entries (density matrix)
template <typename E>
__global__ void exp_state_kernel(const E* __restrict__ lpsi, E* __restrict__ out, int64_t i0, int64_t n, int N, int doubled,
                                 double* __restrict__ bre, double* __restrict__ bim) {
    __shared__ double sh[32];
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double ar = 0.0, ai = 0.0;
    if (i < n) {
        E v = e_exp(lpsi[i]);
        out[i0 + i] = v;
        cxd z = to_cxd(v);
        if (doubled) {
            uint64_t idx = (uint64_t)(i0 + i);
            if ((idx & ((1ull << N) - 1)) == (idx >> N)) { ar = z.re; ai = z.im; }
        } else {
            ar = z.re * z.re + z.im * z.im;
        }
    }
    double sr = block_reduce(ar, false, sh), si = block_reduce(ai, false, sh);
    if (threadIdx.x == 0) { bre[blockIdx.x] = sr; bim[blockIdx.x] = si; }
}

// out *= 1 / norm   (ket: norm = sqrt(sum |psi|^2); density matrix: norm = trace, complex)
template <typename E>
__global__ void scale_state_kernel(E* __restrict__ out, int64_t n, const double* __restrict__ tot, int doubled) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    cxd z = to_cxd(out[i]);
    if (doubled) {
        cxd t(tot[0], tot[1]);
        z = cx_div(z, t);
    } else {
        double inv = 1.0 / sqrt(tot[0]);
        z = cxd(z.re * inv, z.im * inv);
    }
    out[i] = from_cxd<E>(z);
}

// log psi of the basis numbers [i0, i0 + n) into SL_LOGPSI
int eval_chunk(nq_machine_t m, int64_t i0, int64_t n, void** lpsi_out) {
    nq_ctx_t ctx = m->ctx;
    uint64_t* pr = (uint64_t*)nq_scratch(ctx, SL_PROW, (size_t)FS_CHUNK * 8);
    uint64_t* pc = m->doubled() ? (uint64_t*)nq_scratch(ctx, SL_PCOL, (size_t)FS_CHUNK * 8) : nullptr;
    void* lp = nq_scratch(ctx, SL_LOGPSI, (size_t)FS_CHUNK * nq_dtype_size(m->out_dtype));
    if (!pr || (m->doubled() && !pc) || !lp) return NQ_ERR_ALLOC;
    NQ_LAUNCH(ctx, iota_kernel, (unsigned)((n + 255) / 256), 256, 0, pr, pc, i0, n, m->N);
    NQ_CHECK(nq_machine_eval_device(m, pr, pc, n, lp, nullptr, 0));
    *lpsi_out = lp;
    return NQ_OK;
}

int space_size(nq_machine_t m, int64_t* size) {
    const int bits = m->doubled() ? 2 * m->N : m->N;
    if (bits > 30) return nq_fail(m->ctx, NQ_ERR_UNSUPPORTED, "full-space tools need an indexable space of <= 2^30 entries (%d bits)", bits);
    *size = (int64_t)1 << bits;
    return NQ_OK;
}

template <typename E>
int fullspace_state(nq_machine_t m, int norm, E* dout, int64_t size) {
    nq_ctx_t ctx = m->ctx;
    const int64_t nchunks = (size + FS_CHUNK - 1) / FS_CHUNK;
    const int64_t bpc = (FS_CHUNK + 255) / 256, nbt = nchunks * bpc;
    double* bsum = (double*)nq_scratch(ctx, SL_W0, (size_t)nbt * 2 * 8);      // [2][nchunks * bpc]: re, im partial sums
    double* tot = (double*)nq_scratch(ctx, SL_W1, 64);
    if (!bsum || !tot) return NQ_ERR_ALLOC;
    NQ_CUDA(ctx, cudaMemsetAsync(bsum, 0, (size_t)nbt * 2 * 8, ctx->stream));
    for (int64_t c = 0; c < nchunks; c++) {
        const int64_t i0 = c * FS_CHUNK, n = size - i0 < FS_CHUNK ? size - i0 : FS_CHUNK;
        void* lp;
        NQ_CHECK(eval_chunk(m, i0, n, &lp));
        NQ_LAUNCH(ctx, exp_state_kernel<E>, (unsigned)((n + 255) / 256), 256, 0, (const E*)lp, dout, i0, n, m->N,
                  m->doubled() ? 1 : 0, bsum + c * bpc, bsum + nbt + c * bpc);
    }
    if (!norm) return NQ_OK;
    NQ_LAUNCH(ctx, final_reduce_kernel, 1, 1024, 0, bsum, nbt, 0, tot);
    NQ_LAUNCH(ctx, final_reduce_kernel, 1, 1024, 0, bsum + nbt, nbt, 0, tot + 1);
    NQ_LAUNCH(ctx, scale_state_kernel<E>, (unsigned)((size + 255) / 256), 256, 0, dout, size, tot, m->doubled() ? 1 : 0);
    return NQ_OK;
}

}  // namespace

extern "C" int nq_fullspace_size(nq_machine_t m, int64_t* size) {
    if (!m || !size) return NQ_ERR_ARG;
    return space_size(m, size);
}

extern "C" int nq_fullspace_state(nq_machine_t m, int norm, void* out) {
    if (!m || !out) return NQ_ERR_ARG;
    nq_ctx_t ctx = m->ctx;
    int64_t size;
    NQ_CHECK(space_size(m, &size));
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    void* dout = st.out(SL_OUT0, out, (size_t)size * nq_dtype_size(m->out_dtype));
    if (st.status != NQ_OK) return st.status;
    switch (m->out_dtype) {
        case NQ_F32: NQ_CHECK(fullspace_state<float>(m, norm, (float*)dout, size)); break;
        case NQ_F64: NQ_CHECK(fullspace_state<double>(m, norm, (double*)dout, size)); break;
        case NQ_C64: NQ_CHECK(fullspace_state<cxf>(m, norm, (cxf*)dout, size)); break;
        default: NQ_CHECK(fullspace_state<cxd>(m, norm, (cxd*)dout, size)); break;
    }
    return st.finish();
}

extern "C" int nq_exact_table(nq_machine_t m, double* cdf) {
    if (!m || !cdf) return NQ_ERR_ARG;
    nq_ctx_t ctx = m->ctx;
    int64_t size;
    NQ_CHECK(space_size(m, &size));
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    double* d = (double*)st.out(SL_OUT0, cdf, (size_t)size * 8);
    if (st.status != NQ_OK) return st.status;
    const int64_t nb256 = (size + 255) / 256, nb = (size + FS_BLK - 1) / FS_BLK;
    double* bmax = (double*)nq_scratch(ctx, SL_W0, (size_t)nb256 * 8);
    double* bsum = (double*)nq_scratch(ctx, SL_W1, (size_t)nb * 8);
    double* sc = (double*)nq_scratch(ctx, SL_W2, 64);
    if (!bmax || !bsum || !sc) return NQ_ERR_ALLOC;
    for (int64_t i0 = 0; i0 < size; i0 += FS_CHUNK) {
        const int64_t n = size - i0 < FS_CHUNK ? size - i0 : FS_CHUNK;
        void* lp;
        NQ_CHECK(eval_chunk(m, i0, n, &lp));
        const unsigned grid = (unsigned)((n + 255) / 256);
        double* bm = bmax + i0 / 256;
        switch (m->out_dtype) {
            case NQ_F32: NQ_LAUNCH(ctx, logprob_kernel<float>, grid, 256, 0, (const float*)lp, d + i0, n, bm); break;
            case NQ_F64: NQ_LAUNCH(ctx, logprob_kernel<double>, grid, 256, 0, (const double*)lp, d + i0, n, bm); break;
            case NQ_C64: NQ_LAUNCH(ctx, logprob_kernel<cxf>, grid, 256, 0, (const cxf*)lp, d + i0, n, bm); break;
            default: NQ_LAUNCH(ctx, logprob_kernel<cxd>, grid, 256, 0, (const cxd*)lp, d + i0, n, bm); break;
        }
    }
    NQ_LAUNCH(ctx, final_reduce_kernel, 1, 1024, 0, bmax, nb256, 1, sc);
    NQ_LAUNCH(ctx, prob_scan_kernel, (unsigned)nb, FS_BLK, 0, d, size, sc, bsum);
    NQ_LAUNCH(ctx, scan_sums_kernel, 1, 32, 0, bsum, nb, sc + 1);
    NQ_LAUNCH(ctx, scan_finish_kernel, (unsigned)nb, FS_BLK, 0, d, size, bsum, sc + 1, bmax);
    NQ_LAUNCH(ctx, maxscan_sums_kernel, 1, 32, 0, bmax, nb);
    NQ_LAUNCH(ctx, apply_max_kernel, (unsigned)nb, FS_BLK, 0, d, size, bmax);
    return st.finish();
}

extern "C" int nq_exact_sample(nq_machine_t m, const double* cdf, uint64_t seed, int64_t chain_offset, uint64_t draw_base,
                               int64_t B, int64_t L, const double* uniforms, uint64_t* prow, uint64_t* pcol,
                               int64_t* indices) {
    if (!m || !cdf || !prow || B < 0 || L < 0) return NQ_ERR_ARG;
    nq_ctx_t ctx = m->ctx;
    if (m->doubled() != (pcol != nullptr)) return nq_fail(ctx, NQ_ERR_ARG, "row/col configuration mismatch");
    int64_t size;
    NQ_CHECK(space_size(m, &size));
    if (B * L == 0) return NQ_OK;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const int64_t n = B * L;
    const double* dc = (const double*)st.in(SL_IN0, cdf, (size_t)size * 8);
    const double* du = uniforms ? (const double*)st.in(SL_IN1, uniforms, (size_t)n * 8) : nullptr;
    uint64_t* dr = (uint64_t*)st.out(SL_OUT0, prow, (size_t)n * 8);
    uint64_t* dcl = pcol ? (uint64_t*)st.out(SL_OUT1, pcol, (size_t)n * 8) : nullptr;
    int64_t* di = indices ? (int64_t*)st.out(SL_OUT2, indices, (size_t)n * 8) : nullptr;
    if (st.status != NQ_OK) return st.status;
    NQ_LAUNCH(ctx, exact_draw_kernel, (unsigned)((n + 255) / 256), 256, 0, dc, size, m->N, m->doubled() ? 1 : 0, seed,
              chain_offset, draw_base, B, L, du, dr, dcl, di);
    return st.finish();
}
