// nq_symm.cu -- symmetrised machines (NDMSymm): a "bare" machine whose parameters are tied by a list of site
// permutations, and the symmetrisation of its gradient rows.
//
// ref: Networks/MixedDensityMatrix/NDMSymm.jl:27-34 (update! -> set_bare_params!), :79-128 (set_bare_params!),
//      :130-181 (construct_grad_matrices: 0/1 gather matrices, 1/n for the local biases),
//      NDMSymmBatched.jl:22-36 (symmetrize_grad_NDM_batched!: grad_symm = G grad_bare, one mul! per field).
//
// The reference multiplies every gradient field by a dense 0/1 matrix.  Here the map is a CSR gather list
// (symm parameter p <- scale_p * sum of the bare rows idx[ptr[p] .. ptr[p+1])) applied per sample: one CTA stages the
// bare row of a sample in shared memory with coalesced loads, gathers from there and streams the symmetrised row out
// coalesced -- HBM traffic = one read of the bare rows + one write of the symmetrised rows.
#include "nq_internal.cuh"

struct nq_symm_s {
    nq_ctx_t ctx;
    nq_machine_t bare;
    int64_t Ps, Pb;
    int64_t* ptr;      // device [Ps + 1]
    int32_t* idx;      // device [ptr[Ps]]
    double* scale;     // device [Ps]
    int32_t* src;      // device [Pb]: bare parameter q = symm parameter src[q]
    int n_avg;
    int64_t* avg;      // device [n_avg][2]: ranges of the symm vector replaced by their mean (local biases)
    void* w;           // device [Ps] symm parameters, machine dtype
};

namespace {

template <typename E>
__global__ void symm_gather_kernel(const E* __restrict__ Ob, int64_t ldb, int64_t Pb, E* __restrict__ Os, int64_t lds,
                                   int64_t Ps, int64_t Ns, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                   const double* __restrict__ scale, int staged) {
    typedef typename elem_traits<E>::real T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E* row = (E*)smem_raw;
    for (int64_t s = blockIdx.x; s < Ns; s += gridDim.x) {
        const E* __restrict__ src = Ob + s * ldb;
        if (staged) {
            __syncthreads();
            for (int64_t q = threadIdx.x; q < Pb; q += blockDim.x) row[q] = src[q];
            __syncthreads();
        }
        for (int64_t p = threadIdx.x; p < Ps; p += blockDim.x) {
            E acc = make_zero<E>();
            for (int64_t e = ptr[p]; e < ptr[p + 1]; e++) acc += staged ? row[idx[e]] : src[idx[e]];
            Os[s * lds + p] = rscale((T)scale[p], acc);
        }
    }
}

// w[a..b) <- mean(w[a..b)) for every range (one block per range), then bare[q] = w[src[q]]
template <typename T>
__global__ void symm_average_kernel(T* __restrict__ w, const int64_t* __restrict__ avg) {
    __shared__ double sh[32];
    const int64_t a = avg[2 * blockIdx.x], b = avg[2 * blockIdx.x + 1];
    double s = 0.0;
    for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) s += (double)w[i];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) r += __shfl_xor_sync(0xffffffffu, r, m);
        if (threadIdx.x == 0) sh[0] = r / (double)(b - a);
    }
    __syncthreads();
    const T mean = (T)sh[0];
    for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) w[i] = mean;
}

template <typename T>
__global__ void symm_expand_kernel(const T* __restrict__ w, const int32_t* __restrict__ src, T* __restrict__ bare, int64_t Pb) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q < Pb) bare[q] = w[src[q]];
}

template <typename T>
__global__ void symm_axpy_kernel(T* __restrict__ w, const T* __restrict__ dw, T eta, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) w[i] -= eta * dw[i];
}

template <typename T>
int expand(nq_symm_t g) {
    nq_ctx_t ctx = g->ctx;
    if (g->n_avg > 0) NQ_LAUNCH(ctx, symm_average_kernel<T>, (unsigned)g->n_avg, 256, 0, (T*)g->w, g->avg);
    NQ_LAUNCH(ctx, symm_expand_kernel<T>, (unsigned)((g->Pb + 255) / 256), 256, 0, (const T*)g->w, g->src, (T*)g->bare->params, g->Pb);
    g->bare->etab_valid = false;
    return NQ_OK;
}

int expand_any(nq_symm_t g) { return g->bare->dtype == NQ_F64 ? expand<double>(g) : expand<float>(g); }

template <typename E>
int gather(nq_symm_t g, const void* Ob, int64_t ldb, int64_t Ns, void* Os, int64_t lds) {
    nq_ctx_t ctx = g->ctx;
    size_t smem = (size_t)g->Pb * sizeof(E);
    int staged = smem <= ctx->smem_optin ? 1 : 0;
    if (!staged) smem = 0;
    auto kern = symm_gather_kernel<E>;
    if (staged) NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    NQ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    int64_t grid = (int64_t)ctx->num_sms * (per_sm > 0 ? per_sm : 1);
    if (grid > Ns) grid = Ns;
    NQ_LAUNCH(ctx, kern, (unsigned)grid, 256, smem, (const E*)Ob, ldb, g->Pb, (E*)Os, lds, g->Ps, Ns, g->ptr, g->idx, g->scale, staged);
    return NQ_OK;
}

template <typename T> bool upload(T** dst, const T* src, size_t n) {
    if (cudaMalloc((void**)dst, (n ? n : 1) * sizeof(T)) != cudaSuccess) return false;
    return n == 0 || cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

}  // namespace

extern "C" int nq_symm_destroy(nq_symm_t g) {
    if (!g) return NQ_ERR_ARG;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    cudaFree(g->ptr); cudaFree(g->idx); cudaFree(g->scale); cudaFree(g->src); cudaFree(g->avg); cudaFree(g->w);
    delete g;
    return NQ_OK;
}

extern "C" int nq_symm_create(nq_machine_t bare, int64_t Ps, const int64_t* ptr, const int32_t* idx, const double* scale,
                              const int32_t* src, int n_avg, const int64_t* avg_ranges, nq_symm_t* out) {
    if (!bare || !ptr || !idx || !scale || !src || !out || Ps <= 0 || n_avg < 0) return NQ_ERR_ARG;
    *out = nullptr;
    nq_ctx_t ctx = bare->ctx;
    if (nq_dtype_is_complex(bare->dtype)) return nq_fail(ctx, NQ_ERR_UNSUPPORTED, "symmetrised machines take real parameters (NDMSymm.jl:21)");
    const int64_t Pb = bare->P;
    if (ptr[0] != 0) return nq_fail(ctx, NQ_ERR_ARG, "ptr[0] != 0");
    for (int64_t p = 0; p < Ps; p++) if (ptr[p + 1] < ptr[p]) return nq_fail(ctx, NQ_ERR_ARG, "ptr not monotone");
    for (int64_t e = 0; e < ptr[Ps]; e++) if (idx[e] < 0 || idx[e] >= Pb) return nq_fail(ctx, NQ_ERR_ARG, "bare index %d outside 0..%lld", idx[e], (long long)Pb - 1);
    for (int64_t q = 0; q < Pb; q++) if (src[q] < 0 || src[q] >= Ps) return nq_fail(ctx, NQ_ERR_ARG, "symm index %d outside 0..%lld", src[q], (long long)Ps - 1);
    for (int i = 0; i < n_avg; i++)
        if (avg_ranges[2 * i] < 0 || avg_ranges[2 * i + 1] > Ps || avg_ranges[2 * i] >= avg_ranges[2 * i + 1]) return nq_fail(ctx, NQ_ERR_ARG, "bad averaging range %d", i);
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    nq_symm_t g = new nq_symm_s();
    g->ctx = ctx; g->bare = bare; g->Ps = Ps; g->Pb = Pb; g->n_avg = n_avg;
    g->ptr = nullptr; g->idx = nullptr; g->scale = nullptr; g->src = nullptr; g->avg = nullptr; g->w = nullptr;
    bool ok = upload(&g->ptr, ptr, (size_t)Ps + 1) && upload(&g->idx, idx, (size_t)ptr[Ps]) && upload(&g->scale, scale, (size_t)Ps) &&
              upload(&g->src, src, (size_t)Pb) && upload(&g->avg, avg_ranges, (size_t)2 * n_avg) &&
              cudaMalloc(&g->w, (size_t)Ps * nq_dtype_size(bare->dtype)) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        nq_symm_destroy(g);
        return nq_fail(ctx, NQ_ERR_ALLOC, "symmetry map allocation failed");
    }
    cudaMemsetAsync(g->w, 0, (size_t)Ps * nq_dtype_size(bare->dtype), ctx->stream);
    *out = g;
    return NQ_OK;
}

extern "C" int nq_symm_nparams(nq_symm_t g, int64_t* Ps) {
    if (!g || !Ps) return NQ_ERR_ARG;
    *Ps = g->Ps;
    return NQ_OK;
}

extern "C" int nq_symm_set_params(nq_symm_t g, const void* w, int64_t Ps) {
    if (!g || !w) return NQ_ERR_ARG;
    nq_ctx_t ctx = g->ctx;
    if (Ps != g->Ps) return nq_fail(ctx, NQ_ERR_SHAPE, "expected %lld parameters, got %lld", (long long)g->Ps, (long long)Ps);
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NQ_CUDA(ctx, cudaMemcpyAsync(g->w, w, (size_t)Ps * nq_dtype_size(g->bare->dtype), cudaMemcpyDefault, ctx->stream));
    NQ_CHECK(expand_any(g));
    if (!nq_is_device_ptr(w)) NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NQ_OK;
}

extern "C" int nq_symm_get_params(nq_symm_t g, void* w, int64_t Ps) {
    if (!g || !w) return NQ_ERR_ARG;
    nq_ctx_t ctx = g->ctx;
    if (Ps != g->Ps) return nq_fail(ctx, NQ_ERR_SHAPE, "expected %lld parameters, got %lld", (long long)g->Ps, (long long)Ps);
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NQ_CUDA(ctx, cudaMemcpyAsync(w, g->w, (size_t)Ps * nq_dtype_size(g->bare->dtype), cudaMemcpyDefault, ctx->stream));
    if (!nq_is_device_ptr(w)) NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NQ_OK;
}

extern "C" int nq_symm_update(nq_symm_t g, const void* dw, double eta) {
    if (!g || !dw) return NQ_ERR_ARG;
    nq_ctx_t ctx = g->ctx;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const void* d = st.in(SL_IN0, dw, (size_t)g->Ps * nq_dtype_size(g->bare->dtype));
    if (st.status != NQ_OK) return st.status;
    const unsigned grid = (unsigned)((g->Ps + 255) / 256);
    if (g->bare->dtype == NQ_F64) NQ_LAUNCH(ctx, symm_axpy_kernel<double>, grid, 256, 0, (double*)g->w, (const double*)d, eta, g->Ps);
    else NQ_LAUNCH(ctx, symm_axpy_kernel<float>, grid, 256, 0, (float*)g->w, (const float*)d, (float)eta, g->Ps);
    NQ_CHECK(expand_any(g));
    if (!nq_is_device_ptr(dw)) NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NQ_OK;
}

extern "C" int nq_symm_gradient(nq_symm_t g, const void* Obare, int64_t ldb, int64_t Ns, nq_dtype dtype, void* Osymm,
                                int64_t lds) {
    if (!g || !Obare || !Osymm || Ns < 0) return NQ_ERR_ARG;
    nq_ctx_t ctx = g->ctx;
    if (ldb < g->Pb || lds < g->Ps) return nq_fail(ctx, NQ_ERR_SHAPE, "leading dimension smaller than the row");
    if (Ns == 0) return NQ_OK;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const size_t es = nq_dtype_size(dtype);
    const void* db = st.in(SL_HOSTO, Obare, (size_t)ldb * Ns * es);
    void* ds = st.out2d(SL_HOSTG, Osymm, (size_t)g->Ps * es, (size_t)lds * es, (size_t)Ns);
    if (st.status != NQ_OK) return st.status;
    switch (dtype) {
        case NQ_F32: NQ_CHECK(gather<float>(g, db, ldb, Ns, ds, lds)); break;
        case NQ_F64: NQ_CHECK(gather<double>(g, db, ldb, Ns, ds, lds)); break;
        case NQ_C64: NQ_CHECK(gather<cxf>(g, db, ldb, Ns, ds, lds)); break;
        default: NQ_CHECK(gather<cxd>(g, db, ldb, Ns, ds, lds)); break;
    }
    return st.finish();
}
