// nq_comm.cu -- parallel backend: NCCL all-reduce over NVLink (C1-C5 of SURVEY.md section 2.1).
// ref: Parallel/not_parallel.jl:1-19, Parallel/MPI/mpi.jl:21-74 (workers_sum!, workers_mean!).
// libnccl is resolved at run time (dlopen of the NCCL already loaded by torch, or libnccl.so.2) so that
// libnqcuda itself loads on machines without NCCL; without a communicator every collective is the identity
// (NotParallel).
#include "nq_internal.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& api() {
    static NcclApi a;
    if (a.handle) return a;
    const char* env = getenv("NQ_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
    return a;
}

template <typename T>
__global__ void scale_kernel(T* __restrict__ x, int64_t n, T s) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] *= s;
}
}  // namespace

static_assert(sizeof(ncclUniqueId) == NQ_UNIQUE_ID_BYTES, "ncclUniqueId size");

extern "C" int nq_comm_unique_id(uint8_t id[NQ_UNIQUE_ID_BYTES]) {
    if (!id) return NQ_ERR_ARG;
    NcclApi& a = api();
    if (!a.ok) return NQ_ERR_NCCL;
    ncclUniqueId u;
    if (a.GetUniqueId(&u) != ncclSuccess) return NQ_ERR_NCCL;
    memcpy(id, &u, NQ_UNIQUE_ID_BYTES);
    return NQ_OK;
}

extern "C" int nq_comm_init(nq_ctx_t ctx, int nranks, int rank, const uint8_t id[NQ_UNIQUE_ID_BYTES]) {
    if (!ctx || !id || nranks <= 0 || rank < 0 || rank >= nranks) return NQ_ERR_ARG;
    NcclApi& a = api();
    if (!a.ok) return nq_fail(ctx, NQ_ERR_NCCL, "libnccl not found (set NQ_NCCL_LIB)");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) nq_comm_destroy(ctx);
    ncclUniqueId u;
    memcpy(&u, id, NQ_UNIQUE_ID_BYTES);
    ncclComm_t comm;
    ncclResult_t r = a.CommInitRank(&comm, nranks, u, rank);
    if (r != ncclSuccess) return nq_fail(ctx, NQ_ERR_NCCL, "ncclCommInitRank: %s", a.GetErrorString(r));
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return NQ_OK;
}

extern "C" int nq_comm_destroy(nq_ctx_t ctx) {
    if (!ctx) return NQ_ERR_ARG;
    if (ctx->nccl_comm) {
        cudaStreamSynchronize(ctx->stream);
        api().CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->nranks = 1;
    ctx->rank = 0;
    return NQ_OK;
}

extern "C" int nq_comm_size(nq_ctx_t ctx, int* nranks, int* rank) {
    if (!ctx) return NQ_ERR_ARG;
    if (nranks) *nranks = ctx->nranks;
    if (rank) *rank = ctx->rank;
    return NQ_OK;
}

extern "C" int nq_comm_set_global_samples(nq_ctx_t ctx, int64_t ns_total) {
    if (!ctx || ns_total < 0) return NQ_ERR_ARG;
    ctx->ns_total = ns_total;
    return NQ_OK;
}

int nq_allreduce_device(nq_ctx_t ctx, void* buf, int64_t n, nq_dtype dtype, bool mean) {
    if (n <= 0) return NQ_OK;
    const bool dbl = nq_dtype_is_double(dtype);
    const int64_t nreal = n * (nq_dtype_is_complex(dtype) ? 2 : 1);
    if (ctx->nccl_comm) {
        NcclApi& a = api();
        ncclResult_t r = a.AllReduce(buf, buf, (size_t)nreal, dbl ? ncclDouble : ncclFloat, ncclSum,
                                     (ncclComm_t)ctx->nccl_comm, ctx->stream);
        if (r != ncclSuccess) return nq_fail(ctx, NQ_ERR_NCCL, "ncclAllReduce: %s", a.GetErrorString(r));
        ctx->launches++;
    }
    if (mean && ctx->nranks > 1) {
        unsigned g = (unsigned)((nreal + 255) / 256);
        if (dbl) NQ_LAUNCH(ctx, scale_kernel<double>, g, 256, 0, (double*)buf, nreal, 1.0 / ctx->nranks);
        else NQ_LAUNCH(ctx, scale_kernel<float>, g, 256, 0, (float*)buf, nreal, 1.0f / ctx->nranks);
    }
    return NQ_OK;
}

extern "C" int nq_allreduce_sum(nq_ctx_t ctx, void* buf, int64_t n, nq_dtype dtype) {
    if (!ctx || !buf || n < 0) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!nq_is_device_ptr(buf)) return nq_fail(ctx, NQ_ERR_ARG, "all-reduce works in place on device buffers");
    return nq_allreduce_device(ctx, buf, n, dtype, false);
}

extern "C" int nq_allreduce_mean(nq_ctx_t ctx, void* buf, int64_t n, nq_dtype dtype) {
    if (!ctx || !buf || n < 0) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!nq_is_device_ptr(buf)) return nq_fail(ctx, NQ_ERR_ARG, "all-reduce works in place on device buffers");
    return nq_allreduce_device(ctx, buf, n, dtype, true);
}
