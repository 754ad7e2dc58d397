// nq_ctx.cu -- context, error reporting, scratch memory, configuration packing.
#include "nq_common.cuh"
#include <cstdarg>

int nq_fail(nq_ctx_t ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

bool nq_is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

void* nq_scratch(nq_ctx_t ctx, int slot, size_t bytes) {
    auto& s = ctx->slots[slot];
    if (bytes <= s.cap && s.p) return s.p;
    if (s.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(s.p);
        s.p = nullptr; s.cap = 0;
    }
    size_t cap = bytes + bytes / 4 + 256;
    if (cudaMalloc(&s.p, cap) != cudaSuccess) {
        cudaGetLastError();
        nq_fail(ctx, NQ_ERR_ALLOC, "scratch slot %d: cudaMalloc(%zu) failed", slot, cap);
        s.p = nullptr;
        return nullptr;
    }
    s.cap = cap;
    return s.p;
}

extern "C" int nq_version(void) { return NQ_VERSION; }

extern "C" const char* nq_status_string(int s) {
    switch (s) {
        case NQ_OK: return "ok";
        case NQ_ERR_ARG: return "bad argument";
        case NQ_ERR_SHAPE: return "shape mismatch";
        case NQ_ERR_CUDA: return "CUDA error";
        case NQ_ERR_NCCL: return "NCCL error";
        case NQ_ERR_NOT_POSDEF: return "matrix not positive definite";
        case NQ_ERR_NOT_CONVERGED: return "iterative solver did not converge";
        case NQ_ERR_UNSUPPORTED: return "unsupported configuration";
        case NQ_ERR_ALLOC: return "device allocation failed";
        default: return "unknown status";
    }
}

extern "C" int nq_ctx_create(int device, void* stream, nq_ctx_t* out) {
    if (!out) return NQ_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return NQ_ERR_CUDA; }
    if (device < 0 || device >= n) return NQ_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return NQ_ERR_CUDA; }
    nq_ctx_t c = new nq_ctx_s();
    c->device = device;
    // the handle is used as given; NULL is CUDA's default stream (which is also torch's default current stream)
    c->stream = (cudaStream_t)stream;
    c->own_stream = false;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) == cudaSuccess) {
        c->num_sms = p.multiProcessorCount;
        c->smem_optin = p.sharedMemPerBlockOptin;
    }
    *out = c;
    return NQ_OK;
}

extern "C" int nq_comm_destroy(nq_ctx_t ctx);

extern "C" int nq_ctx_destroy(nq_ctx_t ctx) {
    if (!ctx) return NQ_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nq_comm_destroy(ctx);
    for (auto& s : ctx->slots) if (s.p) cudaFree(s.p);
    if (ctx->rowmax) cudaFree(ctx->rowmax);
    if (ctx->shift) cudaFree(ctx->shift);
    for (auto& e : ctx->side_ev) if (e) cudaEventDestroy(e);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return NQ_OK;
}

extern "C" const char* nq_last_error(nq_ctx_t ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int nq_ctx_sync(nq_ctx_t ctx) {
    if (!ctx) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NQ_OK;
}

extern "C" int nq_ctx_launch_count(nq_ctx_t ctx, uint64_t* out) {
    if (!ctx || !out) return NQ_ERR_ARG;
    *out = ctx->launches;
    return NQ_OK;
}

extern "C" int nq_ctx_last_info(nq_ctx_t ctx, int64_t* out) {
    if (!ctx || !out) return NQ_ERR_ARG;
    *out = ctx->info;
    return NQ_OK;
}

// --------------------------------------------------------------------------------------
// pack / unpack: reference float configurations [N,B] <-> uint64 words [B][W64]
// one warp per (sample, word): ballot over 32 sites twice
// --------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_kernel(const T* __restrict__ sigma, uint64_t* __restrict__ packed, int N, int64_t B,
                            int W64, int hilb) {
    int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (wid >= B * W64) return;
    int64_t b = wid / W64;
    int w = (int)(wid % W64);
    const T* s = sigma + b * N;
    uint64_t word = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        int j = w * 64 + h * 32 + lane;
        int bit = 0;
        if (j < N) {
            T v = s[j];
            bit = hilb == NQ_SPIN ? (v > T(0)) : (v > T(0.5));
        }
        unsigned m = __ballot_sync(0xffffffffu, bit);
        word |= (uint64_t)m << (32 * h);
    }
    if (lane == 0) packed[wid] = word;
}

template <typename T>
__global__ void unpack_kernel(const uint64_t* __restrict__ packed, T* __restrict__ sigma, int N, int64_t B,
                              int W64, int hilb) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    int64_t b = i / N;
    int j = (int)(i % N);
    int d = (int)((packed[b * W64 + (j >> 6)] >> (j & 63)) & 1ull);
    sigma[i] = hilb == NQ_SPIN ? T(2 * d - 1) : T(d);
}

extern "C" int nq_states_words(int N) { return nq_words(N); }

int nq_pack_device(nq_ctx_t ctx, nq_hilbert h, int N, int64_t B, const void* dsigma, nq_dtype sdtype,
                   uint64_t* dpacked) {
    if (B == 0) return NQ_OK;
    int W64 = nq_words(N);
    int64_t warps = B * W64;
    int block = 256;
    int64_t grid = (warps * 32 + block - 1) / block;
    if (sdtype == NQ_F64) NQ_LAUNCH(ctx, pack_kernel<double>, (unsigned)grid, block, 0, (const double*)dsigma, dpacked, N, B, W64, (int)h);
    else if (sdtype == NQ_F32) NQ_LAUNCH(ctx, pack_kernel<float>, (unsigned)grid, block, 0, (const float*)dsigma, dpacked, N, B, W64, (int)h);
    else return nq_fail(ctx, NQ_ERR_ARG, "state arrays must be NQ_F32 or NQ_F64");
    return NQ_OK;
}

int nq_unpack_device(nq_ctx_t ctx, nq_hilbert h, int N, int64_t B, const uint64_t* dpacked, void* dsigma,
                     nq_dtype sdtype) {
    if (B == 0) return NQ_OK;
    int W64 = nq_words(N);
    int block = 256;
    int64_t grid = (B * N + block - 1) / block;
    if (sdtype == NQ_F64) NQ_LAUNCH(ctx, unpack_kernel<double>, (unsigned)grid, block, 0, dpacked, (double*)dsigma, N, B, W64, (int)h);
    else if (sdtype == NQ_F32) NQ_LAUNCH(ctx, unpack_kernel<float>, (unsigned)grid, block, 0, dpacked, (float*)dsigma, N, B, W64, (int)h);
    else return nq_fail(ctx, NQ_ERR_ARG, "state arrays must be NQ_F32 or NQ_F64");
    return NQ_OK;
}

extern "C" int nq_pack_states(nq_ctx_t ctx, nq_hilbert h, int N, int64_t B, const void* sigma, nq_dtype sdtype,
                              uint64_t* packed) {
    if (!ctx || !sigma || !packed || N <= 0 || B < 0) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const void* ds = st.in(0, sigma, (size_t)B * N * nq_dtype_size(sdtype));
    uint64_t* dp = (uint64_t*)st.out(1, packed, (size_t)B * nq_words(N) * 8);
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(nq_pack_device(ctx, h, N, B, ds, sdtype, dp));
    return st.finish();
}

extern "C" int nq_unpack_states(nq_ctx_t ctx, nq_hilbert h, int N, int64_t B, const uint64_t* packed,
                                void* sigma, nq_dtype sdtype) {
    if (!ctx || !sigma || !packed || N <= 0 || B < 0) return NQ_ERR_ARG;
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    NqStage st(ctx);
    const uint64_t* dp = (const uint64_t*)st.in(0, packed, (size_t)B * nq_words(N) * 8);
    void* ds = st.out(1, sigma, (size_t)B * N * nq_dtype_size(sdtype));
    if (st.status != NQ_OK) return st.status;
    NQ_CHECK(nq_unpack_device(ctx, h, N, B, dp, ds, sdtype));
    return st.finish();
}
