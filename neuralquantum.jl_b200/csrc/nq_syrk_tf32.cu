// nq_syrk_tf32.cu -- K7, FP32 mode: S = Oc Oc^H on the 5th-generation tensor cores.
//
// tcgen05.mma kind::tf32 with the accumulator in TMEM, 3xTF32 error compensation:
//     x = hi + lo  (hi = top 19 bits, lo = x - hi),   x y ~= lo_x hi_y + hi_x lo_y + hi_x hi_y
// which keeps ~21 mantissa bits per product.  The tensor core's FP32 accumulation truncates (measured: a
// one-sided 2^-24-per-accumulation bias, 7e-5 on a 4096-sample diagonal), so the TMEM accumulator only
// ever holds WIN*8 samples; it is then drained into FP32 registers with round-to-nearest adds, and the
// split-K partials are summed in FP64 (syrk32_finalize_kernel).  Result: <= 1e-5 against the FP64 oracle.
//
// Operands: the gradient rows as they lie in HBM ([P, Ns], parameter fastest).  kind::tf32 reads K-major
// operands from shared memory (the MN-major no-swizzle form returned zeros in a probe, see
// profiles/probe/), so the transpose happens in the staging stores: 8 warps load 16-byte pieces along the
// parameter axis (coalesced), split hi / lo, de-interleave (re, im) into separate component planes -- all-zero
// planes are skipped, the real-parameter NDM has purely real and purely imaginary rows -- and scatter
// 4-byte words into the K-major core matrices (8 rows x 16 bytes; LBO = 144 keeps the scatter conflict-free).
// Thread 0 issues the MMAs (M = N = 128, K = 8); tcgen05.commit releases the stage through an mbarrier while
// the other threads already convert the next chunk.  2 CTAs per SM overlap one CTA's drain with the other's MMAs.
#include "nq_internal.cuh"
#include <algorithm>

namespace {

constexpr int TS = 128;          // tile rows (M = N = 128)
constexpr int KC = 8;            // samples per chunk = K of one kind::tf32 MMA
// K-major no-swizzle core matrices: 8 rows x 16 bytes (4 samples); the two K halves of a row group lie LBO
// apart, row groups SBO apart.  LBO = 144 (not 128) makes the transposing 4-byte stores bank-conflict free.
constexpr int LBO = 144;
constexpr int SBO = 2 * LBO;
constexpr int CB = (TS / 8) * SBO;   // bytes of one component plane (128 rows x 8 samples)
constexpr int NTHR = 256;
constexpr int WIN = 16;           // chunks (x8 samples) between accumulator drains

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // start address [0,14), LBO [16,30), SBO [32,46) (all >> 4), version = 1 at [46,48), no swizzle
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((LBO >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((SBO >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = 128
__host__ __device__ constexpr uint32_t make_idesc(bool neg_b) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((neg_b ? 1u : 0u) << 14) |
           ((uint32_t)(TS >> 3) << 17) | ((uint32_t)(TS >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

template <int NC>
struct Tile {
    static constexpr int BYTES = NC * CB;                  // one of hi / lo of one operand
    static constexpr int NLD = TS * NC * KC / 4 / NTHR;     // float4 loads per thread per operand chunk
};

// byte offset of (row m, sample s) inside a component plane
__device__ __forceinline__ int kmaj_off(int m, int s) { return (m >> 3) * SBO + (s >> 2) * LBO + (m & 7) * 16 + (s & 3) * 4; }

// One operand chunk [128 rows x 8 samples]: coalesced 16-byte loads along the parameter axis (as O lies in
// HBM), then the transpose into the K-major core matrices happens in the shared-memory stores.
template <int NC>
struct Chunk {
    float4 v[Tile<NC>::NLD];
    // lane -> (quad of 4 consecutive reals, sample): a warp covers 16 reals x 8 samples per load
    __device__ __forceinline__ void load(const float* __restrict__ Xr, int64_t ldr, int64_t PR, int64_t Ns, int64_t row0, int64_t s0) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int s = lane & 7, q = lane >> 3;
        const int64_t smp = s0 + s;
#pragma unroll
        for (int i = 0; i < Tile<NC>::NLD; i++) {
            const int64_t r = row0 + 16 * (warp + (NTHR / 32) * i) + 4 * q;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (smp < Ns) {
                if (r + 3 < PR) x = __ldg(reinterpret_cast<const float4*>(Xr + r + ldr * smp));
                else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int j = 0; j < 4; j++) if (r + j < PR) t[j] = Xr[r + j + ldr * smp];
                    x = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            v[i] = x;
        }
    }
    __device__ __forceinline__ void store(unsigned char* hi, unsigned char* lo, unsigned need) const {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int s = lane & 7, q = lane >> 3;
#pragma unroll
        for (int i = 0; i < Tile<NC>::NLD; i++) {
            const int r = 16 * (warp + (NTHR / 32) * i) + 4 * q;     // real-row index inside the tile
            const float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int comp = NC == 2 ? (j & 1) : 0;
                if (!((need >> comp) & 1u)) continue;
                const int m = NC == 2 ? ((r + j) >> 1) : (r + j);
                const float h = __uint_as_float(__float_as_uint(x[j]) & 0xffffe000u);
                const int off = comp * CB + kmaj_off(m, s);
                *reinterpret_cast<float*>(hi + off) = h;
                *reinterpret_cast<float*>(lo + off) = x[j] - h;
            }
        }
    }
};

// read 32 accumulator columns of this thread's TMEM lane and add them to acc (round-to-nearest FP32)
__device__ __forceinline__ void tmem_add32(uint32_t taddr, float* acc) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] += __uint_as_float(r[i]);
}

// 8 warps: all stage; thread 0 issues the MMAs; warp w owns TMEM lanes 32 (w % 4) .. +31 and the column
// half w / 4 of the accumulator.  The tensor core adds into the FP32 TMEM accumulator with truncation (a
// one-sided error that grows linearly with the number of accumulations), so every WIN chunks the
// accumulator is drained into registers (round-to-nearest adds) and restarted.
template <typename T, int NC>
__global__ void __launch_bounds__(NTHR, 2)
syrk_tf32_kernel(const float* __restrict__ Xr, int64_t ldr, int64_t P, int64_t Ns, int ntile, int nsplit, int mode,
                 const unsigned* __restrict__ tflags, float* __restrict__ Wk /* [nsplit][Ppad*Ppad] */) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int TB = Tile<NC>::BYTES;
    // stage layout: [A_hi | A_lo | B_hi | B_lo] x 2 stages, then barriers and the TMEM base
    unsigned char* stage[2] = {smem, smem + 4 * TB};
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8 * TB);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    int t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
    while (ti * (ti + 1) / 2 > t) ti--;
    const int tj = t - ti * (ti + 1) / 2;
    const bool diag = ti == tj;
    const int split = blockIdx.y;
    const int64_t nchunk_tot = (Ns + KC - 1) / KC;
    const int64_t cper = (nchunk_tot + nsplit - 1) / nsplit;
    const int64_t c_begin = split * cper, c_end = std::min<int64_t>(nchunk_tot, c_begin + cper);

    const unsigned fa = NC == 2 ? tflags[ti] : 1u, fb = NC == 2 ? tflags[tj] : 1u;
    unsigned act = 0;                       // bit c: K plane c of A is used
    for (int c = 0; c < NC; c++) {
        unsigned bc = mode == 0 ? c : 1 - c;
        if (((fa >> c) & 1u) && ((fb >> bc) & 1u)) act |= 1u << c;
    }
    unsigned needA = act, needB = 0;
    for (int c = 0; c < NC; c++) if ((act >> c) & 1u) needB |= 1u << (mode == 0 ? c : 1 - c);
    if (diag) { needA |= needB; needB = needA; }

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmine = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);

    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; i++) acc[i] = 0.f;

    const int64_t PR = P * NC, rowA = (int64_t)ti * TS * NC, rowB = (int64_t)tj * TS * NC;
    uint32_t phase[2] = {0, 0};
    bool pending[2] = {false, false};       // a commit on this stage has not been waited for yet
    int in_window = 0;                      // chunks accumulated in TMEM since the last drain
    if (act) {
        int64_t n_it = 0;
        for (int64_t c = c_begin; c < c_end; c++, n_it++) {
            const int buf = (int)(n_it & 1);
            unsigned char* Ahi = stage[buf];
            unsigned char* Alo = Ahi + TB;
            unsigned char* Bhi = diag ? Ahi : Ahi + 2 * TB;
            unsigned char* Blo = diag ? Alo : Ahi + 3 * TB;
            Chunk<NC> ca, cb;
            ca.load(Xr, ldr, PR, Ns, rowA, c * KC);
            if (!diag) cb.load(Xr, ldr, PR, Ns, rowB, c * KC);
            if (pending[buf]) { mbar_wait(&bars[buf], phase[buf]); phase[buf] ^= 1; pending[buf] = false; }   // stage is free
            ca.store(Ahi, Alo, needA);
            if (!diag) cb.store(Bhi, Blo, needB);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(Ahi), a_lo = smem_u32(Alo), b_hi = smem_u32(Bhi), b_lo = smem_u32(Blo);
                bool first = in_window == 0;
#pragma unroll
                for (int comp = 0; comp < NC; comp++) {
                    if (!((act >> comp) & 1u)) continue;
                    const int bcomp = mode == 0 ? comp : 1 - comp;
                    const uint32_t idesc = make_idesc(mode == 1 && comp == 1);
                    const uint32_t ao = (uint32_t)(comp * CB), bo = (uint32_t)(bcomp * CB);
                    // small cross terms first, the leading term last
                    mma_tf32(tmem, make_desc(a_lo + ao), make_desc(b_hi + bo), idesc, first ? 0u : 1u);
                    mma_tf32(tmem, make_desc(a_hi + ao), make_desc(b_lo + bo), idesc, 1u);
                    mma_tf32(tmem, make_desc(a_hi + ao), make_desc(b_hi + bo), idesc, 1u);
                    first = false;
                }
                umma_commit(&bars[buf]);       // arrives when every MMA issued so far has completed
            }
            pending[buf] = true;
            if (++in_window == WIN || c + 1 == c_end) {
                // drain: the last commit covers all earlier MMAs
                mbar_wait(&bars[buf], phase[buf]); phase[buf] ^= 1; pending[buf] = false;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_add32(tmine, acc);
                tmem_add32(tmine + 32, acc + 32);
                in_window = 0;                 // next MMA overwrites; ordered behind these loads by the next __syncthreads
            }
        }
    }

    const int64_t Ppad = (int64_t)ntile * TS;
    float* W = Wk + (size_t)split * Ppad * Ppad;
    const int64_t row = (int64_t)ti * TS + (warp & 3) * 32 + lane;
    const int64_t col0 = (int64_t)tj * TS + (warp >> 2) * 64;
#pragma unroll
    for (int i = 0; i < 64; i++) W[row + Ppad * (col0 + i)] = acc[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u) : "memory");
}

template <typename T, int NC>
__global__ void tile_activity32_kernel(const T* __restrict__ Xr, int64_t ldr, int64_t P, int64_t Ns, unsigned* __restrict__ flags) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= P * NC) return;
    int64_t per = (Ns + gridDim.y - 1) / gridDim.y;
    int64_t s0 = blockIdx.y * per, s1 = s0 + per < Ns ? s0 + per : Ns;
    bool nz = false;
    for (int64_t s = s0; s < s1; s++) nz |= Xr[r + ldr * s] != T(0);
    if (nz) atomicOr(&flags[(r / NC) / TS], 1u << (r % NC));
}

// S from the FP32 split-K partials: FP64 sum over the splits, scale, mirror
template <typename TS_>
__global__ void syrk32_finalize_kernel(const float* __restrict__ Wre, const float* __restrict__ Wim, int nsplit,
                                       int64_t Ppad, int64_t P, double scale, int out_complex, TS_* __restrict__ S) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t l = blockIdx.y;
    if (k >= P || l >= P) return;
    // element-wise lower triangle (also inside diagonal tiles: hi*lo and lo*hi are not bitwise symmetric), so
    // that S is exactly Hermitian
    bool lower = k >= l;
    int64_t r = lower ? k : l, c = lower ? l : k;
    double re = 0.0, im = 0.0;
    for (int s = 0; s < nsplit; s++) {
        re += (double)Wre[(size_t)s * Ppad * Ppad + r + Ppad * c];
        if (Wim) im += (double)Wim[(size_t)s * Ppad * Ppad + r + Ppad * c];
    }
    if (!lower) im = -im;
    if (k == l) im = 0.0;
    if (out_complex) { S[2 * (k + P * l)] = (TS_)(re * scale); S[2 * (k + P * l) + 1] = (TS_)(im * scale); }
    else S[k + P * l] = (TS_)(re * scale);
}

template <int NC>
int run_tf32(nq_ctx_t ctx, const float* X, int64_t ldr, int64_t P, int64_t Ns, int64_t Ns_total, bool out_complex, float* dS) {
    const int ntile = (int)((P + TS - 1) / TS);
    const int64_t Ppad = (int64_t)ntile * TS, ntri = (int64_t)ntile * (ntile + 1) / 2;
    const int64_t nchunk = (Ns + KC - 1) / KC;
    // split K for whole waves of 2 CTAs per SM
    int nsplit = 1;
    {
        double best = 1e30;
        const double slots = 2.0 * ctx->num_sms;
        for (int ns = 1; ns <= 64; ns++) {
            if (ns > 1 && nchunk / ns < 4 * WIN) break;
            double waves = (double)ntri * ns / slots;
            double cost = std::ceil(waves) / waves + 0.004 * ns;
            if (waves >= 1.0 || ns == 1) { if (cost < best) { best = cost; nsplit = ns; } }
        }
    }
    const size_t plane = (size_t)Ppad * Ppad * sizeof(float);
    while (nsplit > 1 && plane * nsplit * (out_complex ? 2 : 1) > ((size_t)3 << 30)) nsplit--;
    float* Wre = (float*)nq_scratch(ctx, SL_W0, plane * nsplit);
    float* Wim = out_complex ? (float*)nq_scratch(ctx, SL_W1, plane * nsplit) : nullptr;
    unsigned* flags = (unsigned*)nq_scratch(ctx, SL_W3, (size_t)ntile * sizeof(unsigned) + 16);
    if (!Wre || (out_complex && !Wim) || !flags) return NQ_ERR_ALLOC;
    if (NC == 2) {
        NQ_CUDA(ctx, cudaMemsetAsync(flags, 0, (size_t)ntile * sizeof(unsigned), ctx->stream));
        dim3 g((unsigned)((P * NC + 255) / 256), (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, Ns / 64)));
        NQ_LAUNCH(ctx, (tile_activity32_kernel<float, NC>), g, 256, 0, X, ldr, P, Ns, flags);
    }
    const size_t smem = (size_t)8 * Tile<NC>::BYTES + 64;
    auto kern = syrk_tf32_kernel<float, NC>;
    NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ntri, (unsigned)nsplit);
    NQ_LAUNCH(ctx, kern, grid, NTHR, smem, X, ldr, P, Ns, ntile, nsplit, 0, (const unsigned*)flags, Wre);
    if (out_complex) NQ_LAUNCH(ctx, kern, grid, NTHR, smem, X, ldr, P, Ns, ntile, nsplit, 1, (const unsigned*)flags, Wim);
    dim3 fg((unsigned)((P + 127) / 128), (unsigned)P);
    NQ_LAUNCH(ctx, syrk32_finalize_kernel<float>, fg, 128, 0, (const float*)Wre, (const float*)Wim, nsplit, Ppad, P,
              1.0 / (double)Ns_total, (int)out_complex, dS);
    return NQ_OK;
}

}  // namespace

// FP32-mode S assembly (O of dtype F32 or C64); S is written as float (real) or interleaved complex float.
int nq_syrk_tf32_device(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, int64_t Ns_total, bool o_complex,
                        bool out_complex, void* dS) {
    if (o_complex) return run_tf32<2>(ctx, (const float*)Oc, ldO * 2, P, Ns, Ns_total, out_complex, (float*)dS);
    return run_tf32<1>(ctx, (const float*)Oc, ldO, P, Ns, Ns_total, false, (float*)dS);
}
