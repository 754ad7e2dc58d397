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
// (This staging kernel is the fallback; the default since round 2 is the TMA-fed pipeline at the end of the file.)
#include "nq_internal.cuh"
#include <algorithm>
#include <cuda.h>
#include <cudaTypedefs.h>

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


// =====================================================================================================================
// TMA-fed variant (round 2).  The staging kernel above spends its time converting: per 8-sample chunk the 8 warps move
// 32 KB through registers into shared memory for ~200 tensor cycles (ncu: tensor pipe 23 % active, long-scoreboard 3.0 per
// issue).  Here the conversion happens ONCE, in a streaming pre-pass that writes the operands the way the tensor core wants
// them -- K-major (sample-contiguous), component planes de-interleaved, split into tf32 hi / lo -- and the SYRK itself is a
// warp-specialised TMA + tcgen05 pipeline:
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor 2D boxes of [128 rows x 32 samples] with the 128-byte swizzle
//                      straight into the UMMA layout, 3 stages of (A_hi, A_lo, B_hi, B_lo) = 64 KB
//   warp 1 (one lane)  MMA issuer: per stage 4 k-steps x (lo*hi, hi*lo, hi*hi); tcgen05.commit frees the stage
//   warps 4-7          epilogue: the FP32 accumulation of the tensor core truncates, so a TMEM accumulator only ever holds
//                      WINS stages (128 samples); two accumulators (2 x 128 columns) alternate, the finished one is drained
//                      with round-to-nearest adds into 128 registers per thread while the MMAs fill the other
// The (re, im) planes of complex rows enter as extra K steps (items = (chunk, component)).
// =====================================================================================================================
constexpr int KB_T = 32;                 // samples per stage: one 128-byte swizzle row of floats
constexpr int OPB = TS * KB_T * 4;       // bytes of one operand box
constexpr int NSTG = 3;
constexpr int WINS = 4;                  // stages per accumulation window

// out[hl][c][k][s] (s contiguous, k padded to Ppad, s padded to Nspad with zeros) from X[(k NC + c) + ldr s]
template <int NC>
__global__ void tf32_split_kernel(const float* __restrict__ Xr, int64_t ldr, int64_t P, int64_t Ns, int64_t Ppad, int64_t Nspad,
                                  float* __restrict__ out) {
    __shared__ float tile[32 * NC][33];
    const int tx = threadIdx.x, ty = threadIdx.y;                    // 32 x 8
    const int64_t k0 = (int64_t)blockIdx.x * 32, s0 = (int64_t)blockIdx.y * 32;
    for (int i = ty; i < 32; i += 8) {                               // sample s0 + i, real rows (k0 NC) + tx + 32 j
        const int64_t smp = s0 + i;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const int64_t r = k0 * NC + tx + 32 * j;
            tile[tx + 32 * j][i] = (smp < Ns && r < P * NC) ? Xr[r + ldr * smp] : 0.f;
        }
    }
    __syncthreads();
    const size_t plane = (size_t)Ppad * Nspad;
    for (int i = ty; i < 32; i += 8) {                               // row k0 + i, sample s0 + tx
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const float x = tile[i * NC + c][tx];
            const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
            const size_t at = (size_t)(k0 + i) * Nspad + (s0 + tx);
            out[(size_t)c * plane + at] = h;
            out[(size_t)(NC + c) * plane + at] = x - h;
        }
    }
}

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    // K-major, 128-byte swizzle: 8-row atoms of 1024 bytes (SBO), LBO unused (1), version 1, layout type 2 = SWIZZLE_128B
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

template <int NC>
__global__ void __launch_bounds__(256, 1)
syrk_tf32_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t Ppad, int64_t Nspad, int ntile, int nsplit, int mode,
                     const unsigned* __restrict__ tflags, float* __restrict__ Wk /* [nsplit][Ppad*Ppad] */) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + NSTG * 4 * OPB);
    uint64_t* empty = full + NSTG;
    uint64_t* accfull = empty + NSTG;
    uint64_t* accempty = accfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    int t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
    while (ti * (ti + 1) / 2 > t) ti--;
    const int tj = t - ti * (ti + 1) / 2;
    const int split = blockIdx.y;
    const int64_t nchunk_tot = Nspad / KB_T;
    const int64_t cper = (nchunk_tot + nsplit - 1) / nsplit;
    const int64_t c_begin = split * cper, c_end = std::min<int64_t>(nchunk_tot, c_begin + cper);

    const unsigned fa = NC == 2 ? tflags[ti] : 1u, fb = NC == 2 ? tflags[tj] : 1u;
    int comps[2], ncomp = 0;                // K planes of A that are used
    for (int c = 0; c < NC; c++) {
        const unsigned bc = mode == 0 ? c : 1 - c;
        if (((fa >> c) & 1u) && ((fb >> bc) & 1u)) comps[ncomp++] = c;
    }
    const bool same = ti == tj && mode == 0;          // B is A itself: one pair of boxes per item
    const int64_t nit = (c_end > c_begin ? c_end - c_begin : 0) * ncomp;
    const int64_t nwin = (nit + WINS - 1) / WINS;

    if (tid == 0) {
        for (int i = 0; i < NSTG; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&accfull[i], 1); mbar_init(&accempty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ---- TMA producer
        const int64_t rowsA = (int64_t)ti * TS, rowsB = (int64_t)tj * TS;
        for (int64_t it = 0; it < nit; it++) {
            const int s = (int)(it % NSTG);
            mbar_wait(&empty[s], (uint32_t)(((it / NSTG) & 1) ^ 1));
            const int comp = comps[it % ncomp], bcomp = mode == 0 ? comp : 1 - comp;
            const int32_t k0 = (int32_t)((c_begin + it / ncomp) * KB_T);
            unsigned char* st = base + (size_t)s * 4 * OPB;
            mbar_expect_tx(&full[s], same ? 2u * OPB : 4u * OPB);
            tma_load_2d(st, &tmap, k0, (int32_t)((0 * NC + comp) * Ppad + rowsA), &full[s]);              // A hi
            tma_load_2d(st + OPB, &tmap, k0, (int32_t)((1 * NC + comp) * Ppad + rowsA), &full[s]);        // A lo
            if (!same) {
                tma_load_2d(st + 2 * OPB, &tmap, k0, (int32_t)((0 * NC + bcomp) * Ppad + rowsB), &full[s]);   // B hi
                tma_load_2d(st + 3 * OPB, &tmap, k0, (int32_t)((1 * NC + bcomp) * Ppad + rowsB), &full[s]);   // B lo
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer
        for (int64_t it = 0; it < nit; it++) {
            const int s = (int)(it % NSTG);
            const int64_t w = it / WINS;
            const int b = (int)(w & 1);
            if (it % WINS == 0) {
                mbar_wait(&accempty[b], (uint32_t)(((w >> 1) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            mbar_wait(&full[s], (uint32_t)((it / NSTG) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int comp = comps[it % ncomp];
            const uint32_t idesc = make_idesc(mode == 1 && comp == 1);
            const uint32_t a_hi = smem_u32(base + (size_t)s * 4 * OPB), a_lo = a_hi + OPB;
            const uint32_t b_hi = same ? a_hi : a_hi + 2 * OPB, b_lo = same ? a_lo : a_hi + 3 * OPB;
            const uint32_t acc = tmem + (uint32_t)(b * 128);
#pragma unroll
            for (int kk = 0; kk < KB_T / KC; kk++) {
                const uint32_t ko = (uint32_t)(kk * KC * 4);
                const uint32_t first = (it % WINS == 0 && kk == 0) ? 0u : 1u;
                mma_tf32(acc, make_desc_sw128(a_lo + ko), make_desc_sw128(b_hi + ko), idesc, first);
                mma_tf32(acc, make_desc_sw128(a_hi + ko), make_desc_sw128(b_lo + ko), idesc, 1u);
                mma_tf32(acc, make_desc_sw128(a_hi + ko), make_desc_sw128(b_hi + ko), idesc, 1u);
            }
            umma_commit(&empty[s]);                                   // the stage is free once these MMAs have read it
            if (it % WINS == WINS - 1 || it + 1 == nit) umma_commit(&accfull[b]);
        }
    } else if (warp >= 4) {
        // ---- epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31 (= rows of the tile), all 128 columns
        float acc[128];
#pragma unroll
        for (int i = 0; i < 128; i++) acc[i] = 0.f;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int64_t w = 0; w < nwin; w++) {
            const int b = (int)(w & 1);
            mbar_wait(&accfull[b], (uint32_t)((w >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ta = tmem + lane_base + (uint32_t)(b * 128);
            tmem_add32(ta, acc);
            tmem_add32(ta + 32, acc + 32);
            tmem_add32(ta + 64, acc + 64);
            tmem_add32(ta + 96, acc + 96);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&accempty[b]);
        }
        float* W = Wk + (size_t)split * Ppad * Ppad;
        const int64_t row = (int64_t)ti * TS + (warp & 3) * 32 + lane;
        const int64_t col0 = (int64_t)tj * TS;
#pragma unroll
        for (int i = 0; i < 128; i++) W[row + Ppad * (col0 + i)] = acc[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256u) : "memory");
}

template <int NC>
int run_tf32_tma(nq_ctx_t ctx, const float* X, int64_t ldr, int64_t P, int64_t Ns, int64_t Ns_total, bool out_complex, float* dS,
                 bool* used) {
    *used = false;
    static PFN_cuTensorMapEncodeTiled encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        cudaGetLastError();
        return (PFN_cuTensorMapEncodeTiled)fn;
    }();
    if (!encode) return NQ_OK;
    const int ntile = (int)((P + TS - 1) / TS);
    const int64_t Ppad = (int64_t)ntile * TS, ntri = (int64_t)ntile * (ntile + 1) / 2;
    const int64_t Nspad = (Ns + KB_T - 1) / KB_T * KB_T;
    const int64_t rows_total = 2 * NC * Ppad;
    if (rows_total > 0x7fffffff || Nspad > 0x7fffffff) return NQ_OK;
    const size_t obytes = (size_t)rows_total * Nspad * sizeof(float);
    float* ops = (float*)nq_scratch(ctx, SL_W4, obytes);
    if (!ops) { cudaGetLastError(); return NQ_OK; }                  // not enough memory for the operand copy: staging kernel
    {
        dim3 g((unsigned)(Ppad / 32), (unsigned)(Nspad / 32)), b(32, 8);
        if (g.y > 65535) return NQ_OK;
        NQ_LAUNCH(ctx, tf32_split_kernel<NC>, g, b, 0, X, ldr, P, Ns, Ppad, Nspad, ops);
    }
    CUtensorMap tmap;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)Nspad, (cuuint64_t)rows_total};
        const cuuint64_t gstr[1] = {(cuuint64_t)Nspad * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)KB_T, (cuuint32_t)TS};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ops, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return nq_fail(ctx, NQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    const int64_t nchunk = Nspad / KB_T;
    int nsplit = 1;
    {
        double best = 1e30;
        const double slots = (double)ctx->num_sms;
        for (int ns = 1; ns <= 64; ns++) {
            if (ns > 1 && nchunk / ns < 4 * WINS) break;
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
        if (ctx->hint_P == P && (int)ctx->hint_tile_flags.size() == ntile) {
            NQ_CUDA(ctx, cudaMemcpyAsync(flags, ctx->hint_tile_flags.data(), (size_t)ntile * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
        } else {
            NQ_CUDA(ctx, cudaMemsetAsync(flags, 0, (size_t)ntile * sizeof(unsigned), ctx->stream));
            dim3 g((unsigned)((P * NC + 255) / 256), (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, Ns / 64)));
            NQ_LAUNCH(ctx, (tile_activity32_kernel<float, NC>), g, 256, 0, X, ldr, P, Ns, flags);
        }
    }
    const size_t smem = (size_t)NSTG * 4 * OPB + 1024 + 256;
    auto kern = syrk_tf32_tma_kernel<NC>;
    NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ntri, (unsigned)nsplit);
    NQ_LAUNCH(ctx, kern, grid, 256, smem, tmap, Ppad, Nspad, ntile, nsplit, 0, (const unsigned*)flags, Wre);
    if (out_complex) NQ_LAUNCH(ctx, kern, grid, 256, smem, tmap, Ppad, Nspad, ntile, nsplit, 1, (const unsigned*)flags, Wim);
    dim3 fg((unsigned)((P + 127) / 128), (unsigned)P);
    NQ_LAUNCH(ctx, syrk32_finalize_kernel<float>, fg, 128, 0, (const float*)Wre, (const float*)Wim, nsplit, Ppad, P,
              1.0 / (double)Ns_total, (int)out_complex, dS);
    *used = true;
    return NQ_OK;
}

}  // namespace

// FP32-mode S assembly (O of dtype F32 or C64); S is written as float (real) or interleaved complex float.
int nq_syrk_tf32_device(nq_ctx_t ctx, const void* Oc, int64_t ldO, int64_t P, int64_t Ns, int64_t Ns_total, bool o_complex,
                        bool out_complex, void* dS) {
    // default: the TMA-fed pipeline (operands pre-split into K-major planes); NQ_SR_TF32=stage forces the staging kernel,
    // which is also the fallback when the operand copy (2 x the size of O) cannot be allocated
    static const int want_tma = [] { const char* e = getenv("NQ_SR_TF32"); return e ? (!strcmp(e, "stage") ? 0 : 1) : 1; }();
    if (want_tma && P <= 65535) {
        bool used = false;
        int st = o_complex ? run_tf32_tma<2>(ctx, (const float*)Oc, ldO * 2, P, Ns, Ns_total, out_complex, (float*)dS, &used)
                           : run_tf32_tma<1>(ctx, (const float*)Oc, ldO, P, Ns, Ns_total, false, (float*)dS, &used);
        if (st != NQ_OK || used) return st;
    }
    if (o_complex) return run_tf32<2>(ctx, (const float*)Oc, ldO * 2, P, Ns, Ns_total, out_complex, (float*)dS);
    return run_tf32<1>(ctx, (const float*)Oc, ldO, P, Ns, Ns_total, false, (float*)dS);
}
