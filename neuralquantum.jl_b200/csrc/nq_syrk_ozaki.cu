// nq_syrk_ozaki.cu -- K7, FP64 mode on the INTEGER tensor cores: S = Re(Oc Oc^H) by the Ozaki scheme.
//
// The FP64 tensor instruction (DMMA, 37 TFLOP/s on B200) is the ceiling of syrk_dmma2_kernel (0.78-0.87 of its issue
// rate).  tcgen05 has no FP64 kind, but kind::i8 multiplies int8 operands EXACTLY into int32 accumulators at ~60x the DMMA
// rate, and a double is a short sum of small integers:
//     x_ks = 2^(e_k + 1) * sum_{p=1..7} q^(p)_ks 2^(-7p),     q^(p) in [-64, 64]   (e_k: exponent of max_s |x_ks|)
//     (O O^T)_kl = 2^(e_k + e_l + 2) * sum_{d=2..9} 2^(-7d) sum_{p+q=d} (Q^(p) Q^(q)T)_kl          (34 integer products)
// Terms with p + q > 9 are dropped: relative to the row scales that is < 2^-55 per sample (keeping d <= 8 only, 2^-48, left
// an element-wise excess of 8e-14 on the cfg3 parity test).  The integer sums are exact (|q q'| <= 4096, int32 holds 65 536 samples of 7 products), so the only rounding
// is the final conversion to double -- the result is at least as accurate as the DMMA path.
//
// Pipeline (same skeleton as the TMA-fed tf32 kernel, nq_syrk_tf32.cu):
//   pre-pass 1  row maxima -> exponents                                   (one read of O)
//   pre-pass 2  scale, peel 7 signed 7-bit digits, transpose to K-major    (one read of O, 7 bytes per element written)
//   main        tiles of 128 x 64: warp 0 = TMA producer (boxes of [rows x 64 samples], 64-byte swizzle, 2 stages of 7 A
//               + 7 B slices = 84 KB), warp 1 = MMA issuer (per stage 2 k-steps x 34 products, M = 128, N = 64, K = 32, one
//               TMEM accumulator of 64 columns per diagonal d = p + q: 8 x 64 = all 512 columns), warps 4-7 = epilogue (int32 ->
//               double, 2^(-7d) and the row / column scales, split-K partials in double)
// The (re, im) planes of complex rows enter as extra K items and share one exponent per parameter row; the imaginary part of a
// complex S is a second launch on the planes (re, im) and (-im, re) (the pre-pass also writes the negated imaginary digits).
#include "nq_internal.cuh"
#include <algorithm>
#include <cuda.h>
#include <cudaTypedefs.h>

namespace {

constexpr int TM = 128, TN = 64;         // tile
constexpr int NSL = 7;                   // slices (digits of 7 bits)
constexpr int DMAX = NSL + 2;            // diagonals d = p + q kept: 2 .. 9 -> 8 accumulators x 64 columns = all 512 TMEM columns
constexpr int KBYTES = 64;               // samples per item = bytes per smem row (64-byte swizzle): two K = 32 steps of kind::i8
constexpr int ABOX = TM * KBYTES, BBOX = TN * KBYTES;
constexpr int STAGE = NSL * (ABOX + BBOX);            // 86 016 bytes
constexpr int NSTG = 2;                  // (32-byte rows with 4 stages measured 11 % slower: 6.5 vs 5.8 ms on cfg4)
constexpr int WIN_ITEMS = 1024;          // items per int32 accumulation window: 7 products x 65 536 samples x 4096 < 2^31

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, swizzle = row width: 8-row atoms of 8 * KBYTES bytes (SBO), LBO unused (1), version 1, layout type 6 = SWIZZLE_32B
// (4 = SWIZZLE_64B, 2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    constexpr uint64_t LT = KBYTES == 32 ? 6 : (KBYTES == 64 ? 4 : 2);
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((8 * KBYTES) >> 4) << 32) | ((uint64_t)1 << 46) |
           (LT << 61);
}
// kind::i8: D = int32 (2 at [4,6)), A, B = signed 8 bit (1 at [7,10), [10,13)), K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t IDESC_I8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
__device__ __forceinline__ void mma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(IDESC_I8), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
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
}

// ---- pre-pass 1: mx[k] = max over samples and components of |x| (bit patterns of non-negative doubles order like integers)
template <int NC>
__global__ void oz_rowmax_kernel(const double* __restrict__ Xr, int64_t ldr, int64_t P, int64_t Ns, unsigned long long* __restrict__ mx) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= P * NC) return;
    const int64_t per = (Ns + gridDim.y - 1) / gridDim.y;
    const int64_t s0 = blockIdx.y * per, s1 = s0 + per < Ns ? s0 + per : Ns;
    double m = 0.0;
    for (int64_t s = s0; s < s1; s++) m = fmax(m, fabs(Xr[r + ldr * s]));
    atomicMax(&mx[r / NC], (unsigned long long)__double_as_longlong(m));
}

// which component planes of each 128-row tile are not identically zero (same contract as tile_activity_kernel)
template <int NC>
__global__ void oz_activity_kernel(const double* __restrict__ Xr, int64_t ldr, int64_t P, int64_t Ns, unsigned* __restrict__ flags) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= P * NC) return;
    const int64_t per = (Ns + gridDim.y - 1) / gridDim.y;
    const int64_t s0 = blockIdx.y * per, s1 = s0 + per < Ns ? s0 + per : Ns;
    bool nz = false;
    for (int64_t s = s0; s < s1; s++) nz |= Xr[r + ldr * s] != 0.0;
    if (nz) atomicOr(&flags[(r / NC) / TM], 1u << (r % NC));
}

// ---- pre-pass 2: digits.  out[p][c][k][s] (int8, s contiguous; k padded to Ppad, s to Nspad with zeros), ex[k] = e_k + 1
template <int NC, bool SHIFT>
__global__ void __launch_bounds__(256) oz_split_kernel(const double* __restrict__ Xr, int64_t ldr, int64_t P, int64_t Ns, int64_t Ppad,
                                                       int64_t Nspad, const unsigned long long* __restrict__ mx,
                                                       const double* __restrict__ shift /* (re, im) per row or NULL */,
                                                       signed char* __restrict__ out, int* __restrict__ ex, int neg_plane,
                                                       const unsigned* __restrict__ tflags /* NC = 2: live planes per 128 rows */) {
    extern __shared__ signed char tile[];                 // [NSL][32 NC][128 + 4]
    constexpr int ROWS = 32 * NC, LD = 128 + 4;
    const int tid = threadIdx.x;
    // consecutive CTAs sweep the SAMPLE axis of the same rows: their 128-byte output lines are neighbours in memory (with the
    // row axis fastest every line went to a different DRAM page: 1.74 ms for 4.3 GB)
    const int64_t k0 = (int64_t)blockIdx.y * 32, s0 = (int64_t)blockIdx.x * 128;
    const int rr = tid % ROWS, sl0 = tid / ROWS;           // real row in the tile, first sample lane
    constexpr int SLN = 256 / ROWS;                        // sample lanes
    const int64_t k = k0 + rr / NC;
    int e = 0;
    // deferred centring: the values are x - shift; mx then bounds the UNcentred row, |shift| is added to the bound
    const double sh = (SHIFT && k < P) ? shift[2 * k + (NC == 2 ? rr % NC : 0)] : 0.0;
    if (k < P) {
        double m = __longlong_as_double((long long)mx[k]);
        if (SHIFT) m += fmax(fabs(shift[2 * k]), fabs(shift[2 * k + 1]));
        if (m > 0.0) { frexp(m, &e); }                     // m = f 2^e, f in [0.5, 1)
        e += 1;                                            // |x| 2^-e <= 0.5: every digit fits [-64, 64]
        if (rr % NC == 0 && blockIdx.x == 0 && sl0 == 0) ex[k] = e;
    } else if (rr % NC == 0 && blockIdx.x == 0 && sl0 == 0 && k < Ppad) {
        ex[k] = 0;
    }
    const double sc49 = scalbn(1.0, 7 * NSL - e);          // x 2^-e in (-1/2, 1/2), times 2^49
    const double msh49 = -sh * sc49;
    // a task = 4 consecutive samples of one row: four loads (each coalesced over the 32 rows of the warp), 4 x 7 digits, packed
    // into one 32-bit word per slice (byte stores cost four times the shared-memory instructions)
    for (int j0 = sl0; j0 < 32; j0 += 2 * SLN) {            // two tasks (8 loads) in flight per thread
        double yv[2][4];
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const int j = j0 + u * SLN;
                const int64_t smp = s0 + 4 * j + w;
                yv[u][w] = (k < P && j < 32 && smp < Ns) ? Xr[(k0 * NC + rr) + ldr * smp] : sh;      // raw value; sh -> digit 0
            }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int j = j0 + u * SLN;
            if (j >= 32) continue;
            unsigned word[NSL];
#pragma unroll
            for (int p = 0; p < NSL; p++) word[p] = 0u;
#pragma unroll
            for (int w = 0; w < 4; w++) {
                // y 2^49 as a 64-bit integer (|y| < 1/2: 48 bits), then 7 balanced base-128 digits, least significant first
                // (x - sh) 2^(49-e): the power of two commutes with the rounding of x - sh, so one fma gives the bits of
                // (x - sh) * sc49 and the eight loads of the task pair stay in flight (a subtraction after each load made ptxas
                // recycle the load registers: 1.16 -> 1.81 ms)
                long long v = __double2ll_rn(SHIFT ? fma(yv[u][w], sc49, msh49) : yv[u][w] * sc49);
#pragma unroll
                for (int p = NSL - 1; p >= 1; p--) {
                    const int q = (int)((v + 64) & 127) - 64;
                    v = (v - q) >> 7;
                    word[p] |= ((unsigned)q & 0xffu) << (8 * w);
                }
                word[0] |= ((unsigned)(int)v & 0xffu) << (8 * w);      // top digit, |v| <= 64
            }
#pragma unroll
            for (int p = 0; p < NSL; p++) *reinterpret_cast<unsigned*>(&tile[(p * ROWS + rr) * LD + 4 * j]) = word[p];
        }
    }
    __syncthreads();
    // write: one (slice, component, row) line of 128 bytes per warp instruction
    const int warp = tid >> 5, lane = tid & 31;
    // planes per slice: the NC components, then (neg_plane) the NEGATED imaginary digits -- the imaginary part of S needs
    // A_re B_im^T - A_im B_re^T and the integer MMA has no negate flag
    const size_t plane = (size_t)Ppad * Nspad;
    const int NP = NC + neg_plane;
    // a component plane that is zero over the whole 128-row tile is never loaded by the product kernel (same flags): its
    // digits are not written (NDM gradients: every column is purely real or purely imaginary -- half of the planes)
    const unsigned live = (NC == 2 && tflags) ? (tflags[k0 / 128] & 3u) : (NC == 2 ? 3u : 1u);
    const int ncl = NC == 2 ? __popc(live) : 1;            // live components of this tile: the 8 warps share their lines
    // shifts, not divisions (a runtime divisor doubled the kernel's instruction count: 1.2 -> 1.8 ms).  The order matters: the 8
    // warps write 8 consecutive lines of one slice per step; giving each warp its 8 rows back to back (strength-reduced
    // pointers, unrolled) was 35 % SLOWER -- the write-back order of L2 follows the store order
    const int lsh = ncl == 2 ? 6 : 5, csh = ncl == 2 ? 1 : 0;
    for (int line = warp; line < (NSL << lsh); line += 8) {
        const int p = line >> lsh, rem = line & ((1 << lsh) - 1), kk = rem >> csh;
        const int c = ncl == NC ? (rem & (NC - 1)) : (int)(live >> 1);   // one live component: 0 (flags 01) or 1 (flags 10)
        const int r2 = kk * NC + c;
        const int word = *reinterpret_cast<const int*>(&tile[(p * ROWS + r2) * LD + 4 * lane]);
        const size_t at = (size_t)(k0 + kk) * Nspad + s0 + 4 * lane;
        *reinterpret_cast<int*>(out + ((size_t)(p * NP + c)) * plane + at) = word;
        if (neg_plane && c == 1) *reinterpret_cast<int*>(out + ((size_t)(p * NP + 2)) * plane + at) = (int)__vneg4((unsigned)word);
    }
}

// ---- main kernel
template <int NC>
__global__ void __launch_bounds__(256, 1)
syrk_ozaki_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int64_t Ppad, int64_t Nspad,
                  int ntile, int nsplit, int mode, int NP, const unsigned* __restrict__ tflags, const int* __restrict__ ex,
                  double* __restrict__ Wk /* [nsplit][Ppad*Ppad] col-major */) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + NSTG * STAGE);
    uint64_t* empty = full + NSTG;
    uint64_t* accfull = empty + NSTG;
    uint64_t* accempty = accfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // tile (ti, tj): rows [128 ti, +128), columns [64 tj, +64), tj <= 2 ti + 1; blockIdx.x enumerates them row by row
    int t = blockIdx.x;
    int ti = (int)((sqrt(4.0 * t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) <= t) ti++;
    while (ti * (ti + 1) > t) ti--;
    const int tj = t - ti * (ti + 1);
    const int split = blockIdx.y;
    const int64_t nchunk_tot = Nspad / KBYTES;
    const int64_t cper = (nchunk_tot + nsplit - 1) / nsplit;
    const int64_t c_begin = split * cper, c_end = std::min<int64_t>(nchunk_tot, c_begin + cper);

    const unsigned fa = NC == 2 ? tflags[ti] : 1u, fb = NC == 2 ? tflags[tj >> 1] : 1u;
    // mode 0 (real part): sum_c A_c B_c^T.  mode 1 (imaginary part): A_re B_im^T + (-A_im) B_re^T, planes (0, 1) and (2, 0)
    int comps[2], ncomp = 0;
    for (int c = 0; c < NC; c++) {
        const unsigned bc = mode == 0 ? c : 1 - c;
        if (((fa >> c) & 1u) && ((fb >> bc) & 1u)) comps[ncomp++] = c;
    }
    const bool same = mode == 0 && (tj >> 1) == ti;     // the B rows are a half of the A rows: no B boxes
    const int64_t nit = (c_end > c_begin ? c_end - c_begin : 0) * ncomp;
    const int64_t nwin = (nit + WIN_ITEMS - 1) / WIN_ITEMS;

    if (tid == 0) {
        for (int i = 0; i < NSTG; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(accfull, 1); mbar_init(accempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ---- TMA producer
        for (int64_t it = 0; it < nit; it++) {
            const int s = (int)(it % NSTG);
            mbar_wait(&empty[s], (uint32_t)(((it / NSTG) & 1) ^ 1));
            const int comp = comps[it % ncomp];
            const int pa = mode == 0 ? comp : (comp == 0 ? 0 : 2), pb = mode == 0 ? comp : (comp == 0 ? 1 : 0);
            const int32_t k0 = (int32_t)((c_begin + it / ncomp) * KBYTES);
            const uint32_t st = smem_u32(base + (size_t)s * STAGE);
            mbar_expect_tx(&full[s], (uint32_t)(same ? NSL * ABOX : STAGE));
#pragma unroll
            for (int p = 0; p < NSL; p++)
                tma_load_2d(st + p * ABOX, &mapA, k0, (int32_t)((int64_t)(p * NP + pa) * Ppad + (int64_t)ti * TM), &full[s]);
            if (!same) {
#pragma unroll
                for (int p = 0; p < NSL; p++)
                    tma_load_2d(st + NSL * ABOX + p * BBOX, &mapB, k0, (int32_t)((int64_t)(p * NP + pb) * Ppad + (int64_t)tj * TN), &full[s]);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer: accumulator of diagonal d = p + q (1-based slices) at TMEM columns 64 (d - 2)
        for (int64_t it = 0; it < nit; it++) {
            const int s = (int)(it % NSTG);
            const int64_t w = it / WIN_ITEMS;
            const bool win_first = it % WIN_ITEMS == 0;
            if (win_first) {
                mbar_wait(accempty, (uint32_t)((w & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            mbar_wait(&full[s], (uint32_t)((it / NSTG) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem_u32(base + (size_t)s * STAGE);
            const uint32_t b0 = same ? a0 + (uint32_t)((tj & 1) * BBOX) : a0 + NSL * ABOX;
            const uint32_t bstep = same ? ABOX : BBOX;
#pragma unroll
            for (int kk = 0; kk < KBYTES / 32; kk++) {
#pragma unroll
                for (int d = 2; d <= DMAX; d++) {
#pragma unroll
                    for (int p = 1; p < d; p++) {
                        const int q = d - p;
                        if (p > NSL || q > NSL) continue;
                        const uint32_t accum = (win_first && kk == 0 && p == (d - NSL > 1 ? d - NSL : 1)) ? 0u : 1u;
                        mma_i8(tmem + (uint32_t)((d - 2) * TN), make_desc_sw64(a0 + (p - 1) * ABOX + kk * 32),
                               make_desc_sw64(b0 + (q - 1) * bstep + kk * 32), accum);
                    }
                }
            }
            umma_commit(&empty[s]);
            if (it % WIN_ITEMS == WIN_ITEMS - 1 || it + 1 == nit) umma_commit(accfull);
        }
    } else if (warp >= 4) {
        // ---- epilogue: thread = row of the tile (TMEM lane), 64 columns
        double acc[TN];
#pragma unroll
        for (int i = 0; i < TN; i++) acc[i] = 0.0;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int64_t w = 0; w < nwin; w++) {
            mbar_wait(accfull, (uint32_t)(w & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int d = 2; d <= DMAX; d++) {
                const double sc = scalbn(1.0, -7 * d);
#pragma unroll
                for (int h = 0; h < TN / 32; h++) {
                    uint32_t r[32];
                    tmem_ld32(tmem + lane_base + (uint32_t)((d - 2) * TN + 32 * h), r);
#pragma unroll
                    for (int i = 0; i < 32; i++) acc[32 * h + i] = fma((double)(int)r[i], sc, acc[32 * h + i]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty);
        }
        double* W = Wk + (size_t)split * Ppad * Ppad;
        const int64_t row = (int64_t)ti * TM + (warp & 3) * 32 + lane;
        const int64_t col0 = (int64_t)tj * TN;
        const int er = ex[row];
#pragma unroll
        for (int i = 0; i < TN; i++) W[row + Ppad * (col0 + i)] = nit > 0 ? scalbn(acc[i], er + ex[col0 + i]) : 0.0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}


// =====================================================================================================================
// v2: 128 x 128 tiles in TWO passes over K.  The 128 x 64 kernel above reads 6 KB of operands from shared memory per 34-cycle
// MMA (A 128 x 32 B + B 64 x 32 B): 180 B/clk against the 128 B/clk the SM delivers -- its tensor pipe saturates at ~52 %.
// With N = 128 an MMA reads 8 KB per 68 cycles.  Four accumulators of 128 columns fill the TMEM, so the diagonals are done
// in two groups: pass A = d 2..5 (10 products, slices 1..4 of both operands: 8 boxes = 64 KB per stage, 3 stages), pass B =
// d 6..9 (24 products, all 7 slices: 14 boxes = 112 KB per stage, 2 stages).  The operand traffic from L2 is the same as for
// two 128 x 64 tiles (176 vs 168 KB per 64 samples).  The epilogue adds each pass into the FP64 split-K partial in global
// memory, 32 columns at a time (a whole 128-column row in registers would need 256 of them).
// =====================================================================================================================
constexpr int T2 = 128;
constexpr int BOX2 = T2 * KBYTES;                    // 8 KB
constexpr int STAGE_A = 8 * BOX2, NSTG_A = 3;        // slices 1..4 of A and of B
constexpr int STAGE_B = 2 * NSL * BOX2, NSTG_B = 2;  // slices 1..7 of A and of B
constexpr int SMEM2 = (NSTG_A * STAGE_A > NSTG_B * STAGE_B ? NSTG_A * STAGE_A : NSTG_B * STAGE_B);
constexpr uint32_t IDESC_I8_2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T2 >> 3) << 17) | ((uint32_t)(T2 >> 4) << 24);
__device__ __forceinline__ void mma_i8_2(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(IDESC_I8_2), "r"(accumulate) : "memory");
}

template <int NC>
__global__ void __launch_bounds__(256, 1)
syrk_ozaki2_kernel(const __grid_constant__ CUtensorMap mapA, int64_t Ppad, int64_t Nspad, int ntile, int nsplit, int mode, int NP,
                   const unsigned* __restrict__ tflags, const int* __restrict__ ex, double* __restrict__ Wk) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + SMEM2);
    uint64_t* empty = full + 3;
    uint64_t* accfull = empty + 3;
    uint64_t* accempty = accfull + 1;
    uint64_t* passdone = accempty + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(passdone + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    int t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
    while (ti * (ti + 1) / 2 > t) ti--;
    const int tj = t - ti * (ti + 1) / 2;
    const int split = blockIdx.y;
    const int64_t nchunk_tot = Nspad / KBYTES;
    const int64_t cper = (nchunk_tot + nsplit - 1) / nsplit;
    const int64_t c_begin = split * cper, c_end = std::min<int64_t>(nchunk_tot, c_begin + cper);

    const unsigned fa = NC == 2 ? tflags[ti] : 1u, fb = NC == 2 ? tflags[tj] : 1u;
    int comps[2], ncomp = 0;
    for (int c = 0; c < NC; c++) {
        const unsigned bc = mode == 0 ? c : 1 - c;
        if (((fa >> c) & 1u) && ((fb >> bc) & 1u)) comps[ncomp++] = c;
    }
    const bool same = mode == 0 && ti == tj;            // B is A itself: no B boxes
    const int64_t nit = (c_end > c_begin ? c_end - c_begin : 0) * ncomp;      // items per pass
    const int64_t nwin = (nit + WIN_ITEMS - 1) / WIN_ITEMS;                  // windows per pass

    if (tid == 0) {
        for (int i = 0; i < 3; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(accfull, 1); mbar_init(accempty, 4); mbar_init(passdone, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ---- TMA producer
        uint32_t uses[3] = {0, 0, 0};
        for (int pass = 0; pass < 2; pass++) {
            const int nstg = pass == 0 ? NSTG_A : NSTG_B, nsl = pass == 0 ? 4 : NSL;
            const uint32_t stage_bytes = pass == 0 ? STAGE_A : STAGE_B;
            if (pass == 1) mbar_wait(passdone, 0);                    // every MMA of pass A has read its operands
            for (int64_t it = 0; it < nit; it++) {
                const int s = (int)(it % nstg);
                mbar_wait(&empty[s], (uses[s] & 1u) ^ 1u);
                uses[s]++;
                const int comp = comps[it % ncomp];
                const int pa = mode == 0 ? comp : (comp == 0 ? 0 : 2), pb = mode == 0 ? comp : (comp == 0 ? 1 : 0);
                const int32_t k0 = (int32_t)((c_begin + it / ncomp) * KBYTES);
                const uint32_t st = smem_u32(base) + (uint32_t)s * stage_bytes;
                mbar_expect_tx(&full[s], (uint32_t)((same ? 1 : 2) * nsl * BOX2));
                for (int p = 0; p < nsl; p++)
                    tma_load_2d(st + p * BOX2, &mapA, k0, (int32_t)((int64_t)(p * NP + pa) * Ppad + (int64_t)ti * T2), &full[s]);
                if (!same)
                    for (int p = 0; p < nsl; p++)
                        tma_load_2d(st + (nsl + p) * BOX2, &mapA, k0, (int32_t)((int64_t)(p * NP + pb) * Ppad + (int64_t)tj * T2), &full[s]);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer: accumulator of diagonal d at TMEM columns 128 (d - dlo)
        uint32_t uses[3] = {0, 0, 0};
        int64_t wglob = 0;
        for (int pass = 0; pass < 2; pass++) {
            const int nstg = pass == 0 ? NSTG_A : NSTG_B, nsl = pass == 0 ? 4 : NSL;
            const uint32_t stage_bytes = pass == 0 ? STAGE_A : STAGE_B;
            for (int64_t it = 0; it < nit; it++) {
                const int s = (int)(it % nstg);
                const bool win_first = it % WIN_ITEMS == 0;
                if (win_first) {
                    mbar_wait(accempty, (uint32_t)((wglob & 1) ^ 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(&full[s], uses[s] & 1u);
                uses[s]++;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(base) + (uint32_t)s * stage_bytes;
                const uint32_t b0 = same ? a0 : a0 + (uint32_t)nsl * BOX2;
                if (pass == 0) {
#pragma unroll
                    for (int kk = 0; kk < KBYTES / 32; kk++)
#pragma unroll
                        for (int d = 2; d <= 5; d++)
#pragma unroll
                            for (int p = 1; p < d; p++)
                                mma_i8_2(tmem + (uint32_t)((d - 2) * T2), make_desc_sw64(a0 + (p - 1) * BOX2 + kk * 32),
                                         make_desc_sw64(b0 + (d - p - 1) * BOX2 + kk * 32), (win_first && kk == 0 && p == 1) ? 0u : 1u);
                } else {
#pragma unroll
                    for (int kk = 0; kk < KBYTES / 32; kk++)
#pragma unroll
                        for (int d = 6; d <= DMAX; d++)
#pragma unroll
                            for (int p = 1; p < d; p++) {
                                if (p > NSL || d - p > NSL) continue;
                                mma_i8_2(tmem + (uint32_t)((d - 6) * T2), make_desc_sw64(a0 + (p - 1) * BOX2 + kk * 32),
                                         make_desc_sw64(b0 + (d - p - 1) * BOX2 + kk * 32),
                                         (win_first && kk == 0 && p == (d - NSL > 1 ? d - NSL : 1)) ? 0u : 1u);
                            }
                }
                umma_commit(&empty[s]);
                if (it % WIN_ITEMS == WIN_ITEMS - 1 || it + 1 == nit) { umma_commit(accfull); wglob++; }
            }
            if (pass == 0) umma_commit(passdone);
        }
    } else if (warp >= 4) {
        // ---- epilogue: thread = row of the tile; every window of every pass is added into the FP64 partial, 32 columns at a time
        double* W = Wk + (size_t)split * Ppad * Ppad;
        const int64_t row = (int64_t)ti * T2 + (warp & 3) * 32 + lane;
        const int64_t col0 = (int64_t)tj * T2;
        const int er = ex[row];
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        int64_t wglob = 0;
        bool first_write = true;
        if (nit == 0) {
            for (int i = 0; i < T2; i++) W[row + Ppad * (col0 + i)] = 0.0;
        }
        for (int pass = 0; pass < 2; pass++) {
            const int dlo = pass == 0 ? 2 : 6;
            for (int64_t w = 0; w < nwin; w++, wglob++) {
                mbar_wait(accfull, (uint32_t)(wglob & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int h = 0; h < T2 / 32; h++) {
                    double acc[32];
#pragma unroll
                    for (int i = 0; i < 32; i++) acc[i] = 0.0;
#pragma unroll
                    for (int dd = 0; dd < 4; dd++) {
                        const double sc = scalbn(1.0, -7 * (dlo + dd));
                        uint32_t r[32];
                        tmem_ld32(tmem + lane_base + (uint32_t)(dd * T2 + 32 * h), r);
#pragma unroll
                        for (int i = 0; i < 32; i++) acc[i] = fma((double)(int)r[i], sc, acc[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        double* dst = W + row + Ppad * (col0 + 32 * h + i);
                        const double v = scalbn(acc[i], er + ex[col0 + 32 * h + i]);
                        *dst = first_write ? v : *dst + v;
                    }
                }
                first_write = false;
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(accempty);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

PFN_cuTensorMapEncodeTiled get_encode() {
    static PFN_cuTensorMapEncodeTiled encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        cudaGetLastError();
        return (PFN_cuTensorMapEncodeTiled)fn;
    }();
    return encode;
}

template <int NC>
int run_ozaki(nq_ctx_t ctx, const double* X, int64_t ldr, int64_t P, int64_t Ns, int ntile, int nsplit, double* W, double* Wim,
              const unsigned long long* known_rowmax, const double* shift, bool* used) {
    *used = false;
    PFN_cuTensorMapEncodeTiled encode = get_encode();
    if (!encode) return NQ_OK;
    const int64_t Ppad = (int64_t)ntile * TM;
    const int64_t Nspad = (Ns + 127) / 128 * 128;
    const int neg_plane = (NC == 2 && Wim) ? 1 : 0, NP = NC + neg_plane;
    const int64_t rows_total = (int64_t)NSL * NP * Ppad;
    if (rows_total > 0x7fffffff || Nspad > 0x7fffffff || Ppad / 32 > 65535) return NQ_OK;
    const size_t obytes = (size_t)rows_total * Nspad;
    signed char* ops = (signed char*)nq_scratch(ctx, SL_W4, obytes);
    if (!ops) { cudaGetLastError(); return NQ_OK; }
    unsigned long long* mx = (unsigned long long*)nq_scratch(ctx, SL_W5, (size_t)Ppad * 8 + (size_t)Ppad * 4 + 64);
    if (!mx) return NQ_ERR_ALLOC;
    int* ex = (int*)(mx + Ppad);
    unsigned* flags = (unsigned*)nq_scratch(ctx, SL_W3, (size_t)ntile * sizeof(unsigned) + 16);
    if (!flags) return NQ_ERR_ALLOC;
    {
        dim3 g((unsigned)((P * NC + 255) / 256), (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, Ns / 64)));
        if (known_rowmax) {
            NQ_CUDA(ctx, cudaMemcpyAsync(mx, known_rowmax, (size_t)P * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            NQ_CUDA(ctx, cudaMemsetAsync(mx, 0, (size_t)Ppad * 8, ctx->stream));
            NQ_LAUNCH(ctx, oz_rowmax_kernel<NC>, g, 256, 0, X, ldr, P, Ns, mx);
        }
        if (NC == 2) {
            if (ctx->hint_P == P && (int)ctx->hint_tile_flags.size() == ntile) {
                NQ_CUDA(ctx, cudaMemcpyAsync(flags, ctx->hint_tile_flags.data(), (size_t)ntile * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
            } else if (shift) {
                // deferred centring without a structural hint: a plane that is zero in THIS shard may have a non-zero global
                // mean, so no plane is skipped (0x03 in every byte = both bits of every tile word)
                NQ_CUDA(ctx, cudaMemsetAsync(flags, 0x03, (size_t)ntile * sizeof(unsigned), ctx->stream));
            } else {
                NQ_CUDA(ctx, cudaMemsetAsync(flags, 0, (size_t)ntile * sizeof(unsigned), ctx->stream));
                NQ_LAUNCH(ctx, oz_activity_kernel<NC>, g, 256, 0, X, ldr, P, Ns, flags);
            }
        }
    }
    {
        dim3 g((unsigned)(Nspad / 128), (unsigned)(Ppad / 32));
        const size_t smem = (size_t)NSL * 32 * NC * (128 + 4);
        auto ks = shift ? oz_split_kernel<NC, true> : oz_split_kernel<NC, false>;
        static const bool skip_dead = [] { const char* e = getenv("NQ_OZ_SKIP"); return !(e && e[0] == '0'); }();
        NQ_CUDA(ctx, cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NQ_LAUNCH(ctx, ks, g, 256, smem, X, ldr, P, Ns, Ppad, Nspad, (const unsigned long long*)mx, shift, ops, ex, neg_plane,
                  skip_dead ? (const unsigned*)flags : (const unsigned*)nullptr);
    }
    CUtensorMap mapA, mapB;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)Nspad, (cuuint64_t)rows_total};
        const cuuint64_t gstr[1] = {(cuuint64_t)Nspad};
        const cuuint32_t boxA[2] = {(cuuint32_t)KBYTES, (cuuint32_t)TM}, boxB[2] = {(cuuint32_t)KBYTES, (cuuint32_t)TN};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r1 = encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ops, gdim, gstr, boxA, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             (KBYTES == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ops, gdim, gstr, boxB, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             (KBYTES == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return nq_fail(ctx, NQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
    }
    static const int want_v1 = [] { const char* e = getenv("NQ_OZAKI"); return (e && !strcmp(e, "v1")) ? 1 : 0; }();
    if (!want_v1) {
        CUtensorMap map2;
        const cuuint64_t gdim[2] = {(cuuint64_t)Nspad, (cuuint64_t)rows_total};
        const cuuint64_t gstr[1] = {(cuuint64_t)Nspad};
        const cuuint32_t box2[2] = {(cuuint32_t)KBYTES, (cuuint32_t)T2};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ops, gdim, gstr, box2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return nq_fail(ctx, NQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        const size_t smem2 = (size_t)SMEM2 + 1024 + 256;
        auto k2 = syrk_ozaki2_kernel<NC>;
        NQ_CUDA(ctx, cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        dim3 grid2((unsigned)((int64_t)ntile * (ntile + 1) / 2), (unsigned)nsplit);
        NQ_LAUNCH(ctx, k2, grid2, 256, smem2, map2, Ppad, Nspad, ntile, nsplit, 0, NP, (const unsigned*)flags, (const int*)ex, W);
        if (neg_plane) NQ_LAUNCH(ctx, k2, grid2, 256, smem2, map2, Ppad, Nspad, ntile, nsplit, 1, NP, (const unsigned*)flags, (const int*)ex, Wim);
        *used = true;
        return NQ_OK;
    }
    const size_t smem = (size_t)NSTG * STAGE + 1024 + 256;
    auto kern = syrk_ozaki_kernel<NC>;
    NQ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((int64_t)ntile * (ntile + 1)), (unsigned)nsplit);
    NQ_LAUNCH(ctx, kern, grid, 256, smem, mapA, mapB, Ppad, Nspad, ntile, nsplit, 0, NP, (const unsigned*)flags, (const int*)ex, W);
    if (neg_plane) NQ_LAUNCH(ctx, kern, grid, 256, smem, mapA, mapB, Ppad, Nspad, ntile, nsplit, 1, NP, (const unsigned*)flags, (const int*)ex, Wim);
    *used = true;
    return NQ_OK;
}

}  // namespace

// Split-K partials of conj(O O^H): the real part W and, for complex S (Wim != NULL, NC = 2), the imaginary part Wim --
// [nsplit][Ppad^2] doubles, column-major, lower 128-tiles and the full diagonal tiles -- from the real rows
// Xr [(k NC + c) + ldr s].  *used = false when the path is not available (no driver entry point, not enough memory for the
// digit planes): the caller runs the DMMA kernel instead.
int nq_syrk_ozaki_device(nq_ctx_t ctx, const double* Xr, int64_t ldr, int64_t P, int64_t Ns, int NC, int ntile, int nsplit, double* W,
                         double* Wim, const unsigned long long* known_rowmax, const double* shift, bool* used) {
    if (NC == 2) return run_ozaki<2>(ctx, Xr, ldr, P, Ns, ntile, nsplit, W, Wim, known_rowmax, shift, used);
    return run_ozaki<1>(ctx, Xr, ldr, P, Ns, ntile, nsplit, W, nullptr, known_rowmax, shift, used);
}
