// Probe of tcgen05.mma kind::tf32 shared-memory descriptor semantics: one MMA (M=128, N=128, K=8) on
// host-provided shared-memory images.  Used once to pin the no-swizzle MN-major layout; not product code.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* Aimg, const float* Bimg, int nfloats, uint64_t desc_hi_a, uint64_t desc_hi_b, uint32_t idesc,
             int nmma, uint32_t a_step, uint32_t b_step, float* D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* As = reinterpret_cast<float*>(smem);
    float* Bs = As + nfloats;
    uint64_t* bar = reinterpret_cast<uint64_t*>(Bs + nfloats);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < nfloats; i += blockDim.x) { As[i] = Aimg[i]; Bs[i] = Bimg[i]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (tid == 0) {
        for (int i = 0; i < nmma; i++) {
            uint64_t da = desc_hi_a | (uint64_t)(((smem_u32(As) + i * a_step) >> 4) & 0x3FFF);
            uint64_t db = desc_hi_b | (uint64_t)(((smem_u32(Bs) + i * b_step) >> 4) & 0x3FFF);
            uint32_t acc = i ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n"
        :: "r"(smem_u32(bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tid;
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
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
        for (int i = 0; i < 32; i++) D[row * 128 + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u) : "memory");
}

extern "C" int umma_probe(const float* Aimg, const float* Bimg, int nfloats, uint64_t desc_hi_a, uint64_t desc_hi_b,
                          uint32_t idesc, int nmma, uint32_t a_step, uint32_t b_step, float* D) {
    float *dA, *dB, *dD;
    cudaMalloc(&dA, nfloats * 4); cudaMalloc(&dB, nfloats * 4); cudaMalloc(&dD, 128 * 128 * 4);
    cudaMemcpy(dA, Aimg, nfloats * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bimg, nfloats * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 128 * 128 * 4);
    size_t smem = (size_t)nfloats * 8 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(dA, dB, nfloats, desc_hi_a, desc_hi_b, idesc, nmma, a_step, b_step, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "probe: %s\n", cudaGetErrorString(e)); return (int)e; }
    cudaMemcpy(D, dD, 128 * 128 * 4, cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return 0;
}
