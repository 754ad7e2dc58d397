// Probe: FP64 mma.sync shapes on sm_100a -- fragment layout check for m16n8k16 and issue-rate comparison
// m8n8k4 vs m16n8k4 / k8 / k16.  Not product code.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// layout hypothesis: a_i = A[g + 8 (i & 1)][tig + 4 (i >> 1)], b_i = B[k = tig + 4 i][n = g],
// c0 = C[g][2 tig], c1 = C[g][2 tig + 1], c2 = C[g + 8][2 tig], c3 = C[g + 8][2 tig + 1]
__global__ void layout_kernel(const double* A /*16x16 row-major*/, const double* B /*16(k) x 8(n) row-major*/, double* C /*16x8*/, int K) {
    int lane = threadIdx.x, g = lane >> 2, tig = lane & 3;
    double a[8], b[4], c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 8; i++) a[i] = A[(g + 8 * (i & 1)) * 16 + tig + 4 * (i >> 1)];
    for (int i = 0; i < 4; i++) b[i] = B[(tig + 4 * i) * 8 + g];
    if (K == 16) mma16816(c, a, b);
    else if (K == 8) mma1688(c, a, b);
    else mma1684(c, a, b);
    C[g * 8 + 2 * tig] = c[0]; C[g * 8 + 2 * tig + 1] = c[1]; C[(g + 8) * 8 + 2 * tig] = c[2]; C[(g + 8) * 8 + 2 * tig + 1] = c[3];
}

template <int SHAPE>
__global__ void __launch_bounds__(256, 1) rate_kernel(double* out, int iters, double seed) {
    double acc[16][4];
    for (int i = 0; i < 16; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x;
    for (int i = 0; i < 4; i++) b[i] = seed - i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (SHAPE == 0) { mma884(acc[i][0], acc[i][1], a[0], b[0]); mma884(acc[i][2], acc[i][3], a[1], b[1]); }
            else if (SHAPE == 4) mma1684(acc[i], a, b);
            else if (SHAPE == 8) mma1688(acc[i], a, b);
            else mma16816(acc[i], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < 16; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
    if (s == 12345.678) out[0] = s;
}

template <int SHAPE>
void rate(const char* name, double fma_per_inst_pair) {
    double* d; cudaMalloc(&d, 8);
    const int iters = 20000;
    rate_kernel<SHAPE><<<148, 256>>>(d, 100, 1.0);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<SHAPE><<<148, 256>>>(d, iters, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)148 * 8 * iters * 16 * fma_per_inst_pair;
    printf("%-10s %.3f ms  %.2f TFLOP/s  (err %s)\n", name, ms, 2 * fma / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    double hA[256], hB[128], hC[128], *dA, *dB, *dC;
    srand(1);
    for (int i = 0; i < 256; i++) hA[i] = rand() % 7 - 3;
    for (int i = 0; i < 128; i++) hB[i] = rand() % 7 - 3;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    for (int K : {4, 8, 16}) {
        layout_kernel<<<1, 32>>>(dA, dB, dC, K);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
        double err = 0;
        for (int m = 0; m < 16; m++) for (int n = 0; n < 8; n++) {
            double r = 0; for (int k = 0; k < K; k++) r += hA[m * 16 + k] * hB[k * 8 + n];
            double d = hC[m * 8 + n] - r; err = d * d > err ? d * d : err;
        }
        printf("m16n8k%d layout hypothesis: max err^2 %.1f (%s)\n", K, err, cudaGetErrorString(e));
    }
    rate<0>("m8n8k4 x2", 2 * 256.0);
    rate<4>("m16n8k4", 512.0);
    rate<8>("m16n8k8", 1024.0);
    rate<16>("m16n8k16", 2048.0);
    return 0;
}
