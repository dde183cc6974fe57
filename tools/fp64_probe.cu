// FP64 pipe probe for B200 (sm_100a): measures DFMA, DMMA.8x8x4 (mma.sync m8n8k4 / m16n8k8 f64)
// issue throughput and cuBLAS DGEMM at square and tall-skinny shapes.  Used to pick the roofline
// denominator for the ao*D / ao^T*M kernels (DESIGN.md "FP64 peak").  Not part of the product path.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters) {
    double a[16];
    double x = 1.0 + threadIdx.x * 1e-9, y = 0.999999;
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = i * 0.5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma884_kernel(double* out, int iters) {
    double c[NACC][2];
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma1688_kernel(double* out, int iters) {
    double c[NACC][4];
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = 0.3, a2 = 0.7, a3 = 0.2, b0 = 0.5, b1 = 0.25;
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; c[i][2] = 0; c[i][3] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        int blocks_per_sm = (warps * 32) / threads;
        int grid = sms * blocks_per_sm;
        float ms = time_ms([&] { dfma_kernel<<<grid, threads>>>(out, iters); });
        double flops = 2.0 * 16 * iters * (double)grid * threads;
        printf("DFMA      warps/SM %2d : %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
    }
    for (int warps : {4, 8, 16}) {
        int threads = warps * 32, grid = sms;
        float ms = time_ms([&] { dmma884_kernel<8><<<grid, threads>>>(out, iters); });
        double flops = 2.0 * 256 * 8 * iters * (double)grid * warps;
        printf("DMMA884   warps/SM %2d acc 8 : %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
        ms = time_ms([&] { dmma884_kernel<2><<<grid, threads>>>(out, iters); });
        flops = 2.0 * 256 * 2 * iters * (double)grid * warps;
        printf("DMMA884   warps/SM %2d acc 2 : %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
        ms = time_ms([&] { dmma1688_kernel<4><<<grid, threads>>>(out, iters); });
        flops = 2.0 * 1024 * 4 * iters * (double)grid * warps;
        printf("DMMA1688  warps/SM %2d acc 4 : %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
    }
    // cuBLAS DGEMM
    cublasHandle_t h; cublasCreate(&h);
    struct Shape { int m, n, k; const char* what; };
    std::vector<Shape> shapes = {
        {8192, 8192, 8192, "square 8192^3"},
        {800, 262144, 400, "fwd  C[N=262144,2n=800] = ao[N,400] D[400,800] (col-major m=800)"},
        {800, 400, 262144, "bwd  C[400,800] = ao^T[400,N] M[N,800], K=262144"},
        {528, 262144, 264, "fwd benzene n=264"},
        {528, 264, 262144, "bwd benzene n=264 K=262144"},
    };
    for (auto& s : shapes) {
        double *A, *B, *C;
        size_t sa = (size_t)s.m * s.k, sb = (size_t)s.k * s.n, sc = (size_t)s.m * s.n;
        CK(cudaMalloc(&A, sa * 8)); CK(cudaMalloc(&B, sb * 8)); CK(cudaMalloc(&C, sc * 8));
        CK(cudaMemset(A, 0, sa * 8)); CK(cudaMemset(B, 0, sb * 8));
        double one = 1.0, zero = 0.0;
        float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, s.m, s.n, s.k, &one, A, s.m, B, s.k, &zero, C, s.m); });
        printf("cuBLAS DGEMM %-70s : %.3f ms  %.2f TFLOP/s\n", s.what, ms, 2.0 * s.m * s.n * s.k / ms * 1e-9);
        // transposed-A variant for the bwd shape (ao^T)
        if (s.k > 100000) {
            ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, s.m, s.n, s.k, &one, A, s.m, B, s.n, &zero, C, s.m); });
            printf("cuBLAS DGEMM (NT) %-65s : %.3f ms  %.2f TFLOP/s\n", s.what, ms, 2.0 * s.m * s.n * s.k / ms * 1e-9);
        }
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    // sustained DGEMM for 3 s (power-capped figure)
    {
        int n = 8192; double *A, *B, *C; CK(cudaMalloc(&A, (size_t)n * n * 8)); CK(cudaMalloc(&B, (size_t)n * n * 8)); CK(cudaMalloc(&C, (size_t)n * n * 8));
        CK(cudaMemset(A, 0, (size_t)n * n * 8)); CK(cudaMemset(B, 0, (size_t)n * n * 8));
        double one = 1.0, zero = 0.0;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int reps = 60;
        for (int r = 0; r < reps; r++) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("cuBLAS DGEMM sustained 8192^3 x%d : %.2f TFLOP/s (%.1f ms total)\n", reps, 2.0 * n * n * (double)n * reps / ms * 1e-9, ms);
    }
    return 0;
}
