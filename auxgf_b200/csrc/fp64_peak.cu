// fp64_peak: measures the FP64 roofline denominators on this GPU (MEASURED_PEAKS.json has none):
//   * cublasDgemm 8192^3, best of 10 (burst) and ~4 s back-to-back (sustained)
//   * register-resident DMMA.8x8x4 issue rate (tensor path), DFMA issue rate (vector path), and
//     both interleaved (do they share a pipe?)
// Prints one JSON object.  Tool, not part of the product path.
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dmma_rate(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int k = 0; k < NACC; ++k) { c[k][0] = 0.0; c[k][1] = 0.0; }
    double a = 1e-3 * threadIdx.x, b = 1e-3 * (threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NACC; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NACC; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_rate(double* out, int iters) {
    double c[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) c[k] = 1e-3 * k;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[k]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NACC; ++k) s += c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// interleaved: per iteration NMMA DMMAs (256 FMA each per warp) + NFMA DFMAs (32 FMA each per warp)
template <int NMMA, int NFMA>
__global__ void __launch_bounds__(256) mixed_rate(double* out, int iters) {
    double c[NMMA][2], f[NFMA];
#pragma unroll
    for (int k = 0; k < NMMA; ++k) { c[k][0] = 0.0; c[k][1] = 0.0; }
#pragma unroll
    for (int k = 0; k < NFMA; ++k) f[k] = 1e-3 * k;
    double a = 1e-3 * threadIdx.x, b = 1e-3 * (threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NMMA; ++k) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int j = 0; j < NFMA / NMMA; ++j)
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[k * (NFMA / NMMA) + j]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NMMA; ++k) s += c[k][0] + c[k][1];
#pragma unroll
    for (int k = 0; k < NFMA; ++k) s += f[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F launch, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 8192;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 8));
    const int iters = 20000;

    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    // ---- register-resident pipe rates: `wps` warps per SM sub-partition ----
    for (int warps_per_cta : {4, 8}) {
        for (int ctas_per_sm : {1, 2, 4}) {
            const int grid = sms * ctas_per_sm, block = warps_per_cta * 32;
            const double warps = (double)grid * warps_per_cta;
            double ms = time_ms([&] { dmma_rate<16><<<grid, block>>>(out, iters); }, 3);
            printf(", \"dmma_tflops_w%d_c%d\": %.2f", warps_per_cta, ctas_per_sm, warps * iters * 16.0 * 512.0 / ms * 1e-9);
        }
    }
    {
        const int grid = sms * 2, block = 256;
        const double warps = (double)grid * 8;
        double ms = time_ms([&] { dfma_rate<16><<<grid, block>>>(out, iters); }, 3);
        printf(", \"dfma_tflops\": %.2f", warps * iters * 16.0 * 64.0 / ms * 1e-9);
        ms = time_ms([&] { mixed_rate<8, 8><<<grid, block>>>(out, iters); }, 3);
        printf(", \"mixed_8dmma_8dfma_tflops\": %.2f", warps * iters * (8.0 * 512.0 + 8.0 * 64.0) / ms * 1e-9);
        ms = time_ms([&] { mixed_rate<8, 32><<<grid, block>>>(out, iters); }, 3);
        printf(", \"mixed_8dmma_32dfma_tflops\": %.2f", warps * iters * (8.0 * 512.0 + 32.0 * 64.0) / ms * 1e-9);
        ms = time_ms([&] { mixed_rate<8, 64><<<grid, block>>>(out, iters); }, 3);
        printf(", \"mixed_8dmma_64dfma_tflops\": %.2f", warps * iters * (8.0 * 512.0 + 64.0 * 64.0) / ms * 1e-9);
    }
    // ---- cuBLAS DGEMM ----
    {
        double *A, *B, *C;
        const size_t nn = (size_t)n * n;
        CK(cudaMalloc(&A, nn * 8)); CK(cudaMalloc(&B, nn * 8)); CK(cudaMalloc(&C, nn * 8));
        std::vector<double> h(nn);
        for (size_t i = 0; i < nn; ++i) h[i] = (double)((i * 2654435761u) % 1000) * 1e-3 - 0.5;
        CK(cudaMemcpy(A, h.data(), nn * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B, h.data(), nn * 8, cudaMemcpyHostToDevice));
        cublasHandle_t hd; cublasCreate(&hd);
        const double one = 1.0, zero = 0.0;
        auto gemm = [&] { cublasDgemm(hd, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n); };
        const double flops = 2.0 * n * (double)n * n;
        double ms = time_ms(gemm, 10);
        printf(", \"dgemm_n\": %d, \"dgemm_tflops_burst\": %.2f", n, flops / ms * 1e-9);
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        int reps = (int)(4000.0 / ms) + 1;
        CK(cudaEventRecord(a));
        for (int r = 0; r < reps; ++r) gemm();
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float tot; CK(cudaEventElapsedTime(&tot, a, b));
        printf(", \"dgemm_tflops_sustained\": %.2f, \"sustained_seconds\": %.2f", flops * reps / tot * 1e-9, tot * 1e-3);
    }
    printf("}\n");
    return 0;
}
