// Measured FP64 peaks on this GPU ("absent from MEASURED_PEAKS.json -> measure a DGEMM/DFMA peak"):
//   * DFMA issue peak: 8 independent FMA chains per thread, 148 x 8 CTAs x 256 threads, CUDA events;
//   * cuBLAS DGEMM 8192^3.
// build + run (on the GPU box):  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_peak.cu -lcublas -o /tmp/fp64_peak && /tmp/fp64_peak
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdio>

__global__ void dfma_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, ctas = sms * 8, thr = 256, iters = 1 << 16;
  double *out;
  cudaMalloc(&out, sizeof(double) * ctas * thr);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dfma_kernel<<<ctas, thr>>>(out, 1024, 1.0000001, 1e-9);
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    dfma_kernel<<<ctas, thr>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double fmas = (double)ctas * thr * 8.0 * iters;
  const double dfma_per_s = fmas / (best * 1e-3);
  // DGEMM
  const int n = 8192;
  double *A, *B, *C;
  cudaMalloc(&A, sizeof(double) * n * n);
  cudaMalloc(&B, sizeof(double) * n * n);
  cudaMalloc(&C, sizeof(double) * n * n);
  cudaMemset(A, 0, sizeof(double) * n * n);
  cudaMemset(B, 0, sizeof(double) * n * n);
  cublasHandle_t h;
  cublasCreate(&h);
  const double one = 1.0, zero = 0.0;
  cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
  float bestg = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < bestg) bestg = ms;
  }
  const double gemm_tflops = 2.0 * n * (double)n * n / (bestg * 1e-3) / 1e12;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %d, \"dfma_lane_ops_per_s\": %.4e, \"dfma_tflops\": %.2f, "
         "\"dgemm_8192_tflops\": %.2f, \"how\": \"tools/fp64_peak.cu: 8 independent DFMA chains per thread, best of 5; "
         "cublasDgemm 8192^3, best of 3; CUDA events\"}\n",
         p.name, sms, p.clockRate / 1000, dfma_per_s, 2.0 * dfma_per_s / 1e12, gemm_tflops);
  return 0;
}
