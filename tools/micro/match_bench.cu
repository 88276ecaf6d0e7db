// Microbenchmark: issue cost of MATCH.ANY / SHFL / REDUX relative to an integer op (tuning aid for the dedup design).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int OP>
__global__ void k(uint32_t *out, int iters) {
  uint32_t v = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (OP == 0) acc += __match_any_sync(0xffffffffu, v + u);
      if (OP == 1) acc += __shfl_xor_sync(0xffffffffu, v + u, 1 + u);
      if (OP == 2) acc += __reduce_max_sync(0xffffffffu, v + u);
      if (OP == 3) acc = acc * 2654435761u + v + u;
      if (OP == 4) acc += __ballot_sync(0xffffffffu, (v + u) & 1);
    }
    v += acc;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int OP>
float run(uint32_t *d, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<148 * 4, 256>>>(d, 10);
  cudaEventRecord(a);
  k<OP><<<148 * 4, 256>>>(d, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms;
}
int main() {
  uint32_t *d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  const int iters = 20000;
  const char *names[] = {"match_any", "shfl_xor", "reduce_max", "imad", "ballot"};
  float ms[5] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters)};
  for (int i = 0; i < 5; ++i) {
    // warp-instructions of the op per SM per cycle: 32 warps/SM * iters * 8 ops / (ms * 1.965e6 cycles)
    double per_cycle = 32.0 * iters * 8 / (ms[i] * 1.965e6);
    printf("%-10s %8.3f ms  -> %.3f warp-ops / cycle / SM (incl. the dependent add)\n", names[i], ms[i], per_cycle);
  }
  return 0;
}
