// Write-bandwidth probe, three output streams as the scheduling kernel has them: A [B][N][36], B [B][N][12], S [B][N][6] doubles.
// One thread per QP; the records of CH consecutive stages are written together (CH = 1: what the one-thread-per-QP roll-out
// does stage by stage; CH = 2, 4, 8: the same bytes buffered over CH stages first), a fixed delay per stage in between.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/write_pattern3 tools/write_pattern3.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ void st32(double *p, double v) { asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void st16(double *p, double v) { asm volatile("st.global.v2.f64 [%0], {%1, %1};" ::"l"(p), "d"(v) : "memory"); }
template <int CH>
__global__ void k3(double *A, double *Bm, double *S, int B, int N, int mask, int delay) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int i = 0; i < N; i += CH) {
    for (int s = 0; s < CH; ++s) { long long t0 = clock64(); while (clock64() - t0 < delay) {} }
    if (mask & 1) { double *d = A + ((size_t)b * N + i) * 36;
#pragma unroll
      for (int e = 0; e < 36 * CH; e += 4) st32(d + e, 1.0); }
    if (mask & 2) { double *d = Bm + ((size_t)b * N + i) * 12;
#pragma unroll
      for (int e = 0; e < 12 * CH; e += 4) st32(d + e, 1.0); }
    if (mask & 4) { double *d = S + ((size_t)b * N + i) * 6;
      if ((6 * CH) % 4 == 0) {
#pragma unroll
        for (int e = 0; e < 6 * CH; e += 4) st32(d + e, 1.0);
      } else {
#pragma unroll
        for (int e = 0; e < 6 * CH; e += 2) st16(d + e, 1.0);
      } }
  }
}
int main(int argc, char **argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 65536, N = 8;
  const int delay = argc > 2 ? atoi(argv[2]) : 0;
  double *A, *Bm, *S; char *flush;
  cudaMalloc(&A, (size_t)B * N * 36 * 8); cudaMalloc(&Bm, (size_t)B * N * 12 * 8); cudaMalloc(&S, (size_t)B * N * 6 * 8); cudaMalloc(&flush, 1ull << 30);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int masks[5] = {1, 3, 5, 7, 6};
  for (int mi = 0; mi < 5; ++mi) for (int ch = 1; ch <= 8; ch *= 2) {
    const int mask = masks[mi];
    const size_t bytes = (size_t)B * N * 8 * ((mask & 1 ? 36 : 0) + (mask & 2 ? 12 : 0) + (mask & 4 ? 6 : 0));
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaMemset(flush, rep, 1ull << 30);
      cudaEventRecord(a);
      const int g = (B + 63) / 64;
      if (ch == 1) k3<1><<<g, 64>>>(A, Bm, S, B, N, mask, delay);
      else if (ch == 2) k3<2><<<g, 64>>>(A, Bm, S, B, N, mask, delay);
      else if (ch == 4) k3<4><<<g, 64>>>(A, Bm, S, B, N, mask, delay);
      else k3<8><<<g, 64>>>(A, Bm, S, B, N, mask, delay);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("{\"streams\": \"%s%s%s\", \"stages_per_write\": %d, \"B\": %d, \"delay_clk\": %d, \"MB\": %.1f, \"us\": %.1f, \"GBps\": %.0f}\n", mask & 1 ? "A" : "", mask & 2 ? "B" : "",
           mask & 4 ? "S" : "", ch, B, delay, bytes / 1e6, best * 1e3, bytes / (best * 1e-3) / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
