// Write-bandwidth probe for the scheduling kernel's output pattern: records of REC bytes, one per (QP, stage), the array
// laid out [B][N][REC]; a warp owns 32 consecutive QPs.  pattern 0: linear stream over the whole array; pattern 1:
// stage-major (all QPs write stage 0, then stage 1, ... as the one-thread-per-QP roll-out does), each lane storing its own
// record with 256-bit stores; pattern 2: stage-major, but the warp's 32 records of a stage leave as consecutive 16-byte pieces.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/write_pattern tools/write_pattern.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__global__ void k_linear(double4 *dst, size_t n4, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(dst + i), "d"(v) : "memory");
  }
}
template <int REC8>   // record length in doubles
__global__ void k_stage_major(double *dst, int B, int N, double v, int delay) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int i = 0; i < N; ++i) {
    double *d = dst + ((size_t)b * N + i) * REC8;
#pragma unroll
    for (int e = 0; e < REC8; e += 4) asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(d + e), "d"(v) : "memory");
    // stand-in for a stage of arithmetic
    long long t0 = clock64();
    while (clock64() - t0 < delay) {}
  }
}
template <int REC8>
__global__ void k_stage_major_coalesced(double *dst, int B, int N, double v, int delay) {
  const int lane = threadIdx.x & 31;
  const int b0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31;
  if (b0 >= B) return;
  for (int i = 0; i < N; ++i) {
    // 32 records of REC8 doubles = 32 * REC8 / 2 pieces of 16 bytes; piece p belongs to QP p / (REC8 / 2)
    for (int p = lane; p < 32 * REC8 / 2; p += 32) {
      const int q = p / (REC8 / 2), e = p - q * (REC8 / 2);
      double *d = dst + ((size_t)(b0 + q) * N + i) * REC8 + 2 * e;
      asm volatile("st.global.v2.f64 [%0], {%1, %1};" ::"l"(d), "d"(v) : "memory");
    }
    long long t0 = clock64();
    while (clock64() - t0 < delay) {}
  }
}
int main(int argc, char **argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 65536, N = 8, REC8 = 36;
  const int delay = argc > 2 ? atoi(argv[2]) : 0;
  const size_t bytes = (size_t)B * N * REC8 * 8;
  double *dst; char *flush;
  cudaMalloc(&dst, bytes); cudaMalloc(&flush, 1ull << 30);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int pat = 0; pat < 3; ++pat) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaMemset(flush, rep, 1ull << 30);
      cudaEventRecord(a);
      if (pat == 0) k_linear<<<148 * 8, 256>>>((double4 *)dst, bytes / 32, 1.0);
      else if (pat == 1) k_stage_major<REC8><<<(B + 63) / 64, 64>>>(dst, B, N, 1.0, delay);
      else k_stage_major_coalesced<REC8><<<(B + 63) / 64, 64>>>(dst, B, N, 1.0, delay);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("{\"pattern\": %d, \"B\": %d, \"delay_clk\": %d, \"MB\": %.1f, \"us\": %.1f, \"GBps\": %.0f}\n", pat, B, delay, bytes / 1e6, best * 1e3, bytes / (best * 1e-3) / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
