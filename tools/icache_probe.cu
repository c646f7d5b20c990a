// Micro-probe: issue rate of a straight-line loop body as a function of its size (instruction-cache reach) and of
// the number of warps per SM sub-partition.  Body = independent FFMA chains (16 B per SASS instruction).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/icache_probe tools/icache_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define F4 asm volatile("fma.rn.f32 %0,%0,%4,%5; fma.rn.f32 %1,%1,%4,%5; fma.rn.f32 %2,%2,%4,%5; fma.rn.f32 %3,%3,%4,%5;" : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3) : "f"(m), "f"(c));
#define F16 F4 F4 F4 F4
#define F64 F16 F16 F16 F16
#define F256 F64 F64 F64 F64
#define F1K F256 F256 F256 F256
template <int KINSTR>
__global__ void probe(float *out, int iters, float m, float c) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  for (int it = 0; it < iters; ++it) {
    if (KINSTR >= 1) { F256 }            // 256 instr = 4 KB
    if (KINSTR >= 2) { F256 }
    if (KINSTR >= 4) { F256 F256 }
    if (KINSTR >= 8) { F1K }
    if (KINSTR >= 16) { F1K F1K }
    if (KINSTR >= 32) { F1K F1K F1K F1K }
    if (KINSTR >= 64) { F1K F1K F1K F1K F1K F1K F1K F1K }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
template <int K>
void run(int warps_per_sm, float *d) {
  int ninstr = 256 * K;                       // FFMA per iteration
  int iters = (1 << 22) / ninstr;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<K><<<148, 32 * warps_per_sm>>>(d, 8, 1.0f, 0.0f);
  cudaEventRecord(e0);
  probe<K><<<148, 32 * warps_per_sm>>>(d, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cycles = ms * 1e-3 * clk * 1e3;
  double per_smsp_instr = (double)ninstr * iters * (warps_per_sm / 4.0 < 1 ? 1 : warps_per_sm / 4.0);
  printf("{\"bench\": \"icache_probe\", \"body_kb\": %d, \"warps_per_sm\": %d, \"ipc_per_smsp_at_max_clock\": %.3f}\n", K * 4, warps_per_sm,
         per_smsp_instr / cycles);
}
int main() {
  float *d; cudaMalloc(&d, 148 * 1024 * sizeof(float));
  int ws[] = {4, 8, 16};
  for (int w : ws) { run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); run<16>(w, d); run<32>(w, d); run<64>(w, d); }
  return 0;
}
