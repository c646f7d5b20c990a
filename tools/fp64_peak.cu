// fp64 FMA / DMMA / shuffle / smem micro-benchmarks: the roofline denominators
// for the ADMM kernel (MEASURED_PEAKS.json carries no fp64 figure).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// dependent-chain latency of DFMA: one warp, one chain
__global__ void dfma_latency(double* out, long long* cyc, int iters, double a, double b) {
  double acc = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b);
    acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
  if (acc == 12345.678) out[0] = acc;
}

__global__ void shfl_latency(double* out, long long* cyc, int iters) {
  double acc = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = __shfl_xor_sync(0xffffffffu, acc, 1 + (j & 3));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (acc == 12345.678) out[0] = acc;
}

__global__ void lds_latency(double* out, long long* cyc, int iters) {
  __shared__ double buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (double)((i * 37 + 11) & 1023);
  __syncthreads();
  double acc = threadIdx.x;
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc = buf[idx]; idx = (int)acc; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (acc == 12345.678) out[0] = acc;
}

// DMMA m8n8k4 throughput
template <int ILP>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters) {
  double c0[ILP], c1[ILP];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"smem_per_sm\": %zu, \"smem_optin\": %zu, \"regs_per_sm\": %d, \"clock_khz\": %d, \"l2\": %d}\n",
         p.name, p.multiProcessorCount, p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, p.regsPerMultiprocessor, clk_khz, p.l2CacheSize);
  double* out; long long* cyc; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, 64));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int sms = p.multiProcessorCount;
  // throughput: 8 CTAs x 256 thr per SM
  for (int rep = 0; rep < 2; ++rep) {
    int iters = 20000;
    const int ILP = 8;
    dim3 grid(sms * 8), block(256);
    dfma_kernel<ILP><<<grid, block>>>(out, 1000, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int t = 0; t < 5; ++t) {
      cudaEventRecord(e0);
      dfma_kernel<ILP><<<grid, block>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 2.0 * (double)grid.x * 256 * ILP * iters;
    printf("{\"bench\": \"dfma_throughput\", \"ms\": %.4f, \"tflops\": %.3f, \"fma_per_clk_per_sm_at_max\": %.2f}\n", best, flops / best * 1e-9,
           flops / 2 / (best * 1e-3) / sms / (clk_khz * 1e3));
  }
  {
    // sustained 3 s loop
    int iters = 20000; const int ILP = 8; dim3 grid(sms * 8), block(256);
    cudaEventRecord(e0);
    int n = 0; float ms = 0;
    do { for (int t = 0; t < 20; ++t) dfma_kernel<ILP><<<grid, block>>>(out, iters, 1.0000001, 1e-9); n += 20;
         cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);} while (ms < 3000);
    double flops = 2.0 * (double)grid.x * 256 * ILP * iters * n;
    printf("{\"bench\": \"dfma_sustained\", \"ms\": %.1f, \"tflops\": %.3f}\n", ms, flops / ms * 1e-9);
  }
  {
    int iters = 5000; const int ILP = 8; dim3 grid(sms * 8), block(256);
    dmma_kernel<ILP><<<grid, block>>>(out, 100); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int t = 0; t < 5; ++t) {
      cudaEventRecord(e0); dmma_kernel<ILP><<<grid, block>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 2.0 * 8 * 8 * 4 * (double)grid.x * 8 * ILP * iters;
    printf("{\"bench\": \"dmma_m8n8k4_throughput\", \"ms\": %.4f, \"tflops\": %.3f}\n", best, flops / best * 1e-9);
  }
  {
    long long h;
    dfma_latency<<<1, 32>>>(out, cyc, 1000, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    dfma_latency<<<1, 32>>>(out, cyc, 10000, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"dfma_dependent_latency_cycles\", \"cycles\": %.2f}\n", (double)h / (10000.0 * 8));
    shfl_latency<<<1, 32>>>(out, cyc, 10000); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"shfl_f64_dependent_latency_cycles\", \"cycles\": %.2f}\n", (double)h / (10000.0 * 8));
    lds_latency<<<1, 32>>>(out, cyc, 10000); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"lds_f64_dependent_latency_cycles(incl cvt)\", \"cycles\": %.2f}\n", (double)h / (10000.0 * 8));
  }
  // occupancy-limited DFMA throughput: 1,2,4,8 warps per SM, ILP 1..8
  for (int warps = 1; warps <= 16; warps *= 2) {
    int iters = 20000; dim3 grid(sms), block(32 * warps);
    float ms1, ms4;
    dfma_kernel<1><<<grid, block>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0); dfma_kernel<1><<<grid, block>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms1, e0, e1);
    dfma_kernel<4><<<grid, block>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0); dfma_kernel<4><<<grid, block>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms4, e0, e1);
    printf("{\"bench\": \"dfma_occupancy\", \"warps_per_sm\": %d, \"tflops_ilp1\": %.3f, \"tflops_ilp4\": %.3f}\n", warps,
           2.0 * sms * 32 * warps * 1 * iters / ms1 * 1e-9, 2.0 * sms * 32 * warps * 4 * iters / ms4 * 1e-9);
  }
  return 0;
}
