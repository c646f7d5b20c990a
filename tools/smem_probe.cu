// Micro-probe: shared-memory pipe cost (cycles per warp instruction per SM) of the access patterns the solve kernel uses.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/smem_probe tools/smem_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define REP 64
template <int PAT>
__global__ void probe(double *out, long long *cyc, int iters, int stride_g) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, r = lane & 7;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double acc = 0.0;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + warp * 1024;
  unsigned a;
  if (PAT == 0) a = base + lane * 16;                                   // LDS.128, all lanes distinct, conflict-free (512 B)
  if (PAT == 1) a = base + g * stride_g;                                // LDS.128, broadcast inside each 8-lane group
  if (PAT == 2) a = base + g * stride_g + r * 64;                       // LDS.128 rows with 64-byte stride (4-way conflict)
  if (PAT == 3) a = base + g * stride_g + r * 64 + (((r >> 1) & 3) << 4);  // LDS.128 swizzled rows (conflict-free per group)
  if (PAT == 4) a = base + g * stride_g + r * 8;                        // LDS.64, 8 consecutive doubles per group
  if (PAT == 5) a = base + lane * 8;                                    // LDS.64 fully coalesced
  if (PAT == 6) a = base;                                               // SHFL (no address)
  if (PAT == 7) a = base + g * stride_g + ((r == 0) ? 0 : (r == 6 ? 16 : 32));  // LDS.64 by lanes 0,6,7 of each group
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < REP; ++i) {
      if (PAT <= 3) { double x, y; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a + i * 128 + (it & 1) * 64) : "memory"); acc += x; }
      else if (PAT == 4 || PAT == 5) { double x; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a + i * 128 + (it & 1) * 64) : "memory"); acc += x; }
      else if (PAT == 6) { acc = __shfl_sync(0xffffffffu, acc, (lane + 1) & 7, 8); }
      else if (PAT == 7) { if (r == 0 || r >= 6) { double x; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a + i * 128 + (it & 1) * 64) : "memory"); acc += x; } }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int PAT>
void run(const char *name, int warps, int stride_g, double *d, long long *dc) {
  const int iters = 64;
  cudaFuncSetAttribute(probe<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  probe<PAT><<<1, warps * 32, 65536>>>(d, dc, 4, stride_g);
  probe<PAT><<<1, warps * 32, 65536>>>(d, dc, iters, stride_g);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("{\"bench\": \"smem_probe\", \"pattern\": \"%s\", \"warps\": %d, \"group_stride_B\": %d, \"sm_cycles_per_warp_instr\": %.2f}\n", name, warps, stride_g,
         (double)c / (iters * REP * warps));
}
int main() {
  double *d; long long *dc;
  cudaMalloc(&d, 65536 * sizeof(double)); cudaMalloc(&dc, sizeof(long long));
  for (int w : {1, 4, 8, 16}) {
    run<0>("lds128_distinct", w, 0, d, dc);
    run<1>("lds128_group_broadcast", w, 64, d, dc);
    run<1>("lds128_group_broadcast", w, 16, d, dc);
    run<1>("lds128_group_broadcast", w, 272, d, dc);
    run<2>("lds128_rows_stride64", w, 64 * 9, d, dc);
    run<3>("lds128_rows_swizzled", w, 64 * 9, d, dc);
    run<4>("lds64_group8", w, 64, d, dc);
    run<4>("lds64_group8", w, 128, d, dc);
    run<5>("lds64_coalesced", w, 0, d, dc);
    run<6>("shfl64_width8", w, 0, d, dc);
    run<7>("lds64_3of8_lanes", w, 64, d, dc);
  }
  return 0;
}
