// Micro-probe: tensor memory (TMEM) as a per-thread scratchpad: tcgen05.st / tcgen05.ld .32x32b round trip, latency of a
// dependent load chain and throughput of independent loads with 1, 2 and 4 warps (one per lane quadrant) of one CTA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st16(uint32_t ta, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
               "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t ta, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(ta) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void probe(long long *out, int *ok, int mode, int iters) {
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16);
  // fill all 512 columns of my lane: column c holds (tid << 16) | c
  for (int c = 0; c < 512; c += 16) {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = ((uint32_t)threadIdx.x << 16) | (uint32_t)(c + i);
    tm_st16(tb + c, v);
  }
  tm_wait_st();
  // verify
  int good = 1;
  for (int c = 0; c < 512; c += 16) {
    uint32_t v[16];
    tm_ld16(tb + c, v);
    tm_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) good &= (v[i] == (((uint32_t)threadIdx.x << 16) | (uint32_t)(c + i)));
  }
  if (!good) atomicExch(ok, 0);
  __syncthreads();
  uint32_t acc = 0, col = (lane & 1) * 16;
  long long t0 = clock64();
  if (mode == 0) {  // dependent chain: next column depends on the loaded value
    for (int it = 0; it < iters; ++it) {
      uint32_t v[16];
      tm_ld16(tb + col, v);
      tm_wait_ld();
      col = (v[0] + 16 + (v[3] & 0)) & 0x1f0;  // low 16 bits hold the column
      acc += v[5];
    }
  } else {  // throughput: 8 independent loads in flight, one wait
    for (int it = 0; it < iters; it += 8) {
      uint32_t a[16], b[16], c[16], d[16], e[16], f[16], g[16], h[16];
      const uint32_t cb = (uint32_t)((it * 16) & 0x180);
      tm_ld16(tb + cb, a); tm_ld16(tb + cb + 16, b); tm_ld16(tb + cb + 32, c); tm_ld16(tb + cb + 48, d);
      tm_ld16(tb + cb + 64, e); tm_ld16(tb + cb + 80, f); tm_ld16(tb + cb + 96, g); tm_ld16(tb + cb + 112, h);
      tm_wait_ld();
      acc += a[1] + b[2] + c[3] + d[4] + e[5] + f[6] + g[7] + h[8];
    }
  }
  long long t1 = clock64();
  if (lane == 0) out[warp] = t1 - t0;
  if (acc == 0x12345678u) out[8] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s) : "memory");
}

int main() {
  long long *d; int *ok;
  cudaMalloc(&d, 16 * sizeof(long long)); cudaMalloc(&ok, sizeof(int));
  for (int warps : {1, 2, 4}) for (int mode : {0, 1}) {
    int one = 1; cudaMemcpy(ok, &one, sizeof(int), cudaMemcpyHostToDevice);
    const int iters = 4096;
    probe<<<1, warps * 32>>>(d, ok, mode, 64);
    probe<<<1, warps * 32>>>(d, ok, mode, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[4] = {0, 0, 0, 0}; int good = 0;
    cudaMemcpy(c, d, warps * sizeof(long long), cudaMemcpyDeviceToHost); cudaMemcpy(&good, ok, sizeof(int), cudaMemcpyDeviceToHost);
    printf("{\"bench\": \"tmem_probe\", \"mode\": \"%s\", \"warps\": %d, \"cycles_per_ld_x16\": %.2f, \"roundtrip_ok\": %d, \"cuda\": \"%s\"}\n",
           mode == 0 ? "dependent_chain" : "independent_x8", warps, (double)c[0] / iters, good, cudaGetErrorString(e));
  }
  return 0;
}
