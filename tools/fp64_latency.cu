// Micro-probe: latency (dependent chain) and single-warp issue interval of the fp64 / shuffle / shared-memory operations
// the solve kernel's critical path is made of.  One CTA, clock64() around an unrolled chain.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256

template <int OP>
__global__ void chain(double *out, long long *cyc, double m, double c, int nwarps_active) {
  __shared__ double sm[32 * 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double a = threadIdx.x * 1e-3 + 1.0, b0 = a + 1, b1 = a + 2, b2 = a + 3, b3 = a + 4, b4 = a + 5, b5 = a + 6, b6 = a + 7;
  sm[threadIdx.x & 511] = a;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < REP; ++i) {
    if (OP == 0) a = fma(a, m, c);                       // DFMA latency
    if (OP == 1) a = a + c;                              // DADD latency
    if (OP == 2) a = a * m;                              // DMUL latency
    if (OP == 3) {                                       // DFMA issue interval: 8 independent chains
      a = fma(a, m, c); b0 = fma(b0, m, c); b1 = fma(b1, m, c); b2 = fma(b2, m, c);
      b3 = fma(b3, m, c); b4 = fma(b4, m, c); b5 = fma(b5, m, c); b6 = fma(b6, m, c);
    }
    if (OP == 4) a = __shfl_xor_sync(0xffffffffu, a, 1);  // 64-bit shuffle (2 x SHFL.32)
    if (OP == 5) {                                       // STS.64 -> syncwarp -> LDS.64 of a neighbour's value
      sm[warp * 32 + lane] = a; __syncwarp(); a = sm[warp * 32 + (lane ^ 1)]; __syncwarp();
    }
    if (OP == 6) {                                       // LDS.64 pointer-chase-like dependent load
      a = sm[((int)a) & 255];
    }
    if (OP == 7) a = 1.0 / a;                            // fp64 reciprocal (div.rn)
    if (OP == 8) a = sqrt(a);                            // fp64 sqrt
    if (OP == 9) {                                       // STS.64 -> LDS.128 x4 gather of 8 doubles, then 1 DFMA
      sm[warp * 32 + lane] = a; __syncwarp();
      const double2 *p = reinterpret_cast<const double2 *>(&sm[warp * 32 + (lane & 24)]);
      double2 g0 = p[0], g1 = p[1], g2 = p[2], g3 = p[3]; __syncwarp();
      a = fma(g0.x, m, g0.y) + fma(g1.x, m, g1.y) + fma(g2.x, m, g2.y) + fma(g3.x, m, g3.y);
    }
    if (OP == 10) {                                      // gather of 8 doubles by 8 x 64-bit shuffles (width 8), then same math
      double g[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) g[q] = __shfl_sync(0xffffffffu, a, q, 8);
      a = fma(g[0], m, g[1]) + fma(g[2], m, g[3]) + fma(g[4], m, g[5]) + fma(g[6], m, g[7]);
    }
    if (OP == 11) a = fma(a, m, c) * (float)1.0f;        // (placeholder: same as 0)
  }
  long long t1 = clock64();
  out[threadIdx.x] = a + b0 + b1 + b2 + b3 + b4 + b5 + b6;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char *name, int threads, double *d, long long *dc, int ops_per_rep) {
  chain<OP><<<1, threads>>>(d, dc, 1.0000001, 1e-9, 0);
  chain<OP><<<1, threads>>>(d, dc, 1.0000001, 1e-9, 0);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("{\"bench\": \"fp64_latency\", \"op\": \"%s\", \"warps\": %d, \"cycles_per_rep\": %.2f, \"ops_per_rep\": %d}\n", name, threads / 32,
         (double)c / REP, ops_per_rep);
}

int main() {
  double *d; long long *dc;
  cudaMalloc(&d, 1024 * sizeof(double)); cudaMalloc(&dc, sizeof(long long));
  int ths[] = {32, 128, 256, 512};
  for (int t : ths) {
    run<0>("dfma_chain", t, d, dc, 1);
    run<1>("dadd_chain", t, d, dc, 1);
    run<2>("dmul_chain", t, d, dc, 1);
    run<3>("dfma_x8_independent", t, d, dc, 8);
    run<4>("shfl64_chain", t, d, dc, 1);
    run<5>("sts64_lds64_roundtrip", t, d, dc, 1);
    run<6>("lds64_dependent", t, d, dc, 1);
    run<7>("drcp", t, d, dc, 1);
    run<8>("dsqrt", t, d, dc, 1);
    run<9>("gather8_smem_plus_dot", t, d, dc, 1);
    run<10>("gather8_shfl_plus_dot", t, d, dc, 1);
  }
  return 0;
}
