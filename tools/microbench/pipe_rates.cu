// Micro-benchmark (development aid): latency and throughput of the instructions the Q4_0 row loop is built from, on one SM
// and on the whole chip -- IMMA.16832.U8.S8 (legacy mma.sync path), IDP.4A, FFMA2, LDSM.  nvcc -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1, const int (&c)[4]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
}

template <int ILP>
__global__ void k_imma(int iters, int *out, long long *cyc) {
  int d[ILP][4];
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) d[i][j] = threadIdx.x + i + j;
  uint32_t a = threadIdx.x * 0x01010101u, b = 0x01020304u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) mma(d[i], a, a + 1, a + 2, a + 3, b, b + 1, d[i]);    // dependent through C within a chain
  }
  long long t1 = clock64();
  int s = 0;
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
__global__ void k_dp4a(int iters, int *out, long long *cyc) {
  int d[ILP];
  for (int i = 0; i < ILP; i++) d[i] = threadIdx.x + i;
  int a = threadIdx.x * 0x01010101, b = 0x01020304;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(d[i]) : "r"(a), "r"(b));
  }
  long long t1 = clock64();
  int s = 0;
  for (int i = 0; i < ILP; i++) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
__global__ void k_ffma2(int iters, int *out, long long *cyc) {
  unsigned long long d[ILP];
  for (int i = 0; i < ILP; i++) d[i] = threadIdx.x + i;
  unsigned long long a = 0x3f8000003f800000ull, b = 0x3f0000003f000000ull;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f32x2 %0, %1, %0, %2;" : "+l"(d[i]) : "l"(a), "l"(b));
  }
  long long t1 = clock64();
  unsigned long long s = 0;
  for (int i = 0; i < ILP; i++) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (int) s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K>
void run(const char *name, K kern, int warps, int ilp, int blocks, double ops_per_instr) {
  int *out; long long *cyc;
  cudaMalloc(&out, (size_t) blocks * warps * 32 * 4); cudaMalloc(&cyc, blocks * 8);
  const int iters = 4096;
  kern<<<blocks, warps * 32>>>(iters, out, cyc);
  kern<<<blocks, warps * 32>>>(iters, out, cyc);
  cudaDeviceSynchronize();
  long long h[1024]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < blocks; i++) c += h[i]; c /= blocks;
  const double instr = (double) iters * ilp * warps;          // warp-instructions per SM
  printf("%-8s warps/SM %2d ILP %d: %8.0f cycles, %.2f cycles per warp-instr per SM (%.2f warp-instr/clk/SM), per-chain latency %.1f cycles\n",
         name, warps, ilp, c, c / instr, instr / c, c / iters);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, nsm);
  for (int warps : {1, 4, 8, 16}) {
    run("IMMA", k_imma<1>, warps, 1, nsm, 1);
    run("IMMA", k_imma<4>, warps, 4, nsm, 1);
    run("IMMA", k_imma<8>, warps, 8, nsm, 1);
  }
  for (int warps : {1, 4, 16}) {
    run("IDP.4A", k_dp4a<1>, warps, 1, nsm, 1);
    run("IDP.4A", k_dp4a<8>, warps, 8, nsm, 1);
    run("FFMA2", k_ffma2<1>, warps, 1, nsm, 1);
    run("FFMA2", k_ffma2<8>, warps, 8, nsm, 1);
  }
  return 0;
}
