// Development micro-benchmark: how fast can one CTA per SM pull a contiguous HBM stream into shared memory?
//   mode 0: cp.async.bulk (1-D TMA, UBLKCP), one issuing thread, S stages of C bytes
//   mode 1: cp.async 16 B (LDGSTS) issued by 256 threads, S stages
//   mode 2: ld.global.v4 (LDG.128, L1 no-allocate) by 256 threads into registers only (no smem), unroll 8
//   mode 3: cp.async.bulk issued by 4 threads (4 independent rings)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench tools/stream_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../llama.swift_b200/csrc/ptx.cuh"
using namespace b200;

__global__ void __launch_bounds__(288, 1) k_bulk(const uint8_t *src, size_t per_cta, int C, int S, int nissue, unsigned long long *sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t *full = reinterpret_cast<uint64_t *>(sm + (size_t) S * C);
  uint64_t *empty = full + S;
  const int tid = threadIdx.x;
  if (tid == 0) { for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); } fence_mbar_init(); }
  __syncthreads();
  const uint8_t *base = src + (size_t) blockIdx.x * per_cta;
  const int n = (int) (per_cta / C);
  if (tid >= 256) {
    const int it = tid - 256;
    if (it < nissue) {
      for (int k = it; k < n; k += nissue) {
        const int s = k % S; const int lap = k / S;
        if (lap > 0) mbar_wait(&empty[s], (lap - 1) & 1);
        mbar_arrive_expect_tx(&full[s], C);
        tma_bulk_g2s(sm + (size_t) s * C, base + (size_t) k * C, C, &full[s]);
      }
    }
    return;
  }
  unsigned long long acc = 0;
  for (int k = 0; k < n; k++) {
    const int s = k % S;
    mbar_wait(&full[s], (k / S) & 1);
    acc += reinterpret_cast<const uint32_t *>(sm + (size_t) s * C)[tid];
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
  }
  if (acc == 0x1234567) sink[0] = acc;
}

__global__ void __launch_bounds__(256, 1) k_cpasync(const uint8_t *src, size_t per_cta, int C, int S, unsigned long long *sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int tid = threadIdx.x;
  const uint8_t *base = src + (size_t) blockIdx.x * per_cta;
  const int n = (int) (per_cta / C);
  unsigned long long acc = 0;
  auto issue = [&](int k) {
    if (k < n) {
      const uint8_t *g = base + (size_t) k * C;
      uint8_t *d = sm + (size_t) (k % S) * C;
      for (int o = tid * 16; o < C; o += 256 * 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d + o)), "l"(g + o));
    }
    asm volatile("cp.async.commit_group;");
  };
  for (int k = 0; k < S - 1; k++) issue(k);
  for (int k = 0; k < n; k++) {
    issue(k + S - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(6));
    __syncthreads();
    acc += reinterpret_cast<const uint32_t *>(sm + (size_t) (k % S) * C)[tid];
    __syncthreads();
  }
  if (acc == 0x1234567) sink[0] = acc;
}

__global__ void __launch_bounds__(256, 1) k_ldg(const uint8_t *src, size_t per_cta, unsigned long long *sink) {
  const int tid = threadIdx.x;
  const uint4 *p = reinterpret_cast<const uint4 *>(src + (size_t) blockIdx.x * per_cta);
  const size_t n = per_cta / 16;
  unsigned int acc = 0;
  for (size_t i = tid; i + 7 * 256 < n; i += 8 * 256) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + u * 256));
#pragma unroll
    for (int u = 0; u < 8; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x1234567) sink[0] = acc;
}

int main() {
  const int n_cta = 148;
  const size_t per_cta = 24u << 20;              // 24 MB per CTA -> 3.5 GB total
  uint8_t *src; unsigned long long *sink;
  cudaMalloc(&src, per_cta * n_cta); cudaMalloc(&sink, 8);
  cudaMemset(src, 1, per_cta * n_cta);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char *name, float ms) { printf("%-44s %8.1f GB/s  (%.3f ms)\n", name, per_cta * n_cta / ms / 1e6, ms); fflush(stdout); };
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int C : {8192, 16384, 32768, 65536}) {
    for (int S : {2, 4, 6}) {
      if ((size_t) C * S + 256 > 220 * 1024) continue;
      for (int ni : {1}) {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
          cudaEventRecord(e0);
          k_bulk<<<n_cta, 288, (size_t) C * S + 256>>>(src, per_cta, C, S, ni, sink);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        char nm[96]; snprintf(nm, sizeof nm, "bulk TMA  chunk %6d  stages %d  issuers %d", C, S, ni); report(nm, best);
      }
    }
  }
  for (int C : {16384, 32768}) {
    const int S = 8;
    if ((size_t) C * S > 220 * 1024) continue;
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      k_cpasync<<<n_cta, 256, (size_t) C * S>>>(src, per_cta, C, S, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    char nm[96]; snprintf(nm, sizeof nm, "cp.async 16B  chunk %6d  stages %d (wait 6)", C, S); report(nm, best);
  }
  {
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      k_ldg<<<n_cta, 256>>>(src, per_cta, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    report("LDG.128 x8 unrolled, 256 thr, 1 CTA/SM", best);
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
