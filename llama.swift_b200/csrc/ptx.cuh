// Thin inline-PTX wrappers for sm_100a: packed f32x2 math, mixed-sign dp4a, mbarrier, 1-D TMA bulk copy,
// programmatic dependent launch.  No torch, no CUTLASS.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack_f2(float lo, float hi) {
  u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ u64 pack_i2(int lo, int hi) {
  u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
__device__ __forceinline__ void unpack_f2(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// two IEEE fp32 FMAs in one instruction (SASS FFMA2) -- each half rounds exactly like fmaf
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
// two IEEE fp32 multiplies / adds in one instruction (SASS FMUL2 / FADD2) -- each half rounds like __fmul_rn / __fadd_rn.
// CAUTION (measured, ptxas 12.9): fadd2(fmul2(a, b), c) is CONTRACTED into one FFMA2 despite the .rn modifiers and -fmad=false;
// do not feed a packed product straight into a packed add where the reference rounds twice (kernels_q4_1.cuh).
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
  u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
// unsigned bytes (a) x signed bytes (b) + c  (SASS IDP.4A.U8.S8)
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
  int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

// signed bytes (a) x signed bytes (b) + c  (SASS IDP.4A.S8.S8)
__device__ __forceinline__ int dp4a_ss(int a, int b, int c) {
  int d; asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

// (a & b) ^ c in ONE LOP3 (immLut 0x6A).  Written as C++ the compiler splits it into two LOP3s when b and c are both
// constants (the instruction has a single immediate slot); with register operands one constant stays in a register.
__device__ __forceinline__ uint32_t and_xor(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

// D(16x8, s32) = A(16x32, u8, row) * B(32x8, s8, col) + C on the warp-level tensor path (SASS IMMA.16832.U8.S8).
// Fragment owner: lane = 4*g + t.  a0 = A[g][4t..4t+3], a1 = A[g+8][4t..], a2 = A[g][16+4t..], a3 = A[g+8][16+4t..];
// b0 = B[4t..4t+3][g], b1 = B[16+4t..][g]; d0,d1 = D[g][2t, 2t+1], d2,d3 = D[g+8][2t, 2t+1].
__device__ __forceinline__ void mma_u8s8_16832(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1, int c) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(c));
}
// 8x8 b16 matrices from shared memory in MMA fragment order (SASS LDSM): lane i gets bytes 4(i%4)..4(i%4)+3 of row i/4;
// lanes 0-7 supply the 16-byte row addresses of matrix 0, lanes 8-15 of matrix 1 (x2)
__device__ __forceinline__ void ldmatrix_x2(uint32_t &r0, uint32_t &r1, uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_addr) : "memory");
}
__device__ __forceinline__ void ldmatrix_x1(uint32_t &r0, uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(smem_addr) : "memory");
}
// byte permute: result byte i = byte sel[4i+3:4i] of {b (bytes 4-7), a (bytes 0-3)}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// ---- bounded waits ----------------------------------------------------------------------------------------------------
// A lost arrival must never hang the GPU box, and a `trap` would poison the CUDA context of the whole process.  Instead a
// spin-wait that outlives its budget sets the per-device abort word and FALLS THROUGH: the kernel runs to completion on
// garbage, so every barrier is still reached and every thread exits normally; any other wait that has already spun for
// ~0.5 ms looks at the word and gives up too (a healthy wait is far shorter, so the word is never read in normal
// operation).  The host reads the word after the launch, reports -1001 and resets the exchange state (engine.cu).
__device__ unsigned int g_b200_abort;
__device__ __forceinline__ bool wait_give_up(long long t0, long long limit) {
  const long long dt = clock64() - t0;
  if (dt > limit || (dt > 1000000LL && *reinterpret_cast<volatile unsigned int *>(&g_b200_abort) != 0u)) {
    *reinterpret_cast<volatile unsigned int *>(&g_b200_abort) = 1u;
    return true;
  }
  return false;
}
// returns false when the wait was abandoned.  CONSUMERS may ignore that (they go on with garbage and keep arriving once per
// stage, so no barrier is over- or under-subscribed); a PRODUCER must stop issuing: re-arming a full barrier whose previous
// phase never completed overflows its transaction count.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, long long limit = 4000000000LL) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (wait_give_up(t0, limit)) return false;
  }
  return true;
}
// 1-D bulk async copy global -> shared, completion counted on an mbarrier (SASS UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// The same copy with an L2 eviction-priority hint: the weight stream is read exactly once per token, so it is marked
// evict_first and does not push the per-layer re-used lines (KV cache, lookup tables, activations) out of L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// Programmatic dependent launch: wait for the upstream grid's memory to be visible / let the downstream grid start
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace b200
