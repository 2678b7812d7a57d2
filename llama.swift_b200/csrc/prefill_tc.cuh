// Batched-prefill mat-mul on the 5th-generation tensor cores: tcgen05.mma (kind::f16) with the accumulators in TMEM,
// operands staged in shared memory by TMA -- BASELINE.json configs[2], SURVEY.md section 7 "hard part 4".
//
// What must be reproduced exactly (ggml_compute_forward_mul_mat_q4_0_f32 -> ggml_vec_dot_q4_0, AVX2, ggml.c:1415-1466):
// per (row, token, Q4 block) EIGHT integer sums isum_l, one per AVX accumulator lane l = elements {2l, 2l+1, 16+2l, 17+2l},
// each folded into its own f32 accumulator by acc[l] = fma(d_w * d_x, (float) isum_l, acc[l]), blocks in order.  The scales
// change every 32 elements on BOTH operands, so one MMA accumulator cannot span blocks, and the 8 lanes must stay apart.
//
// Scheme: one MMA per Q4 block (K = 32 elements) with a BLOCK-DIAGONAL activation operand:
//     A  (128 weight rows x 32)  = the block's weights q - 8 as fp16 (exact small integers), K-major
//     B  (32 x 8T)               column (token t, lane l) = token t's quantized activations at lane l's 4 positions, 0 elsewhere
//     D  (128 x 8T, f32 in TMEM) = isum_l of every (row, token): products and 4-term sums of small integers are exact in
//                                  f32, so D IS (float) isum -- no conversion step
// and the epilogue warps drain D after EVERY block (TMEM holds 4 accumulator tiles: the MMAs of the next blocks run under the drain
// of block b) and apply the reference's own per-lane fma chain in registers: 8 fma per (row, token, block), issued as 4
// fma.rn.f32x2.  That chain is the real cost of exact Q4_0 semantics (it is 2/3 of the CUDA-core decode loop as well); the
// tensor cores remove everything else -- nibble handling per token, integer dot products, int -> float conversion.
//
// CTA = one 128-row weight tile x one 16-token tile at a time (persistent over a tile list, token tiles fastest so that the
// CTAs working at the same time share the weight tile through L2).  Warp roles:
//     warps 0-15  epilogue: thread = weight row = TMEM lane (four warps per lane quarter, 4 tokens each: one tcgen05.ld of 32
//                 columns per block step and warp); 4 tokens x 8 lanes of f32 accumulators live in registers
//     warp 16     TMA producer: per (item, quad of blocks) THREE cp.async.bulk boxes -- 10 KB of raw weights (4 blocks x 128
//                 rows x 20 B), 8 KB of zero-interleaved fp16 activations (16 tokens x 4 blocks), 256 B of block scales
//     warp 17     MMA issuer (one elected lane; owns the TMEM allocation)
//     warps 18-21 operand producers: A = 16 nibble bytes -> 32 fp16 per row per block, straight into the UMMA canonical
//                 layout; B = the 4 non-zero fp16 of every (token, lane) column (the zero pattern is written once)
// Pipelines: raw ring (TMA <-> unpack/epilogue), operand buffers (unpack/build <-> MMA via tcgen05.commit), TMEM
// accumulators (MMA <-> epilogue).  All mbarrier based; no __syncthreads in steady state.
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace b200 {

constexpr int TC_M = 128;                 // weight rows per tile = UMMA M = TMEM lanes
constexpr int TC_T = 16;                  // tokens per tile
constexpr int TC_N = TC_T * 8;            // UMMA N: (token, lane) columns
constexpr int TC_EPI_WARPS = 16;          // epilogue warps: 4 per TMEM lane quarter, 4 tokens each
constexpr int TC_THREADS = (TC_EPI_WARPS + 6) * 32;   // + TMA, MMA, 4 operand-producer warps
constexpr int TC_RAW_STAGES = 3;
constexpr int TC_NBUF = 4;                // TMEM accumulator buffers (4 x 128 columns = all of TMEM)
constexpr int TC_ABUF = 8;                // operand buffers (A, B): two quads of blocks, so that the producers fence ONCE per quad
constexpr int TC_QUAD_BYTES = TC_M * 80;  // 4 blocks x 128 rows x 20 B
constexpr int TC_A_BYTES = TC_M * 64;     // 128 rows x 32 fp16
constexpr int TC_B_BYTES = TC_N * 64;     // 128 columns x 32 fp16
constexpr int TC_DX_BYTES = 4 * TC_T * 4; // block scales of the 16 tokens, 4 blocks
constexpr int TC_XH_BYTES = TC_T * 4 * 128; // fp16 activations of the 16 tokens, 4 blocks, zero-interleaved (8 B per (lane, half))
constexpr int TC_STAGE_BYTES = TC_QUAD_BYTES + TC_XH_BYTES + TC_DX_BYTES;
constexpr int TC_SMEM = TC_RAW_STAGES * TC_STAGE_BYTES + TC_ABUF * (TC_A_BYTES + TC_B_BYTES) + 512;

// ---- prefill weight layout: [row tile mt][quad q] -> { [block b < 4][row r < 128][16 raw nibble bytes] | [row][4] f32 d } ---------
__host__ __device__ __forceinline__ size_t tc_weight_bytes(int M, int nb) {
  return (size_t) ((M + TC_M - 1) / TC_M) * ((nb + 3) / 4) * TC_QUAD_BYTES;
}
// src = concatenation of the fused matrices' raw ggml rows; interleave_half as in repack_q4_0_kernel
__global__ void repack_prefill_kernel(const uint8_t *src, uint8_t *dst, int M, int nb, int interleave_half) {
  const int nbq = (nb + 3) >> 2, n_mt = (M + TC_M - 1) / TC_M;
  const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long) n_mt * TC_M * nbq * 4;
  if (idx >= total) return;
  const int gr = (int) (idx / (nbq * 4)), b = (int) (idx % (nbq * 4));
  const int mt = gr / TC_M, r = gr % TC_M, q = b >> 2, bq = b & 3;
  uint8_t *quad = dst + ((size_t) mt * nbq + q) * TC_QUAD_BYTES;
  uint4 *dn = reinterpret_cast<uint4 *>(quad + (size_t) bq * TC_M * 16) + r;
  float *ds = reinterpret_cast<float *>(quad + (size_t) 4 * TC_M * 16) + r * 4 + bq;
  if (gr < M && b < nb) {
    const int sr = interleave_half > 0 ? ((gr & 1) ? interleave_half + gr / 2 : gr / 2) : gr;
    const uint32_t *sp = reinterpret_cast<const uint32_t *>(src + ((size_t) sr * nb + b) * 20);
    *ds = __uint_as_float(sp[0]);
    *dn = make_uint4(sp[1], sp[2], sp[3], sp[4]);
  } else {
    *ds = 0.0f;                                      // padding rows / blocks: scale 0 -> fma(0, isum, acc) = acc
    *dn = make_uint4(0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u);
  }
}

// ---- activation operand: fp16 copy of the quantized activations + block scales, one contiguous box per (token tile, quad) ----
// from batch_prep_kernel's planes (act):
//   xh  [token tile nt][quad q][token t < 16][block b < 4][half h < 2][lane l < 8] 8 bytes
//       = the two elements {2l, 2l+1} (h = 0) or {16+2l, 17+2l} (h = 1) of AVX lane l, already in the form the block-diagonal
//       operand wants them: (x_a, 0, x_b, 0) for l % 4 < 2, (0, x_a, 0, x_b) otherwise (see the K order of the A unpacker);
//       values -127..127, tokens >= N and blocks >= nb are zero
//   dxq [token tile nt][quad q][block b < 4][token t < 16]     float
__global__ void batch_act_tc_kernel(const uint8_t *act, size_t act_stride, __half *xh, float *dxq, int nb, int N, int Npad) {
  const int n = blockIdx.y;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int nbq = (nb + 3) >> 2, nbp = nbq * 4;
  if (b >= nbp) return;
  const int nt = n / TC_T, t = n % TC_T, q = b >> 2, bq = b & 3;
  uint2 *dst = reinterpret_cast<uint2 *>(xh + ((((size_t) nt * nbq + q) * TC_T + t) * 4 + bq) * 64);     // [h][l]
  float *dd = dxq + (((size_t) nt * nbq + q) * 4 + bq) * TC_T + t;
  if (n >= N || b >= nb) {
#pragma unroll
    for (int i = 0; i < 16; i++) dst[i] = make_uint2(0u, 0u);
    *dd = 0.0f;
    return;
  }
  const uint2 *xq = reinterpret_cast<const uint2 *>(act + (size_t) n * act_stride);
  const float *dxs = reinterpret_cast<const float *>(xq + (size_t) nbp * 4);
  // plane p holds {bytes of lane 2p, bytes of lane 2p+1}; lane l's bytes = elements 2l, 2l+1, 16+2l, 17+2l
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const uint2 v = xq[(size_t) p * nbp + b];
    const uint32_t lw[2] = {v.x, v.y};
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int l = 2 * p + s;
      const int sh = (l & 2) ? 16 : 0;                 // odd K positions for l % 4 >= 2
      uint32_t e[4];
#pragma unroll
      for (int i = 0; i < 4; i++) e[i] = (uint32_t) __half_as_ushort(__float2half((float) (int) (int8_t) ((lw[s] >> (8 * i)) & 0xff))) << sh;
      dst[l] = make_uint2(e[0], e[1]);                 // elements 2l, 2l+1
      dst[8 + l] = make_uint2(e[2], e[3]);             // elements 16+2l, 17+2l
    }
  }
  *dd = dxs[b];
}

// ---- tcgen05 / TMEM wrappers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], both K-major fp16, f32 accumulate; issued by ONE thread (SASS UTCHMMA)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices; lbo = bytes between the two 16-byte K
// chunks of one MMA, sbo = bytes between 8-row groups (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t) ((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t) (lbo >> 4) << 16) | ((uint64_t) (sbo >> 4) << 32) | (1ull << 46);
}
// 32 lanes x 32 columns of f32 accumulators -> 32 registers per thread (SASS LDTM)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- lean shared-memory / mbarrier forms on 32-bit shared addresses (the hot loops keep ONE base register) ----------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_u128(uint32_t a) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u64(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// non-blocking probe.  (Experiment, not in the product: issuing the TMA copies from the MMA warp's loop by polling the ring --
// 21 warps instead of 22 -- was 10x slower with try_wait as the poll, which suspends the thread for a system-dependent time
// when the phase is not complete, and still 28% slower with test_wait: the MMA warp is the one role with no slack.)
__device__ __forceinline__ bool mbar_test_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the warp sleeps in the barrier unit (SASS: TRYWAIT, NANOSLEEP.SYNCS, PHASECHK) instead of
// polling through issue slots the critical warps need.  Measured: on EVERY role it is slower than polling (the wake-up adds
// latency to each hand-over: 2-layer 256-token probe 4.22 ms against 3.85 ms), so only roles with slack use it (SLEEP = true).
__device__ __forceinline__ bool mbar_try_wait_sleep_a(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(hint_ns) : "memory");
  return ok != 0;
}
// bounded wait (see mbar_wait in ptx.cuh) that looks at the clock only every 16th wake-up
template <bool SLEEP = false>
__device__ __forceinline__ bool mbar_wait_a(uint32_t bar, uint32_t parity, long long limit) {
  if (mbar_try_wait_a(bar, parity)) return true;
  long long t0 = 0;
  for (uint32_t n = 1; !(SLEEP ? mbar_try_wait_sleep_a(bar, parity, 1000u) : mbar_try_wait_a(bar, parity)); n++) {
    if ((n & 15u) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (wait_give_up(t0, limit)) return false;
    }
  }
  return true;
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Development switch: B200_TC_TRACE=1 stamps clock64() at the hand-over points of CTA 0's first 512 block steps into
// g_tc_trace (read back with b200_debug_tc_trace); tools/tc_trace.py turns it into a per-role timeline.
#ifndef B200_TC_EPI_SLEEP
#define B200_TC_EPI_SLEEP 1     // the epilogue warps' accumulator wait: 1 = sleeping try_wait, 0 = polling
#endif
#ifndef B200_TC_LD16
#define B200_TC_LD16 0          // epilogue: 1 = two tcgen05.ld x16 per block step instead of one x32 (measured: 3.73 ms against 3.61 ms on the 2-layer probe)
#endif
#ifndef B200_TC_TRACE
#define B200_TC_TRACE 0
#endif
#if B200_TC_TRACE
__device__ long long g_tc_trace[12][512];
#define TC_STAMP(slot, idx) do { if (blockIdx.x == 0 && (idx) < 512u) g_tc_trace[slot][idx] = clock64(); } while (0)
#else
#define TC_STAMP(slot, idx) do { } while (0)
#endif

struct GemmTcArgs {
  const uint8_t *w;        // prefill weight layout (repack_prefill_kernel)
  int M, nb;
  const __half *xh;        // [nt][q][16 tokens][4 blocks][32] fp16 quantized activations
  const float *dxT;        // [nt][q][4 blocks][16 tokens] block scales
  float *out;              // [N][ld_out]
  int ld_out, N, Npad;
  long long spin_limit;
};

__global__ void __launch_bounds__(TC_THREADS, 1) q4_gemm_tc_kernel(const GemmTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_tc[];
  uint8_t *smem = smem_tc;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbq = (a.nb + 3) >> 2;
  const int n_mt = (a.M + TC_M - 1) / TC_M, n_nt = a.Npad / TC_T;
  const int n_items = n_mt * n_nt;

  uint8_t *raw = smem;                                                       // [TC_RAW_STAGES][weights quad | xh | dx]
  uint8_t *abuf = smem + TC_RAW_STAGES * TC_STAGE_BYTES;                      // [TC_ABUF][TC_A_BYTES]
  uint8_t *bbuf = abuf + TC_ABUF * TC_A_BYTES;                                // [TC_ABUF][TC_B_BYTES]
  uint64_t *bars = reinterpret_cast<uint64_t *>(bbuf + TC_ABUF * TC_B_BYTES);
  uint64_t *raw_full = bars, *raw_empty = bars + TC_RAW_STAGES;               // TMA <-> unpack / build / epilogue
  uint64_t *ab_full = bars + 2 * TC_RAW_STAGES, *ab_empty = ab_full + TC_ABUF;   // unpack + build <-> MMA
  uint64_t *tm_full = ab_empty + TC_ABUF, *tm_empty = tm_full + TC_NBUF;      // MMA <-> epilogue
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tm_empty + TC_NBUF);

  if (tid == 0) {
    for (int s = 0; s < TC_RAW_STAGES; s++) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], TC_EPI_WARPS + 4); }   // epilogue warps + 4 producer warps
    for (int i = 0; i < TC_ABUF; i++) {
      mbar_init(&ab_full[i], 4);       // the 4 producer warps
      mbar_init(&ab_empty[i], 1);      // tcgen05.commit
    }
    for (int i = 0; i < TC_NBUF; i++) {
      mbar_init(&tm_full[i], 1);       // tcgen05.commit
      mbar_init(&tm_empty[i], TC_EPI_WARPS);   // the epilogue warps
    }
    fence_mbar_init();
  }
  // the zero pattern of the block-diagonal operand is written once; only the non-zero positions are rewritten per block
  for (int i = tid; i < TC_ABUF * TC_B_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4 *>(bbuf)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == TC_EPI_WARPS + 1) tmem_alloc(tmem_slot, TC_NBUF * TC_N);      // TC_NBUF accumulator buffers of 128 columns (all 512 TMEM columns)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long limit = a.spin_limit;
  // Everything below addresses the barriers as 32-bit shared addresses off ONE base register (immediate offsets in the
  // SASS), and walks the rings with running (stage, parity) counters: the first version of this kernel spent most of its
  // issue slots on re-derived addresses, g % 3 divisions and 14-instruction polling loops (profiles/r2_v_tc_timeline.md).
  const uint32_t bar0 = smem_u32(bars);
  constexpr uint32_t RAW_FULL = 0, RAW_EMPTY = 8 * TC_RAW_STAGES, AB_FULL = 16 * TC_RAW_STAGES, AB_EMPTY = AB_FULL + 8 * TC_ABUF;
  constexpr uint32_t TM_FULL = AB_EMPTY + 8 * TC_ABUF, TM_EMPTY = TM_FULL + 8 * TC_NBUF;
  static_assert(TC_NBUF == 4 && TC_ABUF == 8, "a quad of blocks = one round of the TMEM ring = half a round of the operand ring");

  if (warp == TC_EPI_WARPS) {
    // ===== TMA producer: three boxes per (item, quad) -- 10 KB of weights, 8 KB of fp16 activations, 256 B of block scales =====
    if (lane == 0) {
      uint32_t rs = 0, rpar = 1;           // raw ring: stage, parity of the EMPTY barrier to wait for (first round: passes)
      uint32_t g = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int mt = item / n_nt, nt = item % n_nt;
        const uint8_t *wsrc = a.w + (size_t) mt * nbq * TC_QUAD_BYTES;
        const uint8_t *xsrc = reinterpret_cast<const uint8_t *>(a.xh) + (size_t) nt * nbq * TC_XH_BYTES;
        const uint8_t *dsrc = reinterpret_cast<const uint8_t *>(a.dxT) + (size_t) nt * nbq * TC_DX_BYTES;
        for (int q = 0; q < nbq; q++, g++) {
          if (g >= TC_RAW_STAGES && !mbar_wait_a<true>(bar0 + RAW_EMPTY + 8 * rs, rpar, limit)) { item = n_items; break; }   // abandoned: stop issuing
          uint8_t *dst = raw + (size_t) rs * TC_STAGE_BYTES;
          uint64_t *full = raw_full + rs;
          TC_STAMP(0, g);
          mbar_arrive_expect_tx(full, TC_STAGE_BYTES);
          tma_bulk_g2s(dst, wsrc + (size_t) q * TC_QUAD_BYTES, TC_QUAD_BYTES, full);
          tma_bulk_g2s(dst + TC_QUAD_BYTES, xsrc + (size_t) q * TC_XH_BYTES, TC_XH_BYTES, full);
          tma_bulk_g2s(dst + TC_QUAD_BYTES + TC_XH_BYTES, dsrc + (size_t) q * TC_DX_BYTES, TC_DX_BYTES, full);
          if (++rs == TC_RAW_STAGES) { rs = 0; rpar ^= 1; }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ===== MMA issuer: the whole warp walks the loop converged and ONE elected lane issues (no divergent-uniform fix-up code) =====
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = f16, both K-major, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t) (TC_N >> 3) << 17) | ((uint32_t) (TC_M >> 4) << 24);
    // K = 32 elements = 4 chunks of 16 bytes per row; one MMA consumes 2 chunks (K = 16 fp16).  Descriptors differ only in the
    // 14-bit address field: + buffer * 512 (8 KB), + 256 for the second pair of chunks
    const uint64_t a_desc0 = umma_desc(smem_u32(abuf), TC_M * 16, 128), b_desc0 = umma_desc(smem_u32(bbuf), TC_N * 16, 128);
    const bool leader = elect_one_sync();
    uint32_t qg = 0;                      // quads so far: operand half = qg & 1, parities from qg
    bool alive = true;
    for (int item = blockIdx.x; item < n_items && alive; item += gridDim.x) {
      for (int q = 0; q < nbq && alive; q++, qg++) {
        const uint32_t half = (qg & 1) * 4;
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const uint32_t buf = half + b;
          if (!mbar_wait_a(bar0 + AB_FULL + 8 * buf, (qg >> 1) & 1, limit)) { alive = false; break; }
          TC_STAMP(4, qg * 4 + b);
          if (qg >= 1 && !mbar_wait_a(bar0 + TM_EMPTY + 8 * b, (qg - 1) & 1, limit)) { alive = false; break; }
          TC_STAMP(5, qg * 4 + b);
          tc_fence_after();
          if (leader) {
            const uint64_t ad = a_desc0 + (uint64_t) (buf * (TC_A_BYTES >> 4)), bd = b_desc0 + (uint64_t) (buf * (TC_B_BYTES >> 4));
            const uint32_t d_tmem = tmem_base + (uint32_t) b * TC_N;
            umma_f16(d_tmem, ad, bd, idesc, 0u);
            umma_f16(d_tmem, ad + (2 * TC_M * 16 >> 4), bd + (2 * TC_N * 16 >> 4), idesc, 1u);
            umma_commit_a(bar0 + AB_EMPTY + 8 * buf);      // operand buffers free once both MMAs have read them
            umma_commit_a(bar0 + TM_FULL + 8 * b);         // accumulators complete
          }
          __syncwarp();
          TC_STAMP(6, qg * 4 + b);
        }
      }
    }
  } else if (warp >= TC_EPI_WARPS + 2) {
    // ===== operand producers (4 warps).  A: thread = weight row; 16 nibble bytes -> 32 fp16 (q - 8) straight into the UMMA
    // canonical layout.  Any K order is allowed INSIDE an AVX lane's four elements as long as B uses the same one, which buys a
    // 9-instruction unpack per nibble word (SHF, 4 LOP3, 4 HFMA2): chunk c (elements 8c..8c+7 = nibble word c) is stored in the
    // order 0 4 1 5 2 6 3 7 -- (w & 0x000F000F) | 0x6400.. is the fp16 pair (1024 + e0, 1024 + e4), (w & 0x00F000F0) | 0x6400..
    // the pair (1024 + 16 e1, 1024 + 16 e5), and the same on w >> 8 for (e2, e6), (e3, e7).
    // B: lane = one (token, AVX lane) column of this warp's 4 tokens; its two non-zero 8-byte slots per block come zero-interleaved
    // from batch_act_tc_kernel, so the builder is two LDS.64 + two STS.64. =====
    const int pw = warp - (TC_EPI_WARPS + 2);
    const int row = pw * 32 + lane;
    uint32_t rs = 0, rpar = 0, qg = 0;
    const __half2 mul1 = __float2half2_rn(1.0f), add1 = __float2half2_rn(-1032.0f);
    const __half2 mul16 = __float2half2_rn(0.0625f), add16 = __float2half2_rn(-72.0f);
    const uint32_t raw0 = smem_u32(raw) + row * 16, a0 = smem_u32(abuf) + (row >> 3) * 128 + (row & 7) * 16;
    const uint32_t bl = lane & 7, bt = pw * 4 + (lane >> 3);            // column n = bt * 8 + bl
    const uint32_t xs0 = smem_u32(raw) + TC_QUAD_BYTES + bt * 512 + bl * 8;                                  // + b * 128, + 64 for h = 1
    const uint32_t b0 = smem_u32(bbuf) + (bl >> 2) * (TC_N * 16) + (bt * 8 + bl) * 16 + 8 * (bl & 1);        // + 2 * TC_N * 16 for h = 1
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      for (int q = 0; q < nbq; q++, qg++) {
        mbar_wait_a(bar0 + RAW_FULL + 8 * rs, rpar, limit);
        if (row == 0) TC_STAMP(1, qg);
        const uint32_t st = raw0 + rs * TC_STAGE_BYTES, xs = xs0 + rs * TC_STAGE_BYTES;
        const uint32_t half = (qg & 1) * 4;
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const uint4 nib = lds_u128(st + b * TC_M * 16);
          const uint2 x0 = lds_u64(xs + b * 128), x1 = lds_u64(xs + b * 128 + 64);
          const uint32_t ww[4] = {nib.x, nib.y, nib.z, nib.w};
          uint4 out4[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const uint32_t w = ww[c], w8 = w >> 8;
            const uint32_t ta = (w & 0x000F000Fu) | 0x64006400u, tb = (w & 0x00F000F0u) | 0x64006400u;
            const uint32_t tc = (w8 & 0x000F000Fu) | 0x64006400u, td = (w8 & 0x00F000F0u) | 0x64006400u;
            const __half2 ra = __hfma2(*reinterpret_cast<const __half2 *>(&ta), mul1, add1);
            const __half2 rb = __hfma2(*reinterpret_cast<const __half2 *>(&tb), mul16, add16);
            const __half2 rc = __hfma2(*reinterpret_cast<const __half2 *>(&tc), mul1, add1);
            const __half2 rd = __hfma2(*reinterpret_cast<const __half2 *>(&td), mul16, add16);
            out4[c] = make_uint4(*reinterpret_cast<const uint32_t *>(&ra), *reinterpret_cast<const uint32_t *>(&rb),
                                 *reinterpret_cast<const uint32_t *>(&rc), *reinterpret_cast<const uint32_t *>(&rd));
          }
          if (qg >= 2) mbar_wait_a(bar0 + AB_EMPTY + 8 * (half + b), ((qg >> 1) - 1) & 1, limit);
          const uint32_t A = a0 + (half + b) * TC_A_BYTES;
#pragma unroll
          for (int c = 0; c < 4; c++) sts_u128(A + c * (TC_M * 16), out4[c]);
          const uint32_t B = b0 + (half + b) * TC_B_BYTES;
          sts_u64(B, x0);
          sts_u64(B + 2 * TC_N * 16, x1);
        }
        if (row == 0) TC_STAMP(2, qg);
        fence_proxy_async_smem();          // ONE generic -> async proxy fence for the quad's 4 operand buffers
        __syncwarp();
        if (row == 0) TC_STAMP(3, qg);
        if (lane == 0) {
#pragma unroll
          for (int b = 0; b < 4; b++) mbar_arrive_a(bar0 + AB_FULL + 8 * (half + b));
          mbar_arrive_a(bar0 + RAW_EMPTY + 8 * rs);
        }
        if (++rs == TC_RAW_STAGES) { rs = 0; rpar ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps: thread = row = TMEM lane 32*(warp % 4) + lane; warp group (warp / 4) takes TH tokens.  Several
    // warps per TMEM lane quarter, so the others compute while one waits for its tcgen05.ld =====
    constexpr int TH = TC_T / (TC_EPI_WARPS / 4);          // tokens per warp: 4
    static_assert(TH * 8 == 32, "one 32-column tcgen05.ld per block step and warp");
    const int wq = warp & 3, tg = warp >> 2;
    const int row = wq * 32 + lane;
    uint32_t taddr0 = tmem_base + (((uint32_t) (wq * 32)) << 16) + (uint32_t) tg * (TH * 8);
    uint32_t dw0 = smem_u32(raw) + 4 * TC_M * 16 + row * 16, dx0 = smem_u32(raw) + TC_QUAD_BYTES + TC_XH_BYTES + tg * TH * 4;
    // opaque to the compiler: under the 80-register cap it otherwise re-derives these from %tid (S2R + dependent ALU, ~25
    // cycles of exposed latency each) in every block step
    asm volatile("" : "+r"(taddr0), "+r"(dw0), "+r"(dx0));
    uint32_t rs = 0, rpar = 0, qg = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int mt = item / n_nt, nt = item % n_nt;
      u64 acc[TH][4];
#pragma unroll
      for (int t = 0; t < TH; t++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[t][j] = pack_f2(0.0f, 0.0f);
      for (int q = 0; q < nbq; q++, qg++) {
        mbar_wait_a<true>(bar0 + RAW_FULL + 8 * rs, rpar, limit);
        const uint32_t so = rs * TC_STAGE_BYTES;
        // d_w * d_x of a block for the warp's 4 tokens (_mm256_mul_ps(d0, d1), ggml.c:1431): block 0 here, blocks 1..3 one
        // step ahead, under the latency of the previous block's tcgen05.ld
        float sc[TH];
        {
          const float dwb = lds_f32(dw0 + so);
          const float4 dx4 = lds_f32x4(dx0 + so);
          sc[0] = __fmul_rn(dwb, dx4.x); sc[1] = __fmul_rn(dwb, dx4.y); sc[2] = __fmul_rn(dwb, dx4.z); sc[3] = __fmul_rn(dwb, dx4.w);
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
          if (threadIdx.x == 0) TC_STAMP(10, qg * 4 + b);
          mbar_wait_a<B200_TC_EPI_SLEEP != 0>(bar0 + TM_FULL + 8 * b, qg & 1, limit);
          if (threadIdx.x == 0) TC_STAMP(7, qg * 4 + b);
          tc_fence_after();
          float dwn = 0.0f;
          float4 dxn = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#if B200_TC_LD16
          // two 16-column loads: half the live registers of one 32-column load
          uint32_t d[16];
          tmem_ld16(taddr0 + (uint32_t) b * TC_N, d);
          if (b < 3) { dwn = lds_f32(dw0 + so + 4 * (b + 1)); dxn = lds_f32x4(dx0 + so + (b + 1) * TC_T * 4); }
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const u64 s2 = pack_f2(sc[t], sc[t]);
#pragma unroll
            for (int j = 0; j < 4; j++)
              acc[t][j] = ffma2(s2, pack_i2((int) d[t * 8 + 2 * j], (int) d[t * 8 + 2 * j + 1]), acc[t][j]);   // _mm256_fmadd_ps, ggml.c:1457
          }
          tmem_ld16(taddr0 + (uint32_t) b * TC_N + 16u, d);
          tmem_ld_wait();
          if (threadIdx.x == 0) TC_STAMP(8, qg * 4 + b);
          tc_fence_before();
          __syncwarp();
          if (elect_one_sync()) mbar_arrive_a(bar0 + TM_EMPTY + 8 * b);      // the accumulator buffer is free as soon as it is in registers
#pragma unroll
          for (int t = 2; t < 4; t++) {
            const u64 s2 = pack_f2(sc[t], sc[t]);
#pragma unroll
            for (int j = 0; j < 4; j++)
              acc[t][j] = ffma2(s2, pack_i2((int) d[(t - 2) * 8 + 2 * j], (int) d[(t - 2) * 8 + 2 * j + 1]), acc[t][j]);
          }
#else
          uint32_t d[32];                                  // 32 columns = this warp's 4 tokens x 8 lanes
          tmem_ld32(taddr0 + (uint32_t) b * TC_N, d);
          if (b < 3) { dwn = lds_f32(dw0 + so + 4 * (b + 1)); dxn = lds_f32x4(dx0 + so + (b + 1) * TC_T * 4); }
          tmem_ld_wait();
          if (threadIdx.x == 0) TC_STAMP(8, qg * 4 + b);
          tc_fence_before();
          __syncwarp();
          if (elect_one_sync()) mbar_arrive_a(bar0 + TM_EMPTY + 8 * b);      // the accumulator buffer is free as soon as it is in registers
#pragma unroll
          for (int t = 0; t < TH; t++) {
            const u64 s2 = pack_f2(sc[t], sc[t]);
#pragma unroll
            for (int j = 0; j < 4; j++)
              acc[t][j] = ffma2(s2, pack_i2((int) d[t * 8 + 2 * j], (int) d[t * 8 + 2 * j + 1]), acc[t][j]);   // _mm256_fmadd_ps, ggml.c:1457
          }
#endif
          if (b < 3) { sc[0] = __fmul_rn(dwn, dxn.x); sc[1] = __fmul_rn(dwn, dxn.y); sc[2] = __fmul_rn(dwn, dxn.z); sc[3] = __fmul_rn(dwn, dxn.w); }
        }
        if (threadIdx.x == 0) TC_STAMP(9, qg);
        __syncwarp();
        if (elect_one_sync()) mbar_arrive_a(bar0 + RAW_EMPTY + 8 * rs);
        if (++rs == TC_RAW_STAGES) { rs = 0; rpar ^= 1; }
      }
      // horizontal sum exactly as ggml.c:1461-1466: (acc[k] + acc[k+4]) k < 4, then (r0 + r2) + (r1 + r3)
      const int grow = mt * TC_M + row;
#pragma unroll
      for (int t = 0; t < TH; t++) {
        float l[8];
#pragma unroll
        for (int j = 0; j < 4; j++) unpack_f2(acc[t][j], l[2 * j], l[2 * j + 1]);
        const float r0 = __fadd_rn(l[4], l[0]), r1 = __fadd_rn(l[5], l[1]), r2 = __fadd_rn(l[6], l[2]), r3 = __fadd_rn(l[7], l[3]);
        const float res = __fadd_rn(__fadd_rn(r0, r2), __fadd_rn(r1, r3));
        const int n = nt * TC_T + tg * TH + t;
        if (grow < a.M && n < a.N) a.out[(size_t) n * a.ld_out + grow] = res;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_NBUF * TC_N);
  }
}

}  // namespace b200
