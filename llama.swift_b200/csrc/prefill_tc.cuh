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
// and the epilogue warps drain D after EVERY block (TMEM is double-buffered: the MMA of block b+1 runs under the drain of
// block b) and apply the reference's own per-lane fma chain in registers: 8 fma per (row, token, block), issued as 4
// fma.rn.f32x2.  That chain is the real cost of exact Q4_0 semantics (it is 2/3 of the CUDA-core decode loop as well); the
// tensor cores remove everything else -- nibble handling per token, integer dot products, int -> float conversion.
//
// CTA = one 128-row weight tile x one 16-token tile at a time (persistent over a tile list, token tiles fastest so that the
// CTAs working at the same time share the weight tile through L2).  Warp roles:
//     warps 0-15  epilogue: thread = weight row = TMEM lane (four warps per lane quarter, 4 tokens each: one tcgen05.ld of 32
//                 columns per block step and warp); 4 tokens x 8 lanes of f32 accumulators live in registers
//     warp 16     TMA producer: per (item, quad of blocks) THREE cp.async.bulk boxes -- 10 KB of raw weights (4 blocks x 128
//                 rows x 20 B), 4 KB of fp16 activations (16 tokens x 4 blocks), 256 B of block scales
//     warp 17     MMA issuer (one elected lane; owns the TMEM allocation)
//     warp 18     B builder: writes the 4 non-zero fp16 of every (token, lane) column; the zero pattern is written once
//     warps 19-22 A unpacker: 16 nibble bytes -> 32 fp16 per row per block, straight into the UMMA canonical layout
// Pipelines: raw ring (TMA <-> unpack/epilogue), operand buffers (unpack/build <-> MMA via tcgen05.commit), TMEM
// accumulators (MMA <-> epilogue).  All mbarrier based; no __syncthreads in steady state.
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

#ifndef B200_TC_SKIP
#define B200_TC_SKIP 0      // development ceilings (wrong results): 1 = epilogue without the fma chain, 2 = without the tcgen05.ld, 3 = A unpack without the conversion
#endif

namespace b200 {

constexpr int TC_M = 128;                 // weight rows per tile = UMMA M = TMEM lanes
constexpr int TC_T = 16;                  // tokens per tile
constexpr int TC_N = TC_T * 8;            // UMMA N: (token, lane) columns
constexpr int TC_EPI_WARPS = 16;          // epilogue warps: 4 per TMEM lane quarter, 4 tokens each
constexpr int TC_THREADS = (TC_EPI_WARPS + 7) * 32;   // + TMA, MMA, B builder, 4 unpack warps
constexpr int TC_RAW_STAGES = 3;
constexpr int TC_QUAD_BYTES = TC_M * 80;  // 4 blocks x 128 rows x 20 B
constexpr int TC_A_BYTES = TC_M * 64;     // 128 rows x 32 fp16
constexpr int TC_B_BYTES = TC_N * 64;     // 128 columns x 32 fp16
constexpr int TC_DX_BYTES = 4 * TC_T * 4; // block scales of the 16 tokens, 4 blocks
constexpr int TC_XH_BYTES = TC_T * 4 * 64; // fp16 activations of the 16 tokens, 4 blocks
constexpr int TC_STAGE_BYTES = TC_QUAD_BYTES + TC_XH_BYTES + TC_DX_BYTES;
constexpr int TC_SMEM = TC_RAW_STAGES * TC_STAGE_BYTES + 2 * TC_A_BYTES + 2 * TC_B_BYTES + 256;

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
//   xh  [token tile nt][quad q][token t < 16][block b < 4][32] half   (values -7..7; tokens >= N and blocks >= nb are zero)
//   dxq [token tile nt][quad q][block b < 4][token t < 16]     float
__global__ void batch_act_tc_kernel(const uint8_t *act, size_t act_stride, __half *xh, float *dxq, int nb, int N, int Npad) {
  const int n = blockIdx.y;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int nbq = (nb + 3) >> 2, nbp = nbq * 4;
  if (b >= nbp) return;
  const int nt = n / TC_T, t = n % TC_T, q = b >> 2, bq = b & 3;
  __half2 *dst = reinterpret_cast<__half2 *>(xh + ((((size_t) nt * nbq + q) * TC_T + t) * 4 + bq) * 32);
  float *dd = dxq + (((size_t) nt * nbq + q) * 4 + bq) * TC_T + t;
  if (n >= N || b >= nb) {
#pragma unroll
    for (int i = 0; i < 16; i++) dst[i] = __floats2half2_rn(0.0f, 0.0f);
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
      const int e0 = (int) (int8_t) (lw[s] & 0xff), e1 = (int) (int8_t) ((lw[s] >> 8) & 0xff);
      const int e2 = (int) (int8_t) ((lw[s] >> 16) & 0xff), e3 = (int) (int8_t) (lw[s] >> 24);
      dst[l] = __floats2half2_rn((float) e0, (float) e1);            // elements 2l, 2l+1
      dst[8 + l] = __floats2half2_rn((float) e2, (float) e3);        // elements 16+2l, 17+2l
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
  uint8_t *abuf = smem + TC_RAW_STAGES * TC_STAGE_BYTES;                      // [2][TC_A_BYTES]
  uint8_t *bbuf = abuf + 2 * TC_A_BYTES;                                      // [2][TC_B_BYTES]
  uint64_t *bars = reinterpret_cast<uint64_t *>(bbuf + 2 * TC_B_BYTES);
  uint64_t *raw_full = bars, *raw_empty = bars + TC_RAW_STAGES;               // TMA <-> unpack / build / epilogue
  uint64_t *ab_full = bars + 2 * TC_RAW_STAGES, *ab_empty = ab_full + 2;      // unpack + build <-> MMA
  uint64_t *tm_full = ab_empty + 2, *tm_empty = tm_full + 2;                  // MMA <-> epilogue
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tm_empty + 2);

  if (tid == 0) {
    for (int s = 0; s < TC_RAW_STAGES; s++) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], TC_EPI_WARPS + 5); }   // epilogue warps + B builder + 4 unpack warps
    for (int i = 0; i < 2; i++) {
      mbar_init(&ab_full[i], 5);       // 4 unpack warps + the B builder
      mbar_init(&ab_empty[i], 1);      // tcgen05.commit
      mbar_init(&tm_full[i], 1);       // tcgen05.commit
      mbar_init(&tm_empty[i], TC_EPI_WARPS);   // the epilogue warps
    }
    fence_mbar_init();
  }
  // the zero pattern of the block-diagonal operand is written once; only the non-zero positions are rewritten per block
  for (int i = tid; i < 2 * TC_B_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4 *>(bbuf)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == TC_EPI_WARPS + 1) tmem_alloc(tmem_slot, 256);      // 2 accumulator buffers of 128 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long limit = a.spin_limit;

  if (warp == TC_EPI_WARPS) {
    // ===== TMA producer: three boxes per (item, quad) -- 10 KB of weights, 4 KB of fp16 activations, 256 B of block scales =====
    if (lane == 0) {
      uint32_t g = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int mt = item / n_nt, nt = item % n_nt;
        for (int q = 0; q < nbq; q++, g++) {
          const int s = g % TC_RAW_STAGES;
          if (g >= TC_RAW_STAGES && !mbar_wait(&raw_empty[s], ((g / TC_RAW_STAGES) - 1) & 1, limit)) { item = n_items; break; }   // abandoned: stop issuing
          uint8_t *dst = raw + (size_t) s * TC_STAGE_BYTES;
          mbar_arrive_expect_tx(&raw_full[s], TC_STAGE_BYTES);
          tma_bulk_g2s(dst, a.w + ((size_t) mt * nbq + q) * TC_QUAD_BYTES, TC_QUAD_BYTES, &raw_full[s]);
          tma_bulk_g2s(dst + TC_QUAD_BYTES, reinterpret_cast<const uint8_t *>(a.xh) + ((size_t) nt * nbq + q) * TC_XH_BYTES, TC_XH_BYTES, &raw_full[s]);
          tma_bulk_g2s(dst + TC_QUAD_BYTES + TC_XH_BYTES, reinterpret_cast<const uint8_t *>(a.dxT) + ((size_t) nt * nbq + q) * TC_DX_BYTES, TC_DX_BYTES, &raw_full[s]);
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = f16, both K-major, N = 128, M = 128
      const uint32_t idesc = (1u << 4) | ((uint32_t) (TC_N >> 3) << 17) | ((uint32_t) (TC_M >> 4) << 24);
      uint32_t kb_g = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        for (int kb = 0; kb < nbq * 4; kb++, kb_g++) {
          const int buf = kb_g & 1;
          const uint32_t par = (kb_g >> 1) & 1;
          if (!mbar_wait(&ab_full[buf], par, limit) || (kb_g >= 2 && !mbar_wait(&tm_empty[buf], ((kb_g >> 1) - 1) & 1, limit))) { item = n_items; break; }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(abuf + buf * TC_A_BYTES), b_addr = smem_u32(bbuf + buf * TC_B_BYTES);
          const uint32_t d_tmem = tmem_base + (uint32_t) buf * TC_N;
          // K = 32 elements = 4 chunks of 16 bytes per row; one MMA consumes 2 chunks (K = 16 fp16)
          umma_f16(d_tmem, umma_desc(a_addr, TC_M * 16, 128), umma_desc(b_addr, TC_N * 16, 128), idesc, 0u);
          umma_f16(d_tmem, umma_desc(a_addr + 2 * TC_M * 16, TC_M * 16, 128), umma_desc(b_addr + 2 * TC_N * 16, TC_N * 16, 128), idesc, 1u);
          umma_commit(&ab_empty[buf]);      // operand buffers free once both MMAs have read them
          umma_commit(&tm_full[buf]);       // accumulators complete
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 2) {
    // ===== B builder: column (t, l) of block kb gets token t's activations at elements 2l, 2l+1 (K chunk l/4) and 16+2l, 17+2l
    // (K chunk 2 + l/4), read from the staged fp16 copy; every lane handles 4 (t, l) pairs =====
    uint32_t g = 0, kb_g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      for (int q = 0; q < nbq; q++, g++) {
        const int s = g % TC_RAW_STAGES;
        mbar_wait(&raw_full[s], (g / TC_RAW_STAGES) & 1, limit);
        const uint32_t *xs = reinterpret_cast<const uint32_t *>(raw + (size_t) s * TC_STAGE_BYTES + TC_QUAD_BYTES);   // [t][b][16 words]
        for (int b = 0; b < 4; b++, kb_g++) {
          const int buf = kb_g & 1;
          uint32_t v0[4], v1[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int i = lane + 32 * j, t = i >> 3, l = i & 7;
            v0[j] = xs[(t * 4 + b) * 16 + l];
            v1[j] = xs[(t * 4 + b) * 16 + 8 + l];
          }
          if (kb_g >= 2) mbar_wait(&ab_empty[buf], ((kb_g >> 1) - 1) & 1, limit);
          uint8_t *B = bbuf + buf * TC_B_BYTES;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int i = lane + 32 * j, t = i >> 3, l = i & 7;
            uint8_t *p = B + (l >> 2) * (TC_N * 16) + t * 128 + l * 16 + 4 * (l & 3);
            *reinterpret_cast<uint32_t *>(p) = v0[j];
            *reinterpret_cast<uint32_t *>(p + 2 * TC_N * 16) = v1[j];
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ab_full[buf]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[s]);
      }
    }
  } else if (warp >= TC_EPI_WARPS + 3) {
    // ===== A unpacker: thread = weight row; 16 nibble bytes -> 32 fp16 (q - 8), chunk c = elements 8c..8c+7 = nibble word c =====
    const int row = (warp - (TC_EPI_WARPS + 3)) * 32 + lane;
    uint32_t g = 0, kb_g = 0;
    const __half2 mulv = __halves2half2(__float2half(1.0f), __float2half(0.0625f));
    const __half2 addv = __halves2half2(__float2half(-1032.0f), __float2half(-72.0f));
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      for (int q = 0; q < nbq; q++, g++) {
        const int s = g % TC_RAW_STAGES;
        mbar_wait(&raw_full[s], (g / TC_RAW_STAGES) & 1, limit);
        const uint8_t *st = raw + (size_t) s * TC_STAGE_BYTES;
        for (int b = 0; b < 4; b++, kb_g++) {
          const int buf = kb_g & 1;
          const uint4 nib = *reinterpret_cast<const uint4 *>(st + (size_t) b * TC_M * 16 + row * 16);
          const uint32_t ww[4] = {nib.x, nib.y, nib.z, nib.w};
          uint4 out4[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
            uint32_t h[4];
#if B200_TC_SKIP == 3
            h[0] = ww[c]; h[1] = ww[c] >> 1; h[2] = ww[c] >> 2; h[3] = ww[c] >> 3;
#else
#pragma unroll
            for (int i = 0; i < 4; i++) {
              // byte i of the word: low nibble = element 2i, high nibble = element 2i+1 of this chunk.  0x6400 | x is the
              // fp16 1024 + x: low half 1024 + lo, high half 1024 + 16 hi; (x * {1, 1/16}) + {-1032, -72} = {lo - 8, hi - 8}
              const uint32_t by = prmt(ww[c], 0u, 0x4040u | (uint32_t) i | ((uint32_t) i << 8));      // 0x00BB00BB
              const uint32_t t = (by & 0x00F0000Fu) | 0x64006400u;
              const __half2 r = __hfma2(*reinterpret_cast<const __half2 *>(&t), mulv, addv);
              h[i] = *reinterpret_cast<const uint32_t *>(&r);
            }
#endif
            out4[c] = make_uint4(h[0], h[1], h[2], h[3]);
          }
          if (kb_g >= 2) mbar_wait(&ab_empty[buf], ((kb_g >> 1) - 1) & 1, limit);
          uint8_t *A = abuf + buf * TC_A_BYTES + (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
          for (int c = 0; c < 4; c++) *reinterpret_cast<uint4 *>(A + c * (TC_M * 16)) = out4[c];
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ab_full[buf]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[s]);
      }
    }
  } else {
    // ===== epilogue warps: thread = row = TMEM lane 32*(warp % 4) + lane; warp group (warp / 4) takes TH tokens.  Several
    // warps per TMEM lane quarter, so the others compute while one waits for its tcgen05.ld =====
    constexpr int TH = TC_T / (TC_EPI_WARPS / 4);          // tokens per warp: 4
    static_assert(TH * 8 == 32, "one 32-column tcgen05.ld per block step and warp");
    const int wq = warp & 3, tg = warp >> 2;
    const int row = wq * 32 + lane;
    const uint32_t t_lane = ((uint32_t) (wq * 32)) << 16;
    uint32_t g = 0, kb_g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int mt = item / n_nt, nt = item % n_nt;
      u64 acc[TH][4];
#pragma unroll
      for (int t = 0; t < TH; t++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[t][j] = pack_f2(0.0f, 0.0f);
      for (int q = 0; q < nbq; q++, g++) {
        const int s = g % TC_RAW_STAGES;
        mbar_wait(&raw_full[s], (g / TC_RAW_STAGES) & 1, limit);
        const uint8_t *st = raw + (size_t) s * TC_STAGE_BYTES;
        const float4 dw4 = *reinterpret_cast<const float4 *>(st + (size_t) 4 * TC_M * 16 + row * 16);
        const float dw[4] = {dw4.x, dw4.y, dw4.z, dw4.w};
        const float *dxq = reinterpret_cast<const float *>(st + TC_QUAD_BYTES + TC_XH_BYTES) + tg * TH;
#pragma unroll
        for (int b = 0; b < 4; b++, kb_g++) {
          const int buf = kb_g & 1;
          const float4 dx4 = *reinterpret_cast<const float4 *>(dxq + b * TC_T);
          const float dxv[4] = {dx4.x, dx4.y, dx4.z, dx4.w};
          mbar_wait(&tm_full[buf], (kb_g >> 1) & 1, limit);
          tc_fence_after();
          uint32_t d[32];                                  // 32 columns = this warp's 4 tokens x 8 lanes
#if B200_TC_SKIP == 2
#pragma unroll
          for (int i = 0; i < 32; i++) d[i] = kb_g + i;
#else
          tmem_ld32(tmem_base + t_lane + (uint32_t) buf * TC_N + (uint32_t) tg * (TH * 8), d);
          tmem_ld_wait();
#endif
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tm_empty[buf]);      // the accumulator buffer is free as soon as it is in registers
#if B200_TC_SKIP == 1
          acc[0][0] ^= (u64) d[0] ^ (u64) d[31];
#else
#pragma unroll
          for (int t = 0; t < TH; t++) {
            const float sdx = __fmul_rn(dw[b], dxv[t]);                                              // _mm256_mul_ps(d0, d1), ggml.c:1431
            const u64 s2 = pack_f2(sdx, sdx);
#pragma unroll
            for (int j = 0; j < 4; j++)
              acc[t][j] = ffma2(s2, pack_i2((int) d[t * 8 + 2 * j], (int) d[t * 8 + 2 * j + 1]), acc[t][j]);   // _mm256_fmadd_ps, ggml.c:1457
          }
#endif
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[s]);
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
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace b200
