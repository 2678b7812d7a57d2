// Row loop of the Q4_0 mat-vec on the warp-level tensor path (mma.sync m16n8k32 u8 x s8 -> s32, SASS IMMA.16832.U8.S8).
//
// What has to be computed (ggml_vec_dot_q4_0, AVX2 branch, ggml.c:1415-1466): per (row, block) EIGHT integer sums, one
// per AVX accumulator lane l = elements {2l, 2l+1, 16+2l, 17+2l} of the block, each folded into its own f32 accumulator
// with acc[l] = fma(d_w*d_x, (float) isum_l, acc[l]).  In ggml's own byte order lane l is bytes l and l+8 of the block's
// 16 nibble bytes (low nibble = element 2j, high nibble = element 2j+1 of byte j), so the eight lanes of a block are the
// eight COLUMNS of a block-diagonal activation operand and one MMA delivers the 8 lane sums of 16 rows at once:
//
//   A (16 rows x 32 bytes) = [ the block's 16 raw nibble bytes (u8 value 16*hi + lo) | the same bytes & 0xF0 (16*hi) ]
//   B (32 x 8), column l   = 16*x_lo at k = l, l+8 and (x_hi - 16*x_lo) at k = 16+l, 24+l, zero elsewhere
//   D[row][l] = 16 * sum(q_i * x_i over lane l)           (q = raw nibble 0..15, x = quantized activation -7..7)
//
// -- (16 hi + lo) * 16 x_lo + 16 hi * (x_hi - 16 x_lo) = 16 (lo x_lo + hi x_hi): ONE LOP3 per 8 weights replaces the
// shift + 2 LOP3 + 2 dp4a of the CUDA-core loop, the raw ggml bytes are the operand (no nibble permutation at load), and
// all operand bytes fit s8 (|16 x_lo| <= 112, |x_hi - 16 x_lo| <= 119).  The reference's "- 8" zero point is linear too:
// sum((q-8) x) = sum(q x) - 8 S_l with S_l = the lane's activation sum, folded into the int->float step: the accumulator
// is seeded with 0x4B400000, so D's bit pattern IS the float 12582912 + 16 isum', and ONE fma.rn.f32x2 by 1/16 plus
// c_l = -786432 - 8 S_l (a per-(block, lane) constant written by the activation quantizer) yields the exact (float) isum
// for two lanes.  The second fma.rn.f32x2 is the reference's _mm256_fmadd_ps.  Integer part exact, float part the same
// operations in the same order as the reference: results are bit-identical to the CUDA-core loop and to ggml.
//
// Thread mapping (lane = 4 g + t): the thread owns lanes 2t, 2t+1 of rows g and g+8 of its warp's 16-row tile -- the
// same (row, lane pair) ownership as the LP = 1 CUDA-core loop, so the final horizontal sum is unchanged.
//
// Weight stream, per quad of 4 blocks (R = rows of the CTA): [block b < 4][row r < R][16 raw nibble bytes] (R * 64 B),
// then [row r][4] f32 scales (R * 16 B).  The 16 bytes of 8 consecutive rows are 128 contiguous bytes, so ONE
// ldmatrix.x2 (SASS LDSM) delivers a0 / a1 of a 16-row tile straight into the MMA fragment registers, conflict-free.
#pragma once
#include "ptx.cuh"

namespace b200 {

// Quantized activation vector in shared memory, MMA-operand form.
struct ActSmem {
  uint32_t *xw;   // [8 lanes l][nbw]   word {16 x[2l], x[2l+1] - 16 x[2l], 16 x[16+2l], x[17+2l] - 16 x[16+2l]} of block b (s8 bytes)
  float2 *cs;     // [4 t][nbc]         {c_{2t}, c_{2t+1}} of block b, c_l = -786432 - 8 S_l
  float *dxs;     // [nb]               block scale d
  int nbw, nbc;   // plane strides (bank-conflict padding: nbw % 32 == 4, nbc % 16 == 2)
};
__host__ __device__ __forceinline__ int act_nbw(int nb_max) { return nb_max + ((36 - (nb_max & 31)) & 31); }
__host__ __device__ __forceinline__ int act_nbc(int nb_max) { return nb_max + ((18 - (nb_max & 15)) & 15); }
__host__ __device__ __forceinline__ size_t act_smem_bytes(int nb_max) {
  return (size_t) 8 * act_nbw(nb_max) * 4 + (size_t) 4 * act_nbc(nb_max) * 8 + (size_t) nb_max * 4;
}
// carve: xw | cs | dxs  (base 16-byte aligned; every plane stays 16-byte aligned because the strides are even)
__device__ __forceinline__ ActSmem act_carve(uint8_t *base, int nb_max) {
  ActSmem a;
  a.nbw = act_nbw(nb_max);
  a.nbc = act_nbc(nb_max);
  a.xw = reinterpret_cast<uint32_t *>(base);
  a.cs = reinterpret_cast<float2 *>(a.xw + (size_t) 8 * a.nbw);
  a.dxs = reinterpret_cast<float *>(a.cs + (size_t) 4 * a.nbc);
  return a;
}

// {u, v} bytes of one (x_lo, x_hi) element pair: u = 16 x_lo, v = x_hi - 16 x_lo
__device__ __forceinline__ uint32_t uv_pair(int x_lo, int x_hi) {
  const int u = 16 * x_lo, v = x_hi - u;
  return (uint32_t) (u & 0xff) | ((uint32_t) (v & 0xff) << 8);
}
// c_l = -786432 - 8 S_l from the lane's operand word: 16 S_l = 17 u0 + 16 v0 + 17 u1 + 16 v1 (x_lo = u/16, x_hi = v + u)
__device__ __forceinline__ float lane_const(uint32_t w) {
  const int s16 = dp4a_ss((int) w, 0x10111011, 0);
  return fmaf((float) s16, -0.5f, -786432.0f);       // exact: |8 S| <= 224
}

// quantize_row_q4_0 (AVX2 branch, ggml.c:456-523) for one 32-block handled by 4 consecutive lanes (8 values each).
// s = lane-quad index 0..3 of a live block, >= 4 for a padding thread (takes part in the shuffles, stores nothing).
__device__ __forceinline__ void quantize_block_4t_imma(const float v[8], int b, int s, const ActSmem &x) {
  float amax = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
  const float d = __fdiv_rn(amax, 7.0f);
  const float id = (amax != 0.0f) ? __fdiv_rn(7.0f, amax) : 0.0f;
  int q[8];
#pragma unroll
  for (int i = 0; i < 8; i++) q[i] = __float2int_rn(__fmul_rn(v[i], id));      // round-to-nearest-even = stored nibble - 8
  // this thread holds elements 8s..8s+7 = half (s>>1) of lanes 4*(s&1)+j, j = 0..3 (pairs q[2j], q[2j+1])
  const uint32_t m01 = uv_pair(q[0], q[1]) | (uv_pair(q[2], q[3]) << 16);
  const uint32_t m23 = uv_pair(q[4], q[5]) | (uv_pair(q[6], q[7]) << 16);
  const uint32_t o01 = __shfl_xor_sync(0xffffffffu, m01, 2);
  const uint32_t o23 = __shfl_xor_sync(0xffffffffu, m23, 2);
  if (s < 2) {
    // lanes 4s+j: low half-word = my half (elements 2l, 2l+1), high half-word = the partner's (16+2l, 17+2l)
    const uint32_t w0 = (m01 & 0xffffu) | (o01 << 16);
    const uint32_t w1 = (m01 >> 16) | (o01 & 0xffff0000u);
    const uint32_t w2 = (m23 & 0xffffu) | (o23 << 16);
    const uint32_t w3 = (m23 >> 16) | (o23 & 0xffff0000u);
    uint32_t *pw = x.xw + (size_t) (4 * s) * x.nbw + b;
    pw[0] = w0; pw[x.nbw] = w1; pw[2 * x.nbw] = w2; pw[3 * x.nbw] = w3;
    x.cs[(size_t) (2 * s) * x.nbc + b] = make_float2(lane_const(w0), lane_const(w1));
    x.cs[(size_t) (2 * s + 1) * x.nbc + b] = make_float2(lane_const(w2), lane_const(w3));
    if (s == 0) x.dxs[b] = d;
  }
}

// the same for a whole block held by one thread (per-matrix kernels)
__device__ __forceinline__ void quantize_block_full_imma(const float v[32], int b, const ActSmem &x) {
  float amax = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(v[i]));
  const float d = __fdiv_rn(amax, 7.0f);
  const float id = (amax != 0.0f) ? __fdiv_rn(7.0f, amax) : 0.0f;
  int q[32];
#pragma unroll
  for (int i = 0; i < 32; i++) q[i] = __float2int_rn(__fmul_rn(v[i], id));
  float c[8];
#pragma unroll
  for (int l = 0; l < 8; l++) {
    const uint32_t w = uv_pair(q[2 * l], q[2 * l + 1]) | (uv_pair(q[16 + 2 * l], q[17 + 2 * l]) << 16);
    x.xw[(size_t) l * x.nbw + b] = w;
    c[l] = lane_const(w);
  }
#pragma unroll
  for (int t = 0; t < 4; t++) x.cs[(size_t) t * x.nbc + b] = make_float2(c[2 * t], c[2 * t + 1]);
  x.dxs[b] = d;
}

// padding block (partial last quad / unused tail): contributes fma(0, 0, acc) = acc
__device__ __forceinline__ void act_zero_block(int b, const ActSmem &x, int part /* 0..3: which quarter of the stores */) {
  x.xw[(size_t) (2 * part) * x.nbw + b] = 0u;
  x.xw[(size_t) (2 * part + 1) * x.nbw + b] = 0u;
  x.cs[(size_t) part * x.nbc + b] = make_float2(-786432.0f, -786432.0f);
  if (part == 0) x.dxs[b] = 0.0f;
}

// ---- registers of one quad (4 blocks) for a warp that owns NT tiles of RW (8 or 16) rows --------------------------------
template <int RW, int NT>
struct QuadI {
  uint32_t a[NT][4][RW / 8];   // nibble words: [tile][block] {row g, row g+8} = MMA fragment a0 / a1 (ldmatrix)
  float4 sc[NT][RW / 8];       // weight scales of the 4 blocks, rows g (and g+8) of each tile
  float4 dx;               // activation scales of the 4 blocks
  uint4 w;                 // operand words of lane g, 4 blocks
  float4 c[2];             // {c_2t, c_2t+1} of blocks 0-1 | 2-3
};

template <int RW, int NT>
struct QuadIPtrs {
  uint32_t pa[NT];                 // shared-space address this lane feeds to ldmatrix (row lane & (RW-1) of the tile, block 0)
  const uint8_t *ps[NT][RW / 8];   // row's scales
  int bstride;                     // bytes between the blocks of a quad = R * 16
  const uint32_t *pw;              // xw plane g, first block of the quad
  const float2 *pc;                // cs plane t
  const float *pd;                 // dxs
  int qstride;                     // bytes per quad = R * 80
};

template <int RW, int NT>
__device__ __forceinline__ void quadi_load(QuadI<RW, NT> &q, const QuadIPtrs<RW, NT> &p, int qoff) {
#pragma unroll
  for (int t = 0; t < NT; t++) {
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const uint32_t addr = p.pa[t] + (uint32_t) (qoff * p.qstride + b * p.bstride);
      if constexpr (RW == 16) ldmatrix_x2(q.a[t][b][0], q.a[t][b][1], addr);
      else ldmatrix_x1(q.a[t][b][0], addr);
    }
#pragma unroll
    for (int h = 0; h < RW / 8; h++) q.sc[t][h] = *reinterpret_cast<const float4 *>(p.ps[t][h] + qoff * p.qstride);
  }
  q.dx = *reinterpret_cast<const float4 *>(p.pd + qoff * 4);
  q.w = *reinterpret_cast<const uint4 *>(p.pw + qoff * 4);
  q.c[0] = *reinterpret_cast<const float4 *>(p.pc + qoff * 4);
  q.c[1] = *reinterpret_cast<const float4 *>(p.pc + qoff * 4 + 2);
}

__device__ __forceinline__ uint32_t comp4(const uint4 &v, int b) { return b == 0 ? v.x : b == 1 ? v.y : b == 2 ? v.z : v.w; }
__device__ __forceinline__ float comp4(const float4 &v, int b) { return b == 0 ? v.x : b == 1 ? v.y : b == 2 ? v.z : v.w; }

// acc[t][h] (lanes 2t', 2t'+1 of row g + 8h of tile t) += the 4 blocks of the quad, in order (ggml.c:1431-1457)
template <int RW, int NT>
__device__ __forceinline__ void quadi_math(const QuadI<RW, NT> &q, uint32_t sel0, uint32_t sel1, u64 (&acc)[NT][RW / 8]) {
  const u64 cvt_mul = pack_f2(0.0625f, 0.0625f);
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const uint32_t wb = comp4(q.w, b);
    const uint32_t b0 = prmt(wb, 0u, sel0);          // B[4t..4t+3][g]:    16 x_lo of lane g where its byte falls into this thread's k range
    const uint32_t b1 = prmt(wb, 0u, sel1);          // B[16+4t..][g]:     x_hi - 16 x_lo
    const float4 cc = q.c[b >> 1];
    const u64 cb = (b & 1) ? pack_f2(cc.z, cc.w) : pack_f2(cc.x, cc.y);
    const float dxb = comp4(q.dx, b);
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const uint32_t a0 = q.a[t][b][0];
      const uint32_t a1 = RW == 16 ? q.a[t][b][RW / 8 - 1] : 0u;
      int d[4];
      mma_u8s8_16832(d, a0, a1, a0 & 0xF0F0F0F0u, a1 & 0xF0F0F0F0u, b0, b1, 0x4B400000);
      {
        const u64 f = ffma2(pack_i2(d[0], d[1]), cvt_mul, cb);                   // exact (float) isum of lanes 2t, 2t+1, row g
        const float sdx = __fmul_rn(comp4(q.sc[t][0], b), dxb);                  // _mm256_mul_ps(d0, d1), ggml.c:1431
        acc[t][0] = ffma2(pack_f2(sdx, sdx), f, acc[t][0]);                      // _mm256_fmadd_ps, ggml.c:1457
      }
      if (RW == 16) {
        const u64 f = ffma2(pack_i2(d[2], d[3]), cvt_mul, cb);                   // row g + 8
        const float sdx = __fmul_rn(comp4(q.sc[t][RW / 8 - 1], b), dxb);
        acc[t][RW / 8 - 1] = ffma2(pack_f2(sdx, sdx), f, acc[t][RW / 8 - 1]);
      }
    }
  }
}

// rows of a warp's tiles: row[i][h] = row g + 8h of tile tile0 + i (what this thread accumulates), lrow[i] = the row whose
// 16-byte address this lane hands to ldmatrix (lane & (RW-1)); both clamped to the CTA's last row (duplicates are harmless)
template <int RW, int NT>
__device__ __forceinline__ void imma_tile_rows(int tile0, int lane, int R, int (&row)[NT][RW / 8], int (&lrow)[NT]) {
  const int g = lane >> 2;
#pragma unroll
  for (int i = 0; i < NT; i++) {
#pragma unroll
    for (int h = 0; h < RW / 8; h++) row[i][h] = min((tile0 + i) * RW + g + 8 * h, R - 1);
    lrow[i] = min((tile0 + i) * RW + (lane & (RW - 1)), R - 1);
  }
}

// per-thread byte selectors of the block-diagonal activation operand (see the header): lane g's bytes {u0, v0, u1, v1}
// land in byte (g & 3) of b0 / b1 for the two threads t with (t & 1) == (g >> 2) -- t < 2 takes the first half (u0, v0),
// t >= 2 the second (u1, v1); everything else is zero (PRMT index 4 = byte 0 of the zero operand).
__device__ __forceinline__ void imma_selectors(int lane, uint32_t &sel0, uint32_t &sel1) {
  const int g = lane >> 2, t = lane & 3;
  sel0 = sel1 = 0x4444u;
  if ((t & 1) == (g >> 2)) {
    const int sh = 4 * (g & 3), half = t >> 1;
    sel0 = (0x4444u & ~(0xFu << sh)) | ((uint32_t) (2 * half) << sh);
    sel1 = (0x4444u & ~(0xFu << sh)) | ((uint32_t) (2 * half + 1) << sh);
  }
}

// All quads of one chunk (stage `st`, cqk quads, R rows in the CTA).  row[t][h] = this thread's (clamped) rows,
// lrow[t] = the (clamped) row whose address this lane supplies to ldmatrix.
template <int RW, int NT>
__device__ __forceinline__ void gemv_chunk_imma(const uint8_t *st, int cqk, int R, const int (&row)[NT][RW / 8], const int (&lrow)[NT], int lane,
                                                const ActSmem &x, int b0, uint32_t sel0, uint32_t sel1, u64 (&acc)[NT][RW / 8]) {
  const int g = lane >> 2, t = lane & 3;
  QuadIPtrs<RW, NT> p;
  p.qstride = R * 80;
  p.bstride = R * 16;
#pragma unroll
  for (int i = 0; i < NT; i++) {
    p.pa[i] = smem_u32(st) + (uint32_t) lrow[i] * 16u;
#pragma unroll
    for (int h = 0; h < RW / 8; h++) p.ps[i][h] = st + R * 64 + row[i][h] * 16;
  }
  p.pw = x.xw + (size_t) g * x.nbw + b0;
  p.pc = x.cs + (size_t) t * x.nbc + b0;
  p.pd = x.dxs + b0;
  int q = 0;
  if constexpr (NT == 1) {
    // two register sets: the loads of quad q+1 are issued under the math of quad q.  The look-ahead may read one quad
    // past the end of the chunk -- still inside this CTA's shared memory, and the values are never used.
    QuadI<RW, NT> ra, rb;
    quadi_load<RW, NT>(ra, p, 0);
    for (; q + 2 <= cqk; q += 2) {
      quadi_load<RW, NT>(rb, p, 1);
      quadi_math<RW, NT>(ra, sel0, sel1, acc);
      quadi_load<RW, NT>(ra, p, 2);
      quadi_math<RW, NT>(rb, sel0, sel1, acc);
      p.pa[0] += 2 * p.qstride;
#pragma unroll
      for (int h = 0; h < RW / 8; h++) p.ps[0][h] += 2 * p.qstride;
      p.pw += 8; p.pc += 8; p.pd += 8;
    }
    if (q < cqk) quadi_math<RW, NT>(ra, sel0, sel1, acc);
  } else {
    for (; q < cqk; q++) {
      QuadI<RW, NT> rq;
      quadi_load<RW, NT>(rq, p, 0);
      quadi_math<RW, NT>(rq, sel0, sel1, acc);
#pragma unroll
      for (int i = 0; i < NT; i++) {
        p.pa[i] += p.qstride;
#pragma unroll
        for (int h = 0; h < RW / 8; h++) p.ps[i][h] += p.qstride;
      }
      p.pw += 4; p.pc += 4; p.pd += 4;
    }
  }
}

}  // namespace b200
