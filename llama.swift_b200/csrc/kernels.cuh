// Hand-written sm_100a kernels for the llama.swift decode hot path (llama_eval, PO.mm:510-735).
//
// Arithmetic contract: every kernel reproduces the reference's x86/AVX2 build *operation for operation*
// (same rounding points, same accumulation order), so results are bit-identical to the oracle except for the
// double-precision LayerNorm sums (tree order instead of index order: <= 1 ulp of a double before rounding to f32).
// Build with -fmad=false: an FMA appears only where the reference has one (explicit fmaf / fma.rn.f32x2).
//
//   q4_gemv_kernel ....... ggml_compute_forward_mul_mat_q4_0_f32 (ggml.c:5987-6285) = quantize_row_q4_0
//                          (AVX2, ggml.c:456-523) + ggml_vec_dot_q4_0 (AVX2, ggml.c:1415-1466), N = 1, with the
//                          neighbouring graph nodes fused in as prologue/epilogue (norm+mul, rope+cpy, add, silu*mul)
//   attn_kernel .......... mul_mat_f32 K.Q (ggml.c:5579-5618 + vec_dot_f32 1223-1258), scale (6649), soft_max (6982),
//                          V.P transposed branch (ggml.c:5619-5665 + FINALIZE 5553-5577 + vec_mad_f32 1683-1712)
//   embed_kernel ......... get_rows_q4_0 -> dequantize_row_q4_0 (ggml.c:6760-6785, 651-684)
//   argmax_kernel ........ greedy next-token pick for the device-resident decode loop (bench / teacher forcing)
//   repack_q4_0_kernel ... load-time re-layout of ggml's 20-byte AoS blocks into the tile-major stream (same bytes)
#pragma once
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "ptx.cuh"

#ifndef B200_IMMA
#define B200_IMMA 0         // 1 = row loops on the warp-level tensor path (rowloop_imma.cuh): EXPERIMENT, measured slower
#endif
#include "rowloop_imma.cuh"

namespace b200 {

namespace cg = cooperative_groups;

// per-step scalars live in device memory so one CUDA graph can be replayed for every token
struct StepParams {
  int token;    // id of the token being evaluated
  int pos;      // its position = n_past + i
  int p_part;   // n_past + N of the enclosing llama_eval call: the reference partitions V.P columns by it (ggml.c:5628)
  int step;     // running index for the greedy loop's token log
  int forced;   // != 0: the device-resident loop feeds forced_tokens[step] next (teacher forcing) instead of the arg-max
};

// ---- row partition of a (fused) matrix over the grid: granules of 4 rows, contiguous, balanced ------------------
struct RowPart { int row0; int R; };
__host__ __device__ __forceinline__ RowPart row_part(int g_total, int n_cta, int c) {
  const int q = g_total / n_cta, rem = g_total % n_cta;
  RowPart p;
  p.R = 4 * (q + (c < rem ? 1 : 0));
  p.row0 = 4 * (c * q + (c < rem ? c : rem));
  return p;
}
__host__ __device__ __forceinline__ int cta_of_granule(int g_total, int n_cta, int g) {
  const int q = g_total / n_cta, rem = g_total % n_cta;
  if (g < rem * (q + 1)) return g / (q + 1);
  return rem + (g - rem * (q + 1)) / q;
}

// Tile-major weight stream ("quad-major").  CTA c owns rows [row0, row0+R) and its bytes are contiguous at
// row0 * nbq * 80, nbq = ceil(nb / 4) quads of 4 blocks per row (same 20 bytes per 32 weights as ggml, ggml.c:408; a
// partial last quad is padded with zero-scale blocks, which are exact no-ops).  Chunk k (cq quads, the last one shorter)
// is one contiguous piece, so one pipeline stage is ONE cp.async.bulk.  Inside a quad:
//
//     [lane-pair slot j < LP][row r < R][thread-of-row t < 4/LP][block b < 4]  32-bit nibble words   (R * 64 bytes)
//     [row r][block b]                                                         f32 block scales     (R * 16 bytes)
//
// where LP = lane pairs per thread of the matrix's plan and word p = t*LP + j of a block holds AVX2 accumulator lane 2p
// in its low nibbles and lane 2p+1 in its high nibbles (lane l = elements 2l, 2l+1, 16+2l, 17+2l of the block,
// ggml.c:1443-1452).  The thread that owns lane pairs t*LP..t*LP+LP-1 of rows g, g+G, ... (RPT rows per thread, G = R/RPT)
// gets its words of FOUR blocks with one 16-byte shared-memory load per (row, j) and the four scales with another --
// consecutive threads read consecutive 16 bytes (conflict-free), all block offsets are immediates.
//
// The row loop is bound by shared-memory wavefronts and FMA-pipe issue (IDP.4A, FFMA2, IMAD share that pipe), not by
// HBM, so bytes and instructions per (row, block) are what matters:
//   * the quantized activation is stored as 8 bytes per (block, lane pair): the signed bytes of lanes 2p and 2p+1,
//     plane-major [p][block] so one 16-byte load covers two blocks, and it is loaded once per thread for all its rows;
//   * no zero-point seeds: a nibble q is turned into the signed byte 16*(q-8) in place -- `(w & 0xf0f0f0f0) ^ 0x80808080`
//     for the high nibbles, the same on `w << 4` for the low ones (offset-binary -> two's complement is one XOR of the
//     top bit) -- so dp4a.s32.s32 gives 16 * isum exactly; seeded with 0x4B400000 the result IS the bit pattern of the
//     float 12582912 + 16*isum, and one fma.rn.f32x2 by (1/16, 1/16) plus (-786432, -786432) yields exact (float) isum.
__device__ __forceinline__ int lane_elem(int lane, int k) { return (k < 2 ? 0 : 16) + 2 * lane + (k & 1); }

#ifndef B200_LOP3
#define B200_LOP3 1         // nibble -> signed byte with one LOP3 (see and_xor in ptx.cuh)
#endif
#ifndef B200_PIPE
#define B200_PIPE 1         // small register footprints: two quads in registers, loads of quad q+1 issued under the math of quad q
#endif

#ifndef B200_SHF
#define B200_SHF 0          // 1 = the nibble shift as a funnel shift (SASS SHF, ALU pipe) instead of ptxas' IMAD.SHL (FMA pipe, the busy one)
#endif
__device__ __forceinline__ uint32_t shl4(uint32_t w) {
#if B200_SHF
  uint32_t d; asm("shf.l.clamp.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(0u), "r"(w), "r"(4u)); return d;
#else
  return w << 4;
#endif
}

template <int LP, int RPT>
struct QuadRegs {
  uint4 w[RPT][LP];   // nibble words of 4 blocks, per row and lane-pair slot
  float4 sc[RPT];     // weight scales of the 4 blocks, per row
  float4 dx;          // activation scales of the 4 blocks
  uint4 x[LP][2];     // activation bytes: [slot][blocks 0-1 | 2-3] = {lane 2p of b, lane 2p+1 of b, same for b+1}
};

// per-thread addressing of one matrix phase
struct QuadPtrs {
  const uint8_t *pw;   // this thread's first nibble word of the quad (row g, slot 0)
  const uint8_t *ps;   // row g's scales
  const uint8_t *px;   // activation bytes of the quad's first block, plane t*LP
  const float *pd;     // activation scales of the quad's first block
  int jstride;         // bytes between lane-pair slots           = R * UPR * 16
  int rstride;         // bytes between this thread's rows        = G * UPR * 16
  int sstride;         // bytes between this thread's rows' scales = G * 16
  int xstride;         // bytes between activation planes         = nbp * 8
  int qstride;         // bytes per quad                          = R * 80
};

template <int LP, int RPT>
__device__ __forceinline__ void quad_load(QuadRegs<LP, RPT> &q, const QuadPtrs &p, int qoff /* quads ahead */) {
  const uint8_t *pw = p.pw + qoff * p.qstride, *ps = p.ps + qoff * p.qstride;
#pragma unroll
  for (int i = 0; i < RPT; i++) {
#pragma unroll
    for (int j = 0; j < LP; j++) q.w[i][j] = *reinterpret_cast<const uint4 *>(pw + j * p.jstride + i * p.rstride);
    q.sc[i] = *reinterpret_cast<const float4 *>(ps + i * p.sstride);
  }
  q.dx = *reinterpret_cast<const float4 *>(p.pd + qoff * 4);
#pragma unroll
  for (int j = 0; j < LP; j++) {
    q.x[j][0] = *reinterpret_cast<const uint4 *>(p.px + j * p.xstride + qoff * 32);
    q.x[j][1] = *reinterpret_cast<const uint4 *>(p.px + j * p.xstride + qoff * 32 + 16);
  }
}

// acc[i][j] = fma(d_w * d_x, (float) isum, acc[i][j]) for the 4 blocks of the quad in order (ggml.c:1431-1457)
template <int LP, int RPT>
__device__ __forceinline__ void quad_math(const QuadRegs<LP, RPT> &q, u64 (&acc)[RPT][LP]) {
  const u64 cvt_mul = pack_f2(0.0625f, 0.0625f);
  const u64 cvt_sub = pack_f2(-786432.0f, -786432.0f);
  const float dx[4] = {q.dx.x, q.dx.y, q.dx.z, q.dx.w};
#pragma unroll
  for (int b = 0; b < 4; b++) {
#pragma unroll
    for (int i = 0; i < RPT; i++) {
      const float sc = b == 0 ? q.sc[i].x : b == 1 ? q.sc[i].y : b == 2 ? q.sc[i].z : q.sc[i].w;
      const float sdx = __fmul_rn(sc, dx[b]);                                    // _mm256_mul_ps(d0, d1), ggml.c:1431
#pragma unroll
      for (int j = 0; j < LP; j++) {
        const uint32_t w = b == 0 ? q.w[i][j].x : b == 1 ? q.w[i][j].y : b == 2 ? q.w[i][j].z : q.w[i][j].w;
        const uint4 xv = q.x[j][b >> 1];
        const int xlo = (int) ((b & 1) ? xv.z : xv.x), xhi = (int) ((b & 1) ? xv.w : xv.y);
#if B200_LOP3
        const int a_hi = (int) and_xor(w, 0xF0F0F0F0u, 0x80808080u);             // signed bytes 16*(q-8), lane 2p+1
        const int a_lo = (int) and_xor(shl4(w), 0xF0F0F0F0u, 0x80808080u);       // signed bytes 16*(q-8), lane 2p
#else
        const int a_hi = (int) ((w & 0xF0F0F0F0u) ^ 0x80808080u);                // signed bytes 16*(q-8), lane 2p+1
        const int a_lo = (int) (((w << 4) & 0xF0F0F0F0u) ^ 0x80808080u);         // signed bytes 16*(q-8), lane 2p
#endif
        const int ia = dp4a_ss(a_lo, xlo, 0x4B400000);                           // float bits of 12582912 + 16*isum(lane 2p)
        const int ib = dp4a_ss(a_hi, xhi, 0x4B400000);                           // float bits of 12582912 + 16*isum(lane 2p+1)
        const u64 f = ffma2(pack_i2(ia, ib), cvt_mul, cvt_sub);                  // exact (float)isum for both lanes
        acc[i][j] = ffma2(pack_f2(sdx, sdx), f, acc[i][j]);                      // _mm256_fmadd_ps(scale, p, acc), ggml.c:1457
      }
    }
  }
}

// All quads of one chunk (stage `st`, cqk quads).  Thread (g, t) owns lane pairs t*LP .. t*LP+LP-1 of rows g + i*G.
// xq8 = quantized activation planes [4][nbp (plane stride)] of 8 bytes, dxs = activation scales; b0 = first block of the chunk.
template <int LP, int RPT>
__device__ __forceinline__ void gemv_chunk(const uint8_t *st, int cqk, int R, int g, int t, const uint2 *xq8, const float *dxs,
                                           int nbp, int b0, u64 (&acc)[RPT][LP]) {
  constexpr int UPR = 4 / LP;
  const int G = R / RPT;
  QuadPtrs p;
  p.jstride = R * UPR * 16;
  p.rstride = G * UPR * 16;
  p.sstride = G * 16;
  p.xstride = nbp * 8;
  p.qstride = R * 80;
  p.pw = st + (g * UPR + t) * 16;
  p.ps = st + R * 64 + g * 16;
  p.px = reinterpret_cast<const uint8_t *>(xq8 + (size_t) (t * LP) * nbp + b0);
  p.pd = dxs + b0;
  int q = 0;
#if B200_PIPE
  if constexpr (LP == 1 && RPT <= 2) {
    // Two register sets; the look-ahead may read one quad past the end of the chunk -- still inside this CTA's shared
    // memory, and the values are never used.
    QuadRegs<LP, RPT> ra, rb;
    quad_load<LP, RPT>(ra, p, 0);
    for (; q + 2 <= cqk; q += 2) {
      quad_load<LP, RPT>(rb, p, 1);
      quad_math<LP, RPT>(ra, acc);
      quad_load<LP, RPT>(ra, p, 2);
      quad_math<LP, RPT>(rb, acc);
      p.pw += 2 * p.qstride; p.ps += 2 * p.qstride; p.px += 64; p.pd += 8;
    }
    if (q < cqk) quad_math<LP, RPT>(ra, acc);
    return;
  }
#endif
  for (; q < cqk; q++) {
    QuadRegs<LP, RPT> rq;
    quad_load<LP, RPT>(rq, p, 0);
    quad_math<LP, RPT>(rq, acc);
    p.pw += p.qstride; p.ps += p.qstride; p.px += 32; p.pd += 4;
  }
}

// horizontal sum of one row exactly as ggml.c:1461-1466: (acc[k]+acc[k+4]) k<4, then (r0+r2)+(r1+r3); the LP lane pairs
// of the row are in this thread, the others in the neighbouring 4/LP - 1 threads
template <int LP>
__device__ __forceinline__ float row_hsum(const u64 (&acc)[LP]) {
  float lane[2 * LP];
#pragma unroll
  for (int j = 0; j < LP; j++) unpack_f2(acc[j], lane[2 * j], lane[2 * j + 1]);
  if constexpr (LP == 4) {
    const float r0 = __fadd_rn(lane[4], lane[0]), r1 = __fadd_rn(lane[5], lane[1]);
    const float r2 = __fadd_rn(lane[6], lane[2]), r3 = __fadd_rn(lane[7], lane[3]);
    return __fadd_rn(__fadd_rn(r0, r2), __fadd_rn(r1, r3));
  } else if constexpr (LP == 2) {
    float rr[4];
#pragma unroll
    for (int i = 0; i < 4; i++) rr[i] = __fadd_rn(lane[i], __shfl_xor_sync(0xffffffffu, lane[i], 1));
    return __fadd_rn(__fadd_rn(rr[0], rr[2]), __fadd_rn(rr[1], rr[3]));
  } else {
    const float t0 = __fadd_rn(lane[0], __shfl_xor_sync(0xffffffffu, lane[0], 2));
    const float t1 = __fadd_rn(lane[1], __shfl_xor_sync(0xffffffffu, lane[1], 2));
    const float s0 = __fadd_rn(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
    const float s1 = __fadd_rn(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
    return __fadd_rn(s0, s1);
  }
}

enum GemvPrologue { PRO_PLAIN = 0, PRO_NORM = 1 };
enum GemvEpilogue { EPI_STORE = 0, EPI_RESID = 1, EPI_QKV = 2, EPI_SILU_MUL = 3 };

struct GemvArgs {
  const uint8_t *w;       // tile-major stream
  int M;                  // valid (unpadded) fused rows
  int g_total;            // padded rows / 4
  int nb;                 // blocks per row = K/32
  int cb;                 // blocks per chunk (a multiple of 4: whole quads)
  int n_stages;
  int stage_bytes;
  int rmax;
  const float *x;         // input vector [K] (f32)
  const float *norm_w;    // PRO_NORM: LayerNorm weight
  float *out;             // EPI_STORE / EPI_RESID / EPI_SILU_MUL destination
  const float *resid;     // EPI_RESID
  // EPI_QKV
  float *q_out;
  float *k_layer;         // this layer's K cache [n_ctx][n_embd]
  float *v_layer;
  const double2 *rope;    // [n_ctx][head_dim/2] (cos, sin) built on the host with the reference's libm expressions
  const StepParams *sp;
  int n_embd;
  int head_dim;
  const uint16_t *silu_table;   // EPI_SILU_MUL
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum over the compute threads (nt of them, warps 0..nt/32-1) in a fixed order; every thread gets the result.
__device__ __forceinline__ double block_sum_d(double v, double *red, int tid, int nt) {
  v = warp_sum_d(v);
  const int nw = nt >> 5;
  named_bar_sync(1, nt);            // protect red[] reuse
  if ((tid & 31) == 0) red[tid >> 5] = v;
  named_bar_sync(1, nt);
  double s = red[0];
  for (int i = 1; i < nw; i++) s = __dadd_rn(s, red[i]);
  return s;
}

// ---- epilogue shared by the Q4_0 and Q4_1 mat-vec kernels: the graph nodes that consume the row results ------------
template <int EPI>
__device__ __forceinline__ void gemv_epilogue(const GemvArgs &a, const RowPart rp, const float *rowres, int tid, int nt) {
  const int R = rp.R;
  if (EPI == EPI_STORE || EPI == EPI_RESID) {
    for (int i = tid; i < R; i += nt) {
      const int g = rp.row0 + i;
      if (g < a.M) a.out[g] = (EPI == EPI_RESID) ? __fadd_rn(rowres[i], a.resid[g]) : rowres[i];   // ggml_add, PO.mm:654,687
    }
  } else if (EPI == EPI_SILU_MUL) {
    // fused rows 2i = w1 row i, 2i+1 = w3 row i: silu(w1 x) * (w3 x), PO.mm:678-680; silu via the fp16 table (ggml.c:1955-1963)
    for (int i = tid; i < R / 2; i += nt) {
      const int g = rp.row0 / 2 + i;
      if (2 * g < a.M) {
        const uint16_t hx = __half_as_ushort(__float2half_rn(rowres[2 * i]));
        const float sv = __half2float(__ushort_as_half(a.silu_table[hx]));
        a.out[g] = __fmul_rn(sv, rowres[2 * i + 1]);
      }
    }
  } else {
    // fused rows [0,E) = wq, [E,2E) = wk, [2E,3E) = wv.  RoPE (ggml.c:7110-7127) on Q and K pairs in double, K/V
    // stored to the cache row of this position (PO.mm:585-611: cpy then in-place rope == rope then store).
    const int E = a.n_embd;
    const int pos = a.sp->pos;
    for (int i = tid; i < R / 2; i += nt) {
      const int g = rp.row0 + 2 * i;
      if (g >= a.M) continue;
      const int which = g / E, col = g - which * E;
      float y0 = rowres[2 * i], y1 = rowres[2 * i + 1];
      if (which < 2) {
        const double2 cs = a.rope[(size_t) pos * (a.head_dim / 2) + (col % a.head_dim) / 2];
        const double x0 = y0, x1 = y1;
        y0 = (float) __dsub_rn(__dmul_rn(x0, cs.x), __dmul_rn(x1, cs.y));
        y1 = (float) __dadd_rn(__dmul_rn(x0, cs.y), __dmul_rn(x1, cs.x));
      }
      float *dst = which == 0 ? a.q_out + col
                 : which == 1 ? a.k_layer + (size_t) pos * E + col
                              : a.v_layer + (size_t) pos * E + col;
      dst[0] = y0;
      dst[1] = y1;
    }
  }
}

template <int LP, int PRO, int EPI>
__global__ void __launch_bounds__(544, 1) q4_gemv_kernel(const GemvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int UPR = 4 / LP;                       // threads per row
  const int tid = threadIdx.x;
  const int nt = blockDim.x - 32;                   // compute threads; the last warp is the TMA producer
  const int nb = a.nb;
  const RowPart rp = row_part(a.g_total, gridDim.x, blockIdx.x);
  const int R = rp.R;
  const int nbq = (nb + 3) >> 2, cq = a.cb >> 2;       // quads per row / per chunk
  const int nbp = nbq * 4;                              // blocks incl. the zero padding of a partial last quad
  const int nchunks = (nbq + cq - 1) / cq;
  const int S = a.n_stages;

  // shared memory carve-up
  uint8_t *stages = smem;
#if B200_IMMA
  const ActSmem act = act_carve(smem + (size_t) S * a.stage_bytes, nbp);       // quantized activation, MMA-operand form
  float *rowres = act.dxs + nbp;                                                // [rmax]
#else
  const int nbx = nbp + 2;                              // plane stride, padded by 16 bytes against bank conflicts
  uint2 *xq = reinterpret_cast<uint2 *>(smem + (size_t) S * a.stage_bytes);   // [4 planes p][nbx] {bytes of lane 2p, of lane 2p+1}
  float *dxs = reinterpret_cast<float *>(xq + (size_t) nbx * 4);              // [nbp]
  float *rowres = dxs + nbp;                                                    // [rmax]
#endif
  double *red = reinterpret_cast<double *>(rowres + ((a.rmax + 3) & ~3));       // [32]
  uint64_t *full = reinterpret_cast<uint64_t *>(red + 32);                      // [S]
  uint64_t *empty = full + S;                                                   // [S]

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], nt >> 5); }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= nt) {
    // ===== TMA producer: stream this CTA's contiguous weight bytes; does not depend on the upstream kernel =====
    if (tid == nt) {
      const uint8_t *wbase = a.w + (size_t) rp.row0 * nbq * 80;
      for (int k = 0; k < nchunks; k++) {
        const int s = k % S;
        if (k >= S && !mbar_wait(&empty[s], ((k / S) - 1) & 1)) return;
        const int cqk = min(cq, nbq - k * cq);
        const uint32_t bytes = (uint32_t) cqk * R * 80;
        mbar_arrive_expect_tx(&full[s], bytes);
        tma_bulk_g2s(stages + (size_t) s * a.stage_bytes, wbase + (size_t) k * cq * R * 80, bytes, &full[s]);
      }
    }
    return;
  }

  // ===== compute warps =====
  pdl_launch_dependents();
  pdl_wait();                                       // upstream activations are now visible
  const float *__restrict__ x = a.x;
  const int K = nb * 32;

  // ---- prologue: (LayerNorm * weight) and Q4_0 activation quantization, redundantly per CTA ----
  double mean = 0.0;
  float nscale = 1.0f;
  if (PRO == PRO_NORM) {
    // ggml_compute_forward_norm_f32, ggml.c:5363-5381 (sums in double; here in tree order)
    double s = 0.0;
    for (int i = tid; i < K; i += nt) s = __dadd_rn(s, (double) x[i]);
    s = block_sum_d(s, red, tid, nt);
    mean = s / (double) K;
    double s2 = 0.0;
    for (int i = tid; i < K; i += nt) {
      const double v = __dsub_rn((double) x[i], mean);
      s2 = __dadd_rn(s2, __dmul_rn(v, v));
    }
    s2 = block_sum_d(s2, red, tid, nt);
    const double eps = (double) 1e-5f;
    nscale = (float) (1.0 / sqrt(__dadd_rn(s2 / (double) K, eps)));
  }
  for (int b = tid; b < nb; b += nt) {
    // quantize_row_q4_0, AVX2 branch, ggml.c:456-523
    float v[32];
    const float4 *xp = reinterpret_cast<const float4 *>(x + b * 32);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float4 t = xp[i];
      v[4 * i + 0] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
    if (PRO == PRO_NORM) {
      const float4 *wp = reinterpret_cast<const float4 *>(a.norm_w + b * 32);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float4 wv = wp[i];
        const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float y = (float) __dsub_rn((double) v[4 * i + j], mean);       // y[i] = v            ggml.c:5374-5375
          const float ys = __fmul_rn(y, nscale);                                 // ggml_vec_scale_f32  ggml.c:5381
          v[4 * i + j] = __fmul_rn(ww[j], ys);                                   // ggml_mul            PO.mm:573-575
        }
      }
    }
#if B200_IMMA
    quantize_block_full_imma(v, b, act);
  }
  for (int b = nb + tid; b < nbp; b += nt) {        // padding blocks of a partial last quad: scale 0 on both sides = exact no-op
#pragma unroll
    for (int part = 0; part < 4; part++) act_zero_block(b, act, part);
  }
  named_bar_sync(1, nt);
#else
    float amax = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(v[i]));
    const float d = __fdiv_rn(amax, 7.0f);
    const float id = (amax != 0.0f) ? __fdiv_rn(7.0f, amax) : 0.0f;
    int q[32];
#pragma unroll
    for (int i = 0; i < 32; i++) q[i] = __float2int_rn(__fmul_rn(v[i], id));   // round-to-nearest-even, = stored nibble - 8
    uint32_t xs[8];
#pragma unroll
    for (int l = 0; l < 8; l++) {
      const int e0 = q[2 * l], e1 = q[2 * l + 1], e2 = q[16 + 2 * l], e3 = q[17 + 2 * l];
      xs[l] = (uint32_t) (e0 & 0xff) | ((uint32_t) (e1 & 0xff) << 8) | ((uint32_t) (e2 & 0xff) << 16) | ((uint32_t) (e3 & 0xff) << 24);
    }
#pragma unroll
    for (int p = 0; p < 4; p++) xq[(size_t) p * nbx + b] = make_uint2(xs[2 * p], xs[2 * p + 1]);
    dxs[b] = d;
  }
  for (int b = nb + tid; b < nbp; b += nt) {        // padding blocks of a partial last quad: scale 0 on both sides = exact no-op
#pragma unroll
    for (int p = 0; p < 4; p++) xq[(size_t) p * nbx + b] = make_uint2(0u, 0u);
    dxs[b] = 0.0f;
  }
  named_bar_sync(1, nt);
#endif

#if B200_IMMA
  // ---- main loop on the warp-level tensor path: LP selects the warp tile -- 1: 8 rows, 2: 16 rows, 4: 2 x 16 rows ----
  {
    constexpr int RW = LP == 1 ? 8 : 16, NT = LP == 4 ? 2 : 1;
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2;
    const int ntiles = (R + RW - 1) / RW;
    const bool warp_active = warp * NT < ntiles;
    int row[NT][RW / 8], lrow[NT];
    imma_tile_rows<RW, NT>(warp * NT, lane, R, row, lrow);
    uint32_t sel0, sel1;
    imma_selectors(lane, sel0, sel1);
    u64 acc[NT][RW / 8];
#pragma unroll
    for (int i = 0; i < NT; i++)
#pragma unroll
      for (int h = 0; h < RW / 8; h++) acc[i][h] = pack_f2(0.0f, 0.0f);
    for (int k = 0; k < nchunks; k++) {
      const int s = k % S;
      mbar_wait(&full[s], (k / S) & 1);
      if (warp_active) {
        const int cqk = min(cq, nbq - k * cq);
        gemv_chunk_imma<RW, NT>(stages + (size_t) s * a.stage_bytes, cqk, R, row, lrow, lane, act, k * a.cb, sel0, sel1, acc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int i = 0; i < NT; i++)
#pragma unroll
      for (int h = 0; h < RW / 8; h++) {
        const u64 one[1] = {acc[i][h]};
        const float res = row_hsum<1>(one);
        const int r = (warp * NT + i) * RW + g + 8 * h;
        if ((lane & 3) == 0 && r < R) rowres[r] = res;
      }
    named_bar_sync(1, nt);
  }
#else
  // ---- main loop: 8 exact AVX2 lanes per row, LP lane-pairs per thread, one row per thread ----
  const int u = tid;
  const bool active = u < R * UPR;
  const int r = active ? u / UPR : R - 1;
  const int pg = u % UPR;
  u64 acc[1][LP];
#pragma unroll
  for (int j = 0; j < LP; j++) acc[0][j] = pack_f2(0.0f, 0.0f);

  for (int k = 0; k < nchunks; k++) {
    const int s = k % S;
    mbar_wait(&full[s], (k / S) & 1);
    const int cqk = min(cq, nbq - k * cq);
    gemv_chunk<LP, 1>(stages + (size_t) s * a.stage_bytes, cqk, R, r, pg, xq, dxs, nbx, k * a.cb, acc);
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
  }
  const float res = row_hsum<LP>(acc[0]);
  if (active && pg == 0) rowres[r] = res;
  named_bar_sync(1, nt);
#endif

  gemv_epilogue<EPI>(a, rp, rowres, tid, nt);
}

// ---- attention for one token: cluster of 4 CTAs per head ------------------------------------------------------------
struct AttnArgs {
  const float *q;          // [n_embd] roped query
  const float *k_layer;    // [n_ctx][n_embd] (roped keys)
  const float *v_layer;
  float *out;              // [n_embd] = KQV_merged (PO.mm:641-646)
  const StepParams *sp;
  const uint16_t *exp_table;
  int n_embd;
  int n_threads;           // the reference's thread count: selects its V.P partial-sum partition (ggml.c:5628-5665)
  float kq_scale;          // 1.0f/sqrt(float(n_embd)/n_head), PO.mm:620
  int n_ctx;
};

constexpr int ATTN_THREADS = 256;
constexpr int ATTN_CLUSTER = 4;

__global__ void __cluster_dims__(ATTN_CLUSTER, 1, 1) __launch_bounds__(ATTN_THREADS, 1) attn_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  uint8_t *smem = smem_attn;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int) cluster.block_rank();
  const int h = blockIdx.x / ATTN_CLUSTER;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int HD = 128, NW = ATTN_THREADS / 32;
  const int E = a.n_embd;

  pdl_launch_dependents();
  pdl_wait();
  const int pos = a.sp->pos;
  const int p_valid = pos + 1;             // diag_mask_inf: columns > n_past + i are -inf -> probability 0 (ggml.c:6946-6953)
  const int p_part = a.sp->p_part;

  float *sc = reinterpret_cast<float *>(smem);                                    // [p_valid] scores -> probabilities
  double *redd = reinterpret_cast<double *>(smem + (((size_t) a.n_ctx * 4 + 15) & ~(size_t) 15));     // [NW]
  float *redf = reinterpret_cast<float *>(redd + NW);                             // [NW]
  float *part = redf + NW;                                                        // [nth][32]

  // phase 1: K.Q -- ggml_vec_dot_f32 with the AVX mapping: lane t = 8*vec + l owns elements t, t+32, t+64, t+96
  float qv[4];
#pragma unroll
  for (int i = 0; i < 4; i++) qv[i] = a.q[h * HD + lane + 32 * i];
  float *sc_peer = cluster.map_shared_rank(sc, lane & (ATTN_CLUSTER - 1));   // lane c publishes to CTA c of the cluster
  cluster.sync();                           // every CTA of the cluster is resident before remote smem is written
  for (int j = rank * NW + warp; j < p_valid; j += ATTN_CLUSTER * NW) {
    const float *kp = a.k_layer + (size_t) j * E + h * HD + lane;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) s = fmaf(kp[32 * i], qv[i], s);                 // GGML_F32_VEC_FMA(sum, ax, ay), ggml.c:1239
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 8));                        // sum[0]+sum[1], sum[2]+sum[3]   ggml.c:874-876
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 16));                       // (..)+(..)                      ggml.c:877-879
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));                        // lanes k and k+4                ggml.c:883-884
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));                        // hadd                           ggml.c:885
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));                        // hadd                           ggml.c:886
    s = __fmul_rn(s, a.kq_scale);                                                // ggml_scale, PO.mm:617-621
    if (lane < ATTN_CLUSTER) sc_peer[j] = s;
  }
  cluster.sync();

  // phase 2: soft_max, ggml.c:7019-7041 (every CTA of the cluster does the same row)
  float mx = -CUDART_INF_F;
  for (int j = tid; j < p_valid; j += ATTN_THREADS) mx = fmaxf(mx, sc[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) redf[warp] = mx;
  __syncthreads();
  mx = redf[0];
  for (int i = 1; i < NW; i++) mx = fmaxf(mx, redf[i]);
  double sum = 0.0;   // fp16-valued terms: the double sum is exact in any order
  for (int j = tid; j < p_valid; j += ATTN_THREADS) {
    const uint16_t hx = __half_as_ushort(__float2half_rn(__fsub_rn(sc[j], mx)));
    const float e = __half2float(__ushort_as_half(a.exp_table[hx]));
    sc[j] = e;
    sum += (double) e;
  }
  sum = warp_sum_d(sum);
  if (lane == 0) redd[warp] = sum;
  __syncthreads();
  sum = redd[0];
  for (int i = 1; i < NW; i++) sum += redd[i];
  const float inv = (float) (1.0 / sum);
  for (int j = tid; j < p_valid; j += ATTN_THREADS) sc[j] = __fmul_rn(sc[j], inv);   // ggml_vec_scale_f32, ggml.c:7041
  __syncthreads();

  // phase 3: V.P for this CTA's 32 output dims.  Reference thread t accumulates columns [t*dc, (t+1)*dc) with
  // vec_mad_f32 into its own zeroed buffer; FINALIZE adds the buffers in thread order (ggml.c:5570-5574).
  const int nth = a.n_threads;
  const int dc = (p_part + nth - 1) / nth;
  const float *vp = a.v_layer + h * HD + rank * 32 + lane;
  for (int t = warp; t < nth; t += NW) {
    const int j0 = t * dc;
    const int j1 = min(min(j0 + dc, p_part), p_valid);
    float acc = 0.0f;
    int j = j0;
    for (; j + 8 <= j1; j += 8) {
      float vv[8];
#pragma unroll
      for (int i = 0; i < 8; i++) vv[i] = vp[(size_t) (j + i) * E];
#pragma unroll
      for (int i = 0; i < 8; i++) acc = fmaf(vv[i], sc[j + i], acc);            // GGML_F32_VEC_FMA(ay, ax, vx), ggml.c:1696
    }
    for (; j < j1; j++) acc = fmaf(vp[(size_t) j * E], sc[j], acc);
    part[t * 32 + lane] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float o = part[lane];
    for (int t = 1; t < nth; t++) o = __fadd_rn(o, part[t * 32 + lane]);
    a.out[h * HD + rank * 32 + lane] = o;
  }
}

// ---- embedding row: dequantize_row_q4_0 of the raw ggml row (ggml.c:651-684) ----------------------------------------
__global__ void embed_kernel(const uint8_t *tok_emb_raw, const StepParams *sp, float *out, int n_embd) {
  pdl_launch_dependents();
  pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_embd) return;
  const uint8_t *row = tok_emb_raw + (size_t) sp->token * (n_embd / 32) * 20;
  const uint8_t *blk = row + (e / 32) * 20;
  const float d = *reinterpret_cast<const float *>(blk);
  const uint8_t by = blk[4 + (e % 32) / 2];
  const int qn = (e & 1) ? (by >> 4) : (by & 0xf);
  out[e] = __fmul_rn((float) (qn - 8), d);
}

__global__ void set_step_kernel(StepParams *sp, int token, int pos, int p_part, int step, int forced) {
  pdl_launch_dependents();
  pdl_wait();
  sp->token = token; sp->pos = pos; sp->p_part = p_part; sp->step = step; sp->forced = forced;
}

// greedy pick (first maximum, like numpy.argmax) + advance the step scalars; feeds the next graph replay
__global__ void __launch_bounds__(1024, 1) argmax_advance_kernel(const float *logits, int n_vocab, StepParams *sp,
                                                                 int *token_log, const int *forced_tokens) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  pdl_launch_dependents();
  pdl_wait();
  float best = -CUDART_INF_F;
  int idx = 0x7fffffff;
  for (int i = threadIdx.x; i < n_vocab; i += blockDim.x) {
    const float v = logits[i];
    if (v > best || (v == best && i < idx)) { best = v; idx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int) (blockDim.x >> 5); w++)
      if (bv[w] > best || (bv[w] == best && bi[w] < idx)) { best = bv[w]; idx = bi[w]; }
    if (idx < 0 || idx >= n_vocab) idx = 0;          // all-NaN logits: never index the embedding table out of bounds
    const int step = sp->step;
    token_log[step] = idx;
    sp->token = forced_tokens ? forced_tokens[step] : idx;   // teacher forcing when a token stream is supplied
    sp->pos += 1;
    sp->p_part += 1;
    sp->step = step + 1;
  }
}

// ---- load-time repack: ggml rows of 20-byte blocks -> quad-major stream (see the layout comment above) --------------
// src = concatenation of the fused matrices' raw rows.  interleave_half > 0: fused row 2i <- src row i,
// 2i+1 <- src row interleave_half + i (w1/w3).  Rows >= M and blocks >= nb are zero padding (scale 0).
__global__ void repack_q4_0_kernel(const uint8_t *src, uint8_t *dst, int M, int g_total, int nb, int cb, int n_cta,
                                   int interleave_half, int lp) {
  const int nbq = (nb + 3) >> 2, cq = cb >> 2;
  const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long) g_total * 4 * nbq * 4;
  if (idx >= total) return;
  const int gr = (int) (idx / (nbq * 4)), b = (int) (idx % (nbq * 4));
  const int c = cta_of_granule(g_total, n_cta, gr / 4);
  const RowPart rp = row_part(g_total, n_cta, c);
  const int r = gr - rp.row0, R = rp.R;
  const int qd = b >> 2, bq = b & 3;
  const int k = qd / cq, ql = qd % cq;
  uint8_t *quad = dst + (size_t) rp.row0 * nbq * 80 + (size_t) k * cq * R * 80 + (size_t) ql * R * 80;
  uint32_t *dn = reinterpret_cast<uint32_t *>(quad);
  float *ds = reinterpret_cast<float *>(quad + (size_t) R * 64) + r * 4 + bq;
  const int upr = lp ? 4 / lp : 4;
  uint32_t wout[4] = {0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u};
  float d = 0.0f;
  if (gr < M && b < nb) {
    const int sr = interleave_half > 0 ? ((gr & 1) ? interleave_half + gr / 2 : gr / 2) : gr;
    const uint32_t *sp = reinterpret_cast<const uint32_t *>(src + ((size_t) sr * nb + b) * 20);
    d = __uint_as_float(sp[0]);
    const uint32_t by[4] = {sp[1], sp[2], sp[3], sp[4]};
#pragma unroll
    for (int p = 0; p < 4; p++) {
      uint32_t wv = 0;
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const int ea = lane_elem(2 * p, kk), eb = lane_elem(2 * p + 1, kk);
        const uint32_t ba = (by[(ea / 2) / 4] >> (8 * ((ea / 2) % 4))) & 0xff;
        const uint32_t bb = (by[(eb / 2) / 4] >> (8 * ((eb / 2) % 4))) & 0xff;
        const uint32_t na = (ea & 1) ? (ba >> 4) : (ba & 0xf);
        const uint32_t nbv = (eb & 1) ? (bb >> 4) : (bb & 0xf);
        wv |= (na | (nbv << 4)) << (8 * kk);
      }
      wout[p] = wv;
    }
  }
  if (lp == 0) {
    // tensor-path row loop (rowloop_imma.cuh): the block's 16 nibble bytes in ggml's own order, word t = bytes 4t..4t+3
    uint32_t raw[4] = {0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u};
    if (gr < M && b < nb) {
      const uint32_t *sp = reinterpret_cast<const uint32_t *>(src + ((size_t) (interleave_half > 0 ? ((gr & 1) ? interleave_half + gr / 2 : gr / 2) : gr) * nb + b) * 20);
      raw[0] = sp[1]; raw[1] = sp[2]; raw[2] = sp[3]; raw[3] = sp[4];
    }
#pragma unroll
    for (int t = 0; t < 4; t++) dn[((size_t) bq * R + r) * 4 + t] = raw[t];      // [block][row][16 bytes]
    *ds = d;
    return;
  }
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int t = p / lp, j = p % lp;                       // word p belongs to thread t of the row, lane-pair slot j
    dn[(((size_t) j * R + r) * upr + t) * 4 + bq] = wout[p];
  }
  *ds = d;
}

}  // namespace b200
