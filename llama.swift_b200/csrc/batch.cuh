// Prompt batches: llama_eval with N > 1 tokens (PO.mm:510-735 with N = embd_inp.size(); the reference's callers feed
// 4-token probes and 9-token prompt slices, PO.mm:822, 878-889).
//
// The reference computes every mat-mul of a batch as  for (row) for (column ic < N) vec_dot(row, column)  -- each weight
// row is read ONCE for all N columns (ggml.c:6199-6222).  Round 1 ran a batch as N single-token passes (N x 4.13 GB of
// weight traffic); here a batch is evaluated layer by layer:
//
//   batch_embed_kernel ........ get_rows for N tokens                               (ggml.c:6760-6785)
//   batch_prep_kernel ......... per token: [LayerNorm * weight] + quantize_row_q4_0  (ggml.c:5363-5381, 456-523), written
//                               to global memory in the row loop's operand layout (dp4a byte planes + scales)
//   q4_gemm_cols_kernel ....... the mat-mul: this CTA's weight rows are streamed ONCE per group of up to BATCH_NC columns
//                               whose quantized activations sit in shared memory; per (row, block) the nibbles are turned
//                               into signed bytes once and reused by every column.  Same exact per-lane arithmetic as the
//                               single-token loop (kernels.cuh: quad_math), so every column is bit-identical to it.
//   batch_qkv_kernel .......... RoPE + K/V cache rows + Q buffer                     (ggml.c:7110-7127, PO.mm:585-611)
//   batch_attn_kernel ......... (head, token) units; token i sees positions <= n_past + i (diag_mask_inf), the V*P
//                               partition uses n_past + N like the reference (ggml.c:5628)
//   batch_resid_kernel / batch_silu_kernel ... ggml_add / silu(w1 x) * (w3 x)        (PO.mm:654, 678-680, 687)
//
// The logits are needed for the LAST token only (PO.mm:724-725), so the final norm + lm_head run on one column.
// For large N the mat-mul is handed to the tcgen05 / TMEM kernel (prefill_tc.cuh) with the same prep / epilogue kernels.
#pragma once
#include "kernels.cuh"

namespace b200 {

constexpr int BATCH_NC = 8;          // columns per weight pass of the CUDA-core multi-column loop

// ---- embedding rows ------------------------------------------------------------------------------------------------------
__global__ void batch_embed_kernel(const uint8_t *tok_emb_raw, const int *tokens, float *x, int n_embd) {
  const int n = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_embd) return;
  const uint8_t *row = tok_emb_raw + (size_t) tokens[n] * (n_embd / 32) * 20;
  const uint8_t *blk = row + (e / 32) * 20;
  const float d = *reinterpret_cast<const float *>(blk);
  const uint8_t by = blk[4 + (e % 32) / 2];
  const int qn = (e & 1) ? (by >> 4) : (by & 0xf);
  x[(size_t) n * n_embd + e] = __fmul_rn((float) (qn - 8), d);
}

// ---- activation preparation: one CTA per token ---------------------------------------------------------------------------
// Output per token (operand layout of gemv_chunk): xq [4 planes][nbp] uint2, then dxs [nbp] float; nbp = nb rounded up to
// whole quads (padding blocks are zero = exact no-ops).  tok_stride = bytes per token.
__host__ __device__ __forceinline__ size_t batch_act_bytes(int nb) {
  const int nbp = (nb + 3) & ~3;
  return (size_t) nbp * 32 + (size_t) nbp * 4;
}

template <int NORM>
__global__ void __launch_bounds__(256) batch_prep_kernel(const float *x, const float *norm_w, uint8_t *act, int K) {
  __shared__ double red[16];
  const int n = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float *xr = x + (size_t) n * K;
  const int nb = K / 32, nbp = (nb + 3) & ~3;
  uint2 *xq = reinterpret_cast<uint2 *>(act + (size_t) n * batch_act_bytes(nb));
  float *dxs = reinterpret_cast<float *>(xq + (size_t) nbp * 4);
  double mean = 0.0;
  float nscale = 1.0f;
  if (NORM) {
    // ggml_compute_forward_norm_f32, ggml.c:5363-5381 (sums in double; tree order like the other kernels)
    double s = 0.0;
    for (int i = tid; i < K; i += nt) s = __dadd_rn(s, (double) xr[i]);
    s = block_sum_d(s, red, tid, nt);
    mean = s / (double) K;
    double s2 = 0.0;
    for (int i = tid; i < K; i += nt) {
      const double v = __dsub_rn((double) xr[i], mean);
      s2 = __dadd_rn(s2, __dmul_rn(v, v));
    }
    s2 = block_sum_d(s2, red, tid, nt);
    nscale = (float) (1.0 / sqrt(__dadd_rn(s2 / (double) K, (double) 1e-5f)));
  }
  for (int b = tid; b < nbp; b += nt) {
    if (b >= nb) {
#pragma unroll
      for (int p = 0; p < 4; p++) xq[(size_t) p * nbp + b] = make_uint2(0u, 0u);
      dxs[b] = 0.0f;
      continue;
    }
    float v[32];
    const float4 *xp = reinterpret_cast<const float4 *>(xr + b * 32);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float4 t = xp[i];
      v[4 * i + 0] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
    if (NORM) {
      const float4 *wp = reinterpret_cast<const float4 *>(norm_w + b * 32);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float4 wv = wp[i];
        const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float y = (float) __dsub_rn((double) v[4 * i + j], mean);       // ggml.c:5374-5375
          v[4 * i + j] = __fmul_rn(ww[j], __fmul_rn(y, nscale));                 // vec_scale (5381), ggml_mul (PO.mm:573-575)
        }
      }
    }
    float amax = 0.0f;                                                           // quantize_row_q4_0, AVX2 branch, ggml.c:456-523
#pragma unroll
    for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(v[i]));
    const float d = __fdiv_rn(amax, 7.0f);
    const float id = (amax != 0.0f) ? __fdiv_rn(7.0f, amax) : 0.0f;
    uint32_t xs[8];
#pragma unroll
    for (int l = 0; l < 8; l++) {
      const int e0 = __float2int_rn(__fmul_rn(v[2 * l], id)), e1 = __float2int_rn(__fmul_rn(v[2 * l + 1], id));
      const int e2 = __float2int_rn(__fmul_rn(v[16 + 2 * l], id)), e3 = __float2int_rn(__fmul_rn(v[17 + 2 * l], id));
      xs[l] = (uint32_t) (e0 & 0xff) | ((uint32_t) (e1 & 0xff) << 8) | ((uint32_t) (e2 & 0xff) << 16) | ((uint32_t) (e3 & 0xff) << 24);
    }
#pragma unroll
    for (int p = 0; p < 4; p++) xq[(size_t) p * nbp + b] = make_uint2(xs[2 * p], xs[2 * p + 1]);
    dxs[b] = d;
  }
}

// ---- the multi-column mat-mul ---------------------------------------------------------------------------------------------
struct GemmColsArgs {
  const uint8_t *w;       // tile-major stream (same as GemvArgs)
  int M, g_total, nb, cb, n_stages, stage_bytes, rmax;
  const uint8_t *act;     // [N] quantized activations (batch_prep_kernel)
  float *out;             // [N][ld_out] raw mat-mul results
  int ld_out;
  int N;                  // columns in the batch; blockIdx.y picks the group [y*BATCH_NC, ...)
};

// acc[c][j] += the 4 blocks of the quad for column c; the nibble -> signed-byte conversion is shared by all columns
template <int LP>
struct ColQuadW {
  int a_lo[4][LP], a_hi[4][LP];   // [block][lane-pair slot]: signed bytes 16*(q-8) of lanes 2p / 2p+1
  float4 sc;                      // weight scales of the 4 blocks
};

template <int LP>
__device__ __forceinline__ void colquad_load_w(ColQuadW<LP> &q, const uint8_t *pw, const uint8_t *ps, int jstride) {
#pragma unroll
  for (int j = 0; j < LP; j++) {
    const uint4 w4 = *reinterpret_cast<const uint4 *>(pw + j * jstride);
    const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int b = 0; b < 4; b++) {
      q.a_hi[b][j] = (int) and_xor(ww[b], 0xF0F0F0F0u, 0x80808080u);
      q.a_lo[b][j] = (int) and_xor(ww[b] << 4, 0xF0F0F0F0u, 0x80808080u);
    }
  }
  q.sc = *reinterpret_cast<const float4 *>(ps);
}

template <int LP>
__device__ __forceinline__ void colquad_math(const ColQuadW<LP> &q, const uint8_t *px, const float *pd, int xstride, u64 (&acc)[LP]) {
  const u64 cvt_mul = pack_f2(0.0625f, 0.0625f);
  const u64 cvt_sub = pack_f2(-786432.0f, -786432.0f);
  const float4 dx4 = *reinterpret_cast<const float4 *>(pd);
  const float dx[4] = {dx4.x, dx4.y, dx4.z, dx4.w};
  const float sc[4] = {q.sc.x, q.sc.y, q.sc.z, q.sc.w};
  uint4 x[LP][2];
#pragma unroll
  for (int j = 0; j < LP; j++) {
    x[j][0] = *reinterpret_cast<const uint4 *>(px + j * xstride);
    x[j][1] = *reinterpret_cast<const uint4 *>(px + j * xstride + 16);
  }
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const float sdx = __fmul_rn(sc[b], dx[b]);                                   // _mm256_mul_ps(d0, d1), ggml.c:1431
#pragma unroll
    for (int j = 0; j < LP; j++) {
      const uint4 xv = x[j][b >> 1];
      const int xlo = (int) ((b & 1) ? xv.z : xv.x), xhi = (int) ((b & 1) ? xv.w : xv.y);
      const int ia = dp4a_ss(q.a_lo[b][j], xlo, 0x4B400000);
      const int ib = dp4a_ss(q.a_hi[b][j], xhi, 0x4B400000);
      const u64 f = ffma2(pack_i2(ia, ib), cvt_mul, cvt_sub);                    // exact (float) isum of both lanes
      acc[j] = ffma2(pack_f2(sdx, sdx), f, acc[j]);                              // _mm256_fmadd_ps, ggml.c:1457
    }
  }
}

template <int LP>
__global__ void __launch_bounds__(544, 1) q4_gemm_cols_kernel(const GemmColsArgs a) {
  extern __shared__ __align__(128) uint8_t smem_gc[];
  uint8_t *smem = smem_gc;
  constexpr int UPR = 4 / LP;
  const int tid = threadIdx.x;
  const int nt = blockDim.x - 32;
  const int nb = a.nb;
  const RowPart rp = row_part(a.g_total, gridDim.x, blockIdx.x);
  const int R = rp.R;
  const int nbq = (nb + 3) >> 2, cq = a.cb >> 2, nbp = nbq * 4;
  const int nchunks = (nbq + cq - 1) / cq;
  const int S = a.n_stages;
  const int c0 = blockIdx.y * BATCH_NC;
  const int nc = min(BATCH_NC, a.N - c0);

  // shared memory: ring | BATCH_NC activation vectors (padded plane stride) | barriers
  uint8_t *stages = smem;
  const int nbx = nbp + 2;
  const size_t col_bytes = (size_t) nbx * 32 + (size_t) nbp * 4;
  uint8_t *acts = smem + (size_t) S * a.stage_bytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(acts + BATCH_NC * col_bytes);
  uint64_t *empty = full + S;

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], nt >> 5); }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= nt) {
    if (tid == nt) {      // TMA producer: the weight stream does not depend on the upstream kernels
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

  // ---- the columns' quantized activations: global (batch_prep_kernel) -> shared, plane stride padded against conflicts ----
  {
    const size_t src_col = batch_act_bytes(nb);
    for (int c = 0; c < nc; c++) {
      const uint2 *sq = reinterpret_cast<const uint2 *>(a.act + (size_t) (c0 + c) * src_col);
      const float *sd = reinterpret_cast<const float *>(sq + (size_t) nbp * 4);
      uint2 *dq = reinterpret_cast<uint2 *>(acts + c * col_bytes);
      float *dd = reinterpret_cast<float *>(dq + (size_t) nbx * 4);
      for (int i = tid; i < nbp * 4; i += nt) dq[(size_t) (i / nbp) * nbx + (i % nbp)] = sq[i];
      for (int i = tid; i < nbp; i += nt) dd[i] = sd[i];
    }
  }
  named_bar_sync(1, nt);

  const bool active = tid < R * UPR;
  const int r = active ? tid / UPR : R - 1;
  const int t = tid % UPR;
  u64 acc[BATCH_NC][LP];
#pragma unroll
  for (int c = 0; c < BATCH_NC; c++)
#pragma unroll
    for (int j = 0; j < LP; j++) acc[c][j] = pack_f2(0.0f, 0.0f);
  const bool warp_active = (tid & ~31) < R * UPR;

  for (int k = 0; k < nchunks; k++) {
    const int s = k % S;
    mbar_wait(&full[s], (k / S) & 1);
    if (warp_active) {
      const int cqk = min(cq, nbq - k * cq);
      const uint8_t *st = stages + (size_t) s * a.stage_bytes;
      const uint8_t *pw = st + (r * UPR + t) * 16, *ps = st + R * 64 + r * 16;
      const int jstride = R * UPR * 16, qstride = R * 80, xstride = nbx * 8;
      const int b0 = k * a.cb;
      for (int q = 0; q < cqk; q++) {
        ColQuadW<LP> wq;
        colquad_load_w<LP>(wq, pw + q * qstride, ps + q * qstride, jstride);
#pragma unroll
        for (int c = 0; c < BATCH_NC; c++) {
          if (c < nc) {
            const uint8_t *ac = acts + c * col_bytes;
            const uint8_t *px = ac + ((size_t) (t * LP) * nbx + b0 + q * 4) * 8;
            const float *pd = reinterpret_cast<const float *>(ac + (size_t) nbx * 32) + b0 + q * 4;
            colquad_math<LP>(wq, px, pd, xstride, acc[c]);
          }
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
  }
#pragma unroll
  for (int c = 0; c < BATCH_NC; c++) {
    const float res = row_hsum<LP>(acc[c]);           // every lane takes part in the shuffles
    if (c < nc && active && t == 0 && rp.row0 + r < a.M) a.out[(size_t) (c0 + c) * a.ld_out + rp.row0 + r] = res;
  }
}

// ---- epilogues of the batch path --------------------------------------------------------------------------------------------
// fused rows [0,E) = wq, [E,2E) = wk, [2E,3E) = wv.  RoPE on Q and K pairs in double with host-built angles, K/V stored to the
// cache rows n_past + n (PO.mm:585-611: cpy then in-place rope == rope then store)
__global__ void batch_qkv_kernel(const float *qkv, int n_past, float *q_out, float *k_layer, float *v_layer, const double2 *rope,
                                 int n_embd, int head_dim) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // pair index over 3E/2
  const int E = n_embd;
  if (i >= 3 * E / 2) return;
  const int g = 2 * i, which = g / E, col = g - which * E, pos = n_past + n;
  const float *src = qkv + (size_t) n * 3 * E + g;
  float y0 = src[0], y1 = src[1];
  if (which < 2) {
    const double2 cs = rope[(size_t) pos * (head_dim / 2) + (col % head_dim) / 2];
    const double x0 = y0, x1 = y1;
    y0 = (float) __dsub_rn(__dmul_rn(x0, cs.x), __dmul_rn(x1, cs.y));
    y1 = (float) __dadd_rn(__dmul_rn(x0, cs.y), __dmul_rn(x1, cs.x));
  }
  float *dst = which == 0 ? q_out + (size_t) n * E + col
             : which == 1 ? k_layer + (size_t) pos * E + col
                          : v_layer + (size_t) pos * E + col;
  dst[0] = y0;
  dst[1] = y1;
}

// out[n][e] = a[n][e] + b[n][e]   (ggml_add, PO.mm:654, 687)
__global__ void batch_resid_kernel(const float *a, const float *b, float *out, size_t total) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = __fadd_rn(a[i], b[i]);
}

// h[n][i] = silu(o[n][2i]) * o[n][2i+1]   (fused rows 2i = w1 row i, 2i+1 = w3 row i; PO.mm:678-680, ggml.c:1955-1963)
__global__ void batch_silu_kernel(const float *o, float *h, const uint16_t *silu_table, int n_ff, size_t total) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t n = i / n_ff, f = i - n * n_ff;
  const float *p = o + n * 2 * (size_t) n_ff + 2 * f;
  const uint16_t hx = __half_as_ushort(__float2half_rn(p[0]));
  const float sv = __half2float(__ushort_as_half(silu_table[hx]));
  h[i] = __fmul_rn(sv, p[1]);
}

// ---- attention for N query tokens: cluster of 4 CTAs per (head, token) ------------------------------------------------------
struct BatchAttnArgs {
  const float *q;          // [N][n_embd] roped queries
  const float *k_layer;    // [n_ctx][n_embd]
  const float *v_layer;
  float *out;              // [N][n_embd] = KQV_merged (PO.mm:641-646)
  const uint16_t *exp_table;
  int n_embd, n_threads, n_ctx, n_past, N;   // N: the call's tokens from this chunk's first one on (n_past + N = the call's total)
  int N_chunk;                               // tokens in this launch
  float kq_scale;
};

__global__ void __cluster_dims__(ATTN_CLUSTER, 1, 1) __launch_bounds__(ATTN_THREADS, 1) batch_attn_kernel(const BatchAttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem_battn[];
  uint8_t *smem = smem_battn;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int) cluster.block_rank();
  const int h = blockIdx.x / ATTN_CLUSTER;
  const int n = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int HD = 128, NW = ATTN_THREADS / 32;
  const int E = a.n_embd;
  const int pos = a.n_past + n;
  const int p_valid = pos + 1;                 // diag_mask_inf: columns > n_past + i contribute probability 0 (ggml.c:6946-6953)
  const int p_part = a.n_past + a.N;           // the reference partitions V*P columns by the batch's total (ggml.c:5628)

  float *sc = reinterpret_cast<float *>(smem);
  double *redd = reinterpret_cast<double *>(smem + (((size_t) a.n_ctx * 4 + 15) & ~(size_t) 15));
  float *redf = reinterpret_cast<float *>(redd + NW);
  float *part = redf + NW;

  float qv[4];
#pragma unroll
  for (int i = 0; i < 4; i++) qv[i] = a.q[(size_t) n * E + h * HD + lane + 32 * i];
  float *sc_peer = cluster.map_shared_rank(sc, lane & (ATTN_CLUSTER - 1));
  cluster.sync();
  for (int j = rank * NW + warp; j < p_valid; j += ATTN_CLUSTER * NW) {
    const float *kp = a.k_layer + (size_t) j * E + h * HD + lane;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) s = fmaf(kp[32 * i], qv[i], s);                 // ggml_vec_dot_f32 AVX mapping, ggml.c:1223-1258
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 8));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 16));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
    s = __fmul_rn(s, a.kq_scale);
    if (lane < ATTN_CLUSTER) sc_peer[j] = s;
  }
  cluster.sync();

  float mx = -CUDART_INF_F;                                                      // soft_max, ggml.c:7019-7041
  for (int j = tid; j < p_valid; j += ATTN_THREADS) mx = fmaxf(mx, sc[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) redf[warp] = mx;
  __syncthreads();
  mx = redf[0];
  for (int i = 1; i < NW; i++) mx = fmaxf(mx, redf[i]);
  double sum = 0.0;
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
  for (int j = tid; j < p_valid; j += ATTN_THREADS) sc[j] = __fmul_rn(sc[j], inv);
  __syncthreads();

  const int nth = a.n_threads;                                                   // V*P, ggml.c:5619-5665 + FINALIZE 5553-5577
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
      for (int i = 0; i < 8; i++) acc = fmaf(vv[i], sc[j + i], acc);
    }
    for (; j < j1; j++) acc = fmaf(vp[(size_t) j * E], sc[j], acc);
    part[t * 32 + lane] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float o = part[lane];
    for (int t = 1; t < nth; t++) o = __fadd_rn(o, part[t * 32 + lane]);
    a.out[(size_t) n * E + h * HD + rank * 32 + lane] = o;
  }
}


// ---- attention for N query tokens, query-tiled: one CTA per (head, 8 consecutive query tokens) ------------------------------
// The per-(head, token) kernel above re-reads the whole K / V history of the head for every token: 2 x 4 x (n_past + i)
// cache lines per token and head, i.e. O(N^2) L2 traffic that dominates a 2048-token prefill.  Here a K row is loaded once
// and dotted with 8 queries, a V element once and multiplied into 8 accumulators.  Exactness per (query, position) is kept:
//   K.Q   lane t owns dims t, t+32, t+64, t+96 (ggml_vec_dot_f32, AVX mapping) for all 8 queries; the 32-lane reduction tree
//         of the reference (xor 8, 16, 4, 1, 2 -- ggml.c:872-887) is done for the 8 queries at once by recursive halving:
//         at each of the first three levels a lane keeps half of its queries and sends the other half to its partner, so the
//         SAME pairs of partial sums are added as in the reference with 9 shuffles instead of 40;
//   V.P   thread = output dim; the reference's per-thread column ranges are walked in order (ggml.c:5628-5665), masked
//         positions carry probability 0 exactly as after diag_mask_inf + soft_max (fma(v, 0, acc) = acc).
constexpr int ATT_QT = 8;
constexpr int ATT_TILE_THREADS = 256;

__host__ __device__ __forceinline__ size_t att_tile_smem(int n_ctx) { return (size_t) n_ctx * ATT_QT * 4 + 64; }

__global__ void __launch_bounds__(ATT_TILE_THREADS, 3) batch_attn_tile_kernel(const BatchAttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem_batt[];
  const int h = blockIdx.x, i0 = blockIdx.y * ATT_QT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int HD = 128, NW = ATT_TILE_THREADS / 32;
  const int E = a.n_embd;
  const int nq = min(ATT_QT, a.N_chunk - i0);              // live queries of this tile
  const int pmax = a.n_past + i0 + nq;                     // positions any live query may see: [0, pmax)
  const int p_part = a.n_past + a.N;                       // the reference partitions V*P columns by the call's total (ggml.c:5628)
  float *pT = reinterpret_cast<float *>(smem_batt);        // [position][8 queries]: scores, then probabilities (67 KB at n_ctx 2100: 3 CTAs per SM)

  // ---- K.Q ----
  float qv[ATT_QT][4];
#pragma unroll
  for (int qi = 0; qi < ATT_QT; qi++)
#pragma unroll
    for (int i = 0; i < 4; i++) qv[qi][i] = qi < nq ? a.q[(size_t) (i0 + qi) * E + h * HD + lane + 32 * i] : 0.0f;
  const bool bA = (lane >> 3) & 1, bB = (lane >> 4) & 1, bC = (lane >> 2) & 1;
  const int q_mine = 4 * (int) bA + 2 * (int) bB + (int) bC;               // the query whose sum this lane ends up with
  for (int j = warp; j < pmax; j += NW) {
    const float *kp = a.k_layer + (size_t) j * E + h * HD + lane;
    float kk[4];
#pragma unroll
    for (int i = 0; i < 4; i++) kk[i] = kp[32 * i];
    float p[ATT_QT];
#pragma unroll
    for (int qi = 0; qi < ATT_QT; qi++) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; i++) s = fmaf(kk[i], qv[qi][i], s);                // GGML_F32_VEC_FMA, ggml.c:1239
      p[qi] = s;
    }
    float a4[4], b2[2];
#pragma unroll
    for (int k = 0; k < 4; k++) {                                               // sum[0]+sum[1], sum[2]+sum[3]   (xor 8)
      const float send = bA ? p[k] : p[k + 4];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
      a4[k] = __fadd_rn(bA ? p[k + 4] : p[k], recv);
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {                                               // (..)+(..)                      (xor 16)
      const float send = bB ? a4[k] : a4[k + 2];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
      b2[k] = __fadd_rn(bB ? a4[k + 2] : a4[k], recv);
    }
    float c;
    {                                                                           // lanes k and k+4                (xor 4)
      const float send = bC ? b2[0] : b2[1];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
      c = __fadd_rn(bC ? b2[1] : b2[0], recv);
    }
    c = __fadd_rn(c, __shfl_xor_sync(0xffffffffu, c, 1));                       // hadd
    c = __fadd_rn(c, __shfl_xor_sync(0xffffffffu, c, 2));                       // hadd
    if ((lane & 3) == 0) pT[(size_t) j * ATT_QT + q_mine] = __fmul_rn(c, a.kq_scale);   // ggml_scale, PO.mm:617-621
  }
  __syncthreads();

  // ---- soft_max (ggml.c:7019-7041): warp w owns query w; masked positions (diag_mask_inf) get probability 0 ----
  {
    const int qi = warp;
    const int p_valid = qi < nq ? a.n_past + i0 + qi + 1 : 0;
    float *col = pT + qi;
    float mx = -CUDART_INF_F;
    for (int j = lane; j < p_valid; j += 32) mx = fmaxf(mx, col[(size_t) j * ATT_QT]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double sum = 0.0;        // fp16-valued terms: the double sum is exact in any order
    for (int j = lane; j < p_valid; j += 32) {
      const uint16_t hx = __half_as_ushort(__float2half_rn(__fsub_rn(col[(size_t) j * ATT_QT], mx)));
      const float e = __half2float(__ushort_as_half(a.exp_table[hx]));
      col[(size_t) j * ATT_QT] = e;
      sum += (double) e;
    }
    sum = warp_sum_d(sum);
    const float inv = (float) (1.0 / sum);
    for (int j = lane; j < pmax; j += 32)
      col[(size_t) j * ATT_QT] = j < p_valid ? __fmul_rn(col[(size_t) j * ATT_QT], inv) : 0.0f;   // ggml_vec_scale_f32, ggml.c:7041
  }
  __syncthreads();

  // ---- V.P: thread = (output dim d, query group qg of 4); reference thread t owns columns [t*dc, (t+1)*dc), FINALIZE adds
  // the nth buffers in order (ggml.c:5570-5574) ----
  {
    const int d = tid & (HD - 1), qg = tid >> 7;
    const int nth = a.n_threads;
    const int dc = (p_part + nth - 1) / nth;
    const float *vp = a.v_layer + h * HD + d;
    float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int t = 0; t < nth; t++) {
      const int j0 = t * dc;
      const int j1 = min(min(j0 + dc, p_part), pmax);
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      int j = j0;
      for (; j + 8 <= j1; j += 8) {
        float vv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) vv[u] = vp[(size_t) (j + u) * E];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const float4 p4 = *reinterpret_cast<const float4 *>(pT + (size_t) (j + u) * ATT_QT + 4 * qg);
          acc[0] = fmaf(vv[u], p4.x, acc[0]);                                   // ggml_vec_mad_f32, ggml.c:1696
          acc[1] = fmaf(vv[u], p4.y, acc[1]);
          acc[2] = fmaf(vv[u], p4.z, acc[2]);
          acc[3] = fmaf(vv[u], p4.w, acc[3]);
        }
      }
      for (; j < j1; j++) {
        const float v = vp[(size_t) j * E];
        const float4 p4 = *reinterpret_cast<const float4 *>(pT + (size_t) j * ATT_QT + 4 * qg);
        acc[0] = fmaf(v, p4.x, acc[0]); acc[1] = fmaf(v, p4.y, acc[1]); acc[2] = fmaf(v, p4.z, acc[2]); acc[3] = fmaf(v, p4.w, acc[3]);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) o[k] = t == 0 ? acc[k] : __fadd_rn(o[k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (4 * qg + k < nq) a.out[(size_t) (i0 + 4 * qg + k) * E + h * HD + d] = o[k];      // KQV_merged, PO.mm:641-646
  }
}

}  // namespace b200
