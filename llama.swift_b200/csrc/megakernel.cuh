// decode_token_kernel: one persistent launch evaluates one token through the whole network (llama_eval with N = 1,
// PO.mm:510-735).  It replaces ggml_graph_compute's thread pool (ggml.c:9109-9555) with a B200-shaped schedule:
//
//   * grid = one CTA per SM (148), co-resident (cooperative launch), 16 compute warps + 1 TMA producer warp each;
//   * every Q4_0 weight byte of the token (4.13 GB at 7B) is streamed exactly once through a per-CTA ring of
//     shared-memory stages by cp.async.bulk (1-D TMA).  The producer warp walks the static schedule
//     layer0.{wq|wk|wv, wo, w1|w3, w2}, layer1..., output and runs AHEAD of the compute warps across phase boundaries
//     by up to the ring capacity (~185 KB/SM, ~27 MB chip-wide), so HBM keeps streaming while the compute warps sit
//     in a grid barrier, a LayerNorm prologue or the attention phase;
//   * five grid barriers per layer (qkv | attention | wo | w1w3 | w2) are the only synchronisation; activations
//     cross them through L2 (ld.global.cg), never through stale L1 lines.
//
// The arithmetic is the same operation-for-operation mirror of the reference's AVX2 build as kernels.cuh (the
// multi-kernel path kept for A/B): see the contract there.  -fmad=false; FMAs are explicit.
#pragma once
#include "kernels.cuh"

namespace b200 {

struct MatDesc {
  const uint8_t *w;   // tile-major stream laid out for gridDim.x CTAs and chunk length cb
  int M;              // valid fused rows
  int g_total;        // padded rows / 4
  int nb;             // blocks per row
  int cb;             // blocks per chunk
  int lp;             // lane-pairs per thread (1, 2 or 4)
  int pad;
};

struct LayerDesc {
  MatDesc qkv, wo, w13, w2;
  const float *attn_norm, *ffn_norm;
  float *k_layer, *v_layer;
};

struct TokenArgs {
  const LayerDesc *layers;
  int n_layer;
  MatDesc out;
  const float *final_norm;
  const uint8_t *tok_emb;
  float *inpL, *inpFF, *q, *att, *h, *logits;
  const double2 *rope;
  const uint16_t *silu_table, *exp_table;
  const StepParams *sp;
  unsigned int *bar;        // grid-barrier counter, zeroed before every launch
  int n_embd, n_head, n_ctx, n_ff, n_threads;
  float kq_scale;
  int S, stage_bytes;
  int xs_floats;            // size of the f32 staging area (>= max(n_embd, n_ctx))
};

constexpr int MEGA_COMPUTE_THREADS = 512;
constexpr int MEGA_THREADS = MEGA_COMPUTE_THREADS + 32;
constexpr int MEGA_MAX_ROWS = 512;     // rows per CTA upper bound (rowres[])
constexpr int MEGA_MAX_NTH = 16;       // reference thread counts supported by the V*P partition

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier for the compute warps (the producer warp never joins: it only obeys the ring's empty barriers).
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &phase, int tid) {
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  phase++;
  if (tid == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned int target = phase * gridDim.x;
    if (ld_acquire_u32(bar) < target) {
      const long long t0 = clock64();
      while (ld_acquire_u32(bar) < target) {
        if (clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }   // never hang the box
      }
    }
    __threadfence();
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

struct MegaSmem {
  uint8_t *stages;
  uint4 *xq;        // [nb_max][4]
  float *dxs;       // [nb_max]
  float *xs;        // [xs_floats] f32 staging: LayerNorm input / attention scores
  float *rowres;    // [MEGA_MAX_ROWS]
  double *redd;     // [32]
  float *redf;      // [32]
  float *part;      // [MEGA_MAX_NTH][32]
  uint64_t *full, *empty;
};

__device__ __forceinline__ double block_sum_d512(double v, double *red, int tid) {
  v = warp_sum_d(v);
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  if ((tid & 31) == 0) red[tid >> 5] = v;
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  double s = red[0];
#pragma unroll
  for (int i = 1; i < MEGA_COMPUTE_THREADS / 32; i++) s = __dadd_rn(s, red[i]);
  return s;
}

// quantize_row_q4_0 (AVX2 branch, ggml.c:456-523) for one 32-block handled by 4 consecutive lanes (8 values each).
// Writes the dp4a-ready form: xq[b][p] = {xs(lane 2p), xs(lane 2p+1), seed(lane 2p), seed(lane 2p+1)}, dxs[b] = d.
__device__ __forceinline__ void quantize_block_4t(const float v[8], int b, int s, uint4 *xq, float *dxs) {
  float amax = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
  const float d = __fdiv_rn(amax, 7.0f);
  const float id = (amax != 0.0f) ? __fdiv_rn(7.0f, amax) : 0.0f;
  int q[8];
#pragma unroll
  for (int i = 0; i < 8; i++) q[i] = __float2int_rn(__fmul_rn(v[i], id));
  // this thread holds elements 8s..8s+7 = one half (s>>1) of AVX lanes 4*(s&1)+j, j=0..3 (pairs q[2j], q[2j+1])
  const uint32_t hw01 = (uint32_t) (q[0] & 0xff) | ((uint32_t) (q[1] & 0xff) << 8) | ((uint32_t) (q[2] & 0xff) << 16) | ((uint32_t) (q[3] & 0xff) << 24);
  const uint32_t hw23 = (uint32_t) (q[4] & 0xff) | ((uint32_t) (q[5] & 0xff) << 8) | ((uint32_t) (q[6] & 0xff) << 16) | ((uint32_t) (q[7] & 0xff) << 24);
  const int sum01 = q[0] + q[1], sum11 = q[2] + q[3], sum23 = q[4] + q[5], sum33 = q[6] + q[7];
  const uint32_t o01 = __shfl_xor_sync(0xffffffffu, hw01, 2);
  const uint32_t o23 = __shfl_xor_sync(0xffffffffu, hw23, 2);
  const int os0 = __shfl_xor_sync(0xffffffffu, sum01, 2), os1 = __shfl_xor_sync(0xffffffffu, sum11, 2);
  const int os2 = __shfl_xor_sync(0xffffffffu, sum23, 2), os3 = __shfl_xor_sync(0xffffffffu, sum33, 2);
  if (s < 2) {
    // lanes 4s+0..3: low half-word = my elements (2l, 2l+1), high half-word = partner's (16+2l, 17+2l)
    const uint32_t x0 = (hw01 & 0xffffu) | (o01 << 16);
    const uint32_t x1 = (hw01 >> 16) | (o01 & 0xffff0000u);
    const uint32_t x2 = (hw23 & 0xffffu) | (o23 << 16);
    const uint32_t x3 = (hw23 >> 16) | (o23 & 0xffff0000u);
    const int c0 = 0x4B400000 - 8 * (sum01 + os0);      // even lane: low nibbles
    const int c1 = 0x4B400000 - 128 * (sum11 + os1);    // odd lane: high nibbles carry a factor 16
    const int c2 = 0x4B400000 - 8 * (sum23 + os2);
    const int c3 = 0x4B400000 - 128 * (sum33 + os3);
    xq[b * 4 + 2 * s + 0] = make_uint4(x0, x1, (uint32_t) c0, (uint32_t) c1);
    xq[b * 4 + 2 * s + 1] = make_uint4(x2, x3, (uint32_t) c2, (uint32_t) c3);
    if (s == 0) dxs[b] = d;
  }
}

// ---- activation prologues ------------------------------------------------------------------------------------------
// PLAIN: quantize x[K] (global, L2) -> xq/dxs.
__device__ __forceinline__ void prologue_plain(const float *x, int nb, const MegaSmem &sm, int tid) {
  const int items = nb * 4;
  for (int it = tid; it < ((items + MEGA_COMPUTE_THREADS - 1) / MEGA_COMPUTE_THREADS) * MEGA_COMPUTE_THREADS; it += MEGA_COMPUTE_THREADS) {
    const bool live = it < items;       // keep whole warps in the shuffles
    const int iq = live ? it : items - 1;
    const float4 a = __ldcg(reinterpret_cast<const float4 *>(x) + iq * 2);
    const float4 c = __ldcg(reinterpret_cast<const float4 *>(x) + iq * 2 + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    quantize_block_4t(v, iq >> 2, live ? (iq & 3) : 4 + (iq & 3), sm.xq, sm.dxs);   // s >= 4: shuffle only, no store
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

// NORM: x (already in sm.xs[K]) -> LayerNorm (ggml.c:5363-5381) * weight (PO.mm:573-575) -> quantize.
__device__ __forceinline__ void prologue_norm_from_xs(const float *norm_w, int nb, const MegaSmem &sm, int tid) {
  const int K = nb * 32, items = nb * 4;
  double s = 0.0;
  for (int it = tid; it < items; it += MEGA_COMPUTE_THREADS) {
    const float4 a = *(reinterpret_cast<const float4 *>(sm.xs) + it * 2);
    const float4 c = *(reinterpret_cast<const float4 *>(sm.xs) + it * 2 + 1);
    s = __dadd_rn(s, (double) a.x); s = __dadd_rn(s, (double) a.y); s = __dadd_rn(s, (double) a.z); s = __dadd_rn(s, (double) a.w);
    s = __dadd_rn(s, (double) c.x); s = __dadd_rn(s, (double) c.y); s = __dadd_rn(s, (double) c.z); s = __dadd_rn(s, (double) c.w);
  }
  s = block_sum_d512(s, sm.redd, tid);
  const double mean = s / (double) K;
  double s2 = 0.0;
  for (int it = tid; it < items; it += MEGA_COMPUTE_THREADS) {
    const float4 a = *(reinterpret_cast<const float4 *>(sm.xs) + it * 2);
    const float4 c = *(reinterpret_cast<const float4 *>(sm.xs) + it * 2 + 1);
    const float e[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 8; i++) { const double v = __dsub_rn((double) e[i], mean); s2 = __dadd_rn(s2, __dmul_rn(v, v)); }
  }
  s2 = block_sum_d512(s2, sm.redd, tid);
  const float nscale = (float) (1.0 / sqrt(__dadd_rn(s2 / (double) K, (double) 1e-5f)));
  for (int it = tid; it < ((items + MEGA_COMPUTE_THREADS - 1) / MEGA_COMPUTE_THREADS) * MEGA_COMPUTE_THREADS; it += MEGA_COMPUTE_THREADS) {
    const bool live = it < items;
    const int iq = live ? it : items - 1;
    const float4 a = *(reinterpret_cast<const float4 *>(sm.xs) + iq * 2);
    const float4 c = *(reinterpret_cast<const float4 *>(sm.xs) + iq * 2 + 1);
    const float4 wa = __ldg(reinterpret_cast<const float4 *>(norm_w) + iq * 2);
    const float4 wc = __ldg(reinterpret_cast<const float4 *>(norm_w) + iq * 2 + 1);
    const float e[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    const float w[8] = {wa.x, wa.y, wa.z, wa.w, wc.x, wc.y, wc.z, wc.w};
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float y = (float) __dsub_rn((double) e[i], mean);
      v[i] = __fmul_rn(w[i], __fmul_rn(y, nscale));
    }
    quantize_block_4t(v, iq >> 2, live ? (iq & 3) : 4 + (iq & 3), sm.xq, sm.dxs);
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

// ---- GEMV main loop over this CTA's rows of one matrix; leaves the row results in sm.rowres -------------------------
template <int LP>
__device__ __forceinline__ void gemv_rows(const MatDesc &md, const RowPart rp, const MegaSmem &sm, uint32_t &gchunk,
                                          int S, int stage_bytes, int tid) {
  constexpr int UPR = 4 / LP;
  const int R = rp.R, nb = md.nb;
  const int nchunks = (nb + md.cb - 1) / md.cb;
  const bool active = tid < R * UPR;
  const int r = active ? tid / UPR : R - 1;
  const int pg = tid % UPR;
  u64 acc[LP];
#pragma unroll
  for (int j = 0; j < LP; j++) acc[j] = pack_f2(0.0f, 0.0f);
  const u64 cvt_mul = pack_f2(1.0f, 0.0625f);
  const u64 cvt_sub = pack_f2(-12582912.0f, -786432.0f);
  const bool warp_active = (tid & ~31) < R * UPR;     // warps with no rows skip the math but still release stages

  for (int k = 0; k < nchunks; k++, gchunk++) {
    const int s = gchunk % S;
    mbar_wait(&sm.full[s], (gchunk / S) & 1);
    if (warp_active) {
      const int cbk = min(md.cb, nb - k * md.cb);
      const uint8_t *st = sm.stages + (size_t) s * stage_bytes;
      const uint32_t *nib = reinterpret_cast<const uint32_t *>(st) + (size_t) r * 4 + pg * LP;
      const float *sc = reinterpret_cast<const float *>(st + (size_t) cbk * R * 16) + r;
      const uint4 *xqk = sm.xq + (size_t) k * md.cb * 4 + pg * LP;
      const float *dxk = sm.dxs + k * md.cb;
#pragma unroll 4
      for (int bl = 0; bl < cbk; bl++) {
        uint32_t wv[LP];
        if constexpr (LP == 4) {
          const uint4 t = *reinterpret_cast<const uint4 *>(nib + (size_t) bl * R * 4);
          wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
        } else if constexpr (LP == 2) {
          const uint2 t = *reinterpret_cast<const uint2 *>(nib + (size_t) bl * R * 4);
          wv[0] = t.x; wv[1] = t.y;
        } else {
          wv[0] = nib[(size_t) bl * R * 4];
        }
        const float sdx = __fmul_rn(sc[bl * R], dxk[bl]);
#pragma unroll
        for (int j = 0; j < LP; j++) {
          const uint4 xv = xqk[bl * 4 + j];
          const int ia = dp4a_us(wv[j] & 0x0F0F0F0Fu, (int) xv.x, (int) xv.z);
          const int ib = dp4a_us(wv[j] & 0xF0F0F0F0u, (int) xv.y, (int) xv.w);
          const u64 f = ffma2(pack_i2(ia, ib), cvt_mul, cvt_sub);
          acc[j] = ffma2(pack_f2(sdx, sdx), f, acc[j]);
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sm.empty[s]);
  }

  float lane[2 * LP];
#pragma unroll
  for (int j = 0; j < LP; j++) unpack_f2(acc[j], lane[2 * j], lane[2 * j + 1]);
  float res;
  if constexpr (LP == 4) {
    const float r0 = __fadd_rn(lane[4], lane[0]), r1 = __fadd_rn(lane[5], lane[1]);
    const float r2 = __fadd_rn(lane[6], lane[2]), r3 = __fadd_rn(lane[7], lane[3]);
    res = __fadd_rn(__fadd_rn(r0, r2), __fadd_rn(r1, r3));
  } else if constexpr (LP == 2) {
    float rr[4];
#pragma unroll
    for (int i = 0; i < 4; i++) rr[i] = __fadd_rn(lane[i], __shfl_xor_sync(0xffffffffu, lane[i], 1));
    res = __fadd_rn(__fadd_rn(rr[0], rr[2]), __fadd_rn(rr[1], rr[3]));
  } else {
    const float t0 = __fadd_rn(lane[0], __shfl_xor_sync(0xffffffffu, lane[0], 2));
    const float t1 = __fadd_rn(lane[1], __shfl_xor_sync(0xffffffffu, lane[1], 2));
    const float s0 = __fadd_rn(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
    const float s1 = __fadd_rn(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
    res = __fadd_rn(s0, s1);
  }
  if (active && pg == 0) sm.rowres[r] = res;
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

__device__ __forceinline__ void gemv_dispatch(const MatDesc &md, const RowPart rp, const MegaSmem &sm, uint32_t &gchunk,
                                              int S, int stage_bytes, int tid) {
  if (rp.R == 0) { named_bar_sync(1, MEGA_COMPUTE_THREADS); return; }
  switch (md.lp) {
    case 1: gemv_rows<1>(md, rp, sm, gchunk, S, stage_bytes, tid); break;
    case 2: gemv_rows<2>(md, rp, sm, gchunk, S, stage_bytes, tid); break;
    default: gemv_rows<4>(md, rp, sm, gchunk, S, stage_bytes, tid); break;
  }
}

// producer side of one matrix
__device__ __forceinline__ void stream_matrix(const MatDesc &md, const MegaSmem &sm, uint32_t &gchunk, int S, int stage_bytes) {
  const RowPart rp = row_part(md.g_total, gridDim.x, blockIdx.x);
  if (rp.R == 0) return;
  const int nchunks = (md.nb + md.cb - 1) / md.cb;
  const uint8_t *wbase = md.w + (size_t) rp.row0 * md.nb * 20;
  for (int k = 0; k < nchunks; k++, gchunk++) {
    const int s = gchunk % S;
    if (gchunk >= (uint32_t) S) mbar_wait(&sm.empty[s], ((gchunk / S) - 1) & 1);
    const int cbk = min(md.cb, md.nb - k * md.cb);
    const uint32_t bytes = (uint32_t) cbk * rp.R * 20;
    mbar_arrive_expect_tx(&sm.full[s], bytes);
    tma_bulk_g2s(sm.stages + (size_t) s * stage_bytes, wbase + (size_t) k * md.cb * rp.R * 20, bytes, &sm.full[s]);
  }
}

// ---- attention phase for (head h, output quarter qr): K.Q for all positions, soft_max, V.P for 32 dims --------------
__device__ __forceinline__ void attention_phase(const TokenArgs &a, const LayerDesc &L, const MegaSmem &sm, int h, int qr,
                                                int pos, int p_part, int tid) {
  constexpr int HD = 128, NW = MEGA_COMPUTE_THREADS / 32;
  const int lane = tid & 31, warp = tid >> 5;
  const int E = a.n_embd;
  const int p_valid = pos + 1;
  float *sc = sm.xs;
  float qv[4];
#pragma unroll
  for (int i = 0; i < 4; i++) qv[i] = __ldcg(a.q + h * HD + lane + 32 * i);
  // K.Q: ggml_vec_dot_f32, AVX mapping (lane t = 8*vec + l owns elements t, t+32, t+64, t+96), ggml.c:1223-1258, 872-887
  for (int j0 = warp * 4; j0 < p_valid; j0 += NW * 4) {
    float kk[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = min(j0 + u, p_valid - 1);
      const float *kp = L.k_layer + (size_t) j * E + h * HD + lane;
#pragma unroll
      for (int i = 0; i < 4; i++) kk[u][i] = __ldcg(kp + 32 * i);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; i++) s = fmaf(kk[u][i], qv[i], s);
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 8));
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 16));
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
      s = __fmul_rn(s, a.kq_scale);                                               // ggml_scale, PO.mm:617-621
      if (lane == 0 && j0 + u < p_valid) sc[j0 + u] = s;
    }
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  // soft_max, ggml.c:7019-7041
  float mx = -CUDART_INF_F;
  for (int j = tid; j < p_valid; j += MEGA_COMPUTE_THREADS) mx = fmaxf(mx, sc[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) sm.redf[warp] = mx;
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  mx = sm.redf[0];
#pragma unroll
  for (int i = 1; i < NW; i++) mx = fmaxf(mx, sm.redf[i]);
  double sum = 0.0;   // fp16-valued terms: exact in double in any order
  for (int j = tid; j < p_valid; j += MEGA_COMPUTE_THREADS) {
    const uint16_t hx = __half_as_ushort(__float2half_rn(__fsub_rn(sc[j], mx)));
    const float e = __half2float(__ushort_as_half(__ldg(a.exp_table + hx)));
    sc[j] = e;
    sum += (double) e;
  }
  sum = block_sum_d512(sum, sm.redd, tid);
  const float inv = (float) (1.0 / sum);
  for (int j = tid; j < p_valid; j += MEGA_COMPUTE_THREADS) sc[j] = __fmul_rn(sc[j], inv);
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  // V.P: reference thread t owns columns [t*dc, (t+1)*dc) (ggml.c:5628-5632), FINALIZE adds buffers in order (5570-5574)
  const int nth = a.n_threads;
  const int dc = (p_part + nth - 1) / nth;
  const float *vp = L.v_layer + h * HD + qr * 32 + lane;
  for (int t = warp; t < nth; t += NW) {
    const int j0 = t * dc;
    const int j1 = min(min(j0 + dc, p_part), p_valid);
    float acc = 0.0f;
    int j = j0;
    for (; j + 8 <= j1; j += 8) {
      float vv[8];
#pragma unroll
      for (int i = 0; i < 8; i++) vv[i] = __ldcg(vp + (size_t) (j + i) * E);
#pragma unroll
      for (int i = 0; i < 8; i++) acc = fmaf(vv[i], sc[j + i], acc);            // vec_mad_f32, ggml.c:1696
    }
    for (; j < j1; j++) acc = fmaf(__ldcg(vp + (size_t) j * E), sc[j], acc);
    sm.part[t * 32 + lane] = acc;
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  if (warp == 0) {
    float o = sm.part[lane];
    for (int t = 1; t < nth; t++) o = __fadd_rn(o, sm.part[t * 32 + lane]);
    a.att[h * HD + qr * 32 + lane] = o;                                           // KQV_merged, PO.mm:641-646
  }
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_token_kernel(const TokenArgs a) {
  extern __shared__ __align__(128) uint8_t smem_mega[];
  const int tid = threadIdx.x;
  const int S = a.S, stage_bytes = a.stage_bytes;
  const int nb_max = max(a.n_embd, a.n_ff) / 32;

  MegaSmem sm;
  sm.stages = smem_mega;
  sm.xq = reinterpret_cast<uint4 *>(smem_mega + (size_t) S * stage_bytes);
  sm.dxs = reinterpret_cast<float *>(sm.xq + (size_t) nb_max * 4);
  sm.xs = sm.dxs + ((nb_max + 3) & ~3);
  sm.rowres = sm.xs + a.xs_floats;
  sm.redd = reinterpret_cast<double *>(sm.rowres + MEGA_MAX_ROWS);
  sm.redf = reinterpret_cast<float *>(sm.redd + 32);
  sm.part = sm.redf + 32;
  sm.full = reinterpret_cast<uint64_t *>(sm.part + MEGA_MAX_NTH * 32);
  sm.empty = sm.full + S;

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], MEGA_COMPUTE_THREADS / 32); }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= MEGA_COMPUTE_THREADS) {
    // ===== TMA producer: the whole token's weight stream for this SM, in schedule order =====
    if (tid == MEGA_COMPUTE_THREADS) {
      uint32_t g = 0;
      for (int il = 0; il < a.n_layer; il++) {
        const LayerDesc &L = a.layers[il];
        stream_matrix(L.qkv, sm, g, S, stage_bytes);
        stream_matrix(L.wo, sm, g, S, stage_bytes);
        stream_matrix(L.w13, sm, g, S, stage_bytes);
        stream_matrix(L.w2, sm, g, S, stage_bytes);
      }
      stream_matrix(a.out, sm, g, S, stage_bytes);
    }
    return;
  }

  // ===== compute warps =====
  const int E = a.n_embd, HD = E / a.n_head;
  const int pos = a.sp->pos, p_part = a.sp->p_part, token = a.sp->token;
  uint32_t gchunk = 0;
  unsigned int phase = 0;

  // get_rows: dequantize_row_q4_0 of the token's embedding row (ggml.c:6760-6785, 651-684) straight into the staging area
  {
    const uint8_t *row = a.tok_emb + (size_t) token * (E / 32) * 20;
    for (int e = tid; e < E; e += MEGA_COMPUTE_THREADS) {
      const uint8_t *blk = row + (e / 32) * 20;
      const float d = __ldg(reinterpret_cast<const float *>(blk));
      const uint8_t by = __ldg(blk + 4 + (e % 32) / 2);
      const int qn = (e & 1) ? (by >> 4) : (by & 0xf);
      const float v = __fmul_rn((float) (qn - 8), d);
      sm.xs[e] = v;
      if (blockIdx.x == 0) a.inpL[e] = v;          // residual source for layer 0 (read after two grid barriers)
    }
    named_bar_sync(1, MEGA_COMPUTE_THREADS);
  }

  for (int il = 0; il < a.n_layer; il++) {
    const LayerDesc &L = a.layers[il];
    // ---- phase 1: norm -> wq|wk|wv -> rope -> q buffer + KV cache row (PO.mm:570-611) ----
    {
      if (il > 0) {
        for (int i = tid; i < E / 4; i += MEGA_COMPUTE_THREADS)
          reinterpret_cast<float4 *>(sm.xs)[i] = __ldcg(reinterpret_cast<const float4 *>(a.inpL) + i);
        named_bar_sync(1, MEGA_COMPUTE_THREADS);
      }
      prologue_norm_from_xs(L.attn_norm, E / 32, sm, tid);
      const RowPart rp = row_part(L.qkv.g_total, gridDim.x, blockIdx.x);
      gemv_dispatch(L.qkv, rp, sm, gchunk, S, stage_bytes, tid);
      for (int i = tid; i < rp.R / 2; i += MEGA_COMPUTE_THREADS) {
        const int g = rp.row0 + 2 * i;
        if (g >= L.qkv.M) continue;
        const int which = g / E, col = g - which * E;
        float y0 = sm.rowres[2 * i], y1 = sm.rowres[2 * i + 1];
        if (which < 2) {   // ggml_rope, ggml.c:7110-7127 (double math, host-built angles)
          const double2 cs = a.rope[(size_t) pos * (HD / 2) + (col % HD) / 2];
          const double x0 = y0, x1 = y1;
          y0 = (float) __dsub_rn(__dmul_rn(x0, cs.x), __dmul_rn(x1, cs.y));
          y1 = (float) __dadd_rn(__dmul_rn(x0, cs.y), __dmul_rn(x1, cs.x));
        }
        float *dst = which == 0 ? a.q + col : which == 1 ? L.k_layer + (size_t) pos * E + col : L.v_layer + (size_t) pos * E + col;
        dst[0] = y0;
        dst[1] = y1;
      }
      grid_barrier(a.bar, phase, tid);
    }
    // ---- phase 2: attention (PO.mm:614-646) on the first 4*n_head CTAs ----
    {
      if ((int) blockIdx.x < 4 * a.n_head) attention_phase(a, L, sm, blockIdx.x >> 2, blockIdx.x & 3, pos, p_part, tid);
      grid_barrier(a.bar, phase, tid);
    }
    // ---- phase 3: wo, + inpSA (PO.mm:649-654) ----
    {
      prologue_plain(a.att, E / 32, sm, tid);
      const RowPart rp = row_part(L.wo.g_total, gridDim.x, blockIdx.x);
      gemv_dispatch(L.wo, rp, sm, gchunk, S, stage_bytes, tid);
      for (int i = tid; i < rp.R; i += MEGA_COMPUTE_THREADS) {
        const int g = rp.row0 + i;
        if (g < L.wo.M) a.inpFF[g] = __fadd_rn(sm.rowres[i], __ldcg(a.inpL + g));
      }
      grid_barrier(a.bar, phase, tid);
    }
    // ---- phase 4: norm -> w1|w3 -> silu(w1 x) * (w3 x) (PO.mm:660-680) ----
    {
      for (int i = tid; i < E / 4; i += MEGA_COMPUTE_THREADS)
        reinterpret_cast<float4 *>(sm.xs)[i] = __ldcg(reinterpret_cast<const float4 *>(a.inpFF) + i);
      named_bar_sync(1, MEGA_COMPUTE_THREADS);
      prologue_norm_from_xs(L.ffn_norm, E / 32, sm, tid);
      const RowPart rp = row_part(L.w13.g_total, gridDim.x, blockIdx.x);
      gemv_dispatch(L.w13, rp, sm, gchunk, S, stage_bytes, tid);
      for (int i = tid; i < rp.R / 2; i += MEGA_COMPUTE_THREADS) {
        const int g = rp.row0 / 2 + i;
        if (2 * g < L.w13.M) {
          const uint16_t hx = __half_as_ushort(__float2half_rn(sm.rowres[2 * i]));
          const float sv = __half2float(__ushort_as_half(__ldg(a.silu_table + hx)));
          a.h[g] = __fmul_rn(sv, sm.rowres[2 * i + 1]);
        }
      }
      grid_barrier(a.bar, phase, tid);
    }
    // ---- phase 5: w2, + inpFF (PO.mm:682-687) ----
    {
      prologue_plain(a.h, a.n_ff / 32, sm, tid);
      const RowPart rp = row_part(L.w2.g_total, gridDim.x, blockIdx.x);
      gemv_dispatch(L.w2, rp, sm, gchunk, S, stage_bytes, tid);
      for (int i = tid; i < rp.R; i += MEGA_COMPUTE_THREADS) {
        const int g = rp.row0 + i;
        if (g < L.w2.M) a.inpL[g] = __fadd_rn(sm.rowres[i], __ldcg(a.inpFF + g));
      }
      grid_barrier(a.bar, phase, tid);
    }
  }
  // ---- final norm -> output (PO.mm:694-706) ----
  {
    for (int i = tid; i < E / 4; i += MEGA_COMPUTE_THREADS)
      reinterpret_cast<float4 *>(sm.xs)[i] = __ldcg(reinterpret_cast<const float4 *>(a.inpL) + i);
    named_bar_sync(1, MEGA_COMPUTE_THREADS);
    prologue_norm_from_xs(a.final_norm, E / 32, sm, tid);
    const RowPart rp = row_part(a.out.g_total, gridDim.x, blockIdx.x);
    gemv_dispatch(a.out, rp, sm, gchunk, S, stage_bytes, tid);
    for (int i = tid; i < rp.R; i += MEGA_COMPUTE_THREADS) {
      const int g = rp.row0 + i;
      if (g < a.out.M) a.logits[g] = sm.rowres[i];
    }
  }
}

}  // namespace b200
