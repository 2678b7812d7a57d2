// decode_token_kernel: one persistent launch evaluates one token through the whole network (llama_eval with N = 1,
// PO.mm:510-735).  It replaces ggml_graph_compute's thread pool (ggml.c:9109-9555) with a B200-shaped schedule:
//
//   * grid = one CTA per SM (148), co-resident (cooperative launch): 14 compute warps, 1 TMA loader warp and
//     1 L2-prefetch warp each;
//   * every Q4_0 weight byte of the token (4.13 GB at 7B) is streamed exactly once through a per-CTA ring of
//     shared-memory stages by cp.async.bulk (1-D TMA).  The loader walks the static schedule
//     layer0.{wq|wk|wv, wo, w1|w3, w2}, layer1..., output and runs AHEAD of the compute warps across phase boundaries
//     by up to the ring capacity (~170 KB/SM, ~25 MB chip-wide); the prefetch warp runs further ahead still and pulls
//     the next ~50 MB of the stream from HBM into L2 (cp.async.bulk.prefetch.L2), so HBM keeps streaming while the
//     compute warps sit in a grid barrier, a LayerNorm prologue or the attention phase;
//   * five grid barriers per layer (qkv | attention | wo | w1w3 | w2) are the only synchronisation; activations
//     cross them through L2 (ld.global.cg), never through stale L1 lines.
//
// The arithmetic is the same operation-for-operation mirror of the reference's AVX2 build as kernels.cuh (the
// per-matrix kernels kept for A/B and for reference thread counts > 16): see the contract there.  -fmad=false.
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
  int xs_floats;            // size of the f32 scratch area (attention scores): >= n_ctx
  int l2_ahead;             // chunks the L2-prefetch warp may run ahead of the loader (0 = off)
  long long *prof;          // optional [gridDim.x][prof_marks] globaltimer stamps (development profiler), else null
  int prof_marks;
};

constexpr int MEGA_COMPUTE_WARPS = 14;
constexpr int MEGA_COMPUTE_THREADS = MEGA_COMPUTE_WARPS * 32;   // 448
constexpr int MEGA_THREADS = MEGA_COMPUTE_THREADS + 64;          // + loader warp + L2-prefetch warp = 512 -> 128 regs/thread
constexpr int MEGA_MAX_ROWS = 448;     // rows per CTA upper bound (rowres[])
constexpr int MEGA_MAX_NTH = 16;       // reference thread counts supported by the V*P partition
constexpr int MEGA_NORM_ROUNDS = 3;    // 8-element items per thread held in registers by the LayerNorm prologue (K <= 10752)

__device__ __forceinline__ long long globaltimer_ns() {
  long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
#define PROF_MARK() do { if (a.prof && tid == 0 && pm < a.prof_marks) a.prof[(size_t) blockIdx.x * a.prof_marks + pm] = globaltimer_ns(); pm++; } while (0)

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int *p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Grid barrier for the compute warps (the loader / prefetch warps never join: they only obey the ring).
// bar.sync orders every compute thread's global writes before thread 0's gpu-scope release; the acquire + bar.sync
// order every later ld.global.cg after the other CTAs' releases.
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &phase, int tid) {
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  phase++;
  if (tid == 0) {
    red_release_add_u32(bar, 1u);
    const unsigned int target = phase * gridDim.x;
    if (ld_acquire_u32(bar) < target) {
      const long long t0 = clock64();
      while (ld_acquire_u32(bar) < target) {
        if (clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }   // never hang the box
      }
    }
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

struct MegaSmem {
  uint8_t *stages;
  uint4 *xq;        // [nb_max][4]
  float *dxs;       // [nb_max]
  float *xs;        // [xs_floats] attention scores / probabilities
  float *rowres;    // [MEGA_MAX_ROWS]
  double *redd;     // [2][16]
  float *redf;      // [2][16]
  float *part;      // [MEGA_MAX_NTH][32]
  uint64_t *full, *empty;
  volatile uint32_t *loader_g;   // loader progress, read by the prefetch warp
};

// Block-wide sums over the compute warps with ONE barrier: warp shuffle tree, 14 partials, then every warp folds the
// partials with the same shuffle tree (fixed order => deterministic).  `buf` alternates between consecutive calls.
__device__ __forceinline__ double mega_sum_d(double v, double *redd, int buf, int tid) {
  v = warp_sum_d(v);
  double *r = redd + buf * 16;
  if ((tid & 31) == 0) r[tid >> 5] = v;
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  const int l = tid & 15;
  double s = l < MEGA_COMPUTE_WARPS ? r[l] : 0.0;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
  return s;
}
__device__ __forceinline__ float mega_max_f(float v, float *redf, int buf, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  float *r = redf + buf * 16;
  if ((tid & 31) == 0) r[tid >> 5] = v;
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  const int l = tid & 15;
  float s = l < MEGA_COMPUTE_WARPS ? r[l] : -CUDART_INF_F;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, o));
  return s;
}

// quantize_row_q4_0 (AVX2 branch, ggml.c:456-523) for one 32-block handled by 4 consecutive lanes (8 values each).
// Writes the dp4a-ready form: xq[b][p] = {xs(lane 2p), xs(lane 2p+1), seed(lane 2p), seed(lane 2p+1)}, dxs[b] = d.
// s = lane-quad index 0..3 of a live block, >= 4 for a padding thread (takes part in the shuffles, stores nothing).
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
  for (int i = 0; i < 8; i++) q[i] = __float2int_rn(__fmul_rn(v[i], id));      // round-to-nearest-even = stored nibble - 8
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
    // dp4a accumulator seeds: 0x4B400000 is the bit pattern of 12582912.0f, so (seed + isum) IS the float 12582912+isum
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
// PLAIN: quantize x[K] (global, read through L2) -> xq/dxs.  An "item" is 8 consecutive floats = a quarter block.
__device__ __forceinline__ void prologue_plain(const float *x, int nb, const MegaSmem &sm, int tid) {
  const int items = nb * 4;
  const int rounds = (items + MEGA_COMPUTE_THREADS - 1) / MEGA_COMPUTE_THREADS;
  for (int rd = 0; rd < rounds; rd++) {
    const int it = tid + rd * MEGA_COMPUTE_THREADS;
    const bool live = it < items;       // a block's 4 quarter-items are all live or all padding (448 % 4 == 0)
    const int iq = live ? it : items - 1;
    const float4 a = __ldcg(reinterpret_cast<const float4 *>(x) + iq * 2);
    const float4 c = __ldcg(reinterpret_cast<const float4 *>(x) + iq * 2 + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    quantize_block_4t(v, iq >> 2, live ? (iq & 3) : 4 + (iq & 3), sm.xq, sm.dxs);
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

// NORM: LayerNorm (ggml_compute_forward_norm_f32, ggml.c:5363-5381; double sums, here in tree order) times the norm
// weight (ggml_mul, PO.mm:573-575), then quantize.  The vector lives in registers as doubles: xd[rd][0..8) is item
// tid + rd*448 (caller-filled; padding items must be zero-filled and are excluded from the sums by `items`).
__device__ __forceinline__ void prologue_norm_regs(double (&xd)[MEGA_NORM_ROUNDS][8], const float *norm_w, int nb,
                                                   const MegaSmem &sm, int tid) {
  const int K = nb * 32, items = nb * 4;
  double s = 0.0;
#pragma unroll
  for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
    if (tid + rd * MEGA_COMPUTE_THREADS < items) {
#pragma unroll
      for (int i = 0; i < 8; i++) s = __dadd_rn(s, xd[rd][i]);
    }
  }
  s = mega_sum_d(s, sm.redd, 0, tid);
  const double mean = s / (double) K;
  double s2 = 0.0;
#pragma unroll
  for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
    if (tid + rd * MEGA_COMPUTE_THREADS < items) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        xd[rd][i] = __dsub_rn(xd[rd][i], mean);                                  // v = x - mean      ggml.c:5374
        s2 = __dadd_rn(s2, __dmul_rn(xd[rd][i], xd[rd][i]));                     // sum2 += v*v       ggml.c:5376
      }
    }
  }
  s2 = mega_sum_d(s2, sm.redd, 1, tid);
  const float nscale = (float) (1.0 / sqrt(__dadd_rn(s2 / (double) K, (double) 1e-5f)));   // ggml.c:5379
#pragma unroll
  for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
    if (rd * MEGA_COMPUTE_THREADS < items) {      // CTA-uniform: whole rounds only
      const int it = tid + rd * MEGA_COMPUTE_THREADS;
      const bool live = it < items;
      const int iq = live ? it : items - 1;
      const float4 wa = __ldg(reinterpret_cast<const float4 *>(norm_w) + iq * 2);
      const float4 wc = __ldg(reinterpret_cast<const float4 *>(norm_w) + iq * 2 + 1);
      const float w[8] = {wa.x, wa.y, wa.z, wa.w, wc.x, wc.y, wc.z, wc.w};
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = __fmul_rn(w[i], __fmul_rn((float) xd[rd][i], nscale));   // y = (float)v; y *= scale; w*y
      quantize_block_4t(v, iq >> 2, live ? (iq & 3) : 4 + (iq & 3), sm.xq, sm.dxs);
    }
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

// fill the register copy of a global f32 vector (through L2)
__device__ __forceinline__ void load_items(double (&xd)[MEGA_NORM_ROUNDS][8], const float *x, int nb, int tid) {
  const int items = nb * 4;
#pragma unroll
  for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
    const int it = tid + rd * MEGA_COMPUTE_THREADS;
    if (it < items) {
      const float4 a = __ldcg(reinterpret_cast<const float4 *>(x) + it * 2);
      const float4 c = __ldcg(reinterpret_cast<const float4 *>(x) + it * 2 + 1);
      xd[rd][0] = a.x; xd[rd][1] = a.y; xd[rd][2] = a.z; xd[rd][3] = a.w;
      xd[rd][4] = c.x; xd[rd][5] = c.y; xd[rd][6] = c.z; xd[rd][7] = c.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) xd[rd][i] = 0.0;
    }
  }
}

// ---- GEMV main loop over this CTA's rows of one matrix; leaves the row results in sm.rowres -------------------------
// One "group" = G blocks of this thread's LP lane-pairs.  The loop is software pipelined by hand: the shared-memory
// operands of group g+1 are loaded into a second register set while group g is computed, because with 1-2 resident
// warps per scheduler (small-M matrices) nothing else hides the LDS latency.
template <int LP>
struct GemvGroup {
  static constexpr int G = LP == 1 ? 4 : (LP == 2 ? 2 : 1);
  uint32_t w[G * LP];     // [g][j] -> w[g * LP + j]
  uint4 x[G * LP];
  float sc[G], dx[G];
};

template <int LP>
__device__ __forceinline__ void group_load(GemvGroup<LP> &gr, const uint32_t *nib, const float *sc, const uint4 *xq,
                                           const float *dx, int wstride /*words per block*/, int sstride) {
  constexpr int G = GemvGroup<LP>::G;
#pragma unroll
  for (int g = 0; g < G; g++) {
    if constexpr (LP == 4) {
      const uint4 t = *reinterpret_cast<const uint4 *>(nib + g * wstride);
      gr.w[0] = t.x; gr.w[1] = t.y; gr.w[2] = t.z; gr.w[3] = t.w;
    } else if constexpr (LP == 2) {
      const uint2 t = *reinterpret_cast<const uint2 *>(nib + g * wstride);
      gr.w[g * 2] = t.x; gr.w[g * 2 + 1] = t.y;
    } else {
      gr.w[g] = nib[g * wstride];
    }
    gr.sc[g] = sc[g * sstride];
    gr.dx[g] = dx[g];
#pragma unroll
    for (int j = 0; j < LP; j++) gr.x[g * LP + j] = xq[g * 4 + j];
  }
}

template <int LP>
__device__ __forceinline__ void group_compute(const GemvGroup<LP> &gr, u64 (&acc)[LP], const u64 cvt_mul, const u64 cvt_sub) {
  constexpr int G = GemvGroup<LP>::G;
#pragma unroll
  for (int g = 0; g < G; g++) {
    const float sdx = __fmul_rn(gr.sc[g], gr.dx[g]);                              // _mm256_mul_ps(d0, d1), ggml.c:1431
#pragma unroll
    for (int j = 0; j < LP; j++) {
      const uint32_t wv = gr.w[g * LP + j];
      const uint4 xv = gr.x[g * LP + j];
      const int ia = dp4a_us(wv & 0x0F0F0F0Fu, (int) xv.x, (int) xv.z);          // float bits of 12582912 + isum(lane 2p)
      const int ib = dp4a_us(wv & 0xF0F0F0F0u, (int) xv.y, (int) xv.w);          // float bits of 12582912 + 16*isum(lane 2p+1)
      const u64 f = ffma2(pack_i2(ia, ib), cvt_mul, cvt_sub);                     // exact (float)isum for both lanes
      acc[j] = ffma2(pack_f2(sdx, sdx), f, acc[j]);                               // _mm256_fmadd_ps(scale, p, acc), ggml.c:1457
    }
  }
}

// Out-of-line on purpose: each LP variant gets its own register allocation (inlined 15x into the token kernel the
// pipelined loop spilled).  Shared-memory areas travel as 32-bit shared-window offsets and are turned back into
// pointers here, which keeps the address space visible to the compiler (LDS, not generic LD).
struct GemvCall {
  uint32_t stages, xq, dxs, rowres, full, empty;   // shared-window addresses
  int nb, cb, R, S, stage_bytes;
};

struct GemvSm {
  const uint8_t *stages;
  const uint4 *xq;
  const float *dxs;
  float *rowres;
  uint64_t *full, *empty;
};

template <int LP>
__device__ __noinline__ uint32_t gemv_rows(const GemvCall c, uint32_t gchunk, const int tid) {
  constexpr int UPR = 4 / LP, G = GemvGroup<LP>::G;
  GemvSm sm;
  sm.stages = reinterpret_cast<const uint8_t *>(__cvta_shared_to_generic(c.stages));
  sm.xq = reinterpret_cast<const uint4 *>(__cvta_shared_to_generic(c.xq));
  sm.dxs = reinterpret_cast<const float *>(__cvta_shared_to_generic(c.dxs));
  sm.rowres = reinterpret_cast<float *>(__cvta_shared_to_generic(c.rowres));
  sm.full = reinterpret_cast<uint64_t *>(__cvta_shared_to_generic(c.full));
  sm.empty = reinterpret_cast<uint64_t *>(__cvta_shared_to_generic(c.empty));
  const int R = c.R, nb = c.nb, cb = c.cb, S = c.S, stage_bytes = c.stage_bytes;
  const int nchunks = (nb + cb - 1) / cb;
  const bool active = tid < R * UPR;
  const int r = active ? tid / UPR : R - 1;
  const int pg = tid % UPR;
  u64 acc[LP];
#pragma unroll
  for (int j = 0; j < LP; j++) acc[j] = pack_f2(0.0f, 0.0f);
  const u64 cvt_mul = pack_f2(1.0f, 0.0625f);
  const u64 cvt_sub = pack_f2(-12582912.0f, -786432.0f);
  const bool warp_active = (tid & ~31) < R * UPR;     // warps with no rows skip the math but still release stages
  const int wstride = R * 4;

  for (int k = 0; k < nchunks; k++, gchunk++) {
    const int s = gchunk % S;
    mbar_wait(&sm.full[s], (gchunk / S) & 1);
    if (warp_active) {
      const int cbk = min(cb, nb - k * cb);
      const uint8_t *st = sm.stages + (size_t) s * stage_bytes;
      const uint32_t *nib = reinterpret_cast<const uint32_t *>(st) + r * 4 + pg * LP;
      const float *sc = reinterpret_cast<const float *>(st + cbk * R * 16) + r;
      const uint4 *xqk = sm.xq + k * cb * 4 + pg * LP;
      const float *dxk = sm.dxs + k * cb;
      const int ngroups = cbk / G;
      GemvGroup<LP> ga, gb;
      if (ngroups > 0) group_load<LP>(ga, nib, sc, xqk, dxk, wstride, R);
      int g = 0;
      for (; g + 2 <= ngroups; g += 2) {
        group_load<LP>(gb, nib + (g + 1) * G * wstride, sc + (g + 1) * G * R, xqk + (g + 1) * G * 4, dxk + (g + 1) * G, wstride, R);
        group_compute<LP>(ga, acc, cvt_mul, cvt_sub);
        if (g + 2 < ngroups)
          group_load<LP>(ga, nib + (g + 2) * G * wstride, sc + (g + 2) * G * R, xqk + (g + 2) * G * 4, dxk + (g + 2) * G, wstride, R);
        group_compute<LP>(gb, acc, cvt_mul, cvt_sub);
      }
      if (g < ngroups) group_compute<LP>(ga, acc, cvt_mul, cvt_sub);
      for (int bl = ngroups * G; bl < cbk; bl++) {   // leftover blocks (cbk not a multiple of G)
        const float sdx = __fmul_rn(sc[bl * R], dxk[bl]);
#pragma unroll
        for (int j = 0; j < LP; j++) {
          const uint32_t wv = nib[bl * wstride + j];
          const uint4 xv = xqk[bl * 4 + j];
          const int ia = dp4a_us(wv & 0x0F0F0F0Fu, (int) xv.x, (int) xv.z);
          const int ib = dp4a_us(wv & 0xF0F0F0F0u, (int) xv.y, (int) xv.w);
          const u64 f = ffma2(pack_i2(ia, ib), cvt_mul, cvt_sub);
          acc[j] = ffma2(pack_f2(sdx, sdx), f, acc[j]);
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sm.empty[s]);
  }

  // horizontal sum exactly as ggml.c:1461-1466: (acc[k]+acc[k+4]) k<4, then (r0+r2)+(r1+r3)
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
  return gchunk;
}

__device__ __forceinline__ void gemv_dispatch(const MatDesc &md, const RowPart rp, const MegaSmem &sm, uint32_t &gchunk,
                                              int S, int stage_bytes, int tid) {
  if (rp.R == 0) { named_bar_sync(1, MEGA_COMPUTE_THREADS); return; }
  GemvCall c;
  c.stages = smem_u32(sm.stages); c.xq = smem_u32(sm.xq); c.dxs = smem_u32(sm.dxs); c.rowres = smem_u32(sm.rowres);
  c.full = smem_u32(sm.full); c.empty = smem_u32(sm.empty);
  c.nb = md.nb; c.cb = md.cb; c.R = rp.R; c.S = S; c.stage_bytes = stage_bytes;
  switch (md.lp) {
    case 1: gchunk = gemv_rows<1>(c, gchunk, tid); break;
    case 2: gchunk = gemv_rows<2>(c, gchunk, tid); break;
    default: gchunk = gemv_rows<4>(c, gchunk, tid); break;
  }
}

// loader side of one matrix: ring stages <- this CTA's contiguous bytes, one cp.async.bulk per chunk
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
    *sm.loader_g = gchunk + 1;
  }
}

// prefetch side of one matrix: keep the stream [loader + S, loader + S + ahead) resident in L2
__device__ __forceinline__ void prefetch_matrix(const MatDesc &md, const MegaSmem &sm, uint32_t &pchunk, int S, int ahead) {
  const RowPart rp = row_part(md.g_total, gridDim.x, blockIdx.x);
  if (rp.R == 0) return;
  const int nchunks = (md.nb + md.cb - 1) / md.cb;
  const uint8_t *wbase = md.w + (size_t) rp.row0 * md.nb * 20;
  for (int k = 0; k < nchunks; k++, pchunk++) {
    uint32_t lg = *sm.loader_g;
    if (pchunk < lg + (uint32_t) S) continue;                 // the loader is (nearly) there already: nothing to gain
    if (pchunk >= lg + (uint32_t) (S + ahead)) {
      const long long t0 = clock64();
      while (pchunk >= (lg = *sm.loader_g) + (uint32_t) (S + ahead)) {
        __nanosleep(256);
        if (clock64() - t0 > 4000000000LL) return;            // loader gone (should not happen): stop prefetching
      }
    }
    const int cbk = min(md.cb, md.nb - k * md.cb);
    l2_prefetch_bulk(wbase + (size_t) k * md.cb * rp.R * 20, (uint32_t) cbk * rp.R * 20);
  }
}

// ---- attention phase for (head h, output quarter qr): K.Q for all positions, soft_max, V.P for 32 dims --------------
__device__ __forceinline__ void attention_phase(const TokenArgs &a, const LayerDesc &L, const MegaSmem &sm, int h, int qr,
                                                int pos, int p_part, int tid) {
  constexpr int HD = 128, NW = MEGA_COMPUTE_WARPS;
  const int lane = tid & 31, warp = tid >> 5;
  const int E = a.n_embd;
  const int p_valid = pos + 1;      // diag_mask_inf: columns > n_past + i are -inf -> probability 0 (ggml.c:6946-6953)
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
      for (int i = 0; i < 4; i++) s = fmaf(kk[u][i], qv[i], s);                  // GGML_F32_VEC_FMA, ggml.c:1239
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 8));                       // sum[0]+sum[1], sum[2]+sum[3]  ggml.c:874-876
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 16));                      // (..)+(..)                     ggml.c:877-879
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));                       // lanes k and k+4               ggml.c:883-884
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));                       // hadd                          ggml.c:885
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));                       // hadd                          ggml.c:886
      s = __fmul_rn(s, a.kq_scale);                                               // ggml_scale, PO.mm:617-621
      if (lane == 0 && j0 + u < p_valid) sc[j0 + u] = s;
    }
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  // soft_max, ggml.c:7019-7041
  float mx = -CUDART_INF_F;
  for (int j = tid; j < p_valid; j += MEGA_COMPUTE_THREADS) mx = fmaxf(mx, sc[j]);
  mx = mega_max_f(mx, sm.redf, 0, tid);
  double sum = 0.0;   // fp16-valued terms: exact in double in any order
  for (int j = tid; j < p_valid; j += MEGA_COMPUTE_THREADS) {
    const uint16_t hx = __half_as_ushort(__float2half_rn(__fsub_rn(sc[j], mx)));
    const float e = __half2float(__ushort_as_half(__ldg(a.exp_table + hx)));
    sc[j] = e;                       // same thread re-reads its own entries below: no barrier needed in between
    sum += (double) e;
  }
  sum = mega_sum_d(sum, sm.redd, 0, tid);
  const float inv = (float) (1.0 / sum);
  for (int j = tid; j < p_valid; j += MEGA_COMPUTE_THREADS) sc[j] = __fmul_rn(sc[j], inv);   // ggml_vec_scale_f32, ggml.c:7041
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
  sm.loader_g = reinterpret_cast<volatile uint32_t *>(sm.empty + S);

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], MEGA_COMPUTE_WARPS); }
    *sm.loader_g = 0;
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= MEGA_COMPUTE_THREADS) {
    if (tid == MEGA_COMPUTE_THREADS) {
      // ===== TMA loader: the whole token's weight stream for this SM, in schedule order =====
      uint32_t g = 0;
      for (int il = 0; il < a.n_layer; il++) {
        const LayerDesc &L = a.layers[il];
        stream_matrix(L.qkv, sm, g, S, stage_bytes);
        stream_matrix(L.wo, sm, g, S, stage_bytes);
        stream_matrix(L.w13, sm, g, S, stage_bytes);
        stream_matrix(L.w2, sm, g, S, stage_bytes);
      }
      stream_matrix(a.out, sm, g, S, stage_bytes);
    } else if (tid == MEGA_COMPUTE_THREADS + 32 && a.l2_ahead > 0) {
      // ===== L2 prefetcher: same schedule, a bounded distance ahead of the loader =====
      uint32_t g = 0;
      for (int il = 0; il < a.n_layer; il++) {
        const LayerDesc &L = a.layers[il];
        prefetch_matrix(L.qkv, sm, g, S, a.l2_ahead);
        prefetch_matrix(L.wo, sm, g, S, a.l2_ahead);
        prefetch_matrix(L.w13, sm, g, S, a.l2_ahead);
        prefetch_matrix(L.w2, sm, g, S, a.l2_ahead);
      }
      prefetch_matrix(a.out, sm, g, S, a.l2_ahead);
    }
    return;
  }

  // ===== compute warps =====
  const int E = a.n_embd, HD = E / a.n_head;
  const int pos = a.sp->pos, p_part = a.sp->p_part, token = a.sp->token;
  uint32_t gchunk = 0;
  unsigned int phase = 0;
  int pm = 0;
  PROF_MARK();   // 0: kernel start
  double xd[MEGA_NORM_ROUNDS][8];

  // get_rows: dequantize_row_q4_0 of the token's embedding row (ggml.c:6760-6785, 651-684) straight into registers
  {
    const uint8_t *row = a.tok_emb + (size_t) token * (E / 32) * 20;
    const int items = E / 8;
#pragma unroll
    for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
      const int it = tid + rd * MEGA_COMPUTE_THREADS;
      if (it < items) {
        const uint8_t *blk = row + (it >> 2) * 20;
        const float d = __ldg(reinterpret_cast<const float *>(blk));
        const uint32_t by = __ldg(reinterpret_cast<const uint32_t *>(blk + 4) + (it & 3));   // 4 bytes = 8 nibbles
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int qn = (by >> (4 * i)) & 0xf;                 // element 2j = low nibble of byte j, 2j+1 = high nibble
          const float v = __fmul_rn((float) (qn - 8), d);
          xd[rd][i] = v;
          if (blockIdx.x == 0) a.inpL[it * 8 + i] = v;          // residual source for layer 0 (read after two grid barriers)
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) xd[rd][i] = 0.0;
      }
    }
  }

  for (int il = 0; il < a.n_layer; il++) {
    const LayerDesc &L = a.layers[il];
    // ---- phase 1: norm -> wq|wk|wv -> rope -> q buffer + KV cache row (PO.mm:570-611) ----
    {
      if (il > 0) load_items(xd, a.inpL, E / 32, tid);
      prologue_norm_regs(xd, L.attn_norm, E / 32, sm, tid);
      PROF_MARK();   // +1: qkv prologue done
      const RowPart rp = row_part(L.qkv.g_total, gridDim.x, blockIdx.x);
      gemv_dispatch(L.qkv, rp, sm, gchunk, S, stage_bytes, tid);
      PROF_MARK();   // +2: qkv rows done
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
      PROF_MARK();   // +3: qkv epilogue done
      grid_barrier(a.bar, phase, tid);
      PROF_MARK();   // +4: barrier 1 passed
    }
    // ---- phase 2: attention (PO.mm:614-646) on the first 4*n_head CTAs ----
    {
      if ((int) blockIdx.x < 4 * a.n_head) attention_phase(a, L, sm, blockIdx.x >> 2, blockIdx.x & 3, pos, p_part, tid);
      PROF_MARK();   // +5: attention done
      grid_barrier(a.bar, phase, tid);
      PROF_MARK();   // +6: barrier 2 passed
    }
    // ---- phase 3: wo, + inpSA (PO.mm:649-654) ----
    {
      prologue_plain(a.att, E / 32, sm, tid);
      PROF_MARK();   // +7: wo prologue done
      const RowPart rp = row_part(L.wo.g_total, gridDim.x, blockIdx.x);
      gemv_dispatch(L.wo, rp, sm, gchunk, S, stage_bytes, tid);
      PROF_MARK();   // +8: wo rows done
      for (int i = tid; i < rp.R; i += MEGA_COMPUTE_THREADS) {
        const int g = rp.row0 + i;
        if (g < L.wo.M) a.inpFF[g] = __fadd_rn(sm.rowres[i], __ldcg(a.inpL + g));   // ggml_add, PO.mm:654
      }
      grid_barrier(a.bar, phase, tid);
      PROF_MARK();   // +9: barrier 3 passed
    }
    // ---- phase 4: norm -> w1|w3 -> silu(w1 x) * (w3 x) (PO.mm:660-680) ----
    {
      load_items(xd, a.inpFF, E / 32, tid);
      prologue_norm_regs(xd, L.ffn_norm, E / 32, sm, tid);
      PROF_MARK();   // +10: w13 prologue done
      const RowPart rp = row_part(L.w13.g_total, gridDim.x, blockIdx.x);
      gemv_dispatch(L.w13, rp, sm, gchunk, S, stage_bytes, tid);
      PROF_MARK();   // +11: w13 rows done
      for (int i = tid; i < rp.R / 2; i += MEGA_COMPUTE_THREADS) {
        const int g = rp.row0 / 2 + i;
        if (2 * g < L.w13.M) {   // fused rows 2i = w1 row i, 2i+1 = w3 row i; silu via the fp16 table (ggml.c:1955-1963)
          const uint16_t hx = __half_as_ushort(__float2half_rn(sm.rowres[2 * i]));
          const float sv = __half2float(__ushort_as_half(__ldg(a.silu_table + hx)));
          a.h[g] = __fmul_rn(sv, sm.rowres[2 * i + 1]);
        }
      }
      grid_barrier(a.bar, phase, tid);
      PROF_MARK();   // +12: barrier 4 passed
    }
    // ---- phase 5: w2, + inpFF (PO.mm:682-687) ----
    {
      prologue_plain(a.h, a.n_ff / 32, sm, tid);
      PROF_MARK();   // +13: w2 prologue done
      const RowPart rp = row_part(L.w2.g_total, gridDim.x, blockIdx.x);
      gemv_dispatch(L.w2, rp, sm, gchunk, S, stage_bytes, tid);
      PROF_MARK();   // +14: w2 rows done
      for (int i = tid; i < rp.R; i += MEGA_COMPUTE_THREADS) {
        const int g = rp.row0 + i;
        if (g < L.w2.M) a.inpL[g] = __fadd_rn(sm.rowres[i], __ldcg(a.inpFF + g));   // ggml_add, PO.mm:687
      }
      grid_barrier(a.bar, phase, tid);
      PROF_MARK();   // +15: barrier 5 passed
    }
  }
  // ---- final norm -> output (PO.mm:694-706) ----
  {
    load_items(xd, a.inpL, E / 32, tid);
    prologue_norm_regs(xd, a.final_norm, E / 32, sm, tid);
    PROF_MARK();   // output prologue done
    const RowPart rp = row_part(a.out.g_total, gridDim.x, blockIdx.x);
    gemv_dispatch(a.out, rp, sm, gchunk, S, stage_bytes, tid);
    PROF_MARK();   // output rows done
    for (int i = tid; i < rp.R; i += MEGA_COMPUTE_THREADS) {
      const int g = rp.row0 + i;
      if (g < a.out.M) a.logits[g] = sm.rowres[i];
    }
    PROF_MARK();   // last: logits stored
  }
}

}  // namespace b200
