// Q4_1 variant of the decode mat-vec (ggml_compute_forward_mul_mat_q4_1_f32, ggml.c:6287-6585): activation quantizer
// quantize_row_q4_1 (ggml.c:606-648, scalar) + ggml_vec_dot_q4_1 (ggml.c:1584-1626, scalar only in the reference).
//
// The reference dot is ONE sequential f32 chain per output row,
//     sumf += (d0*q0 + m0)*(d1*q2 + m1) + (d0*q1 + m0)*(d1*q3 + m1)      for every byte of every block, in order,
// compiled without contraction (ISO C mode).  Only the "sumf +=" is a chain: the bracketed TERM of every byte is independent
// work.  So the kernel splits the two: TERM warps (lane = row, one (row group, block) item at a time) compute the 16 terms of
// a block -- 13 instructions per byte -- into a double-buffered shared-memory tile, and CHAIN warps (lane = row) add them up
// in the reference's order, K/2 dependent fadd per row and nothing else on that path.  The floor of a Q4_1 mat-vec is that
// chain: K/2 x the fadd latency (~4.5 cycles) = 4.8 us for K = 4096, 13 us for K = 11008, whatever the bandwidth.
// (Round 1 walked the whole row with one thread: 35-82 us per matrix, one busy warp per SM.)
// Same frame as q4_gemv_kernel: TMA loader warp + mbarrier ring, prologue (LayerNorm) and epilogues shared.
//
// Device layout of a Q4_1 matrix (load-time re-layout of ggml's per-row [nb m][nb d][nb*16 B] rows, same 24 B per 32
// weights): CTA c's rows contiguous; chunk k = [cbk][R][16 B nibbles in ggml order] [cbk][R] f32 m [cbk][R] f32 d.
#pragma once
#include "kernels.cuh"

namespace b200 {

#ifndef B200_Q41_F32X2
#define B200_Q41_F32X2 0       // 1 = term warps with packed mul.rn.f32x2 / add.rn.f32x2 (9 instead of 13 instructions per byte): 4 % faster but NOT
                               // bit-identical to the oracle (tests/test_gpu_parity.py -k q4_1 fails): ptxas 12.9 CONTRACTS the packed pair
                               // mul.rn.f32x2 -> add.rn.f32x2 into one FFMA2 (SASS: FFMA2 R, d0, q, m0) although both carry .rn and the build
                               // uses -fmad=false -- the scalar mul.rn.f32 / add.rn.f32 pair is left alone.  2 = the same with the "+ m0" as two scalar
                               // adds: bit-identical again, and no faster than the scalar form (20.5 vs 19.9 us, 38.7 vs 40.9 us).  So 0.
#endif

// Few rows per CTA (wo, w2: 28): term / chain split.  Many rows per CTA (fused wq|wk|wv 84, w1|w3 152, output 220): there are
// enough rows to fill the SM with one thread per row, and the split's per-chunk hand-overs cost more than they hide (measured:
// 32000 x 4096 52 us per-row vs 94 us split; 4096 x 11008 82 us per-row vs 43 us split).
__host__ __device__ __forceinline__ bool q41_term_mode(int rmax) { return ((rmax + 31) & ~31) <= 64; }

template <int PRO, int EPI>
__global__ void __launch_bounds__(544, 1) q4_1_gemv_kernel(const GemvArgs a) {
  extern __shared__ __align__(128) uint8_t smem_q41[];
  uint8_t *smem = smem_q41;
  const int tid = threadIdx.x;
  const int nt = blockDim.x - 32;
  const int nb = a.nb;
  const RowPart rp = row_part(a.g_total, gridDim.x, blockIdx.x);
  const int R = rp.R;
  const int nchunks = (nb + a.cb - 1) / a.cb;
  const int S = a.n_stages;
  const int K = nb * 32;

  uint8_t *stages = smem;
  float *ys = reinterpret_cast<float *>(smem + (size_t) S * a.stage_bytes);     // [K] dequantized activation d1*q + m1
  float *rowres = ys + K;                                                        // [rmax]
  double *red = reinterpret_cast<double *>(rowres + ((a.rmax + 3) & ~3));        // [32]
  const int r_pad = (a.rmax + 31) & ~31;                                         // rows incl. the idle lanes of the last row group
  const int tstride = a.cb * 16 + 4;                                             // floats per row of a term tile; tstride / 4 odd: LDS.128 / STS.128 by 32 rows conflict-free
  const bool term_mode = q41_term_mode(a.rmax);
  float *terms = reinterpret_cast<float *>(red + 32);                            // [2][r_pad][tstride] (term mode only)
  uint64_t *full = reinterpret_cast<uint64_t *>(terms + (term_mode ? (size_t) 2 * r_pad * tstride : (size_t) 0));
  uint64_t *empty = full + S;
  uint64_t *tfull = empty + S, *tempty = tfull + 2;                              // term tiles: term warps <-> chain warps
  const int n_cw = r_pad >> 5;                                                   // chain warps = row groups (per-row mode: the only busy warps)
  const int n_tw = term_mode ? (nt >> 5) - n_cw : n_cw;                          // warps that read the weight stages

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], n_tw); }
    for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], n_tw); mbar_init(&tempty[i], n_cw); }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= nt) {
    if (tid == nt) {
      const uint8_t *wbase = a.w + (size_t) rp.row0 * nb * 24;
      for (int k = 0; k < nchunks; k++) {
        const int s = k % S;
        if (k >= S && !mbar_wait(&empty[s], ((k / S) - 1) & 1)) return;
        const int cbk = min(a.cb, nb - k * a.cb);
        const uint32_t bytes = (uint32_t) cbk * R * 24;
        mbar_arrive_expect_tx(&full[s], bytes);
        tma_bulk_g2s(stages + (size_t) s * a.stage_bytes, wbase + (size_t) k * a.cb * R * 24, bytes, &full[s]);
      }
    }
    return;
  }

  pdl_launch_dependents();
  pdl_wait();
  const float *__restrict__ x = a.x;

  // ---- prologue: (LayerNorm * weight), then quantize_row_q4_1 and immediate dequantization of the activation ----
  double mean = 0.0;
  float nscale = 1.0f;
  if (PRO == PRO_NORM) {   // ggml_compute_forward_norm_f32, ggml.c:5363-5381
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
    nscale = (float) (1.0 / sqrt(__dadd_rn(s2 / (double) K, (double) 1e-5f)));
  }
  for (int b = tid; b < nb; b += nt) {
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
          const float y = (float) __dsub_rn((double) v[4 * i + j], mean);
          v[4 * i + j] = __fmul_rn(ww[j], __fmul_rn(y, nscale));
        }
      }
    }
    // quantize_row_q4_1, ggml.c:617-646
    float mn = 3.402823466e+38F, mx = -3.402823466e+38F;
#pragma unroll
    for (int i = 0; i < 32; i++) { mn = v[i] < mn ? v[i] : mn; mx = v[i] > mx ? v[i] : mx; }
    const float d = __fdiv_rn(__fsub_rn(mx, mn), 15.0f);
    const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i++) {
      const float t = __fmul_rn(__fsub_rn(v[i], mn), id);
      const int q = (int) (uint8_t) roundf(t);                       // C round(): half away from zero
      ys[b * 32 + i] = __fadd_rn(__fmul_rn(d, (float) q), mn);       // the activation factor d1*q + m1 of ggml.c:1617-1618
    }
  }
  named_bar_sync(1, nt);

  // ---- row loop, ggml.c:1600-1622 ----
  const int warp = tid >> 5, lane = tid & 31;
  if (!term_mode) {
    // one thread per row walks its row in order
    if (warp < n_cw) {
      const bool active = tid < R;
      const int r = active ? tid : R - 1;
      float sumf = 0.0f;
      for (int k = 0; k < nchunks; k++) {
        const int s = k % S;
        mbar_wait(&full[s], (k / S) & 1);
        const int cbk = min(a.cb, nb - k * a.cb);
        const uint8_t *st = stages + (size_t) s * a.stage_bytes;
        const uint4 *nib = reinterpret_cast<const uint4 *>(st) + r;
        const float *pm = reinterpret_cast<const float *>(st + (size_t) cbk * R * 16) + r;
        const float *pd = pm + (size_t) cbk * R;
        const float *yk = ys + (size_t) k * a.cb * 32;
        for (int bl = 0; bl < cbk; bl++) {
          const uint4 wv = nib[(size_t) bl * R];
          const float m0 = pm[bl * R], d0 = pd[bl * R];
          const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const uint32_t by = (ww[j >> 2] >> (8 * (j & 3))) & 0xffu;
            const float f0 = __fadd_rn(__fmul_rn(d0, (float) (by & 0xfu)), m0);
            const float f1 = __fadd_rn(__fmul_rn(d0, (float) (by >> 4)), m0);
            const float f2 = yk[bl * 32 + 2 * j], f3 = yk[bl * 32 + 2 * j + 1];
            sumf = __fadd_rn(sumf, __fadd_rn(__fmul_rn(f0, f2), __fmul_rn(f1, f3)));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      if (active) rowres[r] = sumf;
    }
  } else if (warp < n_cw) {
    // CHAIN warp: lane = row; sumf += term, byte by byte, block by block -- the reference's own order
    const int r = warp * 32 + lane;
    float sumf = 0.0f;
    const float *trow = terms + (size_t) r * tstride;
    for (int k = 0; k < nchunks; k++) {
      const int tb = k & 1;
      mbar_wait(&tfull[tb], (k >> 1) & 1);
      const int cbk = min(a.cb, nb - k * a.cb);
      const float4 *t4 = reinterpret_cast<const float4 *>(trow + (size_t) tb * r_pad * tstride);
      for (int i = 0; i < cbk * 4; i += 4) {          // 16 terms in flight ahead of the chain
        const float4 v0 = t4[i], v1 = t4[i + 1], v2 = t4[i + 2], v3 = t4[i + 3];
        sumf = __fadd_rn(sumf, v0.x); sumf = __fadd_rn(sumf, v0.y); sumf = __fadd_rn(sumf, v0.z); sumf = __fadd_rn(sumf, v0.w);
        sumf = __fadd_rn(sumf, v1.x); sumf = __fadd_rn(sumf, v1.y); sumf = __fadd_rn(sumf, v1.z); sumf = __fadd_rn(sumf, v1.w);
        sumf = __fadd_rn(sumf, v2.x); sumf = __fadd_rn(sumf, v2.y); sumf = __fadd_rn(sumf, v2.z); sumf = __fadd_rn(sumf, v2.w);
        sumf = __fadd_rn(sumf, v3.x); sumf = __fadd_rn(sumf, v3.y); sumf = __fadd_rn(sumf, v3.z); sumf = __fadd_rn(sumf, v3.w);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[tb]);
    }
    if (r < R) rowres[r] = sumf;
  } else {
    // TERM warp: items (row group, block) of the chunk, lane = row of the group
    const int tw = warp - n_cw;
    for (int k = 0; k < nchunks; k++) {
      const int s = k % S, tb = k & 1;
      mbar_wait(&full[s], (k / S) & 1);
      if (k >= 2) mbar_wait(&tempty[tb], ((k >> 1) - 1) & 1);
      const int cbk = min(a.cb, nb - k * a.cb);
      const uint8_t *st = stages + (size_t) s * a.stage_bytes;
      const float *yk = ys + (size_t) k * a.cb * 32;
      float *tt = terms + (size_t) tb * r_pad * tstride;
      for (int it = tw; it < n_cw * cbk; it += n_tw) {
        const int rg = it / cbk, bl = it - rg * cbk;
        const int r = rg * 32 + lane;
        const int rr = r < R ? r : R - 1;              // idle lanes of the last group recompute its last row (never read)
        const uint4 wv = reinterpret_cast<const uint4 *>(st)[(size_t) bl * R + rr];
        const float m0 = reinterpret_cast<const float *>(st + (size_t) cbk * R * 16)[bl * R + rr];
        const float d0 = reinterpret_cast<const float *>(st + (size_t) cbk * R * 20)[bl * R + rr];
        const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
        const float4 *y4 = reinterpret_cast<const float4 *>(yk + bl * 32);
        float4 *dst = reinterpret_cast<float4 *>(tt + (size_t) r * tstride + bl * 16);
#if B200_Q41_F32X2 != 0
        const u64 d2 = pack_f2(d0, d0), m2 = pack_f2(m0, m0), magic = pack_f2(8388608.0f, 8388608.0f);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          float t[4];
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const float4 y = y4[c * 2 + h];            // activation factors d1*q + m1 of 2 bytes (same address for the whole warp)
            const u64 yy[2] = {pack_f2(y.x, y.y), pack_f2(y.z, y.w)};
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int j = h * 2 + e;                 // byte j of word c
              // (float) of both nibbles without the conversion unit (0x4B000000 | q is 8388608 + q), then the reference's
              // (d0*q + m0) * y per nibble as packed multiplies / adds, and the sum of the two products
              const u64 qb = pack_i2((int) (((ww[c] >> (8 * j)) & 0xfu) | 0x4B000000u), (int) (((ww[c] >> (8 * j + 4)) & 0xfu) | 0x4B000000u));
              const u64 q2 = fadd2(qb, pack_f2(-8388608.0f, -8388608.0f));
#if B200_Q41_F32X2 == 2
              // the "+ m0" as two scalar adds: a packed add behind the packed multiply gets contracted into one FFMA2 by ptxas
              float g0, g1;
              unpack_f2(fmul2(d2, q2), g0, g1);
              const u64 f2 = pack_f2(__fadd_rn(g0, m0), __fadd_rn(g1, m0));
#else
              const u64 f2 = fadd2(fmul2(d2, q2), m2);
#endif
              float p0, p1;
              unpack_f2(fmul2(f2, yy[e]), p0, p1);
              t[j] = __fadd_rn(p0, p1);
            }
          }
          dst[c] = make_float4(t[0], t[1], t[2], t[3]);
        }
        (void) magic;
#else
#pragma unroll
        for (int c = 0; c < 4; c++) {
          float t[4];
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const float4 y = y4[c * 2 + h];            // activation factors d1*q + m1 of 2 bytes (same address for the whole warp)
            const float yy[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int j = h * 2 + e;                 // byte j of word c
              // (float) nibble without the conversion unit: 0x4B000000 | q is 8388608 + q
              const float q0 = __fsub_rn(__uint_as_float(((ww[c] >> (8 * j)) & 0xfu) | 0x4B000000u), 8388608.0f);
              const float q1 = __fsub_rn(__uint_as_float(((ww[c] >> (8 * j + 4)) & 0xfu) | 0x4B000000u), 8388608.0f);
              const float f0 = __fadd_rn(__fmul_rn(d0, q0), m0);
              const float f1 = __fadd_rn(__fmul_rn(d0, q1), m0);
              t[j] = __fadd_rn(__fmul_rn(f0, yy[2 * e]), __fmul_rn(f1, yy[2 * e + 1]));
            }
          }
          dst[c] = make_float4(t[0], t[1], t[2], t[3]);
        }
#endif
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(&empty[s]); mbar_arrive(&tfull[tb]); }
    }
  }
  named_bar_sync(1, nt);
  gemv_epilogue<EPI>(a, rp, rowres, tid, nt);
}

// dequantize_row_q4_1 of the raw ggml embedding row ([nb m][nb d][nb*16 B]), ggml.c:686-717: y = q*d + m (mul, add)
__global__ void embed_q4_1_kernel(const uint8_t *tok_emb_raw, const StepParams *sp, float *out, int n_embd) {
  pdl_launch_dependents();
  pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_embd) return;
  const int nb = n_embd / 32;
  const uint8_t *row = tok_emb_raw + (size_t) sp->token * nb * 24;
  const float *pm = reinterpret_cast<const float *>(row);
  const float *pd = pm + nb;
  const uint8_t *pb = reinterpret_cast<const uint8_t *>(pd + nb);
  const int b = e / 32;
  const uint8_t by = pb[b * 16 + (e % 32) / 2];
  const int qn = (e & 1) ? (by >> 4) : (by & 0xf);
  out[e] = __fadd_rn(__fmul_rn((float) qn, pd[b]), pm[b]);
}

// load-time re-layout: ggml Q4_1 rows -> per-CTA chunked layout (see the header comment)
__global__ void repack_q4_1_kernel(const uint8_t *src, uint8_t *dst, int M, int g_total, int nb, int cb, int n_cta,
                                   int interleave_half) {
  const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long) g_total * 4 * nb;
  if (idx >= total) return;
  const int gr = (int) (idx / nb), b = (int) (idx % nb);
  const int c = cta_of_granule(g_total, n_cta, gr / 4);
  const RowPart rp = row_part(g_total, n_cta, c);
  const int r = gr - rp.row0, R = rp.R;
  const int k = b / cb, bl = b % cb;
  const int cbk = min(cb, nb - k * cb);
  uint8_t *chunk = dst + (size_t) rp.row0 * nb * 24 + (size_t) k * cb * R * 24;
  uint4 *dn = reinterpret_cast<uint4 *>(chunk) + (size_t) bl * R + r;
  float *dm = reinterpret_cast<float *>(chunk + (size_t) cbk * R * 16) + (size_t) bl * R + r;
  float *dd = dm + (size_t) cbk * R;
  if (gr >= M) { *dn = make_uint4(0, 0, 0, 0); *dm = 0.0f; *dd = 0.0f; return; }
  const int sr = interleave_half > 0 ? ((gr & 1) ? interleave_half + gr / 2 : gr / 2) : gr;
  const uint8_t *row = src + (size_t) sr * nb * 24;
  const float *pm = reinterpret_cast<const float *>(row);
  const float *pd = pm + nb;
  const uint32_t *pb = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(pd + nb) + (size_t) b * 16);
  *dn = make_uint4(pb[0], pb[1], pb[2], pb[3]);
  *dm = pm[b];
  *dd = pd[b];
}

}  // namespace b200
