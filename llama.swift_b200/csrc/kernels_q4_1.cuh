// Q4_1 variant of the decode mat-vec (ggml_compute_forward_mul_mat_q4_1_f32, ggml.c:6287-6585): activation quantizer
// quantize_row_q4_1 (ggml.c:606-648, scalar) + ggml_vec_dot_q4_1 (ggml.c:1584-1626, scalar only in the reference).
//
// The reference dot is ONE sequential f32 chain per output row,
//     sumf += (d0*q0 + m0)*(d1*q2 + m1) + (d0*q1 + m0)*(d1*q3 + m1)      for every byte of every block, in order,
// compiled without contraction (ISO C mode) -- so the exact-parity kernel has one thread per row walking its row in
// order (K/2 dependent adds).  That makes Q4_1 latency-bound by construction; it exists for coverage and parity
// (BASELINE.json config 5), the throughput path is Q4_0.  Same frame as q4_gemv_kernel: TMA loader warp + mbarrier
// ring, prologue (LayerNorm) and epilogues shared.
//
// Device layout of a Q4_1 matrix (load-time re-layout of ggml's per-row [nb m][nb d][nb*16 B] rows, same 24 B per 32
// weights): CTA c's rows contiguous; chunk k = [cbk][R][16 B nibbles in ggml order] [cbk][R] f32 m [cbk][R] f32 d.
#pragma once
#include "kernels.cuh"

namespace b200 {

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
  uint64_t *full = reinterpret_cast<uint64_t *>(red + 32);
  uint64_t *empty = full + S;

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], nt >> 5); }
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

  // ---- row loop: one sequential chain per row, ggml.c:1600-1622 ----
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
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
  }
  if (active) rowres[r] = sumf;
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
