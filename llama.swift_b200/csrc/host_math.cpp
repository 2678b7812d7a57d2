// Host-side constants that must come from the *same libm expressions* the reference evaluates on the host
// (SURVEY.md Appendix A): the fp16 SiLU / exp tables (ggml.c:2377-2389), the RoPE angles (ggml.c:7114-7117) and the
// attention scale (PO.mm:620).  Compiled by g++ (not nvcc) so the math is the host compiler's and glibc's, as in the
// reference build; the results are uploaded once at load time and only looked up on the GPU.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <immintrin.h>
#include <cmath>
#include <cstdint>

namespace b200 {

// GGML_COMPUTE_FP32_TO_FP16 / FP16_TO_FP32 on x86 with F16C, ggml.c:159-162 (round-to-nearest-even)
static inline uint16_t f32_to_f16(float f) { return _cvtss_sh(f, 0); }
static inline float f16_to_f32(uint16_t h) { return _cvtsh_ss(h); }

// ggml_silu_f32, ggml.c:1944-1946: float in, double exp, result rounded to float
static inline float silu_f32(float x) { return x / (1.0 + exp(-x)); }

void host_build_tables(uint16_t *table_silu_f16, uint16_t *table_exp_f16) {
  for (int i = 0; i < (1 << 16); ++i) {
    const float f = f16_to_f32((uint16_t) i);
    table_silu_f16[i] = f32_to_f16(silu_f32(f));          // ggml.c:2387
    table_exp_f16[i] = f32_to_f16((float) exp(f));         // ggml.c:2388
  }
}

// cs[(p * head_dim/2 + j) * 2 + {0,1}] = {cos, sin}(p * 10000^(-2j/head_dim)), ggml.c:7113-7117.
// The reference object code calls sincos() (gcc merges the cos/sin pair), so this does too.
void host_build_rope(double *cs, int n_ctx, int head_dim) {
  const int n_dims = head_dim;
  for (int p = 0; p < n_ctx; p++) {
    for (int i0 = 0; i0 < n_dims; i0 += 2) {
      const double theta = pow(10000.0, ((double) -i0) / n_dims);
      double s, c;
      sincos(p * theta, &s, &c);
      cs[((size_t) p * (head_dim / 2) + i0 / 2) * 2 + 0] = c;
      cs[((size_t) p * (head_dim / 2) + i0 / 2) * 2 + 1] = s;
    }
  }
}

float host_kq_scale(int n_embd, int n_head) { return 1.0f / sqrt(float(n_embd) / n_head); }   // PO.mm:620

}  // namespace b200
