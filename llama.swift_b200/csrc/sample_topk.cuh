// Candidate stage of llama_sample_top_p_top_k (utils.cpp:345-386) on the GPU: repetition penalty + the top_k best scaled
// logits, so that (top_k + 1) (value, id) pairs cross PCIe per token instead of the n_vocab logits (128 KB for LLaMA).
//
// What the reference does with every logit l (utils.cpp:359-374, doubles, products formed left to right):
//     v = l * scale                       scale = 1.0 / temp
//     v = l * scale * repeat_penalty      id in last_n_tokens and l < 0
//     v = l * scale / repeat_penalty      id in last_n_tokens and l >= 0
// then std::partial_sort by v descending, keeps top_k (utils.cpp:333-343, 376).  The order partial_sort leaves EQUAL values in
// is a property of the library's heap, so this kernel only answers when the order is unambiguous: it returns the top_k + 1
// best pairs and a flag that is set when any two of them compare equal (or a value is NaN); the host then falls back to
// the full logits and the reference's own library call.  Otherwise "sorted by v descending, all distinct" has one answer.
//
// One CTA of 1024 threads (the logits are 128 KB and L2-resident right after the output mat-vec):
//   1. every thread scans its strided share of the logits and keeps the largest order-preserving 32-bit key among its
//      NOT-penalised ids (their v is strictly monotone in l: distinct floats stay distinct after the double multiply by a
//      positive scale)
//   2. radix select (4 passes x 8 bits, per-warp shared-memory histograms) of the (top_k + 1)-th largest of the 1024 thread
//      maxima: at least top_k + 1 logits are >= that key, so every one of the best top_k + 1 is too
//   3. candidates = every not-penalised logit with a key >= it (top_k + 1 plus the few that share a thread with a larger one)
//      + every penalised id (<= n_last of them); a second pass over the L2-resident logits
//   4. their doubles, ranked by counting: rank = number of candidates with a greater value
// (A first version radix-selected over all n_vocab keys in shared memory with __match_any_sync-aggregated histograms: 137 us.)
#pragma once
#include <cstdint>

namespace b200 {

constexpr int TOPK_MAX_K = 64;            // top_k <= 64 (the reference's default is 40, PO.mm:852)
constexpr int TOPK_MAX_LAST = 256;        // repeat_last_n <= 256 (default 64)
constexpr int TOPK_MAX_CAND = 512;        // candidate list capacity; more (massive ties) -> ambiguous
constexpr int TOPK_THREADS = 1024;

struct TopkOut {
  double values[TOPK_MAX_K + 1];
  int32_t ids[TOPK_MAX_K + 1];
  int32_t n;                              // pairs written (min(top_k + 1, n_vocab))
  int32_t ambiguous;                      // 1: equal values / NaN among them, or too many candidates: use the host path
};

// dynamic shared memory: the bitmap of the repetition window
__host__ __device__ __forceinline__ size_t topk_smem_bytes(int n_vocab) { return (size_t) ((n_vocab + 31) / 32) * 4; }

__global__ void __launch_bounds__(TOPK_THREADS) sample_topk_kernel(const float *__restrict__ logits, int n_vocab,
                                                                   const int32_t *__restrict__ last_n, int n_last, double scale,
                                                                   double penalty, int top_k, TopkOut *out) {
  extern __shared__ __align__(16) uint8_t smem_topk[];
  uint32_t *recent = reinterpret_cast<uint32_t *>(smem_topk);          // bitmap of last_n_tokens
  const int n_words = (n_vocab + 31) / 32;
  __shared__ uint32_t hist[32 * 256];                                  // [warp][bin]
  __shared__ uint32_t tot[256];
  __shared__ double cval[TOPK_MAX_CAND];
  __shared__ int32_t cid[TOPK_MAX_CAND];
  __shared__ uint32_t s_prefix, s_want, s_ncand;
  __shared__ int s_amb;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < n_words; i += TOPK_THREADS) recent[i] = 0;
  if (tid == 0) { s_ncand = 0; s_amb = 0; }
  __syncthreads();
  for (int i = tid; i < n_last; i += TOPK_THREADS) {
    const int id = last_n[i];
    if (id >= 0 && id < n_vocab) atomicOr(&recent[id >> 5], 1u << (id & 31));
  }
  __syncthreads();
  // keys: larger float <-> larger unsigned; 0 = "not part of the selection" (penalised ids; a real key is 0 only for a NaN)
  uint32_t best = 0;
  bool nan_seen = false;
  for (int i = tid; i < n_vocab; i += TOPK_THREADS) {
    const uint32_t u = __float_as_uint(__ldg(logits + i));
    nan_seen |= (u & 0x7fffffffu) > 0x7f800000u;
    uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    if ((recent[i >> 5] >> (i & 31)) & 1u) key = 0;
    best = max(best, key);
  }
  // A NaN logit anywhere: the host path decides.  This flag is needed, not belt and braces: measured on B200, a NaN never
  // makes it into the candidate list below although its key is the largest (+Inf and 3e38 do) -- the compiled
  // "key >= threshold" behaves like the float comparison it is equivalent to for every non-NaN value.
  if (nan_seen) s_amb = 1;
  {
    const int have = __syncthreads_count(best != 0);                   // threads that own a selectable logit
    if (tid == 0) { s_want = (uint32_t) min(top_k + 1, have); s_prefix = 0; }
    __syncthreads();
  }
  uint32_t mask = 0;
  if (s_want > 0) {
    for (int pass = 3; pass >= 0; pass--) {
      const int sh = pass * 8;
      for (int i = tid; i < 32 * 256; i += TOPK_THREADS) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      if (best != 0 && (best & mask) == prefix) atomicAdd(&hist[warp * 256 + ((best >> sh) & 255u)], 1u);
      __syncthreads();
      if (tid < 256) {
        uint32_t c = 0;
#pragma unroll 8
        for (int w = 0; w < 32; w++) c += hist[w * 256 + tid];
        tot[tid] = c;
      }
      __syncthreads();
      // warp 0: bins from the top until the wanted count is covered
      if (warp == 0) {
        uint32_t loc[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          loc[j] = tot[255 - (lane * 8 + j)];                          // lane 0 owns the 8 highest bins
          sum += loc[j];
        }
        uint32_t incl = sum;                                           // inclusive scan over lanes (from the top)
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const uint32_t want = s_want, before = incl - sum;
        if (before < want && incl >= want) {                           // exactly one lane
          uint32_t cum = before;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (cum < want && cum + loc[j] >= want) {
              s_prefix = prefix | ((uint32_t) (255 - (lane * 8 + j)) << sh);
              s_want = want - cum;                                     // still wanted among the keys of this bin
            }
            cum += loc[j];
          }
        }
      }
      mask |= 255u << sh;
      __syncthreads();
    }
  }
  // candidates: every selectable key >= threshold, and every penalised id
  const uint32_t thr = s_prefix;
  const bool any = s_want > 0;
  for (int i = tid; i < n_vocab; i += TOPK_THREADS) {
    const float l = __ldg(logits + i);
    const uint32_t u = __float_as_uint(l);
    const uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    const bool pen = (recent[i >> 5] >> (i & 31)) & 1u;
    if (pen || (any && key != 0 && key >= thr)) {
      const uint32_t pos = atomicAdd(&s_ncand, 1u);
      if (pos < (uint32_t) TOPK_MAX_CAND) {
        double v;
        if (pen) v = (l < 0.0) ? __dmul_rn(__dmul_rn((double) l, scale), penalty) : __ddiv_rn(__dmul_rn((double) l, scale), penalty);
        else v = __dmul_rn((double) l, scale);
        cval[pos] = v;
        cid[pos] = i;
      }
    }
  }
  __syncthreads();
  const int nc = (int) min(s_ncand, (uint32_t) TOPK_MAX_CAND);
  if (s_ncand > (uint32_t) TOPK_MAX_CAND && tid == 0) s_amb = 1;
  const int n_out = min(top_k + 1, nc);
  for (int c = tid; c < nc; c += TOPK_THREADS) {
    const double v = cval[c];
    int rank = 0, eq = 0;
    for (int j = 0; j < nc; j++) {
      const double w = cval[j];
      rank += w > v;
      eq += (j != c) && (w == v);
    }
    const bool nan = !(v == v);
    if (nan) s_amb = 1;
    if (rank < n_out) {
      if (eq > 0) s_amb = 1;
      out->values[rank] = v;
      out->ids[rank] = cid[c];
    }
  }
  __syncthreads();
  if (tid == 0) {
    out->n = n_out;
    out->ambiguous = s_amb;
  }
}

}  // namespace b200
