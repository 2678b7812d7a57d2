// Host-side callees of the reference's token loop (-[LlamaPredictOperation main], PO.mm:810, 865) -- SURVEY.md section 8f,
// rows N2 (sampler) and N4 (tokenizer).  They are not GPU work; they are here because once a decode step takes 1.6 ms the
// reference's own versions (675 us per sampled token, 530 us per prompt token at n_vocab = 32000) would take a third
// of the end-to-end time through LlamaRunner.  Same results, different data structures:
//
//   b200_llama_tokenize            == llama_tokenize (utils.cpp:275-311): greedy longest match.  The reference scans the
//                                     whole vocabulary map at every text position; here the vocabulary is a byte trie
//                                     and a position costs one walk of at most max-token-length steps.
//   b200_llama_sample_top_p_top_k  == llama_sample_top_p_top_k (utils.cpp:345-428): repetition penalty, top-k, softmax
//                                     in double, top-p, one draw.  The reference tests "is token i among the last n"
//                                     with a linear std::find per logit (n_vocab x 64 comparisons); here that is one
//                                     bitmap lookup.  Everything that decides the result -- the expression order of the
//                                     penalty, std::partial_sort with the same comparator on the same sequence, libm
//                                     exp, std::discrete_distribution on std::mt19937 -- is the same library call, so
//                                     the drawn ids are identical (tests/test_host_text.py checks them against the
//                                     compiled reference).
//
// Plain C ABI (include/b200_llama.h); no CUDA in this file.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200_llama.h"

struct b200_tokenizer {
  // byte trie in struct-of-arrays form: child[node * 256 + byte] (0 = none; node 0 is the root), id_at[node] = the
  // token that ends here (-1 = none).  32000 LLaMA pieces make ~90 k nodes.
  std::vector<int32_t> child;
  std::vector<int32_t> id_at;
  bool has_empty = false;       // the vocabulary contains an empty piece (it can never be emitted, see tokenize)

  int add_node() {
    child.resize(child.size() + 256, 0);
    id_at.push_back(-1);
    return (int) id_at.size() - 1;
  }
  void insert(const char *s, int len, int32_t id) {
    if (len == 0) { has_empty = true; return; }
    int node = 0;
    for (int i = 0; i < len; i++) {
      const unsigned char c = (unsigned char) s[i];
      int next = child[(size_t) node * 256 + c];
      if (next == 0) {
        next = add_node();
        child[(size_t) node * 256 + c] = next;
      }
      node = next;
    }
    // The reference walks id_to_token in ascending id order and replaces its candidate whenever a piece of the same or
    // greater length matches (`size() < l` skips only shorter ones, utils.cpp:293): of two identical pieces the larger
    // id wins.
    id_at[node] = std::max(id_at[node], id);
  }
};

struct b200_rng {
  std::mt19937 engine;          // the generator of PO.mm:773
};

extern "C" {

b200_tokenizer *b200_tokenizer_create_from(const char *const *pieces, const int *lens, int n_vocab) {
  if (!pieces || !lens || n_vocab < 0) return nullptr;
  b200_tokenizer *t = new b200_tokenizer();
  t->add_node();   // root
  for (int id = 0; id < n_vocab; id++) t->insert(pieces[id], lens[id], id);
  return t;
}

b200_tokenizer *b200_tokenizer_create(const b200_llama *m) {
  if (!m) return nullptr;
  const int n = b200_llama_n_vocab(m);
  std::vector<const char *> pieces(n);
  std::vector<int> lens(n);
  for (int id = 0; id < n; id++) pieces[id] = b200_llama_token_str(m, id, &lens[id]);
  return b200_tokenizer_create_from(pieces.data(), lens.data(), n);
}

void b200_tokenizer_free(b200_tokenizer *t) { delete t; }

int b200_llama_tokenize(const b200_tokenizer *t, const char *text, size_t text_len, int bos, int32_t *out, int cap) {
  if (!t || (!text && text_len)) return -1;
  int n = 0;
  auto emit = [&](int32_t id) { if (out && n < cap) out[n] = id; n++; };
  if (bos) emit(1);                                           // utils.cpp:284-286
  size_t pos = 0;
  for (;;) {
    // longest piece that is a prefix of text[pos:]
    int node = 0, best_len = 0;
    int32_t best_id = 0;
    for (size_t i = pos; i < text_len; i++) {
      node = t->child[(size_t) node * 256 + (unsigned char) text[i]];
      if (node == 0) break;
      if (t->id_at[node] >= 0) { best_len = (int) (i - pos) + 1; best_id = t->id_at[node]; }
    }
    if (best_len == 0) break;    // nothing matches (or only an empty piece does): the reference stops here, utils.cpp:301-303
    emit(best_id);
    pos += (size_t) best_len;
  }
  return n;                      // > cap: the caller's buffer was too small, only the first cap ids were written
}

b200_rng *b200_rng_create(int seed) {
  b200_rng *r = new b200_rng();
  r->engine.seed((std::mt19937::result_type) seed);           // std::mt19937 rng(params.seed), PO.mm:773 (seed -1 wraps the same way)
  return r;
}

void b200_rng_free(b200_rng *r) { delete r; }

uint32_t b200_rng_next_u32(b200_rng *r) { return (uint32_t) r->engine(); }   // one raw draw (gpt_random_prompt, utils.cpp:103)

}  // extern "C"

// utils.cpp:388-428: soft-max over the kept candidates, nucleus cut, one draw
static int32_t draw_from_candidates(std::vector<std::pair<double, int32_t>> &cand, double top_p, b200_rng *rng) {
  double top = -INFINITY;
  for (const auto &c : cand) top = std::max(top, c.first);
  std::vector<double> p;
  p.reserve(cand.size());
  double total = 0.0;
  for (const auto &c : cand) {
    const double e = exp(c.first - top);
    p.push_back(e);
    total += e;
  }
  for (double &x : p) x /= total;

  if (top_p < 1.0f) {                                         // nucleus cut, utils.cpp:399-413
    double run = 0.0f;
    for (int i = 0; i < (int) p.size(); i++) {
      run += p[i];
      if (run >= top_p) {
        p.resize((size_t) i + 1);
        cand.resize((size_t) i + 1);
        break;
      }
    }
    run = 1.0 / run;
    for (double &x : p) x *= run;
  }

  std::discrete_distribution<> pick(p.begin(), p.end());
  return cand[(size_t) pick(rng->engine)].second;
}

extern "C" {

int32_t b200_llama_sample_top_p_top_k(int n_vocab, const float *logits, const int32_t *last_n_tokens, int n_last,
                                      double repeat_penalty, int top_k, double top_p, double temp, b200_rng *rng) {
  if (!logits || !rng || n_vocab <= 0 || top_k <= 0) return -1;
  if (top_k > n_vocab) top_k = n_vocab;                       // (the reference would run off the end of its vector)

  std::vector<uint8_t> recent((size_t) n_vocab, 0);
  for (int i = 0; i < n_last; i++) {
    const int32_t id = last_n_tokens[i];
    if (id >= 0 && id < n_vocab) recent[id] = 1;
  }

  // scaled, penalised logits in vocabulary order (utils.cpp:359-374: the products are formed left to right in double)
  std::vector<std::pair<double, int32_t>> cand;
  cand.reserve((size_t) n_vocab);
  const double scale = 1.0 / temp;
  for (int i = 0; i < n_vocab; i++) {
    const float l = logits[i];
    double v;
    if (recent[i]) v = (l < 0.0) ? l * scale * repeat_penalty : l * scale / repeat_penalty;
    else v = l * scale;
    cand.emplace_back(v, i);
  }

  // the k best, in the order the reference's own library call leaves them (ties included): utils.cpp:333-343
  std::partial_sort(cand.begin(), cand.begin() + top_k, cand.end(),
                    [](const std::pair<double, int32_t> &a, const std::pair<double, int32_t> &b) { return a.first > b.first; });
  cand.resize((size_t) top_k);

  return draw_from_candidates(cand, top_p, rng);
}

int32_t b200_llama_sample_from_candidates(const double *cand_values, const int32_t *cand_ids, int n_cand, double top_p,
                                          b200_rng *rng) {
  if (!cand_values || !cand_ids || !rng || n_cand <= 0) return -1;
  std::vector<std::pair<double, int32_t>> cand;
  cand.reserve((size_t) n_cand);
  for (int i = 0; i < n_cand; i++) cand.emplace_back(cand_values[i], cand_ids[i]);
  return draw_from_candidates(cand, top_p, rng);
}

}  // extern "C"
