// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
//
// C-ABI wrapper around the *unmodified* reference: Sources/cpp/ggml.c and
// Sources/cpp/utils.cpp are compiled where they lie under /root/reference, and
// the reference driver (llama_model_load / llama_eval, PO.mm:98-735) comes in
// through the generated oracle/_ref/llama_ref_tu.inc (see gen_ref_tu.py).
// Everything in this file is glue written for this repository: it only
// forwards to the reference's own functions so that tests/ and bench.py's
// reference arm can call them through ctypes.  Built into oracle/_ref/libllama_ref.so
// by oracle/Makefile; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.

#include <cassert>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "ggml.h"
#include "utils.h"

// ---- stand-ins for the Foundation types the driver mentions -----------------
enum LlamaErrorCode {
  LlamaErrorCodeUnknown = -1,
  LlamaErrorCodeFailedToLoadModel = -1000,   // headers/LlamaError.h:17
  LlamaErrorCodePredictionFailed = -1001,    // headers/LlamaError.h:18
};

struct NSError {
  int code;
  std::string message;
};

static NSError *ref_make_error(LlamaErrorCode code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return new NSError{(int) code, buf};
}

#include "llama_ref_tu.inc"   // PO.mm:32-735, generated

// ---- C ABI --------------------------------------------------------------------
struct ref_llama {
  llama_model model;
  gpt_vocab vocab;
  size_t mem_per_token = 0;
  std::vector<float> logits;
};

static void put_err(NSError *e, char *err, size_t errlen) {
  if (err && errlen) {
    snprintf(err, errlen, "%s", e ? e->message.c_str() : "unknown error");
  }
  delete e;
}

extern "C" {

ref_llama *ref_llama_load(const char *path, int n_ctx, char *err, size_t errlen) {
  ggml_time_init();
  ref_llama *h = new ref_llama();
  NSError *e = nullptr;
  bool ok = false;
  try {
    ok = llama_model_load(path, h->model, h->vocab, n_ctx, &e);
  } catch (const std::exception &ex) {   // LLAMA_N_PARTS.at() throws for unknown n_embd (PO.mm:136)
    e = new NSError{LlamaErrorCodeFailedToLoadModel, ex.what()};
  }
  if (!ok) {
    put_err(e, err, errlen);
    delete h;
    return nullptr;
  }
  return h;
}

int ref_llama_eval(ref_llama *h, int n_threads, int n_past, const int32_t *tokens, int n_tokens,
                   float *logits_out, char *err, size_t errlen) {
  std::vector<gpt_vocab::id> embd(tokens, tokens + n_tokens);
  NSError *e = nullptr;
  if (!llama_eval(h->model, n_threads, n_past, embd, h->logits, h->mem_per_token, &e)) {
    put_err(e, err, errlen);
    return LlamaErrorCodePredictionFailed;
  }
  memcpy(logits_out, h->logits.data(), sizeof(float) * h->model.hparams.n_vocab);
  return 0;
}

void ref_llama_free(ref_llama *h) {
  if (!h) return;
  ggml_free(h->model.ctx);   // PO.mm:900
  delete h;
}

int ref_llama_n_vocab(const ref_llama *h) { return h->model.hparams.n_vocab; }
int ref_llama_n_ctx(const ref_llama *h)   { return h->model.hparams.n_ctx; }
int ref_llama_n_embd(const ref_llama *h)  { return h->model.hparams.n_embd; }
int ref_llama_n_layer(const ref_llama *h) { return h->model.hparams.n_layer; }
int ref_llama_n_head(const ref_llama *h)  { return h->model.hparams.n_head; }

// KV cache rows [0, n_rows) of one layer, reference layout [n_ctx][n_embd] f32 (PO.mm:300-301,586-587).
void ref_llama_kv_export(const ref_llama *h, int layer, int which /*0=K,1=V*/, int n_rows, float *out) {
  const auto &hp = h->model.hparams;
  const ggml_tensor *t = which == 0 ? h->model.memory_k : h->model.memory_v;
  const float *base = (const float *) t->data + (size_t) layer * hp.n_ctx * hp.n_embd;
  memcpy(out, base, sizeof(float) * (size_t) n_rows * hp.n_embd);
}

void ref_llama_kv_import(ref_llama *h, int layer, int which, int n_rows, const float *in) {
  const auto &hp = h->model.hparams;
  ggml_tensor *t = which == 0 ? h->model.memory_k : h->model.memory_v;
  float *base = (float *) t->data + (size_t) layer * hp.n_ctx * hp.n_embd;
  memcpy(base, in, sizeof(float) * (size_t) n_rows * hp.n_embd);
}

const char *ref_llama_token_str(const ref_llama *h, int id, int *len) {
  auto it = h->vocab.id_to_token.find(id);
  if (it == h->vocab.id_to_token.end()) { *len = 0; return ""; }
  *len = (int) it->second.size();
  return it->second.data();
}

// ---- single-op graphs through the reference's own ggml_graph_compute ----------
// type: 2 = Q4_0, 3 = Q4_1 (file ftype numbering, PO.mm:172-173).  W is M rows of
// K/32 blocks in ggml's own layout; x is N columns of K floats; out is [N][M].
int ref_mul_mat_q4(int type, const void *W, int M, int K, const float *x, int N, float *out, int n_threads) {
  const ggml_type wt = type == 2 ? GGML_TYPE_Q4_0 : GGML_TYPE_Q4_1;
  const size_t wbytes = (size_t) M * K / 32 * (type == 2 ? 20 : 24);
  const size_t need = wbytes + (size_t) N * K * 4 * 2 + (size_t) N * M * 4 + (64u << 20);
  std::vector<uint8_t> buf(need);
  ggml_init_params ip = { need, buf.data() };
  ggml_context *ctx = ggml_init(ip);
  if (!ctx) return -1;
  ggml_tensor *a = ggml_new_tensor_2d(ctx, wt, K, M);
  memcpy(a->data, W, wbytes);
  ggml_tensor *b = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, K, N);
  memcpy(b->data, x, (size_t) N * K * 4);
  ggml_tensor *c = ggml_mul_mat(ctx, a, b);
  ggml_cgraph gf = {};
  gf.n_threads = n_threads;
  ggml_build_forward_expand(&gf, c);
  ggml_graph_compute(ctx, &gf);
  memcpy(out, c->data, (size_t) N * M * 4);
  ggml_free(ctx);
  return 0;
}

// op: 0 norm, 1 silu, 2 soft_max (rows of ne0), 3 rope mode 0 (x = [n_dims, n_head, rows]), 4 rope mode 1
int ref_unary_op(int op, const float *x, int ne0, int ne1, int ne2, int n_past, float *out, int n_threads) {
  const size_t n = (size_t) ne0 * ne1 * ne2;
  const size_t need = n * 4 * 4 + (16u << 20);
  std::vector<uint8_t> buf(need);
  ggml_init_params ip = { need, buf.data() };
  ggml_context *ctx = ggml_init(ip);
  if (!ctx) return -1;
  ggml_tensor *a = ggml_new_tensor_3d(ctx, GGML_TYPE_F32, ne0, ne1, ne2);
  memcpy(a->data, x, n * 4);
  ggml_tensor *r = nullptr;
  switch (op) {
    case 0: r = ggml_norm(ctx, a); break;
    case 1: r = ggml_silu(ctx, a); break;
    case 2: r = ggml_soft_max(ctx, a); break;
    case 3: r = ggml_rope(ctx, a, n_past, ne0, 0); break;
    case 4: r = ggml_rope(ctx, a, n_past, ne0, 1); break;
    default: ggml_free(ctx); return -2;
  }
  ggml_cgraph gf = {};
  gf.n_threads = n_threads;
  ggml_build_forward_expand(&gf, r);
  ggml_graph_compute(ctx, &gf);
  memcpy(out, r->data, n * 4);
  ggml_free(ctx);
  return 0;
}

// row (de)quantizers are external symbols of ggml.c (ggml.c:404,606,651,686)
void quantize_row_q4_0(const float *x, void *y, int k);
void quantize_row_q4_1(const float *x, void *y, int k);
void dequantize_row_q4_0(const void *x, float *y, int k);
void dequantize_row_q4_1(const void *x, float *y, int k);

void ref_quantize_row(int type, const float *x, void *y, int k) {
  if (type == 2) quantize_row_q4_0(x, y, k); else quantize_row_q4_1(x, y, k);
}
void ref_dequantize_row(int type, const void *x, float *y, int k) {
  if (type == 2) dequantize_row_q4_0(x, y, k); else dequantize_row_q4_1(x, y, k);
}

// the offline quantizer used by quantize.cpp:171-185 (utils.cpp:431-544); n = total elements, k = row length
size_t ref_quantize_weights(int type, float *src, void *dst, int n, int k) {
  std::vector<int64_t> hist(16, 0);
  return type == 2 ? ggml_quantize_q4_0(src, dst, n, k, 32, hist.data())
                   : ggml_quantize_q4_1(src, dst, n, k, 32, hist.data());
}

// host-side pieces of the token loop (utils.cpp:275-311, 345-428) for the "next" rows N2/N4
int ref_tokenize(const ref_llama *h, const char *text, int bos, int32_t *out, int cap) {
  std::vector<gpt_vocab::id> t = ::llama_tokenize(h->vocab, text, bos != 0);
  int n = (int) t.size();
  for (int i = 0; i < n && i < cap; i++) out[i] = t[i];
  return n;
}

struct ref_sampler { std::mt19937 rng; };
ref_sampler *ref_sampler_new(int seed) { return new ref_sampler{std::mt19937(seed)}; }
void ref_sampler_free(ref_sampler *s) { delete s; }
int ref_sample_top_p_top_k(const ref_llama *h, ref_sampler *s, const float *logits,
                           const int32_t *last_n, int n_last,
                           double repeat_penalty, int top_k, double top_p, double temp) {
  std::vector<gpt_vocab::id> last(last_n, last_n + n_last);
  return llama_sample_top_p_top_k(h->vocab, logits, last, repeat_penalty, top_k, top_p, temp, s->rng);
}

}  // extern "C"
