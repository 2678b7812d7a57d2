// Host side above the C ABI, in C++: the token loop of -[LlamaPredictOperation main] (PO.mm:768-901) and the event
// stream it posts (_LlamaEvent, headers/LlamaEvent.h:12-28), with the GPU library as the evaluator.  The reference's own
// host language (Swift / Objective-C++) has no toolchain in this image; this file is the compiled-code mirror of that
// layer: same order of events, same prompt batching (n_batch + 1 tokens per llama_eval, PO.mm:878-889), same probe call
// (PO.mm:822), same repetition window, same sampler draws -- so that a prompt run through b200_llama_run produces the
// token sequence the reference produces for the same model file, prompt, seed and parameters.
//
// The loop itself (b200_llama_run_loop) takes its evaluator as a function pointer: the product passes b200_llama_eval
// on a resident model; the CPU tests pass the oracle and compare the emitted ids with a run assembled from the
// reference's own tokenizer, sampler and llama_eval (tests/test_host_runner.py).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200_llama.h"

namespace {

struct EventSink {
  b200_event_fn fn;
  void *user;
  void post(int kind, const char *text = nullptr, int len = 0, int code = 0) const {
    if (fn) fn(user, kind, text, len, code);
  }
};

// gpt_random_prompt(rng), utils.cpp:102-119: the reference replaces an empty prompt by one of ten openers chosen with
// the FIRST draw of the run's generator (which therefore also shifts every later sampling draw).
const char *const kOpeners[10] = {"So", "Once upon a time", "When", "The", "After", "If", "import", "He", "She", "They"};

}  // namespace

extern "C" {

void b200_run_params_default(b200_run_params *p) {
  if (!p) return;
  // gpt_params, utils.h:15-37; the Swift Config overrides n_threads (8) and n_predict (512), LlamaRunner.swift:12-31
  p->seed = -1;
  p->n_threads = 8;
  p->n_predict = 512;
  p->repeat_last_n = 64;
  p->top_k = 40;
  p->top_p = 0.95f;
  p->temp = 0.80f;
  p->repeat_penalty = 1.30f;
  p->n_batch = 8;
  p->n_ctx = 512;          // the literal at PO.mm:790
  p->device = 0;
}

// gpu_model != nullptr: the evaluation that is followed by a sampling step runs as b200_llama_eval_topk, i.e. the penalty /
// top_k stage of the sampler happens on the GPU and only top_k (value, id) pairs come back (320 B + ids instead of 128 KB).
static thread_local int g_gpu_sampled = 0, g_host_sampled = 0;

static int run_loop_impl(b200_eval_fn eval, void *eval_ctx, b200_llama *gpu_model, int n_vocab, int n_ctx, const b200_tokenizer *tok,
                         const char *const *pieces, const int *piece_lens, const char *prompt, size_t prompt_len,
                         const char *antiprompt, size_t antiprompt_len, const b200_run_params *params,
                         b200_event_fn on_event, void *user) {
  g_gpu_sampled = g_host_sampled = 0;
  if (!eval || !tok || !params || !pieces || !piece_lens || n_vocab < 4 || n_ctx < 4) return B200_LLAMA_ERR_PREDICT;
  const EventSink sink{on_event, user};
  char err[512] = {0};

  b200_rng *rng = b200_rng_create(params->seed);                     // std::mt19937 rng(_params.seed), PO.mm:773
  struct RngGuard { b200_rng *r; ~RngGuard() { b200_rng_free(r); } } guard{rng};

  std::string text(prompt ? prompt : "", prompt ? prompt_len : 0);
  if (text.empty()) text = kOpeners[b200_rng_next_u32(rng) % 10];    // PO.mm:774-776

  sink.post(B200_EVENT_STARTED_GENERATING_OUTPUT);                   // PO.mm:799

  // tokenize the prompt (with BOS) and the reverse prompt (without; tokenized, never consulted: PO.mm:810, 815)
  std::vector<int32_t> embd_inp(text.size() + 2);
  embd_inp.resize((size_t) b200_llama_tokenize(tok, text.data(), text.size(), 1, embd_inp.data(), (int) embd_inp.size()));
  std::vector<int32_t> anti((antiprompt ? antiprompt_len : 0) + 2);
  anti.resize((size_t) std::max(0, b200_llama_tokenize(tok, antiprompt ? antiprompt : "", antiprompt ? antiprompt_len : 0, 0, anti.data(), (int) anti.size())));

  const int n_predict = std::min(params->n_predict, n_ctx - (int) embd_inp.size());     // PO.mm:812

  std::vector<float> logits((size_t) n_vocab);
  {
    // "determine the required inference memory per token" (PO.mm:819-825): a 4-token evaluation at position 0 whose
    // KV rows are overwritten afterwards.  It is kept because it is part of the call sequence the model sees.
    const int32_t probe[4] = {0, 1, 2, 3};
    const int rc = eval(eval_ctx, params->n_threads, 0, probe, 4, logits.data(), err, sizeof err);
    if (rc != B200_LLAMA_OK) { sink.post(B200_EVENT_FAILED, err, (int) strlen(err), B200_LLAMA_ERR_PREDICT); return rc; }
  }

  std::vector<int32_t> last_n((size_t) std::max(0, params->repeat_last_n), 0);           // PO.mm:827-829
  std::vector<int32_t> embd;
  int n_past = 0, remaining = n_predict;
  size_t consumed = 0;

  // The reference routes top_k / top_p / temp / repeat_penalty through `const float` locals (PO.mm:852-855) before they widen
  // to the sampler's int / double parameters.
  const float top_k = (float) params->top_k, top_p = params->top_p, temp = params->temp, penalty = params->repeat_penalty;
  std::vector<double> cand_v((size_t) std::max(1, (int) top_k));
  std::vector<int32_t> cand_i((size_t) std::max(1, (int) top_k));

  while (remaining > 0) {                                                                  // PO.mm:834
    int n_cand = 0;
    bool logits_on_host = true;
    if (!embd.empty()) {
      int rc;
      if (gpu_model && embd_inp.size() <= consumed) {
        // this evaluation is followed by a sampling step: candidate stage on the GPU.  last_n already holds embd (pushed when
        // the tokens were appended), which is the window the sampler will see.
        rc = b200_llama_eval_topk(gpu_model, params->n_threads, n_past, embd.data(), (int) embd.size(), last_n.data(), (int) last_n.size(),
                                  (double) penalty, (double) temp, (int) top_k, cand_v.data(), cand_i.data(), &n_cand, err, sizeof err);
        logits_on_host = false;
      } else {
        rc = eval(eval_ctx, params->n_threads, n_past, embd.data(), (int) embd.size(), logits.data(), err, sizeof err);
      }
      if (rc != B200_LLAMA_OK) { sink.post(B200_EVENT_FAILED, err, (int) strlen(err), B200_LLAMA_ERR_PREDICT); return rc; }
    }
    n_past += (int) embd.size();
    embd.clear();

    if (embd_inp.size() <= consumed) {
      // out of prompt: sample
      int32_t id;
      if (n_cand > 0) {
        id = b200_llama_sample_from_candidates(cand_v.data(), cand_i.data(), n_cand, (double) top_p, rng);
        g_gpu_sampled++;
      } else {
        if (!logits_on_host) {    // the values alone do not determine the reference's candidate order: its own library call decides
          const int rc = b200_llama_last_logits(gpu_model, logits.data(), err, sizeof err);
          if (rc != B200_LLAMA_OK) { sink.post(B200_EVENT_FAILED, err, (int) strlen(err), B200_LLAMA_ERR_PREDICT); return rc; }
        }
        id = b200_llama_sample_top_p_top_k(n_vocab, logits.data(), last_n.data(), (int) last_n.size(),
                                           (double) penalty, (int) top_k, (double) top_p, (double) temp, rng);
        g_host_sampled++;
      }
      if (!last_n.empty()) { last_n.erase(last_n.begin()); last_n.push_back(id); }         // PO.mm:867-868
      embd.push_back(id);
      --remaining;
    } else {
      // forward the next slice of the prompt: the loop breaks AFTER the size exceeds n_batch, so n_batch + 1 tokens go out
      while (embd_inp.size() > consumed) {                                                 // PO.mm:879-889
        embd.push_back(embd_inp[consumed]);
        if (!last_n.empty()) { last_n.erase(last_n.begin()); last_n.push_back(embd_inp[consumed]); }
        ++consumed;
        if ((int) embd.size() > params->n_batch) break;
      }
    }
    for (const int32_t id : embd) {                                                        // prompt tokens are echoed too, PO.mm:892-895
      const bool ok = id >= 0 && id < n_vocab;
      sink.post(B200_EVENT_OUTPUT_TOKEN, ok ? pieces[id] : "", ok ? piece_lens[id] : 0, id);
    }
  }
  sink.post(B200_EVENT_COMPLETED);                                                         // PO.mm:898
  return B200_LLAMA_OK;
}

int b200_llama_run_loop(b200_eval_fn eval, void *eval_ctx, int n_vocab, int n_ctx, const b200_tokenizer *tok,
                        const char *const *pieces, const int *piece_lens, const char *prompt, size_t prompt_len,
                        const char *antiprompt, size_t antiprompt_len, const b200_run_params *params,
                        b200_event_fn on_event, void *user) {
  return run_loop_impl(eval, eval_ctx, nullptr, n_vocab, n_ctx, tok, pieces, piece_lens, prompt, prompt_len, antiprompt, antiprompt_len,
                       params, on_event, user);
}

void b200_llama_run_sampler_stats(int *gpu_sampled, int *host_sampled) {
  if (gpu_sampled) *gpu_sampled = g_gpu_sampled;
  if (host_sampled) *host_sampled = g_host_sampled;
}

static int eval_on_model(void *ctx, int n_threads, int n_past, const int32_t *tokens, int n_tokens, float *logits, char *err,
                         size_t errlen) {
  return b200_llama_eval(static_cast<b200_llama *>(ctx), n_threads, n_past, tokens, n_tokens, logits, err, errlen);
}

int b200_llama_run(const char *model_path, const char *prompt, size_t prompt_len, const char *antiprompt, size_t antiprompt_len,
                   const b200_run_params *params, b200_event_fn on_event, void *user) {
  b200_run_params dflt;
  if (!params) { b200_run_params_default(&dflt); params = &dflt; }
  const EventSink sink{on_event, user};
  char err[512] = {0};

  sink.post(B200_EVENT_STARTED_LOADING_MODEL);                                             // PO.mm:785
  b200_llama *model = nullptr;
  // resident between runs (the reference re-reads the file every time, PO.mm:790)
  const int rc = b200_llama_acquire(model_path, params->n_ctx, params->device, &model, err, sizeof err);
  if (rc != B200_LLAMA_OK) { sink.post(B200_EVENT_FAILED, err, (int) strlen(err), B200_LLAMA_ERR_LOAD); return rc; }
  sink.post(B200_EVENT_FINISHED_LOADING_MODEL);                                            // PO.mm:797

  const int n_vocab = b200_llama_n_vocab(model);
  std::vector<const char *> pieces((size_t) n_vocab);
  std::vector<int> lens((size_t) n_vocab);
  for (int i = 0; i < n_vocab; i++) pieces[(size_t) i] = b200_llama_token_str(model, i, &lens[(size_t) i]);
  const b200_tokenizer *tok = b200_llama_shared_tokenizer(model);      // built once per resident model, not per run
  const char *host_only = getenv("B200_HOST_SAMPLER");                  // A/B switch: the whole sampler on the host, 128 KB of logits per token
  b200_llama *gpu_model = host_only && host_only[0] == '1' ? nullptr : model;
  const int out = run_loop_impl(eval_on_model, model, gpu_model, n_vocab, b200_llama_n_ctx(model), tok, pieces.data(), lens.data(), prompt,
                                prompt_len, antiprompt, antiprompt_len, params, on_event, user);
  b200_llama_release(model);                                                               // was ggml_free(model.ctx), PO.mm:900
  return out;
}

}  // extern "C"
