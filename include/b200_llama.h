/* b200_llama.h -- C ABI of the B200-native replacement for llama.swift's decode hot path.
 *
 * This is the drop-in boundary: the two C++ functions that -[LlamaPredictOperation main]
 * (Sources/llamaObjCxx/bridge/LlamaPredictOperation.mm, "PO.mm") calls -- llama_model_load (PO.mm:98) and
 * llama_eval (PO.mm:510-518) -- plus the bits of llama_model / gpt_vocab the token loop reads
 * (hparams.n_ctx PO.mm:812, hparams.n_vocab PO.mm:858, vocab.id_to_token PO.mm:893) and ggml_free (PO.mm:900).
 * Plain pointers and sizes only; no exceptions cross the boundary.  INTEGRATION.md shows the Obj-C++ binding.
 *
 * Error convention (headers/LlamaError.h:14-19): 0 on success, otherwise the reference's LlamaErrorCode
 * (-1000 failed to load model, -1001 prediction failed) with a message in `err`.
 */
#ifndef B200_LLAMA_H
#define B200_LLAMA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_LLAMA_OK 0
#define B200_LLAMA_ERR_LOAD (-1000)    /* LlamaErrorCodeFailedToLoadModel, headers/LlamaError.h:17 */
#define B200_LLAMA_ERR_PREDICT (-1001) /* LlamaErrorCodePredictionFailed,  headers/LlamaError.h:18 */

typedef struct b200_llama b200_llama;

/* == llama_model_load(fname, model, vocab, n_ctx, &err), PO.mm:98-498.
 * Reads the "ggml"-magic file (and its .1 .. .n-1 part files, PO.mm:312-322), merges column/row splits
 * (PO.mm:358-388), re-lays the Q4 blocks out for the GPU and allocates the f32 KV cache (PO.mm:290-304) in HBM.
 * device: CUDA device ordinal to place the model on. */
int b200_llama_load(const char *path, int n_ctx, int device, b200_llama **out, char *err, size_t errlen);

/* ---- tensor-parallel groups: one model over the 2 / 4 / 8 GPUs of an NVSwitch box ---------------------------------------
 * The reference has no multi-GPU path; what it does have is the Megatron-style split of its multi-part model files
 * (LLAMA_N_PARTS, PO.mm:33-38; merged back at PO.mm:358-388).  Here every weight matrix is split by ROWS over the
 * group (wq/wk/wv by attention head), so each GPU computes complete reference dot products -- bit-identical to the
 * single-GPU result -- and the finished activation slices are all-gathered by peer-to-peer stores over NVLink from
 * inside the token kernel.  Two ways to form a group:
 *
 *  (a) one process driving all GPUs -- what a LlamaRunner app needs: b200_llama_load_group returns ONE handle that
 *      b200_llama_eval / b200_llama_decode_device / b200_llama_free accept like a single-GPU handle;
 *  (b) one process per GPU (torchrun-style launchers): every rank calls b200_llama_load_shard, publishes its
 *      b200_llama_tp_ipc_handle (64 bytes) to the other ranks by any means, and calls b200_llama_tp_connect_ipc
 *      with all ranks' handles (handle i at handles + i * handle_stride).  After that every rank calls
 *      b200_llama_eval / _decode_device with identical arguments, in lock-step; each gets the full logits. */
int b200_llama_load_group(const char *path, int n_ctx, const int *devices, int n_devices, b200_llama **out,
                          char *err, size_t errlen);
int b200_llama_load_shard(const char *path, int n_ctx, int device, int tp_rank, int tp_size, b200_llama **out,
                          char *err, size_t errlen);
#define B200_LLAMA_IPC_HANDLE_BYTES 64
int b200_llama_tp_ipc_handle(const b200_llama *m, void *handle_out, size_t handle_bytes);
int b200_llama_tp_connect_ipc(b200_llama *m, const void *handles, size_t handle_stride, char *err, size_t errlen);

/* == llama_eval(model, n_threads, n_past, embd_inp, embd_w, mem_per_token, &err), PO.mm:510-735.
 * Evaluates n_tokens tokens at positions [n_past, n_past + n_tokens), appending their K/V rows, and writes the
 * LAST token's logits (n_vocab floats, PO.mm:724-725) to host memory `logits_out`.  May be called again with a
 * smaller n_past (the probe call at PO.mm:822 is later overwritten).  n_threads is the reference's thread count:
 * no host threads are used here, but the value selects the reference's thread-partitioned summation order of the
 * V*softmax(KQ) product (ggml.c:5553-5577, 5619-5665) so results match the reference run with that many threads. */
int b200_llama_eval(b200_llama *m, int n_threads, int n_past, const int32_t *tokens, int n_tokens,
                    float *logits_out, char *err, size_t errlen);

/* == ggml_free(model.ctx), PO.mm:900 */
void b200_llama_free(b200_llama *m);

/* Resident models.  The reference re-reads the model file on every run() (PO.mm:790);
 * b200_llama_acquire returns the model already resident in HBM for the same (file, n_ctx, device) when no other
 * operation holds it, else loads it (same errors as b200_llama_load); b200_llama_release gives it back without
 * freeing; b200_llama_cache_clear frees every idle resident model (e.g. when the LlamaRunner goes away). */
int b200_llama_acquire(const char *path, int n_ctx, int device, b200_llama **out, char *err, size_t errlen);
void b200_llama_release(b200_llama *m);
void b200_llama_cache_clear(void);

/* llama_hparams fields the caller reads (PO.mm:41-50) */
int b200_llama_n_vocab(const b200_llama *m);
int b200_llama_n_ctx(const b200_llama *m);
int b200_llama_n_embd(const b200_llama *m);
int b200_llama_n_layer(const b200_llama *m);
int b200_llama_n_head(const b200_llama *m);
int b200_llama_ftype(const b200_llama *m);

/* gpt_vocab::id_to_token (utils.h:49-55; used at PO.mm:893).  Returns a pointer owned by the model; *len = bytes. */
const char *b200_llama_token_str(const b200_llama *m, int id, int *len);

/* ---- host-side callees of the token loop -------------------------------------
 * Same results as the reference's utils.cpp functions, without their O(n_vocab) scans per token; no GPU involved.
 *
 * == llama_tokenize(vocab, text, bos), utils.cpp:275-311 (called at PO.mm:810, 815): greedy longest match over the
 * vocabulary, BOS id 1 first when bos != 0; stops at the first position where no piece matches.  Writes at most `cap`
 * ids and returns the number of ids the text produces (> cap: buffer too small), or < 0 on bad arguments. */
typedef struct b200_tokenizer b200_tokenizer;
b200_tokenizer *b200_tokenizer_create(const b200_llama *m);      /* from the model's vocabulary (gpt_vocab, PO.mm:140-164) */
b200_tokenizer *b200_tokenizer_create_from(const char *const *pieces, const int *lens, int n_vocab);
void b200_tokenizer_free(b200_tokenizer *t);
/* The model's own tokenizer, built on first use and kept with the handle (so a resident model, b200_llama_acquire, does not
 * rebuild the ~90k-node trie of a 32000-piece vocabulary on every run).  Owned by the handle: do not free. */
const b200_tokenizer *b200_llama_shared_tokenizer(b200_llama *m);
int b200_llama_tokenize(const b200_tokenizer *t, const char *text, size_t text_len, int bos, int32_t *out, int cap);

/* == llama_sample_top_p_top_k(vocab, logits, last_n_tokens, repeat_penalty, top_k, top_p, temp, rng),
 * utils.cpp:345-428 (called at PO.mm:865); b200_rng is the std::mt19937 of PO.mm:773.  Returns the sampled id. */
typedef struct b200_rng b200_rng;
b200_rng *b200_rng_create(int seed);
void b200_rng_free(b200_rng *r);
int32_t b200_llama_sample_top_p_top_k(int n_vocab, const float *logits, const int32_t *last_n_tokens, int n_last,
                                      double repeat_penalty, int top_k, double top_p, double temp, b200_rng *rng);

uint32_t b200_rng_next_u32(b200_rng *r);

/* The same sampler with its candidate stage on the GPU (utils.cpp:359-386, called at PO.mm:865):
 * b200_llama_eval_topk == b200_llama_eval followed by the repetition penalty and the top_k selection of
 * llama_sample_top_p_top_k ON THE DEVICE, so that top_k (value, id) pairs cross PCIe instead of n_vocab logits.
 * On return *n_cand == top_k and (cand_values, cand_ids) are the reference's logits_id after sample_top_k, in its
 * order -- or *n_cand == 0 when that order is not determined by the values alone (two of the best top_k + 1 compare
 * equal, NaN) or the request is outside what the kernel covers (top_k > 64, n_last > 256): the evaluation has still
 * run, and the caller fetches the logits with b200_llama_last_logits and calls b200_llama_sample_top_p_top_k.
 * b200_llama_sample_from_candidates == the rest of llama_sample_top_p_top_k (utils.cpp:388-428): soft-max over the
 * candidates, top_p cut, std::discrete_distribution draw.  cand_values / cand_ids: room for top_k entries. */
int b200_llama_eval_topk(b200_llama *m, int n_threads, int n_past, const int32_t *tokens, int n_tokens,
                         const int32_t *last_n_tokens, int n_last, double repeat_penalty, double temp, int top_k,
                         double *cand_values, int32_t *cand_ids, int *n_cand, char *err, size_t errlen);
int b200_llama_last_logits(b200_llama *m, float *logits_out, char *err, size_t errlen);
/* kernel-level entry of the candidate stage on host logits (tests, timing); kernel_ms (optional) = best-of-5 device time */
int b200_sample_topk(int device, const float *logits, int n_vocab, const int32_t *last_n_tokens, int n_last, double repeat_penalty,
                     double temp, int top_k, double *cand_values, int32_t *cand_ids, int *n_cand, float *kernel_ms,
                     char *err, size_t errlen);
int32_t b200_llama_sample_from_candidates(const double *cand_values, const int32_t *cand_ids, int n_cand, double top_p,
                                          b200_rng *rng);

/* ---- the token loop itself: -[LlamaPredictOperation main], PO.mm:768-901, as one call ---------------------------------
 * The compiled-code mirror of the reference's host layer (csrc/host_runner.cpp): model load (resident between runs),
 * prompt tokenization with BOS, the 4-token probe evaluation, prompt slices of n_batch + 1 tokens, sampling with the
 * repetition window, and the event stream of _LlamaEvent (headers/LlamaEvent.h:12-28) delivered through a callback:
 *   kind STARTED_LOADING_MODEL / FINISHED_LOADING_MODEL / STARTED_GENERATING_OUTPUT / COMPLETED: text = NULL
 *   kind OUTPUT_TOKEN: text/len = the piece (prompt tokens are echoed too, PO.mm:892-895), code = its id
 *   kind FAILED: text/len = the message, code = LlamaErrorCode (-1000 load, -1001 predict); the call returns that code. */
enum { B200_EVENT_STARTED_LOADING_MODEL = 0, B200_EVENT_FINISHED_LOADING_MODEL = 1, B200_EVENT_STARTED_GENERATING_OUTPUT = 2,
       B200_EVENT_OUTPUT_TOKEN = 3, B200_EVENT_COMPLETED = 4, B200_EVENT_FAILED = 5 };
typedef void (*b200_event_fn)(void *user, int kind, const char *text, int len, int code);
typedef struct b200_run_params {   /* gpt_params (utils.h:15-37) as _LlamaRunnerBridge fills it (LlamaRunnerBridge.mm:34-43) */
  int seed, n_threads, n_predict, repeat_last_n, top_k;
  float top_p, temp, repeat_penalty;
  int n_batch;
  int n_ctx;      /* the reference hard-codes 512 (PO.mm:790) */
  int device;
} b200_run_params;
void b200_run_params_default(b200_run_params *p);
int b200_llama_run(const char *model_path, const char *prompt, size_t prompt_len, const char *antiprompt, size_t antiprompt_len,
                   const b200_run_params *params, b200_event_fn on_event, void *user);
/* The loop with an injected evaluator (same contract as b200_llama_eval): what b200_llama_run drives after loading. */
typedef int (*b200_eval_fn)(void *ctx, int n_threads, int n_past, const int32_t *tokens, int n_tokens, float *logits_out,
                            char *err, size_t errlen);
int b200_llama_run_loop(b200_eval_fn eval, void *eval_ctx, int n_vocab, int n_ctx, const b200_tokenizer *tok,
                        const char *const *pieces, const int *piece_lens, const char *prompt, size_t prompt_len,
                        const char *antiprompt, size_t antiprompt_len, const b200_run_params *params,
                        b200_event_fn on_event, void *user);
/* How the sampling steps of the calling thread's last b200_llama_run were served: with the candidate stage on the GPU
 * (b200_llama_eval_topk) or, for an ambiguous candidate order / B200_HOST_SAMPLER=1, by the host sampler on the full logits. */
void b200_llama_run_sampler_stats(int *gpu_sampled, int *host_sampled);

/* Device-resident decode loop (no reference equivalent; used by bench.py's `value` leg and by teacher-forced
 * parity runs).  Starting with `first_token` at position n_past, runs n_steps single-token evaluations entirely on
 * the GPU: after each step the next input token is forced_tokens[i] if forced_tokens != NULL, else the arg-max of
 * the logits (greedy).  tokens_out[i] receives the arg-max of step i.  If logits_all != NULL it receives
 * n_steps * n_vocab floats.  *elapsed_ms (optional) = CUDA-event time of the loop on the model's stream. */
int b200_llama_decode_device(b200_llama *m, int n_threads, int n_past, int first_token, int n_steps,
                             const int32_t *forced_tokens, int32_t *tokens_out, float *logits_all,
                             float *elapsed_ms, char *err, size_t errlen);

/* Test-only: KV cache rows [0, n_rows) of one layer in the reference layout [n_ctx][n_embd] f32, K already roped
 * (PO.mm:300-301, 586-587, 604-611).  which: 0 = K, 1 = V.  On a tensor-parallel handle only the columns of the
 * rank's own heads are meaningful (a group handle exports rank 0's copy). */
int b200_llama_kv_export(const b200_llama *m, int layer, int which, int n_rows, float *out);
int b200_llama_kv_import(b200_llama *m, int layer, int which, int n_rows, const float *in);

/* Introspection for bench.py / profiles: number of kernel launches issued by the last eval/decode call (on a group:
 * per GPU), bytes of quantized weights resident on this handle's GPU, and run-time options: "graph" (CUDA-graph replay),
 * "mega" (1 = whole-token kernel, 0 = per-matrix kernels; tensor-parallel handles always use the whole-token kernel),
 * "pdl" (programmatic dependent launch between the per-matrix kernels), "time_kernel". */
long long b200_llama_last_launches(const b200_llama *m);
long long b200_llama_weight_bytes(const b200_llama *m);
/* With option "time_kernel" = 1, b200_llama_decode_device brackets every token-kernel launch with CUDA events on the
 * model's stream; this returns the sum of those per-launch durations (ms) for the last call. */
double b200_llama_last_kernel_ms(const b200_llama *m);
/* Device time (ms, CUDA events on the model's stream) of the last b200_llama_eval call that took the batch path
 * (n_tokens > 1): embedding .. last token's logits, without the host copies. */
double b200_llama_last_eval_ms(const b200_llama *m);
int b200_llama_set_option(b200_llama *m, const char *key, int value);

/* Development profiler of the whole-token kernel: evaluates one token at position pos and returns per-CTA
 * %globaltimer stamps (ns) at every phase boundary: out[cta * marks + i].  Returns marks (< 0 on error). */
int b200_llama_profile_token(b200_llama *m, int n_threads, int token, int pos, long long *out, int cap, int *n_cta);

/* Stand-alone kernels behind the same ABI, for kernel-level parity tests (level 1):
 * out[M] = W[M x K] (Q4_0, ggml row layout) * x[K] computed exactly as ggml_compute_forward_mul_mat_q4_0_f32 does. */
int b200_q4_0_matvec(int device, const void *w_ggml, int M, int K, const float *x, float *out,
                     int lane_pairs, float *kernel_ms, char *err, size_t errlen);
/* The same for N columns x[N][K] -> out[N][M] (the row-major mat-mul of a prompt batch, ggml.c:6199-6222): path 0 = CUDA-core
 * multi-column loop (every weight row read once per 8 columns), path 1 = tcgen05 / TMEM kernel.  Every column is bit-identical
 * to the single-column result.  kernel_ms (optional): best-of-5 device time of the mat-mul kernel alone. */
int b200_q4_0_matmul(int device, const void *w_ggml, int M, int K, const float *x, int N, float *out, int path,
                     float *kernel_ms, char *err, size_t errlen);

/* Same for Q4_1 (ggml per-row layout [nb f32 min][nb f32 d][nb*16 B]): ggml_compute_forward_mul_mat_q4_1_f32,
 * ggml.c:6287-6585 (quantize_row_q4_1 ggml.c:606-648 + ggml_vec_dot_q4_1 ggml.c:1584-1626). */
int b200_q4_1_matvec(int device, const void *w_ggml, int M, int K, const float *x, float *out, float *kernel_ms,
                     char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* B200_LLAMA_H */
