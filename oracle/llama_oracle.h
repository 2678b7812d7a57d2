/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.  See llama_oracle.c. */
#ifndef LLAMA_ORACLE_H
#define LLAMA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_model ora_model;

/* llama_model_load (PO.mm:98-498): reads fname[.i] part files, merges column/row splits. */
ora_model *ora_load(const char *path, int n_ctx, char *err, size_t errlen);
/* llama_eval (PO.mm:510-735): appends KV rows [n_past, n_past+n_tokens), returns last token's logits.
 * n_threads selects the reference's thread-partitioned summation order of the V*P product
 * (ggml.c:5553-5577, 5619-5665) -- the only place where the reference's result depends on it. */
int ora_eval(ora_model *m, int n_threads, int n_past, const int32_t *tokens, int n_tokens,
             float *logits, char *err, size_t errlen);
void ora_free(ora_model *m);
int ora_n_vocab(const ora_model *m);
int ora_n_ctx(const ora_model *m);
int ora_n_embd(const ora_model *m);
int ora_n_layer(const ora_model *m);
int ora_n_head(const ora_model *m);
int ora_ftype(const ora_model *m);
void ora_kv_export(const ora_model *m, int layer, int which, int n_rows, float *out);
void ora_kv_import(ora_model *m, int layer, int which, int n_rows, const float *in);

/* op-level restatements (each cites the reference in llama_oracle.c) */
void ora_quantize_row_q4_0(const float *x, void *y, int k);
void ora_quantize_row_q4_1(const float *x, void *y, int k);
void ora_dequantize_row_q4_0(const void *x, float *y, int k);
void ora_dequantize_row_q4_1(const void *x, float *y, int k);
float ora_vec_dot_q4_0(int n, const void *x, const void *y);
float ora_vec_dot_q4_1(int n, const void *x, const void *y);
float ora_vec_dot_f32(int n, const float *x, const float *y);
void ora_vec_mad_f32(int n, float *y, const float *x, float v);
void ora_mul_mat_q4(int type, const void *W, int M, int K, const float *x, int N, float *out);
void ora_norm(const float *x, float *y, int n);
void ora_rope(float *x, int n_head, int head_dim, int pos);
void ora_soft_max(float *p, int n);
void ora_silu(const float *x, float *y, int n);
uint16_t ora_fp32_to_fp16(float f);
float ora_fp16_to_fp32(uint16_t h);
const uint16_t *ora_table_silu_f16(void);
const uint16_t *ora_table_exp_f16(void);

#ifdef __cplusplus
}
#endif
#endif
