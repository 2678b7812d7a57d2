// Host engine + C ABI (include/b200_llama.h) for the B200-native llama.swift decode path.
//
//   b200_llama_load  == llama_model_load  (PO.mm:98-498)   file -> HBM (tile-major Q4_0 streams, f32 KV cache)
//   b200_llama_eval  == llama_eval        (PO.mm:510-735)  per-token kernel pipeline, replayed as a CUDA graph --
//                       the replacement for ggml_graph_compute's pthread INIT/COMPUTE/FINALIZE scheduler
//                       (ggml.c:9109-9555): 5 kernels per layer instead of 36 graph nodes, no host threads.
//
// There is deliberately no CPU fallback: every entry point fails with an error if CUDA is unavailable.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200_llama.h"
#include "kernels.cuh"
#include "kernels_q4_1.cuh"
#include "megakernel.cuh"
#include "batch.cuh"
#include "prefill_tc.cuh"
#include "sample_topk.cuh"

namespace b200 {
void host_build_tables(uint16_t *table_silu_f16, uint16_t *table_exp_f16);
void host_build_rope(double *cs, int n_ctx, int head_dim);
float host_kq_scale(int n_embd, int n_head);
}  // namespace b200

using namespace b200;

namespace {

constexpr int kSmemBudget = 227 * 1024;

void set_err(char *err, size_t errlen, const char *fmt, ...) {
  if (!err || !errlen) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, errlen, fmt, ap);
  va_end(ap);
}

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess) {                                                                            \
      set_err(err, errlen, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #expr); \
      return fail_code;                                                                                  \
    }                                                                                                    \
  } while (0)

struct GemvPlan {
  int qtype = 2;            // 2 = Q4_0 (20 B / 32 weights), 3 = Q4_1 (24 B / 32 weights)
  uint8_t *d_w = nullptr;
  size_t bytes = 0;
  uint8_t *d_wtc = nullptr;   // the same weights in the tensor-core prefill layout (prefill_tc.cuh), or null
  int M = 0, g_total = 0, nb = 0, n_cta = 0, cb = 0, lp = 0, rpt = 1, rmax = 0, S = 0, stage_bytes = 0, threads = 0, half_rows = 0;
  size_t smem = 0;
};

int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
}

int stage_bytes_cfg() { return env_int("B200_STAGE_BYTES", 57344) & ~127; }   // 56 KB measured best on B200 (3 ring stages of 56 KB)

// Partition + pipeline geometry for one fused matrix of M rows x K columns on n_sm SMs.
GemvPlan make_plan(int M, int K, int n_sm, int lp_override, int qtype = 2) {
  GemvPlan p;
  p.qtype = qtype;
  p.M = M;
  p.nb = K / 32;
  const int Mpad = (M + 3) & ~3;
  p.g_total = Mpad / 4;
  p.n_cta = std::min(n_sm, p.g_total);
  p.rmax = 4 * ((p.g_total + p.n_cta - 1) / p.n_cta);
  // (lane pairs per thread, rows per thread) of the whole-token kernel's row loop.  The loop is bound by shared-memory
  // wavefronts, so the activation operands are shared by several rows of a thread where there are rows to spare;
  // few-row matrices keep one row and one lane pair per thread (4 warps, one per SM sub-partition).
  // lp_override: 1, 2, 4 = lane pairs (rows per thread 1); 12, 14, 22 = (1,2), (1,4), (2,2).
  // Measured on B200 (7B): one row per thread and as many threads as fit wins -- the row phases then run at 5-6.5 TB/s,
  // i.e. at HBM speed; sharing activation operands between 2 or 4 rows of a thread (fewer, fatter threads) was slower.
  int lp = p.rmax <= 128 ? 1 : 2;
  int rpt = 1;
  if (lp_override == 1 || lp_override == 2 || lp_override == 4) { lp = lp_override; rpt = 1; }
  if (lp_override == 12) { lp = 1; rpt = 2; }
  if (lp_override == 14) { lp = 1; rpt = 4; }
  if (lp_override == 22) { lp = 2; rpt = 2; }
  while (p.rmax * (4 / lp) > MEGA_COMPUTE_THREADS && lp < 4) lp *= 2;     // (the per-matrix kernels run one row per thread)
  if (lp == 4) rpt = 1;
  p.rpt = rpt;
  if (qtype == 3) {
    // Q4_1 (kernels_q4_1.cuh): 16 compute warps = one CHAIN warp per group of 32 rows + TERM warps.  A chunk is cb blocks of
    // every row; its terms (16 floats per block and row) live in a double-buffered tile of r_pad x (cb * 16 + 4) floats, so
    // cb is sized for ~512 (row, block) items per chunk.  smem also holds the dequantized activation (K floats).
    p.lp = 4;
    p.threads = 16 * 32 + 32;
    const int r_pad = (p.rmax + 31) & ~31;
    const bool terms = q41_term_mode(p.rmax);      // few rows per CTA: term / chain split; many: one thread per row
    const size_t ys_etc = (size_t) K * 4 + (size_t) ((p.rmax + 3) & ~3) * 4 + 32 * 8 + 4 * 8;
    auto tile_bytes = [&](int cb) { return terms ? (size_t) 2 * r_pad * (cb * 16 + 4) * 4 : (size_t) 0; };
    p.cb = terms ? std::max(1, std::min(p.nb, 1024 / r_pad)) : std::max(1, std::min(p.nb, stage_bytes_cfg() / (p.rmax * 24)));
    // at least 3 weight stages next to the activation and the term tiles
    while (p.cb > 1 && ys_etc + tile_bytes(p.cb) + (size_t) 3 * (((size_t) p.cb * p.rmax * 24 + 127) & ~(size_t) 127) + 512 > kSmemBudget) p.cb--;
    p.stage_bytes = (p.cb * p.rmax * 24 + 127) & ~127;
    const int nch = (p.nb + p.cb - 1) / p.cb;
    const size_t fx = ys_etc + tile_bytes(p.cb);
    int S1 = (int) ((kSmemBudget - fx - 256) / (p.stage_bytes + 16));
    p.S = std::max(1, std::min(std::min(S1, nch), 12));
    p.smem = (size_t) p.S * p.stage_bytes + fx + (size_t) 2 * p.S * 8;
    p.bytes = (size_t) p.g_total * 4 * p.nb * 24;
    return p;
  }
  p.lp = lp;
  p.threads = ((p.rmax * (4 / lp) + 31) & ~31) + 32;
  // few-row matrices: the whole-token kernel walks the LP = 1 stream with EIGHT threads per row (megakernel.cuh:
  // gemv_rows_half); encoded as half_rows = 1 on top of lp = 1 (the per-matrix kernels and the batch kernels use lp)
  p.half_rows = (lp == 1 && rpt == 1 && p.rmax * 8 <= MEGA_COMPUTE_THREADS && env_int("B200_HALF_ROWS", 0) != 0) ? 1 : 0;   // measured on 1 GPU: no gain (1618 vs 1604 us/token), off by default
#if B200_IMMA
  // Tensor-path row loop (rowloop_imma.cuh): lp = rows per warp tile (8 for few-row matrices, so that 4+ warps share the
  // work; else 16), rpt = tiles per warp (2 when a CTA owns more than 16 tiles).  lp_override 8 / 16 forces the tile.
  {
    int rw = p.rmax <= 64 ? 8 : 16;
    if (lp_override == 8 || lp_override == 16) rw = lp_override;
    if ((p.rmax + rw - 1) / rw > MEGA_COMPUTE_WARPS) rw = 16;
    const int tiles = (p.rmax + rw - 1) / rw;
    p.lp = rw;
    p.rpt = tiles > MEGA_COMPUTE_WARPS ? 2 : 1;
    p.threads = MEGA_COMPUTE_THREADS + 32;
  }
#endif
  // One ring stage holds one chunk of whole quads (4 blocks x rmax rows x 80 B) in both the per-matrix kernels and the
  // whole-token kernel; an even number of quads per chunk where possible (the LP = 1 row loop works on quad pairs).
  const int nbq = (p.nb + 3) / 4;
  int cq = std::max(1, std::min(nbq, stage_bytes_cfg() / (p.rmax * 80)));
  if (cq >= 2) cq -= cq % 2;
  p.cb = cq * 4;
  p.stage_bytes = (cq * p.rmax * 80 + 127) & ~127;
  const int nchunks = (nbq + cq - 1) / cq;
#if B200_IMMA
  const size_t fixed = act_smem_bytes(nbq * 4) + (size_t) ((p.rmax + 3) & ~3) * 4 + 32 * 8;
#else
  const size_t fixed = (size_t) nbq * 4 * 64 + (size_t) nbq * 4 * 4 + (size_t) ((p.rmax + 3) & ~3) * 4 + 32 * 8;
#endif
  int S = (int) ((kSmemBudget - (long) fixed - 256) / (p.stage_bytes + 16));
  S = std::max(1, std::min(S, nchunks));
  p.S = S;
  p.smem = (size_t) S * p.stage_bytes + fixed + (size_t) 2 * S * 8;
  p.bytes = (size_t) p.g_total * 4 * nbq * 80;
  return p;
}

template <int LP, int PRO, int EPI>
cudaError_t launch_gemv_t(const GemvPlan &p, const GemvArgs &a, cudaStream_t st, bool pdl) {
  auto kern = q4_gemv_kernel<LP, PRO, EPI>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_cta);
  cfg.blockDim = dim3(p.threads);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

template <int PRO, int EPI>
cudaError_t launch_gemv_lp(const GemvPlan &p, const GemvArgs &a, cudaStream_t st, bool pdl) {
#if B200_IMMA
  // template slot: 1 = 8-row warp tiles, 2 = 16-row tiles, 4 = two 16-row tiles per warp
  if (p.lp == 8) return launch_gemv_t<1, PRO, EPI>(p, a, st, pdl);
  if (p.rpt == 1) return launch_gemv_t<2, PRO, EPI>(p, a, st, pdl);
  return launch_gemv_t<4, PRO, EPI>(p, a, st, pdl);
#else
  switch (p.lp) {
    case 1: return launch_gemv_t<1, PRO, EPI>(p, a, st, pdl);
    case 2: return launch_gemv_t<2, PRO, EPI>(p, a, st, pdl);
    default: return launch_gemv_t<4, PRO, EPI>(p, a, st, pdl);
  }
#endif
}

template <int PRO, int EPI>
cudaError_t launch_gemv_q41(const GemvPlan &p, const GemvArgs &a, cudaStream_t st, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_cta);
  cfg.blockDim = dim3(p.threads);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, q4_1_gemv_kernel<PRO, EPI>, a);
}

// Q4_0 or Q4_1 by the plan's type
template <int PRO, int EPI>
cudaError_t launch_gemv(const GemvPlan &p, const GemvArgs &a, cudaStream_t st, bool pdl) {
  return p.qtype == 3 ? launch_gemv_q41<PRO, EPI>(p, a, st, pdl) : launch_gemv_lp<PRO, EPI>(p, a, st, pdl);
}

template <int PRO, int EPI>
cudaError_t configure_gemv() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(q4_gemv_kernel<1, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(q4_gemv_kernel<2, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(q4_1_gemv_kernel<PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(q4_gemv_kernel<4, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
}

// opt every kernel instance on the current device in to 227 KB of dynamic shared memory
cudaError_t configure_kernels() {
  cudaError_t e;
  if ((e = configure_gemv<PRO_NORM, EPI_QKV>()) != cudaSuccess) return e;
  if ((e = configure_gemv<PRO_PLAIN, EPI_RESID>()) != cudaSuccess) return e;
  if ((e = configure_gemv<PRO_NORM, EPI_SILU_MUL>()) != cudaSuccess) return e;
  if ((e = configure_gemv<PRO_NORM, EPI_STORE>()) != cudaSuccess) return e;
  if ((e = configure_gemv<PRO_PLAIN, EPI_STORE>()) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(decode_token_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(q4_gemm_cols_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(q4_gemm_cols_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(q4_gemm_cols_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(batch_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(q4_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(batch_attn_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}

GemvArgs base_args(const GemvPlan &p) {
  GemvArgs a = {};
  a.w = p.d_w; a.M = p.M; a.g_total = p.g_total; a.nb = p.nb; a.cb = p.cb;
  a.n_stages = p.S; a.stage_bytes = p.stage_bytes; a.rmax = p.rmax;
  return a;
}

template <typename... KArgs, typename... Args>
cudaError_t launch_small(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// A tensor of the model file, referenced WHERE IT LIES in the memory-mapped part files (nothing is copied on the host until
// its rows are staged for upload, and a tensor-parallel rank only ever touches the pages of its own row slices).
struct HostTensor {
  int n_dims = 0, ftype = 0;
  int ne[2] = {1, 1};              // merged shape (ne[0] = row length, ne[1] = rows)
  int split = 1;                   // multi-part files: 0 = columns (ne[0]) split over the parts, 1 = rows (PO.mm:358-388)
  int n_parts = 1;
  const uint8_t *part[8] = {};     // start of this tensor's data in part p
  size_t row_bytes = 0;            // bytes of one merged row (1-D tensors: the whole tensor)
};

// read-only mapping of one model part file
struct FileMap {
  const uint8_t *p = nullptr;
  size_t size = 0;
  FileMap() = default;
  FileMap(const FileMap &) = delete;
  FileMap &operator=(const FileMap &) = delete;
  ~FileMap() { if (p) munmap(const_cast<uint8_t *>(p), size); }
  bool open_file(const char *path) {
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { ::close(fd); return false; }
    void *q = mmap(nullptr, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (q == MAP_FAILED) return false;
    p = (const uint8_t *) q; size = (size_t) st.st_size;
    madvise(q, size, MADV_WILLNEED);
    return true;
  }
};

// Two pinned staging buffers: the CPU assembles the next rows (merging the column slices of a multi-part file on the way)
// while the previous buffer is in flight to the GPU.
struct Stager {
  static constexpr size_t kCap = 32u << 20;
  uint8_t *buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int cur = 0;
  cudaError_t init() {
    for (int i = 0; i < 2; i++) {
      cudaError_t e = cudaMallocHost(&buf[i], kCap);
      if (e != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  ~Stager() {
    for (int i = 0; i < 2; i++) { if (buf[i]) cudaFreeHost(buf[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
  }
};

}  // namespace

struct b200_llama {
  // llama_hparams, PO.mm:41-50
  int n_vocab = 0, n_ctx = 0, n_embd = 0, n_mult = 0, n_head = 0, n_layer = 0, n_rot = 0, f16 = 0, n_ff = 0;
  std::vector<std::string> id_to_token;
  b200_tokenizer *shared_tok = nullptr;    // built on first use, lives with the (resident) model: b200_llama_shared_tokenizer

  int device = 0, n_sm = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  struct Layer {
    GemvPlan qkv, wo, w13, w2;
    float *attn_norm = nullptr, *ffn_norm = nullptr;
  };
  std::vector<Layer> layers;
  GemvPlan out;
  float *d_norm = nullptr;
  uint8_t *d_tok_emb = nullptr;     // raw ggml rows (get_rows needs one row per token)
  float *d_k = nullptr, *d_v = nullptr;   // [n_layer][n_ctx][n_embd] f32, PO.mm:290-304
  double2 *d_rope = nullptr;
  uint16_t *d_silu = nullptr, *d_exp = nullptr;
  float *d_inpL = nullptr, *d_inpFF = nullptr, *d_q = nullptr, *d_att = nullptr, *d_h = nullptr, *d_logits = nullptr;
  StepParams *d_sp = nullptr;
  int *d_token_log = nullptr, *d_forced = nullptr;
  float *d_logits_log = nullptr;
  size_t logits_log_cap = 0;
  int log_cap = 0;
  float *h_logits = nullptr;        // pinned
  TopkOut *d_topk = nullptr, *h_topk = nullptr;      // GPU candidate stage of the sampler (sample_topk.cuh); h_topk pinned
  int32_t *d_last_n = nullptr, *h_last_n = nullptr;  // repetition window, h_last_n pinned
  unsigned int *h_abort = nullptr;  // pinned copy of the device abort word (ptx.cuh: bounded waits)
  float kq_scale = 0.f;
  long long weight_bytes = 0;
  long long last_launches = 0;

  // whole-token persistent kernel
  long long *d_prof = nullptr;
  int prof_marks = 0;
  TokenArgs *h_token_args = nullptr;   // host copy of the kernel parameter block (layer descriptors prefilled)
  unsigned int *d_bar = nullptr;
  int mega_S = 0, mega_stage_bytes = 0, mega_xs_floats = 0, mega_ll_stage = 0;
  size_t mega_smem = 0;

  // tensor-parallel group: this handle is rank tp_rank of tp_size; every matrix is split by rows
  int tp_rank = 0, tp_size = 1;
  int e_loc = 0, f_loc = 0, v_loc = 0;      // this rank's share of n_embd / n_ff / n_vocab
  // the exchange area (ONE allocation, so one IPC handle): flagged activations | logits | end-of-token flags
  uint8_t *d_xchg = nullptr;
  size_t xchg_ll_bytes = 0, xchg_bytes = 0;
  uint8_t *peer_xchg[MEGA_MAX_TP] = {};     // every rank's exchange area as mapped into this process / device
  bool peer_ipc[MEGA_MAX_TP] = {};          // opened with cudaIpcOpenMemHandle (to be closed)
  bool tp_connected = false;
  unsigned int *d_epoch = nullptr;
  std::vector<b200_llama *> group;          // single-process group: the leader (rank 0) owns ranks 1..n-1

  int opt_graph = 1, opt_pdl = 0, opt_mega = 1, opt_time_kernel = 0, opt_fold = 1, opt_batch = 1, opt_tc = 1;
  int tc_min_n = 24;                       // batches of at least this many tokens take the tcgen05 mat-mul
  int opt_attn_tile = 1;                   // batches: query-tiled attention kernel (batch.cuh)
  int opt_spin_limit = 0;                  // != 0: overrides the wait budget of the token kernel (clock ticks)
  bool want_tc_copy = false;               // keep a second copy of the weights in the prefill (tcgen05) layout
  __half *b_xh = nullptr;                  // [cap_pad][max K] fp16 quantized activations (tcgen05 operand source)
  float *b_dxT = nullptr;                  // [max nb][cap_pad] their block scales, transposed
  // prompt batches (batch.cuh): per-chunk activation buffers, allocated on first use
  int batch_cap = 0;
  float *b_x = nullptr, *b_ff = nullptr, *b_q = nullptr, *b_att = nullptr, *b_h = nullptr, *b_o = nullptr;
  uint8_t *b_act = nullptr;
  int *b_tok = nullptr;
  int fold_argmax = 0;                     // this launch folds the greedy pick into the token kernel (decode_device)
  float2 *d_am = nullptr;                  // per-CTA arg-max candidates
  double last_kernel_ms = 0.0;             // sum of per-launch token-kernel durations (opt_time_kernel)
  double last_eval_ms = 0.0;               // device time of the last batched b200_llama_eval (CUDA events on the launching stream)
  std::vector<cudaEvent_t> kev;
  cudaGraphExec_t graph_exec = nullptr;   // one token step: embed .. logits
  int graph_threads = -1, graph_pdl = -1;
  const void *graph_log = nullptr, *graph_forced = nullptr;   // buffers baked into the captured kernel parameters
  int graph_spin = 0;
};

namespace {

int attn_smem_bytes(const b200_llama *m, int n_threads) {
  return (int) ((((size_t) m->n_ctx * 4 + 15) & ~(size_t) 15) + 8 * 8 + 8 * 4 + (size_t) n_threads * 32 * 4 + 64);
}

// One token through the network: the kernel sequence that replaces the 36-nodes-per-layer ggml graph.
MatDesc mat_desc(const GemvPlan &p) {
  MatDesc d = {};
  d.w = p.d_w; d.M = p.M; d.g_total = p.g_total; d.nb = p.nb; d.cb = p.cb; d.lp = p.half_rows ? 0 : p.lp; d.rpt = p.rpt;
  return d;
}

bool mega_usable(const b200_llama *m, int n_threads);
bool fold_usable(const b200_llama *m, int n_threads) {
  return m->opt_fold && m->tp_size == 1 && mega_usable(m, n_threads);
}

bool mega_usable(const b200_llama *m, int n_threads) {
  return (m->opt_mega || m->tp_size > 1) && m->f16 == 2 && m->mega_S >= 2 && n_threads <= MEGA_MAX_NTH;
}

// The whole token as ONE cooperative launch of the persistent kernel (megakernel.cuh).
cudaError_t enqueue_token_mega(b200_llama *m, int n_threads, long long *launches) {
  cudaError_t e = cudaMemsetAsync(m->d_bar, 0, 2 * sizeof(unsigned int), m->stream);
  if (e != cudaSuccess) return e;
  TokenArgs &a = *m->h_token_args;     // descriptors were filled in at load time
  a.n_layer = m->n_layer; a.out = mat_desc(m->out); a.final_norm = m->d_norm;
  a.tok_emb = m->d_tok_emb; a.q = m->d_q;
  a.rope = m->d_rope; a.silu_table = m->d_silu; a.exp_table = m->d_exp; a.sp = m->d_sp;
  a.fold_argmax = m->fold_argmax; a.am = m->d_am; a.token_log = m->d_token_log; a.forced_tokens = m->d_forced;
  a.tp.rank = m->tp_rank; a.tp.size = m->tp_size; a.tp.e_loc = m->e_loc; a.tp.f_loc = m->f_loc; a.tp.v_loc = m->v_loc;
  for (int p = 0; p < m->tp_size; p++) {
    uint8_t *base = m->peer_xchg[p];
    a.tp.ll[p] = reinterpret_cast<uint2 *>(base);
    a.tp.logits[p] = reinterpret_cast<float *>(base + m->xchg_ll_bytes);
    a.tp.done[p] = reinterpret_cast<unsigned int *>(base + m->xchg_ll_bytes + (size_t) m->n_vocab * 4);
    a.tp.hint[p] = a.tp.done[p] + MEGA_MAX_TP;
  }
  {
    // producer CTAs per layer over the whole group (every rank has the same shapes): CTAs that own rows of wo / w2
    // (n_embd / tp rows), of w1|w3 (2 n_ff / tp fused rows), and the attention CTAs (4 per head)
    const unsigned int ge = (unsigned int) ((m->e_loc + 3) / 4), gf = (unsigned int) ((2 * m->f_loc + 3) / 4);
    a.tp.p_e = (unsigned int) m->tp_size * std::min<unsigned int>((unsigned int) m->n_sm, ge);
    a.tp.p_f = (unsigned int) m->tp_size * std::min<unsigned int>((unsigned int) m->n_sm, gf);
    a.tp.p_att = 4u * (unsigned int) m->n_head;
  }
  a.epoch = m->d_epoch;
  // a spin-wait that outlives this many clock ticks traps instead of hanging the GPU; a multi-GPU group has to
  // tolerate the launch skew of its processes (graph instantiation, a slow host)
  a.spin_limit = m->opt_spin_limit ? (long long) m->opt_spin_limit : (m->tp_size > 1 ? 40000000000LL : 4000000000LL);
  a.bar = m->d_bar; a.n_embd = m->n_embd; a.n_head = m->n_head; a.n_ctx = m->n_ctx; a.n_ff = m->n_ff;
  a.n_threads = n_threads; a.kq_scale = m->kq_scale; a.S = m->mega_S; a.stage_bytes = m->mega_stage_bytes;
  a.xs_floats = m->mega_xs_floats;
  a.ll_stage = m->mega_ll_stage;
  a.prof = m->d_prof; a.prof_marks = m->prof_marks;
  a.pace_ps_per_byte = env_int("B200_PACE_PS_PER_BYTE", 0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(m->n_sm);
  cfg.blockDim = dim3(MEGA_THREADS);
  cfg.dynamicSmemBytes = m->mega_smem;
  cfg.stream = m->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident: they wait for each other's activation words
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, decode_token_kernel, a);
  if (launches) *launches += 1;
  return e;
}

cudaError_t enqueue_token(b200_llama *m, int n_threads, bool pdl, long long *launches) {
  if (mega_usable(m, n_threads)) return enqueue_token_mega(m, n_threads, launches);
  cudaStream_t st = m->stream;
  cudaError_t e;
  const int E = m->n_embd;
  long long n = 0;
  e = launch_small(m->f16 == 3 ? embed_q4_1_kernel : embed_kernel, dim3((E + 255) / 256), dim3(256), 0, st, pdl,
                   (const uint8_t *) m->d_tok_emb, (const StepParams *) m->d_sp, m->d_inpL, E);
  if (e != cudaSuccess) return e;
  n++;
  for (int il = 0; il < m->n_layer; il++) {
    b200_llama::Layer &L = m->layers[il];
    float *k_layer = m->d_k + (size_t) il * m->n_ctx * E;
    float *v_layer = m->d_v + (size_t) il * m->n_ctx * E;
    {  // norm * attention_norm -> wq|wk|wv -> rope -> Q buffer, KV cache      PO.mm:570-611
      GemvArgs a = base_args(L.qkv);
      a.x = m->d_inpL; a.norm_w = L.attn_norm; a.q_out = m->d_q; a.k_layer = k_layer; a.v_layer = v_layer;
      a.rope = m->d_rope; a.sp = m->d_sp; a.n_embd = E; a.head_dim = E / m->n_head;
      e = launch_gemv<PRO_NORM, EPI_QKV>(L.qkv, a, st, pdl);
      if (e != cudaSuccess) return e;
      n++;
    }
    {  // KQ, scale, mask, soft_max, V*P, merge heads                             PO.mm:614-646
      AttnArgs a = {};
      a.q = m->d_q; a.k_layer = k_layer; a.v_layer = v_layer; a.out = m->d_att; a.sp = m->d_sp;
      a.exp_table = m->d_exp; a.n_embd = E; a.n_threads = n_threads; a.kq_scale = m->kq_scale; a.n_ctx = m->n_ctx;
      e = launch_small(attn_kernel, dim3(m->n_head * ATTN_CLUSTER), dim3(ATTN_THREADS), attn_smem_bytes(m, n_threads), st, pdl, a);
      if (e != cudaSuccess) return e;
      n++;
    }
    {  // wo, + inpSA                                                              PO.mm:649-654
      GemvArgs a = base_args(L.wo);
      a.x = m->d_att; a.out = m->d_inpFF; a.resid = m->d_inpL;
      e = launch_gemv<PRO_PLAIN, EPI_RESID>(L.wo, a, st, pdl);
      if (e != cudaSuccess) return e;
      n++;
    }
    {  // norm * ffn_norm -> w1|w3 -> silu(w1 x) * (w3 x)                         PO.mm:660-680
      GemvArgs a = base_args(L.w13);
      a.x = m->d_inpFF; a.norm_w = L.ffn_norm; a.out = m->d_h; a.silu_table = m->d_silu;
      e = launch_gemv<PRO_NORM, EPI_SILU_MUL>(L.w13, a, st, pdl);
      if (e != cudaSuccess) return e;
      n++;
    }
    {  // w2, + inpFF                                                              PO.mm:682-687
      GemvArgs a = base_args(L.w2);
      a.x = m->d_h; a.out = m->d_inpL; a.resid = m->d_inpFF;
      e = launch_gemv<PRO_PLAIN, EPI_RESID>(L.w2, a, st, pdl);
      if (e != cudaSuccess) return e;
      n++;
    }
  }
  {  // final norm * norm.weight -> output                                       PO.mm:694-706
    GemvArgs a = base_args(m->out);
    a.x = m->d_inpL; a.norm_w = m->d_norm; a.out = m->d_logits;
    e = launch_gemv<PRO_NORM, EPI_STORE>(m->out, a, st, pdl);
    if (e != cudaSuccess) return e;
    n++;
  }
  if (launches) *launches += n;
  return cudaSuccess;
}

// Capture (once per {n_threads, pdl}) and replay the token step as a CUDA graph.
cudaError_t run_token(b200_llama *m, int n_threads, long long *launches) {
  const bool pdl = m->opt_pdl != 0;
  if (!m->opt_graph) return enqueue_token(m, n_threads, pdl, launches);
  const int key_pdl = (int) pdl | (mega_usable(m, n_threads) ? 2 : 0) | (m->fold_argmax ? 4 : 0);
  if (!m->graph_exec || m->graph_threads != n_threads || m->graph_pdl != key_pdl || m->graph_spin != m->opt_spin_limit ||
      (m->fold_argmax && (m->graph_log != m->d_token_log || m->graph_forced != m->d_forced))) {
    if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return e;
    long long dummy = 0;
    e = enqueue_token(m, n_threads, pdl, &dummy);
    cudaError_t e2 = cudaStreamEndCapture(m->stream, &g);
    if (e != cudaSuccess) { if (g) cudaGraphDestroy(g); return e; }
    if (e2 != cudaSuccess) return e2;
    e = cudaGraphInstantiate(&m->graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return e;
    m->graph_threads = n_threads;
    m->graph_pdl = key_pdl;
    m->graph_log = m->d_token_log; m->graph_forced = m->d_forced;
    m->graph_spin = m->opt_spin_limit;
  }
  if (launches) *launches += mega_usable(m, n_threads) ? 1 : 2 + 5LL * m->n_layer;
  return cudaGraphLaunch(m->graph_exec, m->stream);
}

// Upload the concatenated raw rows of the fused matrices and repack them into the tile-major stream.
struct RowSlice { const HostTensor *t; int row0, n_rows; };   // rows [row0, row0 + n_rows) of a (merged) tensor

RowSlice rows_of(const HostTensor &t, int row0, int n_rows) { return RowSlice{&t, row0, n_rows}; }

// rows [row0, row0 + n) of the merged tensor -> contiguous bytes at dst (host)
void gather_rows(const HostTensor &t, int row0, int n, uint8_t *dst) {
  const size_t rb = t.row_bytes;
  if (t.n_parts == 1) { memcpy(dst, t.part[0] + (size_t) row0 * rb, (size_t) n * rb); return; }
  if (t.split == 1) {                                            // part p holds rows [p*R/n_parts, (p+1)*R/n_parts), PO.mm:478-487
    const int per = t.ne[1] / t.n_parts;
    for (int r = row0; r < row0 + n;) {
      const int pp = r / per, k = std::min(row0 + n, (pp + 1) * per) - r;
      memcpy(dst + (size_t) (r - row0) * rb, t.part[pp] + (size_t) (r - pp * per) * rb, (size_t) k * rb);
      r += k;
    }
    return;
  }
  const size_t w = rb / t.n_parts;                               // part p holds columns [p*K/n_parts, ...) of every row, PO.mm:467-477
  for (int r = 0; r < n; r++)
    for (int pp = 0; pp < t.n_parts; pp++)
      memcpy(dst + (size_t) r * rb + pp * w, t.part[pp] + (size_t) (row0 + r) * w, w);
}

// host rows -> device (contiguous at d_dst), through the pinned double buffer
cudaError_t upload_rows(Stager &sg, cudaStream_t st, const RowSlice &rs, uint8_t *d_dst) {
  const size_t rb = rs.t->row_bytes;
  const int rows_per = (int) std::max<size_t>(1, Stager::kCap / rb);
  for (int r = 0; r < rs.n_rows; r += rows_per) {
    const int k = std::min(rows_per, rs.n_rows - r);
    cudaError_t e = cudaEventSynchronize(sg.ev[sg.cur]);         // the copy that last used this buffer has finished
    if (e != cudaSuccess) return e;
    gather_rows(*rs.t, rs.row0 + r, k, sg.buf[sg.cur]);
    if ((e = cudaMemcpyAsync(d_dst + (size_t) r * rb, sg.buf[sg.cur], (size_t) k * rb, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(sg.ev[sg.cur], st)) != cudaSuccess) return e;
    sg.cur ^= 1;
  }
  return cudaSuccess;
}

cudaError_t upload_matrix(b200_llama *m, Stager &sg, GemvPlan &p, const std::vector<RowSlice> &parts, int interleave_half,
                          uint8_t *d_stage) {
  cudaError_t e = cudaMalloc(&p.d_w, p.bytes);
  if (e != cudaSuccess) return e;
  size_t off = 0;
  for (const RowSlice &t : parts) {
    if ((e = upload_rows(sg, m->stream, t, d_stage + off)) != cudaSuccess) return e;
    off += (size_t) t.n_rows * t.t->row_bytes;
  }
  const int threads = 256;
  if (p.qtype == 3) {
    const long long total = (long long) p.g_total * 4 * p.nb;
    repack_q4_1_kernel<<<(unsigned) ((total + threads - 1) / threads), threads, 0, m->stream>>>(
        d_stage, p.d_w, p.M, p.g_total, p.nb, p.cb, p.n_cta, interleave_half);
  } else {
    const long long total = (long long) p.g_total * 4 * ((p.nb + 3) / 4) * 4;
    repack_q4_0_kernel<<<(unsigned) ((total + threads - 1) / threads), threads, 0, m->stream>>>(
        d_stage, p.d_w, p.M, p.g_total, p.nb, p.cb, p.n_cta, interleave_half, B200_IMMA ? 0 : p.lp);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (p.qtype == 2 && m->want_tc_copy) {
    // second copy for the tcgen05 prefill mat-mul: 128-row tiles, one 10 KB TMA box per (tile, quad of blocks)
    const size_t tcb = tc_weight_bytes(p.M, p.nb);
    if ((e = cudaMalloc(&p.d_wtc, tcb)) != cudaSuccess) return e;
    const long long total = (long long) ((p.M + TC_M - 1) / TC_M) * TC_M * ((p.nb + 3) / 4) * 4;
    repack_prefill_kernel<<<(unsigned) ((total + threads - 1) / threads), threads, 0, m->stream>>>(d_stage, p.d_wtc, p.M, p.nb, interleave_half);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  m->weight_bytes += (long long) p.M * p.nb * (p.qtype == 3 ? 24 : 20);
  return cudaSuccess;      // stream order protects d_stage: the next matrix's copies queue behind this repack
}

// small f32 tensors (norm weights): part 0 holds the whole tensor (PO.mm:453-457)
cudaError_t upload_f32_tensor(const HostTensor &t, float **dst) {
  cudaError_t e = cudaMalloc(dst, t.row_bytes);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dst, t.part[0], t.row_bytes, cudaMemcpyHostToDevice);
}

void free_model(b200_llama *m) {
  if (!m) return;
  for (size_t i = 1; i < m->group.size(); i++) free_model(m->group[i]);   // a group leader owns the other ranks
  m->group.clear();
  cudaSetDevice(m->device);
  for (int p = 0; p < MEGA_MAX_TP; p++)
    if (m->peer_ipc[p] && m->peer_xchg[p]) cudaIpcCloseMemHandle(m->peer_xchg[p]);
  cudaFree(m->d_xchg); cudaFree(m->d_epoch);
  if (m->graph_exec) cudaGraphExecDestroy(m->graph_exec);
  for (auto &L : m->layers) {
    cudaFree(L.qkv.d_w); cudaFree(L.wo.d_w); cudaFree(L.w13.d_w); cudaFree(L.w2.d_w);
    cudaFree(L.attn_norm); cudaFree(L.ffn_norm);
  }
  cudaFree(m->out.d_w); cudaFree(m->d_norm); cudaFree(m->d_tok_emb); cudaFree(m->d_k); cudaFree(m->d_v);
  cudaFree(m->d_rope); cudaFree(m->d_silu); cudaFree(m->d_exp);
  cudaFree(m->d_inpL); cudaFree(m->d_inpFF); cudaFree(m->d_q); cudaFree(m->d_att); cudaFree(m->d_h);   // d_logits lives inside d_xchg
  if (m->shared_tok) b200_tokenizer_free(m->shared_tok);
  delete m->h_token_args; cudaFree(m->d_bar); cudaFree(m->d_am);
  cudaFree(m->b_x); cudaFree(m->b_ff); cudaFree(m->b_q); cudaFree(m->b_att); cudaFree(m->b_h); cudaFree(m->b_o); cudaFree(m->b_act); cudaFree(m->b_tok);
  cudaFree(m->b_xh); cudaFree(m->b_dxT);
  for (auto &L : m->layers) { cudaFree(L.qkv.d_wtc); cudaFree(L.wo.d_wtc); cudaFree(L.w13.d_wtc); cudaFree(L.w2.d_wtc); }
  cudaFree(m->out.d_wtc);
  cudaFree(m->d_sp); cudaFree(m->d_token_log); cudaFree(m->d_forced); cudaFree(m->d_logits_log);
  if (m->h_logits) cudaFreeHost(m->h_logits);
  cudaFree(m->d_topk); cudaFree(m->d_last_n);
  if (m->h_topk) cudaFreeHost(m->h_topk);
  if (m->h_last_n) cudaFreeHost(m->h_last_n);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  for (cudaEvent_t e : m->kev) cudaEventDestroy(e);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

const std::map<int, int> kNParts = {{4096, 1}, {5120, 2}, {6656, 4}, {8192, 8}};   // LLAMA_N_PARTS, PO.mm:33-38

// llama_model_load for rank tp_rank of a tensor-parallel group of tp_size GPUs (1 = the whole model on one GPU).
int load_impl(const char *path, int n_ctx, int device, int tp_rank, int tp_size, b200_llama **out, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_LOAD;
  if (tp_size < 1 || tp_size > MEGA_MAX_TP || tp_rank < 0 || tp_rank >= tp_size) {
    set_err(err, errlen, "bad tensor-parallel rank %d of %d (at most %d GPUs)", tp_rank, tp_size, MEGA_MAX_TP);
    return fail_code;
  }
  if (!out) { set_err(err, errlen, "null out pointer"); return fail_code; }
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    set_err(err, errlen, "no CUDA device available (this library has no CPU path)");
    return fail_code;
  }
  if (device < 0 || device >= n_dev) { set_err(err, errlen, "bad device ordinal %d", device); return fail_code; }

  // The part files are memory-mapped and parsed in place (PO.mm:98-498 reads them with ifstream into ggml tensors): no host
  // copy of the weights exists, and what a tensor-parallel rank never uploads it never reads.
  std::vector<std::unique_ptr<FileMap>> maps;       // (a FileMap owns its mapping: never copied)
  maps.emplace_back(new FileMap());
  if (!maps[0]->open_file(path)) { set_err(err, errlen, "failed to open '%s'", path); return fail_code; }                // PO.mm:100-104
  size_t cur = 0;
  auto take = [&](const FileMap &fm, size_t &at, void *dst, size_t n) -> bool {
    if (at + n > fm.size) return false;
    if (dst) memcpy(dst, fm.p + at, n);
    at += n;
    return true;
  };
  uint32_t magic = 0;
  if (!take(*maps[0], cur, &magic, 4) || magic != 0x67676d6c) { set_err(err, errlen, "invalid model file '%s' (bad magic)", path); return fail_code; }   // PO.mm:110-114

  b200_llama *m = new b200_llama();
  struct Guard { b200_llama *p; ~Guard() { if (p) free_model(p); } } guard{m};
  m->device = device;
  m->tp_rank = tp_rank; m->tp_size = tp_size;
  int32_t hp[7] = {0};
  const bool hp_ok = take(*maps[0], cur, hp, sizeof(hp));                                                    // PO.mm:124-131
  m->n_vocab = hp[0]; m->n_embd = hp[1]; m->n_mult = hp[2]; m->n_head = hp[3]; m->n_layer = hp[4]; m->n_rot = hp[5]; m->f16 = hp[6];
  m->n_ctx = n_ctx;
  if (!hp_ok || m->n_vocab <= 0 || m->n_vocab > (1 << 24) || m->n_embd <= 0 || m->n_mult <= 0 || m->n_head <= 0 || m->n_layer <= 0 || n_ctx <= 0) {
    set_err(err, errlen, "invalid model file '%s' (bad header)", path);
    return fail_code;
  }
  m->n_ff = ((2 * (4 * m->n_embd) / 3 + m->n_mult - 1) / m->n_mult) * m->n_mult;                             // PO.mm:135
  auto np = kNParts.find(m->n_embd);
  if (np == kNParts.end()) { set_err(err, errlen, "unsupported n_embd %d (LLAMA_N_PARTS has no entry)", m->n_embd); return fail_code; }   // PO.mm:136 throws
  const int n_parts = np->second;
  if (m->f16 == 3 && n_parts != 1) {   // the reference merges column-split Q4_1 parts as if rows were AoS (PO.mm:467-477), which they are not
    set_err(err, errlen, "multi-part Q4_1 model files are not supported (the reference's own merge of them is ill-defined)");
    return fail_code;
  }
  if (m->n_embd % m->n_head != 0 || m->n_embd / m->n_head != 128) {
    set_err(err, errlen, "unsupported head size %d (kernels are built for 128)", m->n_embd / std::max(1, m->n_head));
    return fail_code;
  }
  if (tp_size > 1) {
    if (m->f16 != 2) { set_err(err, errlen, "tensor-parallel groups support Q4_0 model files only (type 2), got type %d", m->f16); return fail_code; }
    if (m->n_head % tp_size != 0 || m->n_ff % (2 * tp_size) != 0 || m->n_vocab % tp_size != 0) {
      set_err(err, errlen, "n_head %d / n_ff %d / n_vocab %d do not split over %d GPUs", m->n_head, m->n_ff, m->n_vocab, tp_size);
      return fail_code;
    }
  }
  m->e_loc = m->n_embd / tp_size; m->f_loc = m->n_ff / tp_size; m->v_loc = m->n_vocab / tp_size;
  m->id_to_token.resize(m->n_vocab);
  for (int i = 0; i < m->n_vocab; i++) {                                                                    // PO.mm:149-163
    uint32_t len = 0;
    if (!take(*maps[0], cur, &len, 4) || len > (1u << 20) || cur + len > maps[0]->size) { set_err(err, errlen, "invalid model file '%s' (bad vocab)", path); return fail_code; }
    m->id_to_token[i].assign((const char *) maps[0]->p + cur, len);
    cur += len;
  }
  switch (m->f16) {                                                                                         // PO.mm:169-180
    case 2: case 3: break;
    case 0: case 1:
      set_err(err, errlen, "model file '%s' has weight type %d; this build accelerates Q4_0 / Q4_1 (types 2, 3) only", path, m->f16);
      return fail_code;
    default:
      set_err(err, errlen, "invalid model file '%s' (bad f16 value %d)", path, m->f16);
      return fail_code;
  }
  const size_t file_offset = cur;

  // expected tensors, PO.mm:246-286
  const int E = m->n_embd, F = m->n_ff, V = m->n_vocab;
  const int QT = m->f16;
  const size_t BPB = QT == 3 ? 24 : 20;    // bytes per 32-weight block (ggml.c:2039-2040)
  std::map<std::string, HostTensor> tensors;
  auto expect = [&](const std::string &name, int ne0, int ne1, int n_dims) {
    HostTensor t; t.n_dims = n_dims; t.ne[0] = ne0; t.ne[1] = ne1; t.n_parts = n_dims == 1 ? 1 : n_parts;
    t.row_bytes = n_dims == 1 ? (size_t) ne0 * 4 : (size_t) (ne0 / 32) * BPB;
    tensors[name] = t;
  };
  expect("tok_embeddings.weight", E, V, 2);
  expect("norm.weight", E, 1, 1);
  expect("output.weight", E, V, 2);
  for (int i = 0; i < m->n_layer; i++) {
    const std::string p = "layers." + std::to_string(i) + ".";
    expect(p + "attention_norm.weight", E, 1, 1);
    expect(p + "attention.wq.weight", E, E, 2);
    expect(p + "attention.wk.weight", E, E, 2);
    expect(p + "attention.wv.weight", E, E, 2);
    expect(p + "attention.wo.weight", E, E, 2);
    expect(p + "ffn_norm.weight", E, 1, 1);
    expect(p + "feed_forward.w1.weight", E, F, 2);
    expect(p + "feed_forward.w2.weight", F, E, 2);
    expect(p + "feed_forward.w3.weight", E, F, 2);
  }
  std::map<std::string, int> seen;

  for (int part = 0; part < n_parts; part++) {                                                              // PO.mm:312-495
    std::string fname = path;
    if (part > 0) fname += "." + std::to_string(part);
    if (part > 0) {
      maps.emplace_back(new FileMap());
      if (!maps[part]->open_file(fname.c_str())) { set_err(err, errlen, "failed to open '%s'", fname.c_str()); return fail_code; }
    }
    const FileMap &fm = *maps[part];
    size_t at = file_offset;
    while (at < fm.size) {
      int32_t n_dims = 0, length = 0, ftype = 0;
      if (!take(fm, at, &n_dims, 4) || !take(fm, at, &length, 4) || !take(fm, at, &ftype, 4) ||
          n_dims < 1 || n_dims > 2 || length <= 0 || length > 255) {
        set_err(err, errlen, "corrupt tensor record in '%s'", fname.c_str());
        return fail_code;
      }
      int32_t ne[2] = {1, 1};
      int64_t nelements = 1;
      for (int i = 0; i < n_dims; i++) {
        if (!take(fm, at, &ne[i], 4) || ne[i] <= 0) { set_err(err, errlen, "corrupt tensor record in '%s'", fname.c_str()); return fail_code; }
        nelements *= ne[i];
      }
      if (at + (size_t) length > fm.size) { set_err(err, errlen, "corrupt tensor record in '%s'", fname.c_str()); return fail_code; }
      std::string name((const char *) fm.p + at, (size_t) length);
      at += (size_t) length;
      auto it = tensors.find(name);
      if (it == tensors.end()) { set_err(err, errlen, "unknown tensor '%s' in model file", name.c_str()); return fail_code; }   // PO.mm:352-356
      HostTensor &t = it->second;
      int split_type = 0;                                                                                   // PO.mm:358-388
      if (name.find("tok_embeddings") != std::string::npos) split_type = 0;
      else if (name.find("layers") != std::string::npos) {
        if (name.find("attention.wo.weight") != std::string::npos) split_type = 0;
        else if (name.find("feed_forward.w2.weight") != std::string::npos) split_type = 0;
        else split_type = 1;
      } else if (name.find("output") != std::string::npos) split_type = 1;

      size_t nbytes = 0;
      if (n_dims == 1) {
        if (t.n_dims != 1 || (int64_t) t.ne[0] != nelements) { set_err(err, errlen, "tensor '%s' has wrong size in model file", name.c_str()); return fail_code; }
        if (ftype != 0) { set_err(err, errlen, "tensor '%s': 1-D tensors must be f32 (ftype %d)", name.c_str(), ftype); return fail_code; }
        nbytes = t.row_bytes;
        if (part == 0) t.part[0] = fm.p + at;                                                               // PO.mm:453-457
      } else {
        if (t.n_dims != 2 || (int64_t) t.ne[0] * t.ne[1] / n_parts != nelements) { set_err(err, errlen, "tensor '%s' has wrong size in model file", name.c_str()); return fail_code; }
        const bool ok = split_type == 0 ? (t.ne[0] / n_parts == ne[0] && t.ne[1] == ne[1]) : (t.ne[0] == ne[0] && t.ne[1] / n_parts == ne[1]);
        if (!ok) {
          set_err(err, errlen, "tensor '%s' has wrong shape in model file: got [%d, %d], expected [%d, %d]", name.c_str(),
                  split_type == 0 ? t.ne[0] / n_parts : t.ne[0], split_type == 0 ? t.ne[1] : t.ne[1] / n_parts, ne[0], ne[1]);
          return fail_code;
        }
        if (ftype != QT) { set_err(err, errlen, "tensor '%s': ftype %d in a type-%d model file", name.c_str(), ftype, QT); return fail_code; }
        if (ne[0] % 64 != 0) { set_err(err, errlen, "tensor '%s': row length %d is not a multiple of 64", name.c_str(), ne[0]); return fail_code; }   // PO.mm:437
        t.split = split_type;
        t.part[part] = fm.p + at;          // merged on the way to the GPU (gather_rows): PO.mm:467-487
        nbytes = t.row_bytes * (size_t) t.ne[1] / (size_t) n_parts;
      }
      if (at + nbytes > fm.size) { set_err(err, errlen, "unexpected end of file in '%s' (tensor '%s')", fname.c_str(), name.c_str()); return fail_code; }
      at += nbytes;
      seen[name]++;
    }
  }
  for (auto &kv : tensors) {
    if (seen[kv.first] != n_parts) { set_err(err, errlen, "tensor '%s' missing from model file", kv.first.c_str()); return fail_code; }
  }

  // ---- device side ----
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { set_err(err, errlen, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return fail_code; }
  m->n_sm = env_int("B200_NUM_CTAS", prop.multiProcessorCount);
  CUDA_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&m->ev0));
  CUDA_TRY(cudaEventCreate(&m->ev1));

  const size_t stage_cap = std::max({(size_t) 3 * E * (E / 32) * BPB, (size_t) 2 * F * (E / 32) * BPB, (size_t) V * (E / 32) * BPB});
  uint8_t *d_stage = nullptr;
  CUDA_TRY(cudaMalloc(&d_stage, stage_cap));
  struct StageGuard { uint8_t *p; ~StageGuard() { cudaFree(p); } } sguard{d_stage};

  Stager stager;
  CUDA_TRY(stager.init());
  auto upload_f32 = [&](const HostTensor &t, float **dst) -> cudaError_t { return upload_f32_tensor(t, dst); };
  // the tensor-core prefill path keeps a second copy of the Q4_0 weights in its own tile layout (+1x weight bytes of HBM)
  m->want_tc_copy = QT == 2 && tp_size == 1 && env_int("B200_PREFILL_COPY", 1) != 0 && !B200_IMMA;
  const int lp_small = env_int("B200_LP_SMALL", 0), lp_qkv = env_int("B200_LP_QKV", 0), lp_w13 = env_int("B200_LP_W13", 0), lp_out = env_int("B200_LP_OUT", 0);
  m->layers.resize(m->n_layer);
  // Row split over the tensor-parallel group: rank r keeps rows [r*n/tp, (r+1)*n/tp) of every matrix (for wq/wk/wv
  // these are its heads).  A row is a complete reference dot product, so the split changes no arithmetic.
  const int El = m->e_loc, Fl = m->f_loc, Vl = m->v_loc;
  const int e0 = tp_rank * El, f0 = tp_rank * Fl, v0 = tp_rank * Vl;
  for (int i = 0; i < m->n_layer; i++) {
    const std::string p = "layers." + std::to_string(i) + ".";
    b200_llama::Layer &L = m->layers[i];
    L.qkv = make_plan(3 * El, E, m->n_sm, lp_qkv, QT);
    CUDA_TRY(upload_matrix(m, stager, L.qkv, {rows_of(tensors[p + "attention.wq.weight"], e0, El), rows_of(tensors[p + "attention.wk.weight"], e0, El),
                                      rows_of(tensors[p + "attention.wv.weight"], e0, El)}, 0, d_stage));
    L.wo = make_plan(El, E, m->n_sm, lp_small, QT);
    CUDA_TRY(upload_matrix(m, stager, L.wo, {rows_of(tensors[p + "attention.wo.weight"], e0, El)}, 0, d_stage));
    L.w13 = make_plan(2 * Fl, E, m->n_sm, lp_w13, QT);
    CUDA_TRY(upload_matrix(m, stager, L.w13, {rows_of(tensors[p + "feed_forward.w1.weight"], f0, Fl), rows_of(tensors[p + "feed_forward.w3.weight"], f0, Fl)}, Fl, d_stage));
    L.w2 = make_plan(El, F, m->n_sm, lp_small, QT);
    CUDA_TRY(upload_matrix(m, stager, L.w2, {rows_of(tensors[p + "feed_forward.w2.weight"], e0, El)}, 0, d_stage));
    CUDA_TRY(upload_f32(tensors[p + "attention_norm.weight"], &L.attn_norm));
    CUDA_TRY(upload_f32(tensors[p + "ffn_norm.weight"], &L.ffn_norm));
  }
  m->out = make_plan(Vl, E, m->n_sm, lp_out, QT);
  CUDA_TRY(upload_matrix(m, stager, m->out, {rows_of(tensors["output.weight"], v0, Vl)}, 0, d_stage));
  CUDA_TRY(upload_f32(tensors["norm.weight"], &m->d_norm));
  {
    const HostTensor &t = tensors["tok_embeddings.weight"];
    CUDA_TRY(cudaMalloc(&m->d_tok_emb, t.row_bytes * (size_t) t.ne[1]));
    CUDA_TRY(upload_rows(stager, m->stream, rows_of(t, 0, t.ne[1]), m->d_tok_emb));
  }
  CUDA_TRY(cudaStreamSynchronize(m->stream));       // every staged copy and repack has finished: the mappings can go
  tensors.clear();
  maps.clear();

  const size_t kv_bytes = (size_t) m->n_layer * n_ctx * E * sizeof(float);                                  // PO.mm:297-301
  CUDA_TRY(cudaMalloc(&m->d_k, kv_bytes));
  CUDA_TRY(cudaMalloc(&m->d_v, kv_bytes));
  CUDA_TRY(cudaMemset(m->d_k, 0, kv_bytes));
  CUDA_TRY(cudaMemset(m->d_v, 0, kv_bytes));

  {
    std::vector<uint16_t> ts(1 << 16), te(1 << 16);
    host_build_tables(ts.data(), te.data());
    CUDA_TRY(cudaMalloc(&m->d_silu, ts.size() * 2));
    CUDA_TRY(cudaMalloc(&m->d_exp, te.size() * 2));
    CUDA_TRY(cudaMemcpy(m->d_silu, ts.data(), ts.size() * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->d_exp, te.data(), te.size() * 2, cudaMemcpyHostToDevice));
    const int hd = E / m->n_head;
    std::vector<double> cs((size_t) n_ctx * hd);
    host_build_rope(cs.data(), n_ctx, hd);
    CUDA_TRY(cudaMalloc(&m->d_rope, cs.size() * 8));
    CUDA_TRY(cudaMemcpy(m->d_rope, cs.data(), cs.size() * 8, cudaMemcpyHostToDevice));
    m->kq_scale = host_kq_scale(E, m->n_head);
  }
  CUDA_TRY(cudaMalloc(&m->d_inpL, E * 4));
  CUDA_TRY(cudaMalloc(&m->d_inpFF, E * 4));
  CUDA_TRY(cudaMalloc(&m->d_q, E * 4));
  CUDA_TRY(cudaMalloc(&m->d_att, E * 4));
  CUDA_TRY(cudaMalloc(&m->d_h, (size_t) F * 4));
  // exchange area: flagged activations inpL[2][E] | inpFF[2][E] | att[2][E] | h[2][F] | qkv[2][3E] (8 B per value), then the logits,
  // then the end-of-token flags and the arrival-hint counters.  One allocation = one IPC handle; zero-filled, so no flag matches a live sequence number.
  m->xchg_ll_bytes = ((size_t) 12 * E + (size_t) 2 * F) * 8;
  m->xchg_bytes = m->xchg_ll_bytes + (size_t) V * 4 + MEGA_MAX_TP * 4 + 4 * 4 + 256;
  CUDA_TRY(cudaMalloc(&m->d_xchg, m->xchg_bytes));
  CUDA_TRY(cudaMemset(m->d_xchg, 0, m->xchg_bytes));
  m->d_logits = reinterpret_cast<float *>(m->d_xchg + m->xchg_ll_bytes);
  m->peer_xchg[tp_rank] = m->d_xchg;
  m->tp_connected = tp_size == 1;
  CUDA_TRY(cudaMalloc(&m->d_epoch, sizeof(unsigned int)));
  {
    const unsigned int one = 1;
    CUDA_TRY(cudaMemcpy(m->d_epoch, &one, sizeof(one), cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMalloc(&m->d_sp, sizeof(StepParams)));
  CUDA_TRY(cudaMemset(m->d_sp, 0, sizeof(StepParams)));
  CUDA_TRY(cudaMallocHost(&m->h_logits, (size_t) V * 4 + 64));    // + the abort word read back with every call
  m->h_abort = reinterpret_cast<unsigned int *>(m->h_logits + V);
  *m->h_abort = 0;
  CUDA_TRY(configure_kernels());
  {
    std::vector<LayerDesc> descs(m->n_layer);
    for (int i = 0; i < m->n_layer; i++) {
      descs[i].qkv = mat_desc(m->layers[i].qkv); descs[i].wo = mat_desc(m->layers[i].wo);
      descs[i].w13 = mat_desc(m->layers[i].w13); descs[i].w2 = mat_desc(m->layers[i].w2);
      descs[i].attn_norm = m->layers[i].attn_norm; descs[i].ffn_norm = m->layers[i].ffn_norm;
      descs[i].k_layer = m->d_k + (size_t) i * n_ctx * E; descs[i].v_layer = m->d_v + (size_t) i * n_ctx * E;
    }
    m->h_token_args = new TokenArgs();
    memset(m->h_token_args, 0, sizeof(TokenArgs));
    if (m->n_layer <= MEGA_MAX_LAYERS) memcpy(m->h_token_args->layers, descs.data(), descs.size() * sizeof(LayerDesc));
    CUDA_TRY(cudaMalloc(&m->d_am, (size_t) std::max(1, m->n_sm) * sizeof(float2)));
    CUDA_TRY(cudaMalloc(&m->d_bar, 2 * sizeof(unsigned int)));
    CUDA_TRY(cudaMemset(m->d_bar, 0, 2 * sizeof(unsigned int)));
    // shared-memory budget of the whole-token kernel: fixed areas first, the rest is the weight ring
    const int nb_max = ((std::max(E, F) / 32) + 3) & ~3;     // whole quads
    m->mega_xs_floats = (n_ctx + 3) & ~3;
    m->mega_stage_bytes = stage_bytes_cfg();
    const size_t act_bytes = B200_IMMA ? act_smem_bytes(nb_max) : (size_t) (nb_max + 2) * 32 + (size_t) ((nb_max + 3) & ~3) * 4;
    const size_t fixed = act_bytes + (size_t) m->mega_xs_floats * 4 +
                         MEGA_MAX_ROWS * 4 + 32 * 8 + 32 * 4 + MEGA_MAX_NTH * 32 * 4 + 64 * 16 + 288 * 4 + MEGA_COMPUTE_WARPS * 4;
    const long ring = (long) kSmemBudget - (long) fixed - 256;
    m->mega_S = ring > 0 ? (int) (ring / (m->mega_stage_bytes + 16)) : 0;
    m->mega_smem = (size_t) m->mega_S * m->mega_stage_bytes + fixed + (size_t) 2 * m->mega_S * 8;
    // what is left after the ring (it holds whole stages only) can stage one flagged n_embd-sized vector for the prologues
    m->mega_ll_stage = 0;
    if (m->mega_S > 0 && m->mega_smem + 16 + (size_t) E * 8 + 64 <= (size_t) kSmemBudget) {
      m->mega_ll_stage = 1;
      m->mega_smem += 16 + (size_t) E * 8 + 64;
    }
    int rmax_all = m->out.rmax;
    for (auto &L : m->layers) rmax_all = std::max({rmax_all, L.qkv.rmax, L.wo.rmax, L.w13.rmax, L.w2.rmax});
    bool fits = rmax_all * 80 <= m->mega_stage_bytes && E / 8 <= MEGA_NORM_ROUNDS * MEGA_COMPUTE_THREADS && m->n_layer <= MEGA_MAX_LAYERS && F / 8 <= 6 * MEGA_COMPUTE_THREADS;
#if B200_IMMA
    auto rows_fit = [&](const GemvPlan &p) { return (p.rmax + p.lp - 1) / p.lp <= MEGA_COMPUTE_WARPS * p.rpt && p.rmax <= MEGA_MAX_ROWS; };
#else
    auto rows_fit = [&](const GemvPlan &p) { return p.rmax / p.rpt * (4 / p.lp) <= MEGA_COMPUTE_THREADS && p.rmax <= MEGA_MAX_ROWS; };
#endif
    fits = fits && rows_fit(m->out);
    for (auto &L : m->layers) fits = fits && rows_fit(L.qkv) && rows_fit(L.wo) && rows_fit(L.w13) && rows_fit(L.w2);
    if (!fits) m->mega_S = 0;   // falls back to the per-matrix kernels
  }
  m->opt_mega = env_int("B200_MEGA", 1);
  m->opt_graph = env_int("B200_GRAPH", 1);
  m->opt_pdl = env_int("B200_PDL", 0);

  if (tp_size > 1 && !(m->f16 == 2 && m->mega_S >= 2)) {
    set_err(err, errlen, "this model's shapes do not fit the whole-token kernel, which a tensor-parallel group requires");
    return fail_code;
  }
  *out = m;
  guard.p = nullptr;
  return B200_LLAMA_OK;
}

// checks shared by the entry points that launch the token kernel
const char *tp_ready(const b200_llama *m, int n_threads) {
  if (m->tp_size > 1 && !m->tp_connected) return "tensor-parallel group is not connected (b200_llama_tp_connect_ipc / b200_llama_load_group)";
  if (m->tp_size > 1 && n_threads > MEGA_MAX_NTH) return "a tensor-parallel group supports n_threads <= 16";
  return nullptr;
}

}  // namespace

extern "C" {

int b200_llama_load(const char *path, int n_ctx, int device, b200_llama **out, char *err, size_t errlen) {
  return load_impl(path, n_ctx, device, 0, 1, out, err, errlen);
}

int b200_llama_load_shard(const char *path, int n_ctx, int device, int tp_rank, int tp_size, b200_llama **out,
                          char *err, size_t errlen) {
  return load_impl(path, n_ctx, device, tp_rank, tp_size, out, err, errlen);
}

int b200_llama_tp_ipc_handle(const b200_llama *m, void *handle_out, size_t handle_bytes) {
  if (!m || !handle_out || handle_bytes < sizeof(cudaIpcMemHandle_t)) return B200_LLAMA_ERR_LOAD;
  if (cudaSetDevice(m->device) != cudaSuccess) return B200_LLAMA_ERR_LOAD;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, m->d_xchg) != cudaSuccess) return B200_LLAMA_ERR_LOAD;
  memcpy(handle_out, &h, sizeof(h));
  return B200_LLAMA_OK;
}

int b200_llama_tp_connect_ipc(b200_llama *m, const void *handles, size_t handle_stride, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_LOAD;
  if (!m || !handles || handle_stride < sizeof(cudaIpcMemHandle_t)) { set_err(err, errlen, "null argument"); return fail_code; }
  CUDA_TRY(cudaSetDevice(m->device));
  for (int p = 0; p < m->tp_size; p++) {
    if (p == m->tp_rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t *) handles + (size_t) p * handle_stride, sizeof(h));
    void *ptr = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    m->peer_xchg[p] = (uint8_t *) ptr;
    m->peer_ipc[p] = true;
  }
  m->tp_connected = true;
  return B200_LLAMA_OK;
}

int b200_llama_load_group(const char *path, int n_ctx, const int *devices, int n_devices, b200_llama **out,
                          char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_LOAD;
  if (!out || !devices || n_devices < 1 || n_devices > MEGA_MAX_TP) { set_err(err, errlen, "bad device list"); return fail_code; }
  *out = nullptr;
  std::vector<b200_llama *> ms(n_devices, nullptr);
  auto drop = [&]() { for (b200_llama *x : ms) free_model(x); };
  for (int r = 0; r < n_devices; r++) {
    const int rc = load_impl(path, n_ctx, devices[r], r, n_devices, &ms[r], err, errlen);
    if (rc != B200_LLAMA_OK) { drop(); return rc; }
  }
  for (int r = 0; r < n_devices; r++) {
    cudaError_t e = cudaSetDevice(devices[r]);
    for (int p = 0; p < n_devices && e == cudaSuccess; p++) {
      if (p == r) continue;
      int can = 0;
      e = cudaDeviceCanAccessPeer(&can, devices[r], devices[p]);
      if (e == cudaSuccess && !can) { set_err(err, errlen, "device %d cannot access device %d (no NVLink / P2P)", devices[r], devices[p]); drop(); return fail_code; }
      if (e == cudaSuccess) {
        e = cudaDeviceEnablePeerAccess(devices[p], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
      }
      ms[r]->peer_xchg[p] = ms[p]->d_xchg;
    }
    if (e != cudaSuccess) { set_err(err, errlen, "CUDA error %s while enabling peer access", cudaGetErrorString(e)); drop(); return fail_code; }
    ms[r]->tp_connected = true;
  }
  ms[0]->group = ms;
  *out = ms[0];
  return B200_LLAMA_OK;
}

// The ranks of a single-process group (b200_llama_load_group) are driven by one host thread: everything that may
// block or synchronise a device (allocation, blocking copies) happens BEFORE the first launch on any rank, then the
// launches of all ranks are enqueued (asynchronous), then all streams are synchronised.  The token kernels of the
// ranks wait for each other on the GPUs, never on the host.
static std::vector<b200_llama *> ranks_of(b200_llama *m) {
  return m->group.empty() ? std::vector<b200_llama *>{m} : m->group;
}



// ---- bounded waits, host side (ptx.cuh): a device wait that gave up set the per-device abort word and the kernels ran to
// completion on garbage.  Every entry point reads the word back with its results; when it is set the call fails with -1001,
// and the exchange state (flag words, arrival counters, launch counter) is reset so that the NEXT call starts clean -- the
// CUDA context stays healthy (the reference's failure path is an NSError from llama_eval, PO.mm:841-846).
cudaError_t abort_fetch_async(b200_llama *m) {
  return cudaMemcpyFromSymbolAsync(m->h_abort, g_b200_abort, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, m->stream);
}

bool abort_check_and_reset(const std::vector<b200_llama *> &ranks) {
  bool aborted = false;
  for (b200_llama *r : ranks) aborted = aborted || (r->h_abort && *r->h_abort != 0);
  if (!aborted) return false;
  for (b200_llama *r : ranks) {
    cudaSetDevice(r->device);
    const unsigned int zero = 0, one = 1;
    cudaMemcpyToSymbol(g_b200_abort, &zero, sizeof(zero));
    cudaMemset(r->d_xchg, 0, r->xchg_bytes);
    cudaMemcpy(r->d_epoch, &one, sizeof(one), cudaMemcpyHostToDevice);
    cudaMemset(r->d_bar, 0, 2 * sizeof(unsigned int));
    *r->h_abort = 0;
  }
  return true;
}

// ---- prompt batches (batch.cuh) ------------------------------------------------------------------------------------------
constexpr int kBatchChunk = 256;      // tokens evaluated together (bounds the activation buffers: ~210 KB per token at 7B)

bool batch_usable(const b200_llama *m, int n_tokens) {
  return m->opt_batch && n_tokens >= 2 && m->f16 == 2 && m->tp_size == 1 && !B200_IMMA;
}

cudaError_t batch_reserve(b200_llama *m, int n) {
  if (m->batch_cap >= n) return cudaSuccess;
  cudaFree(m->b_x); cudaFree(m->b_ff); cudaFree(m->b_q); cudaFree(m->b_att); cudaFree(m->b_h); cudaFree(m->b_o); cudaFree(m->b_act); cudaFree(m->b_tok);
  cudaFree(m->b_xh); cudaFree(m->b_dxT);
  m->b_x = m->b_ff = m->b_q = m->b_att = m->b_h = m->b_o = nullptr; m->b_act = nullptr; m->b_tok = nullptr; m->batch_cap = 0;
  m->b_xh = nullptr; m->b_dxT = nullptr;
  const size_t E = m->n_embd, F = m->n_ff;
  cudaError_t e;
  if ((e = cudaMalloc(&m->b_x, (size_t) n * E * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&m->b_ff, (size_t) n * E * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&m->b_q, (size_t) n * E * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&m->b_att, (size_t) n * E * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&m->b_h, (size_t) n * F * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&m->b_o, (size_t) n * std::max(3 * E, 2 * F) * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&m->b_act, (size_t) n * batch_act_bytes((int) (std::max(E, F) / 32)))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&m->b_tok, (size_t) n * 4)) != cudaSuccess) return e;
  if (m->want_tc_copy) {
    const size_t npad = ((size_t) n + TC_T - 1) / TC_T * TC_T, nbq = (std::max(E, F) / 32 + 3) / 4;
    if ((e = cudaMalloc(&m->b_xh, npad / TC_T * nbq * TC_XH_BYTES)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&m->b_dxT, npad / TC_T * nbq * TC_DX_BYTES)) != cudaSuccess) return e;
  }
  m->batch_cap = n;
  return cudaSuccess;
}

cudaError_t launch_gemm_cols(b200_llama *m, const GemvPlan &p, int n, float *out, int ld_out, long long *launches);

// The mat-mul of a batch on the tensor cores (prefill_tc.cuh): fp16 operand copy of the prepared activations, then the
// persistent tcgen05 / TMEM kernel over (128-row tile, 16-token tile) items
cudaError_t launch_gemm_tc(b200_llama *m, const GemvPlan &p, int n, float *out, int ld_out, long long *launches) {
  const int npad = (n + TC_T - 1) / TC_T * TC_T;
  batch_act_tc_kernel<<<dim3((((p.nb + 3) & ~3) + 63) / 64, npad), 64, 0, m->stream>>>(m->b_act, batch_act_bytes(p.nb), m->b_xh, m->b_dxT, p.nb, n, npad);
  GemmTcArgs a = {};
  a.w = p.d_wtc; a.M = p.M; a.nb = p.nb; a.xh = m->b_xh; a.dxT = m->b_dxT; a.out = out; a.ld_out = ld_out; a.N = n; a.Npad = npad;
  a.spin_limit = 4000000000LL;
  const int n_items = ((p.M + TC_M - 1) / TC_M) * (npad / TC_T);
  q4_gemm_tc_kernel<<<std::min(m->n_sm, n_items), TC_THREADS, TC_SMEM, m->stream>>>(a);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

// tensor-core path for batches when the prefill copy exists; else the CUDA-core multi-column loop
cudaError_t launch_gemm_batch(b200_llama *m, const GemvPlan &p, int n, float *out, int ld_out, long long *launches) {
  // measured on B200 (profiles/r2_e_to_i_prompt_batches.md): one 16-token tile costs ~13 ms for the whole 7B model on the tensor-core
  // kernel (a 128-row tile walks its K loop alone), 8 columns ~5-9 ms on the CUDA-core loop: the crossover is near 20 tokens
  if (m->opt_tc && p.d_wtc && m->b_xh && n >= m->tc_min_n) return launch_gemm_tc(m, p, n, out, ld_out, launches);
  return launch_gemm_cols(m, p, n, out, ld_out, launches);
}

// out[n][ld] = W (plan p) x the n prepared activation vectors in m->b_act: weights streamed once per BATCH_NC columns
cudaError_t launch_gemm_cols(b200_llama *m, const GemvPlan &p, int n, float *out, int ld_out, long long *launches) {
  GemmColsArgs a = {};
  a.w = p.d_w; a.M = p.M; a.g_total = p.g_total; a.nb = p.nb; a.cb = p.cb; a.stage_bytes = p.stage_bytes; a.rmax = p.rmax;
  a.act = m->b_act; a.out = out; a.ld_out = ld_out; a.N = n;
  const int nbp = (p.nb + 3) & ~3;
  const size_t col_bytes = (size_t) (nbp + 2) * 32 + (size_t) nbp * 4;
  const long ring = (long) kSmemBudget - (long) (BATCH_NC * col_bytes) - 256;
  const int nchunks = ((p.nb + 3) / 4 + p.cb / 4 - 1) / (p.cb / 4);
  int S = (int) (ring / (p.stage_bytes + 16));
  if (S < 1) return cudaErrorInvalidConfiguration;
  S = std::min(S, nchunks);
  a.n_stages = S;
  const size_t smem = (size_t) S * p.stage_bytes + BATCH_NC * col_bytes + (size_t) 2 * S * 8;
  const dim3 grid(p.n_cta, (n + BATCH_NC - 1) / BATCH_NC), block(p.threads);
  switch (p.lp) {
    case 1: q4_gemm_cols_kernel<1><<<grid, block, smem, m->stream>>>(a); break;
    case 2: q4_gemm_cols_kernel<2><<<grid, block, smem, m->stream>>>(a); break;
    default: q4_gemm_cols_kernel<4><<<grid, block, smem, m->stream>>>(a); break;
  }
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// llama_eval for tokens [t0, t0 + n) of a call with n_call tokens in total (PO.mm:510-735, N > 1)
cudaError_t enqueue_batch_chunk(b200_llama *m, int n_threads, int n_past_call, int n_call, int t0, int n, const int32_t *tokens,
                                bool last_chunk, long long *launches) {
  cudaStream_t st = m->stream;
  cudaError_t e;
  const int E = m->n_embd, F = m->n_ff, hd = E / m->n_head;
  const int n_past = n_past_call + t0;
  long long nl = 0;
  if ((e = cudaMemcpyAsync(m->b_tok, tokens + t0, (size_t) n * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
  batch_embed_kernel<<<dim3((E + 255) / 256, n), 256, 0, st>>>(m->d_tok_emb, m->b_tok, m->b_x, E);
  nl++;
  auto ew = [](size_t total) { return (unsigned) ((total + 255) / 256); };
  for (int il = 0; il < m->n_layer; il++) {
    b200_llama::Layer &L = m->layers[il];
    float *k_layer = m->d_k + (size_t) il * m->n_ctx * E;
    float *v_layer = m->d_v + (size_t) il * m->n_ctx * E;
    batch_prep_kernel<1><<<n, 256, 0, st>>>(m->b_x, L.attn_norm, m->b_act, E);                       // PO.mm:570-575
    if ((e = launch_gemm_batch(m, L.qkv, n, m->b_o, 3 * E, &nl)) != cudaSuccess) return e;           // PO.mm:579-583
    batch_qkv_kernel<<<dim3((3 * E / 2 + 255) / 256, n), 256, 0, st>>>(m->b_o, n_past, m->b_q, k_layer, v_layer, m->d_rope, E, hd);
    {
      BatchAttnArgs a = {};
      a.q = m->b_q; a.k_layer = k_layer; a.v_layer = v_layer; a.out = m->b_att; a.exp_table = m->d_exp;
      a.n_embd = E; a.n_threads = n_threads; a.n_ctx = m->n_ctx; a.n_past = n_past; a.N = n_past_call + n_call - n_past;
      a.kq_scale = m->kq_scale; a.N_chunk = n;
      const size_t tile_smem = att_tile_smem(m->n_ctx);
      if (m->opt_attn_tile && n >= 2 && tile_smem <= (size_t) kSmemBudget)     // 8 queries share every K / V row they read
        batch_attn_tile_kernel<<<dim3(m->n_head, (n + ATT_QT - 1) / ATT_QT), ATT_TILE_THREADS, tile_smem, st>>>(a);
      else
        batch_attn_kernel<<<dim3(m->n_head * ATTN_CLUSTER, n), ATTN_THREADS, attn_smem_bytes(m, n_threads), st>>>(a);   // PO.mm:614-646
    }
    batch_prep_kernel<0><<<n, 256, 0, st>>>(m->b_att, nullptr, m->b_act, E);
    if ((e = launch_gemm_batch(m, L.wo, n, m->b_o, E, &nl)) != cudaSuccess) return e;                // PO.mm:649-651
    batch_resid_kernel<<<ew((size_t) n * E), 256, 0, st>>>(m->b_o, m->b_x, m->b_ff, (size_t) n * E);    // PO.mm:654
    batch_prep_kernel<1><<<n, 256, 0, st>>>(m->b_ff, L.ffn_norm, m->b_act, E);                        // PO.mm:660-665
    if ((e = launch_gemm_batch(m, L.w13, n, m->b_o, 2 * F, &nl)) != cudaSuccess) return e;           // PO.mm:668-676
    batch_silu_kernel<<<ew((size_t) n * F), 256, 0, st>>>(m->b_o, m->b_h, m->d_silu, F, (size_t) n * F);   // PO.mm:678-680
    batch_prep_kernel<0><<<n, 256, 0, st>>>(m->b_h, nullptr, m->b_act, F);
    if ((e = launch_gemm_batch(m, L.w2, n, m->b_o, E, &nl)) != cudaSuccess) return e;                // PO.mm:682-684
    batch_resid_kernel<<<ew((size_t) n * E), 256, 0, st>>>(m->b_o, m->b_ff, m->b_x, (size_t) n * E);    // PO.mm:687
    nl += 8;
  }
  if (last_chunk) {
    // the caller gets the logits of the LAST token only (PO.mm:724-725): final norm + lm_head on one column
    batch_prep_kernel<1><<<1, 256, 0, st>>>(m->b_x + (size_t) (n - 1) * E, m->d_norm, m->b_act, E);   // PO.mm:694-701
    if ((e = launch_gemm_cols(m, m->out, 1, m->d_logits, m->n_vocab, &nl)) != cudaSuccess) return e;  // PO.mm:705
    nl += 1;
  }
  if (launches) *launches += nl;
  return cudaGetLastError();
}

// One token of a batch on one rank.  The reference evaluates the N columns of every mat-mul independently, so a batch is
// run one token at a time; p_part carries n_past + N, the one place where the batch size enters the arithmetic (V*P
// partition, ggml.c:5628).
static int eval_enqueue_token(b200_llama *m, int n_threads, int token, int pos, int p_part, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  CUDA_TRY(cudaSetDevice(m->device));
  set_step_kernel<<<1, 1, 0, m->stream>>>(m->d_sp, token, pos, p_part, 0, 0);
  CUDA_TRY(cudaGetLastError());
  m->last_launches++;
  CUDA_TRY(run_token(m, n_threads, &m->last_launches));
  return B200_LLAMA_OK;
}

// the candidate stage of the sampler, enqueued behind the evaluation on the leader's stream
struct TopkReq {
  const int32_t *last_n;
  int n_last;
  double scale, penalty;
  int top_k;
};

static cudaError_t enqueue_topk(b200_llama *m, const TopkReq &tk) {
  cudaError_t e;
  if (!m->d_topk) {
    if ((e = cudaMalloc(&m->d_topk, sizeof(TopkOut))) != cudaSuccess) return e;
    if ((e = cudaMallocHost(&m->h_topk, sizeof(TopkOut))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&m->d_last_n, TOPK_MAX_LAST * 4)) != cudaSuccess) return e;
    if ((e = cudaMallocHost(&m->h_last_n, TOPK_MAX_LAST * 4)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sample_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) topk_smem_bytes(m->n_vocab))) != cudaSuccess) return e;
  }
  if (tk.n_last > 0) {
    memcpy(m->h_last_n, tk.last_n, (size_t) tk.n_last * 4);
    if ((e = cudaMemcpyAsync(m->d_last_n, m->h_last_n, (size_t) tk.n_last * 4, cudaMemcpyHostToDevice, m->stream)) != cudaSuccess) return e;
  }
  sample_topk_kernel<<<1, TOPK_THREADS, topk_smem_bytes(m->n_vocab), m->stream>>>(m->d_logits, m->n_vocab, m->d_last_n, tk.n_last, tk.scale,
                                                                                   tk.penalty, tk.top_k, m->d_topk);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  m->last_launches++;
  return cudaMemcpyAsync(m->h_topk, m->d_topk, sizeof(TopkOut), cudaMemcpyDeviceToHost, m->stream);
}

// logits_out == nullptr: the logits stay on the device (tk != nullptr: only the sampler's candidates come back)
static int eval_core(b200_llama *m, int n_threads, int n_past, const int32_t *tokens, int n_tokens, float *logits_out,
                     const TopkReq *tk, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  if (!m || !tokens) { set_err(err, errlen, "null argument"); return fail_code; }
  if (n_tokens < 1 || n_past < 0 || n_past + n_tokens > m->n_ctx) {
    set_err(err, errlen, "n_past %d + n_tokens %d exceeds n_ctx %d", n_past, n_tokens, m->n_ctx);
    return fail_code;
  }
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) { set_err(err, errlen, "n_threads %d > 64 not supported", n_threads); return fail_code; }
  for (int i = 0; i < n_tokens; i++) {
    if (tokens[i] < 0 || tokens[i] >= m->n_vocab) { set_err(err, errlen, "token id %d out of range", tokens[i]); return fail_code; }
  }
  if (const char *why = tp_ready(m, n_threads)) { set_err(err, errlen, "%s", why); return fail_code; }
  const std::vector<b200_llama *> ranks = ranks_of(m);
  for (b200_llama *r : ranks) r->last_launches = 0;
  if (batch_usable(m, n_tokens)) {
    // prompt batch: every weight row is read once per group of columns (batch.cuh), not once per token
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(batch_reserve(m, std::min(n_tokens, kBatchChunk)));
    CUDA_TRY(cudaEventRecord(m->ev0, m->stream));
    for (int t0 = 0; t0 < n_tokens; t0 += kBatchChunk) {
      const int n = std::min(kBatchChunk, n_tokens - t0);
      CUDA_TRY(enqueue_batch_chunk(m, n_threads, n_past, n_tokens, t0, n, tokens, t0 + n == n_tokens, &m->last_launches));
    }
    CUDA_TRY(cudaEventRecord(m->ev1, m->stream));
    if (logits_out) CUDA_TRY(cudaMemcpyAsync(m->h_logits, m->d_logits, (size_t) m->n_vocab * 4, cudaMemcpyDeviceToHost, m->stream));
    if (tk) CUDA_TRY(enqueue_topk(m, *tk));
    CUDA_TRY(abort_fetch_async(m));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    { float ms = 0; if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) m->last_eval_ms = ms; }
    if (abort_check_and_reset(ranks)) { set_err(err, errlen, "a device-side wait timed out; the evaluation was abandoned and the exchange state reset"); return fail_code; }
    if (logits_out) memcpy(logits_out, m->h_logits, (size_t) m->n_vocab * 4);
    return B200_LLAMA_OK;
  }
  // Token-major enqueue: the token kernels of a group wait for each other ON THE GPUS, so every rank must receive
  // token i before any rank receives so much work that the driver's launch queue blocks the host (one host thread
  // drives all ranks of a single-process group).
  for (int i = 0; i < n_tokens; i++) {
    for (b200_llama *r : ranks) {
      const int rc = eval_enqueue_token(r, n_threads, tokens[i], n_past + i, n_past + n_tokens, err, errlen);
      if (rc != B200_LLAMA_OK) return rc;
    }
    if (ranks.size() > 1 && (i & 63) == 63) {      // bound the work in flight per rank
      for (b200_llama *r : ranks) { CUDA_TRY(cudaSetDevice(r->device)); CUDA_TRY(cudaStreamSynchronize(r->stream)); }
    }
  }
  for (b200_llama *r : ranks) {
    CUDA_TRY(cudaSetDevice(r->device));
    if (logits_out && r == m) CUDA_TRY(cudaMemcpyAsync(r->h_logits, r->d_logits, (size_t) r->n_vocab * 4, cudaMemcpyDeviceToHost, r->stream));
    if (tk && r == m) CUDA_TRY(enqueue_topk(m, *tk));
    CUDA_TRY(abort_fetch_async(r));
  }
  for (b200_llama *r : ranks) {
    CUDA_TRY(cudaSetDevice(r->device));
    CUDA_TRY(cudaStreamSynchronize(r->stream));
  }
  if (abort_check_and_reset(ranks)) { set_err(err, errlen, "a device-side wait timed out; the evaluation was abandoned and the exchange state reset"); return fail_code; }
  if (logits_out) memcpy(logits_out, m->h_logits, (size_t) m->n_vocab * 4);      // every rank holds the full logits; the leader's are returned
  return B200_LLAMA_OK;
}

int b200_llama_eval(b200_llama *m, int n_threads, int n_past, const int32_t *tokens, int n_tokens, float *logits_out,
                    char *err, size_t errlen) {
  if (!logits_out) { set_err(err, errlen, "null argument"); return B200_LLAMA_ERR_PREDICT; }
  return eval_core(m, n_threads, n_past, tokens, n_tokens, logits_out, nullptr, err, errlen);
}

int b200_llama_eval_topk(b200_llama *m, int n_threads, int n_past, const int32_t *tokens, int n_tokens,
                         const int32_t *last_n_tokens, int n_last, double repeat_penalty, double temp, int top_k,
                         double *cand_values, int32_t *cand_ids, int *n_cand, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  if (!m || !cand_values || !cand_ids || !n_cand || (n_last > 0 && !last_n_tokens)) { set_err(err, errlen, "null argument"); return fail_code; }
  *n_cand = 0;
  // outside what the kernel covers the evaluation still runs; the caller fetches the logits (b200_llama_last_logits)
  const bool covered = top_k >= 1 && top_k <= TOPK_MAX_K && n_last >= 0 && n_last <= TOPK_MAX_LAST &&
                       topk_smem_bytes(m->n_vocab) <= 40 * 1024;
  TopkReq tk{last_n_tokens, n_last, 1.0 / temp, repeat_penalty, std::min(top_k, m->n_vocab)};   // const double scale = 1.0/temp, utils.cpp:357
  const int rc = eval_core(m, n_threads, n_past, tokens, n_tokens, nullptr, covered ? &tk : nullptr, err, errlen);
  if (rc != B200_LLAMA_OK || !covered) return rc;
  const TopkOut &o = *m->h_topk;
  if (o.ambiguous || o.n < tk.top_k) return B200_LLAMA_OK;       // equal values among the best: only the reference's own partial_sort knows their order
  for (int i = 0; i < tk.top_k; i++) { cand_values[i] = o.values[i]; cand_ids[i] = o.ids[i]; }
  *n_cand = tk.top_k;
  return B200_LLAMA_OK;
}

int b200_llama_last_logits(b200_llama *m, float *logits_out, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  if (!m || !logits_out) { set_err(err, errlen, "null argument"); return fail_code; }
  CUDA_TRY(cudaSetDevice(m->device));
  CUDA_TRY(cudaMemcpyAsync(m->h_logits, m->d_logits, (size_t) m->n_vocab * 4, cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  memcpy(logits_out, m->h_logits, (size_t) m->n_vocab * 4);
  return B200_LLAMA_OK;
}

static int decode_prepare(b200_llama *m, int n_steps, const int32_t *forced_tokens, bool want_logits, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  CUDA_TRY(cudaSetDevice(m->device));
  if (m->log_cap < n_steps) {
    cudaFree(m->d_token_log); cudaFree(m->d_forced);
    m->d_token_log = nullptr; m->d_forced = nullptr;
    CUDA_TRY(cudaMalloc(&m->d_token_log, (size_t) n_steps * 4));
    CUDA_TRY(cudaMalloc(&m->d_forced, (size_t) n_steps * 4));
    m->log_cap = n_steps;
  }
  if (forced_tokens) CUDA_TRY(cudaMemcpy(m->d_forced, forced_tokens, (size_t) n_steps * 4, cudaMemcpyHostToDevice));
  if (want_logits) {
    const size_t need = (size_t) n_steps * m->n_vocab;
    if (m->logits_log_cap < need) {
      cudaFree(m->d_logits_log); m->d_logits_log = nullptr;
      CUDA_TRY(cudaMalloc(&m->d_logits_log, need * 4));
      m->logits_log_cap = need;
    }
  }
  if (m->opt_time_kernel) {
    while ((int) m->kev.size() < 2 * n_steps) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); m->kev.push_back(e); }
  }
  return B200_LLAMA_OK;
}

static int decode_begin(b200_llama *m, int n_past, int first_token, bool forced, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  CUDA_TRY(cudaSetDevice(m->device));
  m->last_launches = 0;
  set_step_kernel<<<1, 1, 0, m->stream>>>(m->d_sp, first_token, n_past, n_past + 1, 0, forced ? 1 : 0);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(m->ev0, m->stream));
  return B200_LLAMA_OK;
}

static int decode_step(b200_llama *m, int n_threads, int i, bool forced, bool want_logits, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  CUDA_TRY(cudaSetDevice(m->device));
  if (m->opt_time_kernel) CUDA_TRY(cudaEventRecord(m->kev[2 * i], m->stream));
  const bool fold = fold_usable(m, n_threads);      // the greedy pick runs inside the token kernel (one launch per token)
  m->fold_argmax = fold ? 1 : 0;
  CUDA_TRY(run_token(m, n_threads, &m->last_launches));
  m->fold_argmax = 0;
  if (m->opt_time_kernel) CUDA_TRY(cudaEventRecord(m->kev[2 * i + 1], m->stream));
  if (want_logits) {
    CUDA_TRY(cudaMemcpyAsync(m->d_logits_log + (size_t) i * m->n_vocab, m->d_logits, (size_t) m->n_vocab * 4, cudaMemcpyDeviceToDevice, m->stream));
  }
  if (!fold) {
    argmax_advance_kernel<<<1, 1024, 0, m->stream>>>(m->d_logits, m->n_vocab, m->d_sp, m->d_token_log, forced ? m->d_forced : nullptr);
    CUDA_TRY(cudaGetLastError());
    m->last_launches++;
  }
  return B200_LLAMA_OK;
}

int b200_llama_decode_device(b200_llama *m, int n_threads, int n_past, int first_token, int n_steps,
                             const int32_t *forced_tokens, int32_t *tokens_out, float *logits_all, float *elapsed_ms,
                             char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  if (!m) { set_err(err, errlen, "null model"); return fail_code; }
  if (n_steps < 1 || n_past < 0 || n_past + n_steps > m->n_ctx) {
    set_err(err, errlen, "n_past %d + n_steps %d exceeds n_ctx %d", n_past, n_steps, m->n_ctx);
    return fail_code;
  }
  if (first_token < 0 || first_token >= m->n_vocab) { set_err(err, errlen, "token id %d out of range", first_token); return fail_code; }
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) { set_err(err, errlen, "n_threads %d > 64 not supported", n_threads); return fail_code; }
  if (forced_tokens) {
    for (int i = 0; i < n_steps; i++)
      if (forced_tokens[i] < 0 || forced_tokens[i] >= m->n_vocab) { set_err(err, errlen, "token id %d out of range", forced_tokens[i]); return fail_code; }
  }
  if (const char *why = tp_ready(m, n_threads)) { set_err(err, errlen, "%s", why); return fail_code; }
  const std::vector<b200_llama *> ranks = ranks_of(m);
  for (b200_llama *r : ranks) {
    r->opt_time_kernel = m->opt_time_kernel;
    const int rc = decode_prepare(r, n_steps, forced_tokens, r == m && logits_all != nullptr, err, errlen);
    if (rc != B200_LLAMA_OK) return rc;
  }
  for (b200_llama *r : ranks) {
    const int rc = decode_begin(r, n_past, first_token, forced_tokens != nullptr, err, errlen);
    if (rc != B200_LLAMA_OK) return rc;
  }
  // step-major enqueue (see b200_llama_eval): rank 0's step i must not be queued behind hundreds of its own later steps
  // while rank 1 has not been given step i yet
  for (int i = 0; i < n_steps; i++) {
    for (b200_llama *r : ranks) {
      const int rc = decode_step(r, n_threads, i, forced_tokens != nullptr, r == m && logits_all != nullptr, err, errlen);
      if (rc != B200_LLAMA_OK) return rc;
    }
    if (ranks.size() > 1 && (i & 63) == 63) {
      for (b200_llama *r : ranks) { CUDA_TRY(cudaSetDevice(r->device)); CUDA_TRY(cudaStreamSynchronize(r->stream)); }
    }
  }
  for (b200_llama *r : ranks) {
    CUDA_TRY(cudaSetDevice(r->device));
    CUDA_TRY(cudaEventRecord(r->ev1, r->stream));
    CUDA_TRY(abort_fetch_async(r));
  }
  for (b200_llama *r : ranks) {
    CUDA_TRY(cudaSetDevice(r->device));
    CUDA_TRY(cudaStreamSynchronize(r->stream));
  }
  if (abort_check_and_reset(ranks)) { set_err(err, errlen, "a device-side wait timed out; the run was abandoned and the exchange state reset"); return fail_code; }
  CUDA_TRY(cudaSetDevice(m->device));
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, m->ev0, m->ev1));
  if (m->opt_time_kernel) {
    m->last_kernel_ms = 0.0;
    for (int i = 0; i < n_steps; i++) { float t = 0; CUDA_TRY(cudaEventElapsedTime(&t, m->kev[2 * i], m->kev[2 * i + 1])); m->last_kernel_ms += t; }
  }
  if (tokens_out) CUDA_TRY(cudaMemcpy(tokens_out, m->d_token_log, (size_t) n_steps * 4, cudaMemcpyDeviceToHost));
  if (logits_all) CUDA_TRY(cudaMemcpy(logits_all, m->d_logits_log, (size_t) n_steps * m->n_vocab * 4, cudaMemcpyDeviceToHost));
  return B200_LLAMA_OK;
}

void b200_llama_free(b200_llama *m);

// ---- resident models ----------------------------------------------------------------------
// The reference reads the model file again on every run() (llama_model_load inside -[LlamaPredictOperation main],
// PO.mm:790): 4 GB from disk + upload for a 7B model, several seconds against 0.8 s for 512 generated tokens.  A
// LlamaRunner-lifetime cache keeps the model in HBM between runs: acquire hands out the resident handle for the same
// (file, n_ctx, device) when no other operation is using it (concurrent operations each get a private model, as they
// each need their own KV cache), release gives it back without freeing.  The KV cache needs no reset: llama_eval only
// ever reads rows that the same run has written (positions < n_past).
namespace {
struct CacheEntry {
  std::string path;
  int n_ctx, device;
  long long size, mtime_ns;
  b200_llama *model;
  bool in_use;
};
std::mutex g_cache_mu;
std::vector<CacheEntry> g_cache;

bool file_identity(const char *path, long long *size, long long *mtime_ns) {
  struct stat st;
  if (stat(path, &st) != 0) return false;
  *size = (long long) st.st_size;
  *mtime_ns = (long long) st.st_mtim.tv_sec * 1000000000LL + st.st_mtim.tv_nsec;
  return true;
}
}  // namespace

int b200_llama_acquire(const char *path, int n_ctx, int device, b200_llama **out, char *err, size_t errlen) {
  if (!out || !path) { set_err(err, errlen, "null argument"); return B200_LLAMA_ERR_LOAD; }
  *out = nullptr;
  long long size = 0, mtime = 0;
  const bool have_id = file_identity(path, &size, &mtime);
  std::vector<b200_llama *> stale;
  if (have_id) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    for (CacheEntry &e : g_cache) {
      if (!e.in_use && e.path == path && e.n_ctx == n_ctx && e.device == device && e.size == size && e.mtime_ns == mtime) {
        e.in_use = true;
        *out = e.model;
        return B200_LLAMA_OK;
      }
    }
    // the file was rewritten since an idle entry of the same path was loaded: that entry can never match again
    for (size_t i = 0; i < g_cache.size();) {
      CacheEntry &e = g_cache[i];
      if (!e.in_use && e.path == path && (e.size != size || e.mtime_ns != mtime)) { stale.push_back(e.model); g_cache.erase(g_cache.begin() + (long) i); }
      else i++;
    }
  }
  for (b200_llama *old_model : stale) free_model(old_model);
  b200_llama *m = nullptr;
  const int rc = b200_llama_load(path, n_ctx, device, &m, err, errlen);     // same errors as an uncached load
  if (rc != B200_LLAMA_OK) return rc;
  if (have_id) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    g_cache.push_back(CacheEntry{path, n_ctx, device, size, mtime, m, true});
  }
  *out = m;
  return B200_LLAMA_OK;
}

void b200_llama_release(b200_llama *m) {
  if (!m) return;
  // a model whose last run failed on the device (a trapped kernel poisons the context) must not be handed out again
  bool healthy = cudaSetDevice(m->device) == cudaSuccess;
  if (healthy) {
    const cudaError_t q = cudaStreamQuery(m->stream);
    healthy = q == cudaSuccess || q == cudaErrorNotReady;
  }
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    size_t n_idle_same = 0;
    for (size_t i = 0; i < g_cache.size(); i++) {
      if (g_cache[i].model != m) continue;
      for (const CacheEntry &o : g_cache)
        if (!o.in_use && o.path == g_cache[i].path && o.n_ctx == g_cache[i].n_ctx && o.device == g_cache[i].device) n_idle_same++;
      // keep at most one idle copy per (file, n_ctx, device): the private models of concurrent runs are freed on release
      if (healthy && n_idle_same == 0) { g_cache[i].in_use = false; return; }
      g_cache.erase(g_cache.begin() + (long) i);
      break;
    }
  }
  free_model(m);      // never cached (no file identity), unhealthy, or a surplus copy: behaves like b200_llama_free
}

void b200_llama_free(b200_llama *m) {
  if (!m) return;
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);      // a handle obtained with acquire may be freed directly: forget it
    for (size_t i = 0; i < g_cache.size(); i++)
      if (g_cache[i].model == m) { g_cache.erase(g_cache.begin() + (long) i); break; }
  }
  free_model(m);
}

void b200_llama_cache_clear(void) {
  std::vector<b200_llama *> idle;
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    for (size_t i = 0; i < g_cache.size();) {
      if (!g_cache[i].in_use) { idle.push_back(g_cache[i].model); g_cache.erase(g_cache.begin() + (long) i); }
      else i++;
    }
  }
  for (b200_llama *m : idle) free_model(m);
}

int b200_llama_n_vocab(const b200_llama *m) { return m->n_vocab; }
int b200_llama_n_ctx(const b200_llama *m) { return m->n_ctx; }
int b200_llama_n_embd(const b200_llama *m) { return m->n_embd; }
int b200_llama_n_layer(const b200_llama *m) { return m->n_layer; }
int b200_llama_n_head(const b200_llama *m) { return m->n_head; }
int b200_llama_ftype(const b200_llama *m) { return m->f16; }

const b200_tokenizer *b200_llama_shared_tokenizer(b200_llama *m) {
  if (!m) return nullptr;
  std::lock_guard<std::mutex> lock(g_cache_mu);
  if (!m->shared_tok) m->shared_tok = b200_tokenizer_create(m);
  return m->shared_tok;
}

const char *b200_llama_token_str(const b200_llama *m, int id, int *len) {
  if (id < 0 || id >= m->n_vocab) { if (len) *len = 0; return ""; }
  if (len) *len = (int) m->id_to_token[id].size();
  return m->id_to_token[id].data();
}

int b200_llama_kv_export(const b200_llama *m, int layer, int which, int n_rows, float *out) {
  if (!m || layer < 0 || layer >= m->n_layer || n_rows < 0 || n_rows > m->n_ctx) return B200_LLAMA_ERR_PREDICT;
  if (cudaSetDevice(m->device) != cudaSuccess) return B200_LLAMA_ERR_PREDICT;
  const float *base = (which == 0 ? m->d_k : m->d_v) + (size_t) layer * m->n_ctx * m->n_embd;
  if (cudaStreamSynchronize(m->stream) != cudaSuccess) return B200_LLAMA_ERR_PREDICT;
  return cudaMemcpy(out, base, (size_t) n_rows * m->n_embd * 4, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : B200_LLAMA_ERR_PREDICT;
}

int b200_llama_kv_import(b200_llama *m, int layer, int which, int n_rows, const float *in) {
  if (!m || layer < 0 || layer >= m->n_layer || n_rows < 0 || n_rows > m->n_ctx) return B200_LLAMA_ERR_PREDICT;
  if (cudaSetDevice(m->device) != cudaSuccess) return B200_LLAMA_ERR_PREDICT;
  float *base = (which == 0 ? m->d_k : m->d_v) + (size_t) layer * m->n_ctx * m->n_embd;
  if (cudaStreamSynchronize(m->stream) != cudaSuccess) return B200_LLAMA_ERR_PREDICT;
  return cudaMemcpy(base, in, (size_t) n_rows * m->n_embd * 4, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : B200_LLAMA_ERR_PREDICT;
}

long long b200_llama_last_launches(const b200_llama *m) { return m->last_launches; }
double b200_llama_last_kernel_ms(const b200_llama *m) { return m->last_kernel_ms; }
double b200_llama_last_eval_ms(const b200_llama *m) { return m->last_eval_ms; }
long long b200_llama_weight_bytes(const b200_llama *m) { return m->weight_bytes; }

/* Development profiler: run ONE token (current step scalars) through the whole-token kernel with per-CTA globaltimer
 * stamps at every phase boundary.  out receives n_cta * marks int64 nanosecond stamps; returns marks (or < 0). */
int b200_llama_profile_token(b200_llama *m, int n_threads, int token, int pos, long long *out, int cap, int *n_cta) {
  if (!m || !mega_usable(m, n_threads) || tp_ready(m, n_threads)) return -1;
  // A tensor-parallel group is profiled as a whole: every rank of a single-process group runs the stamped token here; the
  // ranks of a one-process-per-GPU group must all make this call at the same time.  The leader's stamps are returned.
  const std::vector<b200_llama *> ranks = ranks_of(m);
  const int marks = 2 + 26 * m->n_layer + 12;
  if ((long long) marks * m->n_sm > cap) return -2;
  cudaError_t e = cudaSuccess;
  for (b200_llama *r : ranks) {
    cudaSetDevice(r->device);
    if (cudaMalloc(&r->d_prof, (size_t) marks * r->n_sm * 8) != cudaSuccess) return -3;
    cudaMemset(r->d_prof, 0, (size_t) marks * r->n_sm * 8);
    r->prof_marks = marks;
    set_step_kernel<<<1, 1, 0, r->stream>>>(r->d_sp, token, pos, pos + 1, 0, 0);
  }
  for (b200_llama *r : ranks) {
    cudaSetDevice(r->device);
    long long dummy = 0;
    if (e == cudaSuccess) e = enqueue_token_mega(r, n_threads, &dummy);
  }
  for (b200_llama *r : ranks) {
    cudaSetDevice(r->device);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->stream);
  }
  cudaSetDevice(m->device);
  if (e == cudaSuccess) e = cudaMemcpy(out, m->d_prof, (size_t) marks * m->n_sm * 8, cudaMemcpyDeviceToHost);
  for (b200_llama *r : ranks) {
    cudaSetDevice(r->device);
    cudaFree(r->d_prof);
    r->d_prof = nullptr;
    r->prof_marks = 0;
  }
  if (n_cta) *n_cta = m->n_sm;
  return e == cudaSuccess ? marks : -4;
}

int b200_llama_set_option(b200_llama *m, const char *key, int value) {
  if (!m || !key) return -1;
  if (!strcmp(key, "graph")) { m->opt_graph = value; return 0; }
  if (!strcmp(key, "pdl")) { m->opt_pdl = value; return 0; }
  if (!strcmp(key, "mega")) { m->opt_mega = value; return 0; }
  if (!strcmp(key, "time_kernel")) { m->opt_time_kernel = value; return 0; }
  if (!strcmp(key, "fold_argmax")) { m->opt_fold = value; return 0; }
  if (!strcmp(key, "batch")) { m->opt_batch = value; return 0; }
  if (!strcmp(key, "tc")) { m->opt_tc = value; return 0; }
  if (!strcmp(key, "attn_tile")) { m->opt_attn_tile = value; return 0; }
  if (!strcmp(key, "tc_min_n")) { m->tc_min_n = value; return 0; }
  if (!strcmp(key, "spin_limit_cycles")) { m->opt_spin_limit = value; return 0; }    // test hook: provoke the bounded-wait abort path
  return -1;
}

static int q4_matvec_impl(int qtype, int device, const void *w_ggml, int M, int K, const float *x, float *out, int lane_pairs,
                          float *kernel_ms, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { set_err(err, errlen, "no CUDA device available (this library has no CPU path)"); return fail_code; }
  if (M < 1 || K < 64 || K % 64 != 0) { set_err(err, errlen, "bad shape %d x %d", M, K); return fail_code; }
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  b200_llama tmp;   // only stream / counters are used by upload_matrix
  tmp.device = device;
  CUDA_TRY(cudaStreamCreateWithFlags(&tmp.stream, cudaStreamNonBlocking));
  GemvPlan p = make_plan(M, K, env_int("B200_NUM_CTAS", prop.multiProcessorCount), lane_pairs, qtype);
  HostTensor t;             // the caller's ggml rows, referenced in place
  t.n_dims = 2; t.ne[0] = K; t.ne[1] = M; t.part[0] = (const uint8_t *) w_ggml; t.row_bytes = (size_t) (K / 32) * (qtype == 3 ? 24 : 20);
  Stager stager;
  uint8_t *d_stage = nullptr;
  float *d_x = nullptr, *d_out = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = B200_LLAMA_OK;
  auto cleanup = [&]() {
    cudaFree(d_stage); cudaFree(d_x); cudaFree(d_out); cudaFree(p.d_w);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaStreamDestroy(tmp.stream);
  };
#define MV_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { set_err(err, errlen, "CUDA error %s (%s)", cudaGetErrorString(e__), #expr); cleanup(); return fail_code; } } while (0)
  MV_TRY(configure_kernels());
  MV_TRY(stager.init());
  MV_TRY(cudaMalloc(&d_stage, t.row_bytes * (size_t) M));
  MV_TRY(upload_matrix(&tmp, stager, p, {rows_of(t, 0, M)}, 0, d_stage));
  MV_TRY(cudaStreamSynchronize(tmp.stream));
  MV_TRY(cudaMalloc(&d_x, (size_t) K * 4));
  MV_TRY(cudaMalloc(&d_out, (size_t) M * 4));
  MV_TRY(cudaMemcpy(d_x, x, (size_t) K * 4, cudaMemcpyHostToDevice));
  MV_TRY(cudaEventCreate(&e0));
  MV_TRY(cudaEventCreate(&e1));
  GemvArgs a = base_args(p);
  a.x = d_x; a.out = d_out;
  const int reps = kernel_ms ? 5 : 1;
  float best = 1e30f;
  for (int i = 0; i < reps; i++) {
    MV_TRY(cudaEventRecord(e0, tmp.stream));
    MV_TRY((launch_gemv<PRO_PLAIN, EPI_STORE>(p, a, tmp.stream, false)));
    MV_TRY(cudaEventRecord(e1, tmp.stream));
    MV_TRY(cudaStreamSynchronize(tmp.stream));
    float ms = 0;
    MV_TRY(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  if (kernel_ms) *kernel_ms = best;
  MV_TRY(cudaMemcpy(out, d_out, (size_t) M * 4, cudaMemcpyDeviceToHost));
#undef MV_TRY
  cleanup();
  return rc;
}


/* Kernel-level entry of the sampler's candidate stage (sample_topk.cuh) on host logits: what b200_llama_eval_topk runs behind
 * the evaluation.  *n_cand = top_k, or 0 when the candidate order is ambiguous (see the header).  kernel_ms: best of 5. */
int b200_sample_topk(int device, const float *logits, int n_vocab, const int32_t *last_n_tokens, int n_last, double repeat_penalty,
                     double temp, int top_k, double *cand_values, int32_t *cand_ids, int *n_cand, float *kernel_ms,
                     char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { set_err(err, errlen, "no CUDA device available (this library has no CPU path)"); return fail_code; }
  if (!logits || !cand_values || !cand_ids || !n_cand || n_vocab < 1 || top_k < 1 || top_k > TOPK_MAX_K || n_last < 0 || n_last > TOPK_MAX_LAST ||
      topk_smem_bytes(n_vocab) > 40 * 1024) { set_err(err, errlen, "bad argument"); return fail_code; }
  CUDA_TRY(cudaSetDevice(device));
  float *d_logits = nullptr;
  int32_t *d_last = nullptr;
  TopkOut *d_out = nullptr;
  TopkOut h_out;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&]() { cudaFree(d_logits); cudaFree(d_last); cudaFree(d_out); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); };
#define TK_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_err(err, errlen, "%s: %s", #x, cudaGetErrorString(e_)); cleanup(); return fail_code; } } while (0)
  TK_TRY(cudaMalloc(&d_logits, (size_t) n_vocab * 4));
  TK_TRY(cudaMalloc(&d_last, TOPK_MAX_LAST * 4));
  TK_TRY(cudaMalloc(&d_out, sizeof(TopkOut)));
  TK_TRY(cudaMemcpy(d_logits, logits, (size_t) n_vocab * 4, cudaMemcpyHostToDevice));
  if (n_last) TK_TRY(cudaMemcpy(d_last, last_n_tokens, (size_t) n_last * 4, cudaMemcpyHostToDevice));
  TK_TRY(cudaFuncSetAttribute(sample_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) topk_smem_bytes(n_vocab)));
  TK_TRY(cudaEventCreate(&e0));
  TK_TRY(cudaEventCreate(&e1));
  const int k = std::min(top_k, n_vocab);
  float best = 1e30f;
  for (int rep = 0; rep < (kernel_ms ? 5 : 1); rep++) {
    TK_TRY(cudaEventRecord(e0));
    sample_topk_kernel<<<1, TOPK_THREADS, topk_smem_bytes(n_vocab)>>>(d_logits, n_vocab, d_last, n_last, 1.0 / temp, repeat_penalty, k, d_out);
    TK_TRY(cudaGetLastError());
    TK_TRY(cudaEventRecord(e1));
    TK_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    TK_TRY(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  TK_TRY(cudaMemcpy(&h_out, d_out, sizeof(TopkOut), cudaMemcpyDeviceToHost));
#undef TK_TRY
  cleanup();
  if (kernel_ms) *kernel_ms = best;
  *n_cand = 0;
  if (!h_out.ambiguous && h_out.n >= k) {
    for (int i = 0; i < k; i++) { cand_values[i] = h_out.values[i]; cand_ids[i] = h_out.ids[i]; }
    *n_cand = k;
  }
  return B200_LLAMA_OK;
}

#if B200_TC_TRACE
/* development build only (-DB200_TC_TRACE=1): the hand-over timeline of CTA 0 of the last tcgen05 mat-mul launch */
extern "C" int b200_debug_tc_trace(long long *out, int n_words) {
  const size_t n = std::min<size_t>((size_t) n_words, 12 * 512) * sizeof(long long);
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, b200::g_tc_trace, n) == cudaSuccess ? 0 : -1;
}
#endif

/* out[N][M] = W[M x K] (Q4_0, ggml rows) x the N columns x[N][K], every column computed exactly as
 * ggml_compute_forward_mul_mat_q4_0_f32 does (ggml.c:6199-6222).  path 0: CUDA-core multi-column loop (weights streamed
 * once per 8 columns), path 1: tcgen05 / TMEM kernel.  kernel_ms: best-of-5 device time of the mat-mul kernel(s) alone. */
int b200_q4_0_matmul(int device, const void *w_ggml, int M, int K, const float *x, int N, float *out, int path,
                     float *kernel_ms, char *err, size_t errlen) {
  const int fail_code = B200_LLAMA_ERR_PREDICT;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { set_err(err, errlen, "no CUDA device available (this library has no CPU path)"); return fail_code; }
  if (M < 1 || K < 64 || K % 64 != 0 || N < 1 || N > 4096) { set_err(err, errlen, "bad shape %d x %d, N %d", M, K, N); return fail_code; }
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  b200_llama tmp;
  tmp.device = device;
  tmp.n_sm = env_int("B200_NUM_CTAS", prop.multiProcessorCount);
  tmp.want_tc_copy = path == 1;
  CUDA_TRY(cudaStreamCreateWithFlags(&tmp.stream, cudaStreamNonBlocking));
  GemvPlan p = make_plan(M, K, tmp.n_sm, 0, 2);
  const size_t wbytes = (size_t) M * (K / 32) * 20;
  uint8_t *d_stage = nullptr;
  float *d_x = nullptr, *d_out = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const int npad = (N + TC_T - 1) / TC_T * TC_T, nbq = (K / 32 + 3) / 4;
  auto cleanup = [&]() {
    cudaFree(d_stage); cudaFree(d_x); cudaFree(d_out); cudaFree(p.d_w); cudaFree(p.d_wtc); cudaFree(tmp.b_act); cudaFree(tmp.b_xh); cudaFree(tmp.b_dxT);
    tmp.b_act = nullptr; tmp.b_xh = nullptr; tmp.b_dxT = nullptr;
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaStreamDestroy(tmp.stream);
    tmp.stream = nullptr;
  };
#define MM_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { set_err(err, errlen, "CUDA error %s (%s)", cudaGetErrorString(e__), #expr); cleanup(); return fail_code; } } while (0)
  MM_TRY(configure_kernels());
  MM_TRY(cudaMalloc(&d_stage, wbytes));
  {
    HostTensor t;           // the caller's ggml rows, referenced in place
    t.n_dims = 2; t.ne[0] = K; t.ne[1] = M; t.part[0] = (const uint8_t *) w_ggml; t.row_bytes = (size_t) (K / 32) * 20;
    Stager stager;
    MM_TRY(stager.init());
    MM_TRY(upload_matrix(&tmp, stager, p, {rows_of(t, 0, M)}, 0, d_stage));
    MM_TRY(cudaStreamSynchronize(tmp.stream));
  }
  MM_TRY(cudaMalloc(&d_x, (size_t) N * K * 4));
  MM_TRY(cudaMalloc(&d_out, (size_t) N * M * 4));
  MM_TRY(cudaMalloc(&tmp.b_act, (size_t) N * batch_act_bytes(K / 32)));
  if (path == 1) {
    MM_TRY(cudaMalloc(&tmp.b_xh, (size_t) npad / TC_T * nbq * TC_XH_BYTES));
    MM_TRY(cudaMalloc(&tmp.b_dxT, (size_t) npad / TC_T * nbq * TC_DX_BYTES));
  }
  MM_TRY(cudaMemcpy(d_x, x, (size_t) N * K * 4, cudaMemcpyHostToDevice));
  MM_TRY(cudaEventCreate(&e0));
  MM_TRY(cudaEventCreate(&e1));
  batch_prep_kernel<0><<<N, 256, 0, tmp.stream>>>(d_x, nullptr, tmp.b_act, K);
  MM_TRY(cudaGetLastError());
  float best = 1e30f;
  for (int i = 0; i < (kernel_ms ? 5 : 1); i++) {
    MM_TRY(cudaEventRecord(e0, tmp.stream));
    MM_TRY(path == 1 ? launch_gemm_tc(&tmp, p, N, d_out, M, nullptr) : launch_gemm_cols(&tmp, p, N, d_out, M, nullptr));
    MM_TRY(cudaEventRecord(e1, tmp.stream));
    MM_TRY(cudaStreamSynchronize(tmp.stream));
    float ms = 0;
    MM_TRY(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  if (kernel_ms) *kernel_ms = best;
  MM_TRY(cudaMemcpy(out, d_out, (size_t) N * M * 4, cudaMemcpyDeviceToHost));
#undef MM_TRY
  cleanup();
  return B200_LLAMA_OK;
}

int b200_q4_0_matvec(int device, const void *w_ggml, int M, int K, const float *x, float *out, int lane_pairs,
                     float *kernel_ms, char *err, size_t errlen) {
  return q4_matvec_impl(2, device, w_ggml, M, K, x, out, lane_pairs, kernel_ms, err, errlen);
}

int b200_q4_1_matvec(int device, const void *w_ggml, int M, int K, const float *x, float *out, float *kernel_ms,
                     char *err, size_t errlen) {
  return q4_matvec_impl(3, device, w_ggml, M, K, x, out, 0, kernel_ms, err, errlen);
}

}  // extern "C"
