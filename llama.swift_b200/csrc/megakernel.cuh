// decode_token_kernel: one persistent launch evaluates one token through the whole network (llama_eval with N = 1,
// PO.mm:510-735).  It replaces ggml_graph_compute's thread pool (ggml.c:9109-9555) with a B200-shaped schedule:
//
//   * grid = one CTA per SM (148), co-resident (cooperative launch): 16 compute warps + 1 TMA loader warp each;
//   * every Q4_0 weight byte of the token (4.13 GB at 7B) is streamed exactly once through a per-CTA ring of
//     shared-memory stages by cp.async.bulk (1-D TMA).  The loader walks the static schedule
//     layer0.{wq|wk|wv, wo, w1|w3, w2}, layer1..., output and runs AHEAD of the compute warps across phase boundaries
//     by up to the ring capacity (3 x 56 KB per SM, ~25 MB chip-wide), so HBM keeps streaming while the compute warps
//     wait for activations, run a LayerNorm prologue or the attention phase;
//   * activations travel as FLAGGED words: every f32 is stored as an 8-byte {value, sequence} pair (one
//     st.volatile.v2 -- single-copy atomic) and the consumer verifies the sequence number of (token, layer) on every
//     word it uses.  Data and "ready" signal arrive in ONE store, so all five grid-wide barriers per layer of a
//     conventional schedule disappear (this token's q / K / V reach the attention phase the same way; the f32 KV cache
//     row is written for future tokens only).  Arrival counters tell a consumer when a read will succeed, so one thread
//     polls one word; on a single GPU the two big LayerNorm-free vectors (att, h) travel as plain f32 with that
//     counter as a release/acquire flag instead (half the consumer-side bytes);
//   * the same stores go to every GPU of a tensor-parallel group through peer-mapped memory (NVLink 5 / NVSwitch):
//     the all-gather of finished activation slices IS the epilogue -- no NCCL call, no separate collective kernel.
//
// The arithmetic is the same operation-for-operation mirror of the reference's AVX2 build as kernels.cuh (the
// per-matrix kernels kept for A/B and for reference thread counts > 16): see the contract there.  -fmad=false.
#pragma once
#include "kernels.cuh"

// ---- build switches (development A/B; the defaults are the measured-best settings) ---------------------------------
#ifndef B200_EVICT_FIRST
#define B200_EVICT_FIRST 1  // weight stream marked evict_first in L2 (read once per token)
#endif
#ifndef B200_PLAIN_LOCAL
#define B200_PLAIN_LOCAL 1  // single GPU: att and h (the mat-vec inputs that need no LayerNorm) cross phases as plain f32 +
#endif                      // release/acquire arrival counters; measured 1.0 / 2.3 us per layer faster than flagged words
#ifndef B200_PLAIN_H_GROUP
#define B200_PLAIN_H_GROUP 0  // tensor-parallel group: h as plain f32 + system-scope release -- measured 3 % SLOWER than flagged words (682 vs 705 tok/s on 2 GPUs)
#endif
#ifndef B200_LL_STAGE
#define B200_LL_STAGE 0     // (measured: no gain) flagged n_embd-sized vectors fetched by one TMA bulk copy per CTA, verified from smem
#endif
#ifndef B200_PLAIN_X
#define B200_PLAIN_X 0      // single GPU: the same for the residual stream (inpL, inpFF); measured 0.5-1.0 us SLOWER
#endif
#ifndef B200_ROTATE
#define B200_ROTATE 0       // (measured: 3.5 % SLOWER) every CTA starts its sweep over an activation vector at a different block (spreads the
#endif                      // 148-fold re-read of the same lines over the L2 slices)
#ifndef B200_ATT_KB
#define B200_ATT_KB 8       // K.Q: positions per warp batch (12 and 16 measured slower: register pressure)
#endif
#ifndef B200_IDLE_PREFETCH
#define B200_IDLE_PREFETCH 0  // round-2 experiment (see the loader): L2 bulk prefetch of up to this many chunks while the ring is full
#endif
#ifndef B200_HINTS
#define B200_HINTS 1        // arrival-hint counters in front of the flagged-word reads (polite polling)
#endif
#ifndef B200_JITTER
#define B200_JITTER 0       // stress build: every CTA sleeps a pseudo-random 0-4 us before every phase (soak test of the barrier-free exchange)
#endif
#ifndef B200_PROF_ATT
#define B200_PROF_ATT 0     // development: 5 timeline stamps inside every attention phase (stored in the spare tail of the profile array)
#endif
#ifndef B200_NO_MATH
#define B200_NO_MATH 0      // development: 1 = consume the ring without doing the math (delivery-rate ceiling; wrong results)
#endif

namespace b200 {

struct MatDesc {
  const uint8_t *w;   // tile-major stream laid out for gridDim.x CTAs and chunk length cb
  int M;              // valid fused rows
  int g_total;        // padded rows / 4
  int nb;             // blocks per row
  int cb;             // blocks per chunk
  int lp;             // lane-pairs per thread (1, 2 or 4)
  int rpt;            // rows per thread (1, 2 or 4)
};

struct LayerDesc {
  MatDesc qkv, wo, w13, w2;
  const float *attn_norm, *ffn_norm;
  float *k_layer, *v_layer;
};

constexpr int MEGA_MAX_LAYERS = 80;   // LLaMA-65B; the descriptors ride in the (large, CUDA >= 12.1) kernel parameter block
constexpr int MEGA_MAX_TP = 8;        // GPUs of one NVSwitch box

// Tensor-parallel group.  Every matrix is split by ROWS over the ranks (wq/wk/wv by head), so
// each rank evaluates complete reference dot products -- the same operations in the same order as the unsharded
// reference, bit for bit -- and what crosses NVLink are finished activation slices (an all-gather by peer stores),
// never partial sums.  size == 1 is the single-GPU case and uses the very same code path with one "peer".
struct TpArgs {
  int rank, size;
  int e_loc, f_loc, v_loc;            // this rank's share of n_embd (rows of wq|wk|wv|wo|w2 = its heads), n_ff, n_vocab
  uint2 *ll[MEGA_MAX_TP];             // peer p's flagged activation area (p == rank: the local one)
  float *logits[MEGA_MAX_TP];         // peer p's logits vector [n_vocab]
  unsigned int *done[MEGA_MAX_TP];    // peer p's end-of-token flags [size]
  unsigned int *hint[MEGA_MAX_TP];    // peer p's arrival counters [4]: inpL, inpFF, att, h (monotonic, never reset)
  unsigned int p_e, p_f, p_att;       // producer CTAs per layer over the whole group: wo/w2 rows, w1|w3 rows, attention
};

struct TokenArgs {
  LayerDesc layers[MEGA_MAX_LAYERS];   // in the constant bank: descriptor fields cost no registers and no loads
  int n_layer;
  MatDesc out;
  const float *final_norm;
  const uint8_t *tok_emb;
  float *q;                 // roped Q of this token [n_embd] (rank-local: only this rank's heads are used)
  TpArgs tp;
  unsigned int *epoch;      // launch counter in HBM (>= 1): makes every flag value of every token unique
  long long spin_limit;     // clock64 ticks before a spin-wait gives up (sets the abort word, ptx.cuh) instead of hanging the box
  const double2 *rope;
  const uint16_t *silu_table, *exp_table;
  StepParams *sp;           // read at kernel start; advanced by the folded arg-max (below) at the very end
  // Greedy pick folded into the token kernel (single GPU, device-resident loop): every CTA reduces its own logits rows,
  // the last CTA to finish reduces the per-CTA candidates, logs the token and advances the step scalars -- what the
  // separate arg-max kernel did in one more launch per token (19.5 us under ncu, profiles/r1_f_launches_bench.md).
  int fold_argmax;
  float2 *am;               // [gridDim.x] {value, index bits} candidate of every CTA
  int *token_log;
  const int *forced_tokens;
  unsigned int *bar;        // [0] arg-max CTA counter, [1] end-of-token CTA counter (tensor-parallel groups); zeroed before every launch
  int n_embd, n_head, n_ctx, n_ff, n_threads;
  float kq_scale;
  int S, stage_bytes;
  int xs_floats;            // size of the f32 scratch area (attention scores): >= n_ctx
  int ll_stage;             // != 0: shared memory has room to stage one n_embd-sized flagged vector (TMA bulk copy)
  int pace_ps_per_byte;     // != 0: the loader spaces its bulk copies so that this SM streams at most 1 byte per that many ps (experiment)
  long long *prof;          // optional [gridDim.x][prof_marks] globaltimer stamps (development profiler), else null
  int prof_marks;
};

constexpr int MEGA_COMPUTE_WARPS = 16;
constexpr int MEGA_COMPUTE_THREADS = MEGA_COMPUTE_WARPS * 32;   // 512: one 8-float item per thread at n_embd 4096
constexpr int MEGA_THREADS = MEGA_COMPUTE_THREADS + 32;          // + the TMA loader warp (96 regs/thread)
constexpr int MEGA_MAX_ROWS = 512;     // rows per CTA upper bound (rowres[])
constexpr int MEGA_MAX_NTH = 16;       // reference thread counts supported by the V*P partition
constexpr int MEGA_NORM_ROUNDS = 2;    // 8-element items per thread held in registers by the LayerNorm prologue (K <= 7168)

__device__ __forceinline__ long long globaltimer_ns() {
  long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
#define PROF_MARK() do { if (a.prof && tid == 0 && pm < a.prof_marks) a.prof[(size_t) blockIdx.x * a.prof_marks + pm] = globaltimer_ns(); pm++; } while (0)

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int *p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// ---- flagged activation words ("LL" protocol: value and ready-flag in one 8-byte single-copy-atomic store) -----------
__device__ __forceinline__ uint4 ld_vol_v4(const void *p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint2 ld_vol_v2(const void *p) {
  uint2 r;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint32_t ld_vol_u32(const void *p) {
  uint32_t r;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void ll_store(uint2 *p, float v, uint32_t seq) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(seq) : "memory");
}
// one flagged word: spin until its sequence number is `seq`
__device__ __forceinline__ float ll_wait1(const uint2 *p, uint32_t seq, long long limit) {
  uint2 r = ld_vol_v2(p);
  if (r.y != seq) {
    const long long t0 = clock64();
    do {
      __nanosleep(20);
      r = ld_vol_v2(p);
      if (wait_give_up(t0, limit)) break;                      // never hang the box (ptx.cuh: bounded waits)
    } while (r.y != seq);
  }
  return __uint_as_float(r.x);
}

struct TokenArgs;
// store one activation value into the flagged area of every GPU of the group (the local one included)
__device__ __forceinline__ void ll_bcast(uint2 *const (&ll)[MEGA_MAX_TP], int n_peers, uint32_t off, float v, uint32_t seq) {
  for (int p = 0; p < n_peers; p++) ll_store(ll[p] + off, v, seq);
}

// Arrival hints.  Correctness never depends on them (every flagged word is verified by its reader); they only tell a
// consumer CTA WHEN a full read is likely to succeed, so that one thread polls one word with back-off instead of 512
// threads hammering L2 with re-reads while the producers are still working.  Relaxed atomics, no fences.
enum { HINT_INPL = 0, HINT_INPFF = 1, HINT_ATT = 2, HINT_H = 3 };
__device__ __forceinline__ void hint_arrive(unsigned int *const (&hint)[MEGA_MAX_TP], int n_peers, int which) {
  for (int p = 0; p < n_peers; p++)
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(hint[p] + which) : "memory");
}
// thread 0 waits until the local counter reached `expected` (wrap-safe), then the CTA's compute threads go on together
__device__ __forceinline__ void hint_wait(const unsigned int *cnt, unsigned int expected, long long limit, int tid) {
  if (tid == 0) {
    if ((int) (ld_vol_u32(cnt) - expected) < 0) {
      const long long t0 = clock64();
      while ((int) (ld_vol_u32(cnt) - expected) < 0) {
        __nanosleep(40);
        if (wait_give_up(t0, limit)) break;                      // never hang the box
      }
    }
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

// item index with the CTA's rotation applied (idx >= items: a padding thread, re-reads the last item)
__device__ __forceinline__ int rot_item(int idx, int items, int rot) {
  if (idx >= items) return items - 1;
  const int it = idx + rot;
  return it >= items ? it - items : it;
}

// ---- the single-GPU variant of the exchange: plain f32 words, the arrival counter is authoritative -------------------
// Flagged words cost 8 bytes per value on the consumer side, and every CTA reads every activation vector (an
// all-gather through L2: 148 x 92 KB per layer at 7B).  Inside one GPU a release/acquire counter is cheap, so there the
// values travel as plain f32 (half the L2 read traffic) and the counter that is only a hint for the flagged words
// becomes the synchronisation: producers' stores -> bar.sync -> red.release.gpu; consumer: ld.acquire.gpu -> bar.sync.
__device__ __forceinline__ void plain_arrive(unsigned int *cnt) { red_release_add_u32(cnt, 1u); }
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// sys: the producers are on several GPUs (their stores came over NVLink and were released with a system-scope fence)
__device__ __forceinline__ void plain_wait(const unsigned int *cnt, unsigned int expected, long long limit, int tid, bool sys = false) {
  if (tid == 0) {
    if ((int) ((sys ? ld_acquire_sys_u32(cnt) : ld_acquire_u32(cnt)) - expected) < 0) {
      const long long t0 = clock64();
      while ((int) ((sys ? ld_acquire_sys_u32(cnt) : ld_acquire_u32(cnt)) - expected) < 0) {
        __nanosleep(20);
        if (wait_give_up(t0, limit)) break;                      // never hang the box
      }
    }
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}
template <int N>
__device__ __forceinline__ void plain_read_rounds(const float *src, int items, int it0, int rot, float (&v)[N][8], int tid) {
  float4 va[N], vc[N];
#pragma unroll
  for (int rd = 0; rd < N; rd++) {        // every round's loads are in flight before the first use: one L2 latency
    const int it = rot_item(it0 + tid + rd * MEGA_COMPUTE_THREADS, items, rot);
    va[rd] = __ldcg(reinterpret_cast<const float4 *>(src) + it * 2);
    vc[rd] = __ldcg(reinterpret_cast<const float4 *>(src) + it * 2 + 1);
  }
#pragma unroll
  for (int rd = 0; rd < N; rd++) {
    v[rd][0] = va[rd].x; v[rd][1] = va[rd].y; v[rd][2] = va[rd].z; v[rd][3] = va[rd].w;
    v[rd][4] = vc[rd].x; v[rd][5] = vc[rd].y; v[rd][6] = vc[rd].z; v[rd][7] = vc[rd].w;
  }
}

struct MegaSmem {
  uint8_t *stages;
#if B200_IMMA
  ActSmem act;      // quantized activation vector in MMA-operand form (rowloop_imma.cuh)
#else
  uint2 *xq;        // [4 planes p][nb_max] {signed bytes of lane 2p, of lane 2p+1} of every block
  int nbx;          // plane stride = nb_max + 2 (bank-conflict padding)
  float *dxs;       // [nb_max]
#endif
  float *xs;        // [xs_floats] attention scores / probabilities
  float *rowres;    // [MEGA_MAX_ROWS]
  double *redd;     // [2][16]
  float *redf;      // [2][16]
  float *part;      // [MEGA_MAX_NTH][32]
  double2 *ropev;   // [head_dim/2] (cos, sin) of this token's position, fetched once at kernel start
  float *qkc;       // [288] this token's q (128), k (128) of the head and v (32) of the head quarter, for the attention phase
  uint64_t *full, *empty;
  uint2 *llbuf;     // [n_embd] staging area for a flagged activation vector (null when it does not fit)
  uint64_t *llbar;  // its mbarrier
};

// Block-wide sums over the compute warps with ONE barrier: warp shuffle tree, 14 partials, then every warp folds the
// partials with the same shuffle tree (fixed order => deterministic).  `buf` alternates between consecutive calls.
__device__ __forceinline__ double mega_sum_d(double v, double *redd, int buf, int tid) {
  v = warp_sum_d(v);
  double *r = redd + buf * 16;
  if ((tid & 31) == 0) r[tid >> 5] = v;
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  const int l = tid & 15;
  double s = l < MEGA_COMPUTE_WARPS ? r[l] : 0.0;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
  return s;
}
__device__ __forceinline__ float mega_max_f(float v, float *redf, int buf, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  float *r = redf + buf * 16;
  if ((tid & 31) == 0) r[tid >> 5] = v;
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  const int l = tid & 15;
  float s = l < MEGA_COMPUTE_WARPS ? r[l] : -CUDART_INF_F;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, o));
  return s;
}

// quantize_row_q4_0 (AVX2 branch, ggml.c:456-523) for one 32-block handled by 4 consecutive lanes (8 values each).
// Writes the dp4a-ready form: plane p of xq gets {bytes of lane 2p, bytes of lane 2p+1} of block b, dxs[b] = d.
// s = lane-quad index 0..3 of a live block, >= 4 for a padding thread (takes part in the shuffles, stores nothing).
__device__ __forceinline__ void quantize_block_4t(const float v[8], int b, int s, uint2 *xq, int nbx, float *dxs) {
  float amax = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
  const float d = __fdiv_rn(amax, 7.0f);
  const float id = (amax != 0.0f) ? __fdiv_rn(7.0f, amax) : 0.0f;
  int q[8];
#pragma unroll
  for (int i = 0; i < 8; i++) q[i] = __float2int_rn(__fmul_rn(v[i], id));      // round-to-nearest-even = stored nibble - 8
  // this thread holds elements 8s..8s+7 = one half (s>>1) of AVX lanes 4*(s&1)+j, j=0..3 (pairs q[2j], q[2j+1])
  const uint32_t hw01 = (uint32_t) (q[0] & 0xff) | ((uint32_t) (q[1] & 0xff) << 8) | ((uint32_t) (q[2] & 0xff) << 16) | ((uint32_t) (q[3] & 0xff) << 24);
  const uint32_t hw23 = (uint32_t) (q[4] & 0xff) | ((uint32_t) (q[5] & 0xff) << 8) | ((uint32_t) (q[6] & 0xff) << 16) | ((uint32_t) (q[7] & 0xff) << 24);
  const uint32_t o01 = __shfl_xor_sync(0xffffffffu, hw01, 2);
  const uint32_t o23 = __shfl_xor_sync(0xffffffffu, hw23, 2);
  if (s < 2) {
    // lanes 4s+0..3: low half-word = my elements (2l, 2l+1), high half-word = partner's (16+2l, 17+2l)
    const uint32_t x0 = (hw01 & 0xffffu) | (o01 << 16);
    const uint32_t x1 = (hw01 >> 16) | (o01 & 0xffff0000u);
    const uint32_t x2 = (hw23 & 0xffffu) | (o23 << 16);
    const uint32_t x3 = (hw23 >> 16) | (o23 & 0xffff0000u);
    xq[(2 * s + 0) * nbx + b] = make_uint2(x0, x1);
    xq[(2 * s + 1) * nbx + b] = make_uint2(x2, x3);
    if (s == 0) dxs[b] = d;
  }
}

// ---- activation prologues ------------------------------------------------------------------------------------------
// An "item" is 8 consecutive activation values = a quarter of a Q4 block = 4 x 16 bytes of flagged words.
// Read N rounds of items (item it0 + tid + rd*512 in round rd): all loads are issued before the first flag is
// checked (one L2 latency when the data is already there), then every 16-byte half-pair is re-polled until both of its
// sequence numbers match.  Threads past the end re-read the last item (harmless duplicates).
template <int N>
__device__ __forceinline__ void ll_read_rounds(const uint2 *src, int items, int it0, int rot, uint32_t seq, long long limit,
                                               float (&v)[N][8], int tid) {
  uint4 r[N][4];
#pragma unroll
  for (int rd = 0; rd < N; rd++) {
    const int it = rot_item(it0 + tid + rd * MEGA_COMPUTE_THREADS, items, rot);
    const uint4 *p = reinterpret_cast<const uint4 *>(src + (size_t) it * 8);
#pragma unroll
    for (int i = 0; i < 4; i++) r[rd][i] = ld_vol_v4(p + i);
  }
#pragma unroll
  for (int rd = 0; rd < N; rd++) {
    const int it = rot_item(it0 + tid + rd * MEGA_COMPUTE_THREADS, items, rot);
    const uint4 *p = reinterpret_cast<const uint4 *>(src + (size_t) it * 8);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (r[rd][i].y != seq || r[rd][i].w != seq) {
        const long long t0 = clock64();
        do {
          r[rd][i] = ld_vol_v4(p + i);
          if (wait_give_up(t0, limit)) break;                      // never hang the box
        } while (r[rd][i].y != seq || r[rd][i].w != seq);
      }
      v[rd][2 * i] = __uint_as_float(r[rd][i].x);
      v[rd][2 * i + 1] = __uint_as_float(r[rd][i].z);
    }
  }
}

// blocks nb .. 4*ceil(nb/4)-1 pad a partial last quad: scale 0 on both sides makes them exact no-ops in the row loop
__device__ __forceinline__ void zero_pad_blocks(int nb, const MegaSmem &sm, int tid) {
  const int nbp = (nb + 3) & ~3;
  if (tid < (nbp - nb) * 4) {
#if B200_IMMA
    act_zero_block(nb + (tid >> 2), sm.act, tid & 3);
#else
    sm.xq[(tid & 3) * sm.nbx + nb + (tid >> 2)] = make_uint2(0u, 0u);
    if ((tid & 3) == 0) sm.dxs[nb + (tid >> 2)] = 0.0f;
#endif
  }
}

// The same from a copy of the whole vector that ONE cp.async.bulk put into shared memory (sm.llbuf): 148 CTAs x 512
// threads x 4 separate 16-byte requests for the same 32 KB is what made the flagged read slow (L2 request rate); the TMA
// engine fetches whole lines.  Every word is still verified; a word whose store had not landed when the copy ran is
// re-polled from global memory.
template <int N>
__device__ __forceinline__ void ll_read_rounds_staged(const uint2 *stage, const uint2 *src, int items, uint32_t seq, long long limit,
                                                      float (&v)[N][8], int tid) {
#pragma unroll
  for (int rd = 0; rd < N; rd++) {
    const int it = min(tid + rd * MEGA_COMPUTE_THREADS, items - 1);
    const uint4 *ps = reinterpret_cast<const uint4 *>(stage + (size_t) it * 8);
    const uint4 *pg = reinterpret_cast<const uint4 *>(src + (size_t) it * 8);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint4 r = ps[i];
      if (r.y != seq || r.w != seq) {
        const long long t0 = clock64();
        do {
          r = ld_vol_v4(pg + i);
          if (wait_give_up(t0, limit)) break;                      // never hang the box
        } while (r.y != seq || r.w != seq);
      }
      v[rd][2 * i] = __uint_as_float(r.x);
      v[rd][2 * i + 1] = __uint_as_float(r.z);
    }
  }
}
// issue the copy (thread 0) and wait for it (everybody); `par` is the phase parity of sm.llbar, toggled per use
__device__ __forceinline__ void ll_stage_vector(const MegaSmem &sm, const uint2 *src, int n_words, uint32_t &par, int tid) {
  if (tid == 0) {
    mbar_arrive_expect_tx(sm.llbar, (uint32_t) n_words * 8u);
    tma_bulk_g2s(sm.llbuf, src, (uint32_t) n_words * 8u, sm.llbar);
  }
  mbar_wait(sm.llbar, par);
  par ^= 1u;
}

// PLAIN: quantize x[K] (flagged words, polled) -> xq/dxs.
template <int N>
__device__ __forceinline__ void prologue_plain_batch(const uint2 *src, bool plain, int items, int it0, int rot, uint32_t seq, long long limit,
                                                     const MegaSmem &sm, int tid) {
  float v[N][8];
  if (plain) plain_read_rounds<N>(reinterpret_cast<const float *>(src), items, it0, rot, v, tid);
  else ll_read_rounds<N>(src, items, it0, rot, seq, limit, v, tid);
#pragma unroll
  for (int rd = 0; rd < N; rd++) {
    const int it = it0 + tid + rd * MEGA_COMPUTE_THREADS;
    const bool live = it < items;         // a block's 4 quarter-items are all live or all padding (512 % 4 == 0)
    const int iq = rot_item(it, items, rot);
#if B200_IMMA
    quantize_block_4t_imma(v[rd], iq >> 2, live ? (iq & 3) : 4 + (iq & 3), sm.act);
#else
    quantize_block_4t(v[rd], iq >> 2, live ? (iq & 3) : 4 + (iq & 3), sm.xq, sm.nbx, sm.dxs);
#endif
  }
}

// src: flagged words (plain == false) or the same area used as a plain f32 vector (plain == true)
__device__ __forceinline__ void prologue_plain(const uint2 *src, bool plain, int nb, uint32_t seq, long long limit, const MegaSmem &sm, int tid) {
  const int items = nb * 4;
  const int rounds = (items + MEGA_COMPUTE_THREADS - 1) / MEGA_COMPUTE_THREADS;
  const int rot = B200_ROTATE ? 4 * (int) (((long long) blockIdx.x * nb) / gridDim.x) : 0;   // whole blocks
  for (int r0 = 0; r0 < rounds; r0 += 3) {     // at most 3 rounds (12 x 16 B per thread) in flight at a time
    const int n = rounds - r0;
    const int it0 = r0 * MEGA_COMPUTE_THREADS;
    if (n >= 3) prologue_plain_batch<3>(src, plain, items, it0, rot, seq, limit, sm, tid);
    else if (n == 2) prologue_plain_batch<2>(src, plain, items, it0, rot, seq, limit, sm, tid);
    else prologue_plain_batch<1>(src, plain, items, it0, rot, seq, limit, sm, tid);
  }
  zero_pad_blocks(nb, sm, tid);
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

// NORM: LayerNorm (ggml_compute_forward_norm_f32, ggml.c:5363-5381; double sums, here in tree order) times the norm
// weight (ggml_mul, PO.mm:573-575), then quantize.  The vector lives in registers as doubles: xd[rd][0..8) is item
// tid + rd*448 (caller-filled; padding items must be zero-filled and are excluded from the sums by `items`).
#ifndef B200_PROF_LN
#define B200_PROF_LN 0      // development: 3 extra timeline marks inside every LayerNorm prologue
#endif
#if B200_PROF_LN
#define LN_MARK() do { if (prof && tid == 0) prof[pm] = globaltimer_ns(); pm++; } while (0)
#else
#define LN_MARK() do { } while (0)
#endif
__device__ __forceinline__ void prologue_norm_regs(double (&xd)[MEGA_NORM_ROUNDS][8], const float *norm_w, int nb,
                                                   const MegaSmem &sm, int tid, long long *prof, int &pm) {
  const int K = nb * 32, items = nb * 4;
  // the norm weights do not depend on the upstream phase: fetch them now, under the latency of everything below
  float4 wa[MEGA_NORM_ROUNDS], wc[MEGA_NORM_ROUNDS];
#pragma unroll
  for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
    const int iq = min(tid + rd * MEGA_COMPUTE_THREADS, items - 1);
    wa[rd] = __ldg(reinterpret_cast<const float4 *>(norm_w) + iq * 2);
    wc[rd] = __ldg(reinterpret_cast<const float4 *>(norm_w) + iq * 2 + 1);
  }
  const bool pow2 = (K & (K - 1)) == 0;          // x / 2^k == x * 2^-k exactly: skip the IEEE division routine
  const double invK = 1.0 / (double) K;
  double s = 0.0;
#pragma unroll
  for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
    if (tid + rd * MEGA_COMPUTE_THREADS < items) {
#pragma unroll
      for (int i = 0; i < 8; i++) s = __dadd_rn(s, xd[rd][i]);
    }
  }
  s = mega_sum_d(s, sm.redd, 0, tid);
  LN_MARK();
  const double mean = pow2 ? __dmul_rn(s, invK) : s / (double) K;
  double s2 = 0.0;
#pragma unroll
  for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
    if (tid + rd * MEGA_COMPUTE_THREADS < items) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        xd[rd][i] = __dsub_rn(xd[rd][i], mean);                                  // v = x - mean      ggml.c:5374
        s2 = __dadd_rn(s2, __dmul_rn(xd[rd][i], xd[rd][i]));                     // sum2 += v*v       ggml.c:5376
      }
    }
  }
  s2 = mega_sum_d(s2, sm.redd, 1, tid);
  const double var = pow2 ? __dmul_rn(s2, invK) : s2 / (double) K;
  const float nscale = (float) (1.0 / sqrt(__dadd_rn(var, (double) 1e-5f)));               // ggml.c:5379
#if B200_PROF_LN
  if (nscale == 123.456f) sm.rowres[0] = nscale;   // keeps the mark below after the division (never true in practice)
#endif
  LN_MARK();
#pragma unroll
  for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
    if (rd * MEGA_COMPUTE_THREADS < items) {      // CTA-uniform: whole rounds only
      const int it = tid + rd * MEGA_COMPUTE_THREADS;
      const bool live = it < items;
      const int iq = live ? it : items - 1;
      const float w[8] = {wa[rd].x, wa[rd].y, wa[rd].z, wa[rd].w, wc[rd].x, wc[rd].y, wc[rd].z, wc[rd].w};
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = __fmul_rn(w[i], __fmul_rn((float) xd[rd][i], nscale));   // y = (float)v; y *= scale; w*y
#if B200_IMMA
      quantize_block_4t_imma(v, iq >> 2, live ? (iq & 3) : 4 + (iq & 3), sm.act);
#else
      quantize_block_4t(v, iq >> 2, live ? (iq & 3) : 4 + (iq & 3), sm.xq, sm.nbx, sm.dxs);
#endif
    }
  }
  zero_pad_blocks(nb, sm, tid);
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

// fill the register copy of a flagged activation vector (polled)
__device__ __forceinline__ void load_items(double (&xd)[MEGA_NORM_ROUNDS][8], const uint2 *src, bool plain, int nb, uint32_t seq,
                                           long long limit, const MegaSmem &sm, uint32_t &ll_par, int tid) {
  const int items = nb * 4;
  float v[MEGA_NORM_ROUNDS][8];
  if (!plain && sm.llbuf != nullptr && B200_LL_STAGE) {
    ll_stage_vector(sm, src, nb * 32, ll_par, tid);
    if (items <= MEGA_COMPUTE_THREADS) {
      float v1[1][8];
      ll_read_rounds_staged<1>(sm.llbuf, src, items, seq, limit, v1, tid);
#pragma unroll
      for (int i = 0; i < 8; i++) { v[0][i] = v1[0][i]; v[1][i] = 0.0f; }
    } else {
      ll_read_rounds_staged<MEGA_NORM_ROUNDS>(sm.llbuf, src, items, seq, limit, v, tid);
    }
  } else if (plain) {
    plain_read_rounds<MEGA_NORM_ROUNDS>(reinterpret_cast<const float *>(src), items, 0, 0, v, tid);
  } else if (items <= MEGA_COMPUTE_THREADS) {
    float v1[1][8];
    ll_read_rounds<1>(src, items, 0, 0, seq, limit, v1, tid);
#pragma unroll
    for (int i = 0; i < 8; i++) { v[0][i] = v1[0][i]; v[1][i] = 0.0f; }
  } else {
    ll_read_rounds<MEGA_NORM_ROUNDS>(src, items, 0, 0, seq, limit, v, tid);
  }
#pragma unroll
  for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
    const bool live = tid + rd * MEGA_COMPUTE_THREADS < items;
#pragma unroll
    for (int i = 0; i < 8; i++) xd[rd][i] = live ? (double) v[rd][i] : 0.0;
  }
}

// ---- GEMV main loop over this CTA's rows of one matrix; leaves the row results in sm.rowres -------------------------
// Position in the ring of stages: stage index and the parity of its mbarrier phase, advanced without div/mod.
struct RingPos {
  int s;
  uint32_t par;
  uint32_t g;      // chunks consumed / issued so far (global index of the next one)
  __device__ __forceinline__ void next(int S) { g++; if (++s == S) { s = 0; par ^= 1u; } }
};

#if B200_IMMA
// Tensor-path row loop: warp w owns tiles w*NT .. w*NT+NT-1 of RW rows (RW = 8: rows g of the MMA tile only).
template <int RW, int NT>
__device__ __forceinline__ void gemv_rows_imma(const MatDesc &md, const RowPart rp, const MegaSmem &sm, RingPos &ring,
                                               int S, int stage_bytes, int tid) {
  const int R = rp.R, cb = md.cb;
  const int nbq = (md.nb + 3) >> 2, cq = cb >> 2;
  const int nchunks = (nbq + cq - 1) / cq;
  const int lane = tid & 31, warp = tid >> 5, g = lane >> 2;
  const int ntiles = (R + RW - 1) / RW;
  const bool warp_active = warp * NT < ntiles;         // warps with no rows skip the math but still release stages
  int row[NT][RW / 8], lrow[NT];
  imma_tile_rows<RW, NT>(warp * NT, lane, R, row, lrow);
  uint32_t sel0, sel1;
  imma_selectors(lane, sel0, sel1);
  u64 acc[NT][RW / 8];
#pragma unroll
  for (int i = 0; i < NT; i++)
#pragma unroll
    for (int h = 0; h < RW / 8; h++) acc[i][h] = pack_f2(0.0f, 0.0f);

  for (int k = 0; k < nchunks; k++, ring.next(S)) {
    const int s = ring.s;
    mbar_wait(&sm.full[s], ring.par);
    if (warp_active && !B200_NO_MATH) {
      const int cqk = min(cq, nbq - k * cq);
      gemv_chunk_imma<RW, NT>(sm.stages + (size_t) s * stage_bytes, cqk, R, row, lrow, lane, sm.act, k * cb, sel0, sel1, acc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);
  }
#pragma unroll
  for (int i = 0; i < NT; i++)
#pragma unroll
    for (int h = 0; h < RW / 8; h++) {
      const u64 one[1] = {acc[i][h]};
      const float res = row_hsum<1>(one);
      const int r = (warp * NT + i) * RW + g + 8 * h;
      if ((lane & 3) == 0 && r < R) sm.rowres[r] = res;
    }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

// md.lp = rows per warp tile (8 / 16), md.rpt = tiles per warp (1 / 2) -- chosen by the host plan (engine.cu: make_plan)
__device__ __forceinline__ void gemv_dispatch(const MatDesc &md, const RowPart rp, const MegaSmem &sm, RingPos &ring,
                                              int S, int stage_bytes, int tid) {
  if (md.lp == 8) gemv_rows_imma<8, 1>(md, rp, sm, ring, S, stage_bytes, tid);
  else if (md.rpt == 1) gemv_rows_imma<16, 1>(md, rp, sm, ring, S, stage_bytes, tid);
  else gemv_rows_imma<16, 2>(md, rp, sm, ring, S, stage_bytes, tid);
}
#else
template <int LP, int RPT>
__device__ __forceinline__ void gemv_rows(const MatDesc &md, const RowPart rp, const MegaSmem &sm, RingPos &ring,
                                          int S, int stage_bytes, int tid) {
  constexpr int UPR = 4 / LP;
  const int R = rp.R, cb = md.cb;
  const int G = R / RPT;                                // thread (g, t) owns rows g, g + G, ... (R is a multiple of 4)
  const int nbq = (md.nb + 3) >> 2, cq = cb >> 2;
  const int nchunks = (nbq + cq - 1) / cq;
  const bool active = tid < G * UPR;
  const int g = active ? tid / UPR : G - 1;
  const int t = tid % UPR;
  u64 acc[RPT][LP];
#pragma unroll
  for (int i = 0; i < RPT; i++)
#pragma unroll
    for (int j = 0; j < LP; j++) acc[i][j] = pack_f2(0.0f, 0.0f);
  const bool warp_active = (tid & ~31) < G * UPR;     // warps with no rows skip the math but still release stages

  for (int k = 0; k < nchunks; k++, ring.next(S)) {
    const int s = ring.s;
    mbar_wait(&sm.full[s], ring.par);
    if (warp_active && !B200_NO_MATH) {
      const int cqk = min(cq, nbq - k * cq);
      gemv_chunk<LP, RPT>(sm.stages + (size_t) s * stage_bytes, cqk, R, g, t, sm.xq, sm.dxs, sm.nbx, k * cb, acc);
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sm.empty[s]);
  }
#pragma unroll
  for (int i = 0; i < RPT; i++) {
    const float res = row_hsum<LP>(acc[i]);
    if (active && t == 0) sm.rowres[g + i * G] = res;
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}


// ---- few-row matrices (wo, w2: 28 rows per CTA at 7B; every matrix of a tensor-parallel shard): EIGHT threads per row --------
// With 4 threads per row such a phase keeps only 4 of the 16 warps busy, one per scheduler, and each of them walks its row
// at the latency-bound single-warp rate (profiles/r1_f_phase_profile.md: wo 4.1 TB/s, w2 5.2 TB/s from a pre-filled ring).
// Here thread t8 of a row owns ONE accumulator lane (lane l = t8 = word l/2, low nibbles for even l, high for odd): the same
// LP = 1 stream, the same per-lane operations (dp4a of the lane's 4 elements, exact int -> float, fma with d_w * d_x), twice
// the warps, ~0.7x the instructions per thread.  The lanes are not paired, so the fma is a scalar FFMA.
__device__ __forceinline__ void gemv_rows_half(const MatDesc &md, const RowPart rp, const MegaSmem &sm, RingPos &ring,
                                               int S, int stage_bytes, int tid) {
  const int R = rp.R, cb = md.cb;
  const int nbq = (md.nb + 3) >> 2, cq = cb >> 2;
  const int nchunks = (nbq + cq - 1) / cq;
  const bool active = tid < R * 8;
  const int g = active ? tid >> 3 : R - 1;
  const int t8 = tid & 7, t = t8 >> 1, hi = t8 & 1;
  const uint32_t sh = hi ? 0u : 4u;                      // high nibbles are in place; low nibbles move up by 4
  const bool warp_active = (tid & ~31) < R * 8;
  float acc = 0.0f;
  const int qstride = R * 80;

  for (int k = 0; k < nchunks; k++, ring.next(S)) {
    const int s = ring.s;
    mbar_wait(&sm.full[s], ring.par);
    if (warp_active && !B200_NO_MATH) {
      const int cqk = min(cq, nbq - k * cq);
      const uint8_t *st = sm.stages + (size_t) s * stage_bytes;
      const uint8_t *pw = st + (g * 4 + t) * 16, *ps = st + R * 64 + g * 16;
      const uint8_t *px = reinterpret_cast<const uint8_t *>(sm.xq + (size_t) t * sm.nbx + k * cb) + hi * 4;
      const float *pd = sm.dxs + k * cb;
      uint4 w4 = *reinterpret_cast<const uint4 *>(pw);
      float4 sc4 = *reinterpret_cast<const float4 *>(ps);
      for (int q = 0; q < cqk; q++) {
        // loads of quad q+1 under the math of quad q (one quad past the end of the chunk is still inside this CTA's smem)
        const uint4 w4n = *reinterpret_cast<const uint4 *>(pw + (q + 1) * qstride);
        const float4 sc4n = *reinterpret_cast<const float4 *>(ps + (q + 1) * qstride);
        const float4 dx4 = *reinterpret_cast<const float4 *>(pd + q * 4);
        const uint2 *xp = reinterpret_cast<const uint2 *>(px + q * 32);          // {lane 2t, lane 2t+1} bytes of 4 blocks; px is offset by hi
        const int x0 = (int) reinterpret_cast<const uint32_t *>(xp)[0], x1 = (int) reinterpret_cast<const uint32_t *>(xp)[2];
        const int x2 = (int) reinterpret_cast<const uint32_t *>(xp)[4], x3 = (int) reinterpret_cast<const uint32_t *>(xp)[6];
        const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
        const int xx[4] = {x0, x1, x2, x3};
        const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, dx[4] = {dx4.x, dx4.y, dx4.z, dx4.w};
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const int a8 = (int) and_xor(ww[b] << sh, 0xF0F0F0F0u, 0x80808080u);   // signed bytes 16*(q-8) of this lane
          const int ib = dp4a_ss(a8, xx[b], 0x4B400000);                          // float bits of 12582912 + 16*isum
          const float f = fmaf(__int_as_float(ib), 0.0625f, -786432.0f);          // exact (float) isum
          const float sdx = __fmul_rn(sc[b], dx[b]);                              // _mm256_mul_ps(d0, d1), ggml.c:1431
          acc = fmaf(sdx, f, acc);                                                // _mm256_fmadd_ps, ggml.c:1457
        }
        w4 = w4n; sc4 = sc4n;
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sm.empty[s]);
  }
  // horizontal sum as ggml.c:1461-1466 with the 8 lanes in 8 consecutive threads: (acc[k] + acc[k+4]), (r0 + r2), (r1 + r3), sum
  float r = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
  if (active && t8 == 0) sm.rowres[g] = r;
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
}

__device__ __forceinline__ void gemv_dispatch(const MatDesc &md, const RowPart rp, const MegaSmem &sm, RingPos &ring,
                                              int S, int stage_bytes, int tid) {
  if (md.lp == 0) { gemv_rows_half(md, rp, sm, ring, S, stage_bytes, tid); return; }     // LP = 1 stream, 8 threads per row
  const int rpt = (rp.R % md.rpt == 0) ? md.rpt : 1;   // R is a multiple of 4, so 2 and 4 always divide it
  switch (md.lp * 8 + rpt) {
    case 1 * 8 + 1: gemv_rows<1, 1>(md, rp, sm, ring, S, stage_bytes, tid); break;
    case 1 * 8 + 2: gemv_rows<1, 2>(md, rp, sm, ring, S, stage_bytes, tid); break;
    case 1 * 8 + 4: gemv_rows<1, 4>(md, rp, sm, ring, S, stage_bytes, tid); break;
    case 2 * 8 + 1: gemv_rows<2, 1>(md, rp, sm, ring, S, stage_bytes, tid); break;
    case 2 * 8 + 2: gemv_rows<2, 2>(md, rp, sm, ring, S, stage_bytes, tid); break;
    default: gemv_rows<4, 1>(md, rp, sm, ring, S, stage_bytes, tid); break;
  }
}
#endif

// ---- attention phase for (head h, output quarter qr): K.Q for all positions, soft_max, V.P for 32 dims --------------
#if B200_PROF_ATT
#define ATT_MARK(k) do { if (a.prof && tid == 0) a.prof[(size_t) blockIdx.x * a.prof_marks + 18 * a.n_layer + 8 + 5 * (int) (seq - 1u - ld_vol_u32(a.epoch) * (uint32_t) (a.n_layer + 2)) + (k)] = globaltimer_ns(); } while (0)
#else
#define ATT_MARK(k) do { } while (0)
#endif
__device__ __forceinline__ void attention_phase(const TokenArgs &a, const LayerDesc &L, const MegaSmem &sm, int h, int qr,
                                                int pos, int p_part, uint32_t att_off, const uint2 *qkv_ll, uint32_t seq,
                                                long long limit, int tid) {
  constexpr int HD = 128, NW = MEGA_COMPUTE_WARPS;
  const int lane = tid & 31, warp = tid >> 5;
  const int E = a.n_embd;
  const int p_valid = pos + 1;      // diag_mask_inf: columns > n_past + i are -inf -> probability 0 (ggml.c:6946-6953)
  float *sc = sm.xs;
  ATT_MARK(0);
  // This token's roped Q, its K row and V row arrive as flagged words straight from the CTAs that computed them (no
  // grid barrier between the mat-vec and the attention); rows of earlier positions come from the f32 cache, whose
  // current row is written for FUTURE tokens only.
  // One warp polls (288 words), the other 15 sleep in the barrier: polling threads compete with the weight stream for L2.
  if (warp == 0) {
    const uint2 *src[9];
#pragma unroll
    for (int i = 0; i < 4; i++) { src[i] = qkv_ll + h * HD + lane + 32 * i; src[4 + i] = qkv_ll + E + h * HD + lane + 32 * i; }
    src[8] = qkv_ll + 2 * E + h * HD + qr * 32 + lane;
    uint2 w[9];
#pragma unroll
    for (int i = 0; i < 9; i++) w[i] = ld_vol_v2(src[i]);          // all nine in flight: one L2 round trip when the data is there
#pragma unroll
    for (int i = 0; i < 9; i++) {
      if (w[i].y != seq) {
        const long long t0 = clock64();
        do {
          __nanosleep(20);
          w[i] = ld_vol_v2(src[i]);
          if (wait_give_up(t0, limit)) break;                      // never hang the box
        } while (w[i].y != seq);
      }
      sm.qkc[i < 8 ? (i < 4 ? 0 : 128) + lane + 32 * (i & 3) : 256 + lane] = __uint_as_float(w[i].x);
    }
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  ATT_MARK(1);      // q / k / v of this token have arrived
  float qv[4], kcur[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { qv[i] = sm.qkc[lane + 32 * i]; kcur[i] = sm.qkc[128 + lane + 32 * i]; }
  const float vcur = sm.qkc[256 + lane];
  // K.Q: ggml_vec_dot_f32, AVX mapping (lane t = 8*vec + l owns elements t, t+32, t+64, t+96), ggml.c:1223-1258, 872-887
  constexpr int KB = B200_ATT_KB;     // positions per batch: 4*KB independent loads in flight per lane
  for (int j0 = warp * KB; j0 < p_valid; j0 += NW * KB) {
    float kk[KB][4];
#pragma unroll
    for (int u = 0; u < KB; u++) {
      const int j = min(j0 + u, p_valid - 1);
      const float *kp = L.k_layer + (size_t) j * E + h * HD + lane;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float kc = __ldcg(kp + 32 * i);
        kk[u][i] = j == pos ? kcur[i] : kc;
      }
    }
#pragma unroll
    for (int u = 0; u < KB; u++) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; i++) s = fmaf(kk[u][i], qv[i], s);                  // GGML_F32_VEC_FMA, ggml.c:1239
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 8));                       // sum[0]+sum[1], sum[2]+sum[3]  ggml.c:874-876
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 16));                      // (..)+(..)                     ggml.c:877-879
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));                       // lanes k and k+4               ggml.c:883-884
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));                       // hadd                          ggml.c:885
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));                       // hadd                          ggml.c:886
      s = __fmul_rn(s, a.kq_scale);                                               // ggml_scale, PO.mm:617-621
      if (lane == 0 && j0 + u < p_valid) sc[j0 + u] = s;
    }
  }
  // V rows of this warp's first chain do not depend on the scores: get the first batch on its way before the soft_max
  const int nth = a.n_threads;
  const int dc = (p_part + nth - 1) / nth;
  const float *vp = L.v_layer + h * HD + qr * 32 + lane;
  constexpr int VB = 16;
  float vpre[VB];
  {
    const int j0 = warp * dc;
    const int j1 = warp < nth ? min(min(j0 + dc, p_part), p_valid) : j0;
#pragma unroll
    for (int i = 0; i < VB; i++) vpre[i] = (j0 + i < j1) ? (j0 + i == pos ? vcur : __ldcg(vp + (size_t) (j0 + i) * E)) : 0.0f;
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  ATT_MARK(2);      // K.Q done
  // soft_max, ggml.c:7019-7041
  float mx = -CUDART_INF_F;
  for (int j = tid; j < p_valid; j += MEGA_COMPUTE_THREADS) mx = fmaxf(mx, sc[j]);
  mx = mega_max_f(mx, sm.redf, 0, tid);
  double sum = 0.0;   // fp16-valued terms: exact in double in any order
  for (int j = tid; j < p_valid; j += MEGA_COMPUTE_THREADS) {
    const uint16_t hx = __half_as_ushort(__float2half_rn(__fsub_rn(sc[j], mx)));
    const float e = __half2float(__ushort_as_half(__ldg(a.exp_table + hx)));
    sc[j] = e;                       // same thread re-reads its own entries below: no barrier needed in between
    sum += (double) e;
  }
  sum = mega_sum_d(sum, sm.redd, 0, tid);
  const float inv = (float) (1.0 / sum);
  for (int j = tid; j < p_valid; j += MEGA_COMPUTE_THREADS) sc[j] = __fmul_rn(sc[j], inv);   // ggml_vec_scale_f32, ggml.c:7041
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  ATT_MARK(3);      // soft_max done
  // V.P: reference thread t owns columns [t*dc, (t+1)*dc) (ggml.c:5628-5632), FINALIZE adds buffers in order (5570-5574)
  for (int t = warp; t < nth; t += NW) {
    const int j0 = t * dc;
    const int j1 = min(min(j0 + dc, p_part), p_valid);
    float acc = 0.0f;
    int j = j0;
    if (t == warp) {   // the pre-loaded batch (same order: positions j0, j0+1, ...)
#pragma unroll
      for (int i = 0; i < VB; i++)
        if (j0 + i < j1) acc = fmaf(vpre[i], sc[j0 + i], acc);                  // vec_mad_f32, ggml.c:1696
      j = min(j0 + VB, j1);
    }
    for (; j + 16 <= j1; j += 16) {
      float vv[16];
#pragma unroll
      for (int i = 0; i < 16; i++) { const float vc = __ldcg(vp + (size_t) (j + i) * E); vv[i] = j + i == pos ? vcur : vc; }
#pragma unroll
      for (int i = 0; i < 16; i++) acc = fmaf(vv[i], sc[j + i], acc);           // vec_mad_f32, ggml.c:1696
    }
    if (j + 8 <= j1) {
      float vv[8];
#pragma unroll
      for (int i = 0; i < 8; i++) { const float vc = __ldcg(vp + (size_t) (j + i) * E); vv[i] = j + i == pos ? vcur : vc; }
#pragma unroll
      for (int i = 0; i < 8; i++) acc = fmaf(vv[i], sc[j + i], acc);
      j += 8;
    }
    if (j + 4 <= j1) {
      float vv[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { const float vc = __ldcg(vp + (size_t) (j + i) * E); vv[i] = j + i == pos ? vcur : vc; }
#pragma unroll
      for (int i = 0; i < 4; i++) acc = fmaf(vv[i], sc[j + i], acc);
      j += 4;
    }
    if (j < j1) {     // up to 3 positions left: their loads go out together, the fmaf chain stays in order
      float vv[3];
#pragma unroll
      for (int i = 0; i < 3; i++) vv[i] = j + i < j1 ? (j + i == pos ? vcur : __ldcg(vp + (size_t) (j + i) * E)) : 0.0f;
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (j + i < j1) acc = fmaf(vv[i], sc[j + i], acc);
    }
    sm.part[t * 32 + lane] = acc;
  }
  named_bar_sync(1, MEGA_COMPUTE_THREADS);
  ATT_MARK(4);      // V.P chains done
  if (warp == 0) {
    float o = sm.part[lane];
    for (int t = 1; t < nth; t++) o = __fadd_rn(o, sm.part[t * 32 + lane]);
    if (B200_PLAIN_LOCAL && a.tp.size == 1) {
      __stcg(reinterpret_cast<float *>(a.tp.ll[0] + att_off) + h * HD + qr * 32 + lane, o);   // KQV_merged, PO.mm:641-646
      __syncwarp();
      if (lane == 0) plain_arrive(a.tp.hint[0] + HINT_ATT);
    } else {
      ll_bcast(a.tp.ll, a.tp.size, att_off + h * HD + qr * 32 + lane, o, seq);    // ... -> every GPU of the group
      __syncwarp();
      if (lane == 0) hint_arrive(a.tp.hint, a.tp.size, HINT_ATT);
    }
  }
}

extern __shared__ __align__(128) uint8_t smem_mega[];

__device__ __forceinline__ MegaSmem carve_smem(const TokenArgs &a) {
  const int S = a.S;
  const int nb_max = ((max(a.n_embd, a.n_ff) / 32) + 3) & ~3;   // whole quads
  MegaSmem sm;
  sm.stages = smem_mega;
#if B200_IMMA
  sm.act = act_carve(smem_mega + (size_t) S * a.stage_bytes, nb_max);
  sm.xs = sm.act.dxs + nb_max;
#else
  sm.xq = reinterpret_cast<uint2 *>(smem_mega + (size_t) S * a.stage_bytes);
  sm.nbx = nb_max + 2;      // plane stride padded by 16 bytes: the 4 planes a quarter-warp reads together fall into
                            // different banks (an unpadded power-of-two stride made every activation load a 4-way conflict)
  sm.dxs = reinterpret_cast<float *>(sm.xq + (size_t) sm.nbx * 4);
  sm.xs = sm.dxs + nb_max;
#endif
  sm.rowres = sm.xs + a.xs_floats;
  sm.redd = reinterpret_cast<double *>(sm.rowres + MEGA_MAX_ROWS);
  sm.redf = reinterpret_cast<float *>(sm.redd + 32);
  sm.part = sm.redf + 32;
  sm.ropev = reinterpret_cast<double2 *>(sm.part + MEGA_MAX_NTH * 32);
  sm.qkc = reinterpret_cast<float *>(sm.ropev + 64);
  sm.full = reinterpret_cast<uint64_t *>(sm.qkc + 288);
  sm.empty = sm.full + S;
  sm.llbar = sm.empty + S;
  sm.llbuf = a.ll_stage ? reinterpret_cast<uint2 *>(sm.llbar + 2) : nullptr;     // 16-byte aligned
  return sm;
}

#undef PROF_MARK
#define PROF_MARK() do { if (a.prof && tid == 0 && pm < a.prof_marks) a.prof[(size_t) blockIdx.x * a.prof_marks + pm] = globaltimer_ns(); pm++; } while (0)

enum PhaseKind { PH_QKV = 0, PH_ATTN = 1, PH_WO = 2, PH_W13 = 3, PH_W2 = 4, PH_OUT = 5 };

__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_token_kernel(const __grid_constant__ TokenArgs a) {
  const int tid = threadIdx.x;
  const int S = a.S, stage_bytes = a.stage_bytes;
  const MegaSmem sm = carve_smem(a);

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], MEGA_COMPUTE_WARPS); }
    mbar_init(sm.llbar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= MEGA_COMPUTE_THREADS) {
    // ===== TMA loader warp: the whole token's weight stream for this SM, in schedule order (lane 0).  A CONTINUOUS L2
    // look-ahead of the stream (prefetch.global.L2 or cp.async.bulk.prefetch.L2 a fixed distance down the schedule) was
    // measured three times in different forms and was slower every time.  B200_IDLE_PREFETCH (round-2 experiment, off,
    // not yet measured) prefetches only while the loader would otherwise block on a full ring, i.e. exactly when HBM
    // idles: one bulk prefetch per chunk, at most B200_IDLE_PREFETCH chunks beyond the ring. =====
    if (tid == MEGA_COMPUTE_THREADS) {
      RingPos g = {0, 0u, 0u};
#if B200_EVICT_FIRST
      const uint64_t l2pol = l2_policy_evict_first();
#endif
      const int n_mats = 4 * a.n_layer + 1;
      long long pace_next = 0;
#if B200_IDLE_PREFETCH
      // second cursor over the same schedule: (matrix pf_mi, chunk pf_k) is chunk number pf_g of this CTA's stream
      int pf_mi = 0, pf_k = 0;
      uint32_t pf_g = 0;
      auto prefetch_while_blocked = [&](uint32_t cur_g) {
        const uint32_t lo = cur_g + (uint32_t) S, hi = lo + (uint32_t) B200_IDLE_PREFETCH;   // chunks the ring cannot hold yet
        while (pf_mi < n_mats && pf_g < hi) {
          const MatDesc &pd = pf_mi < 4 * a.n_layer ? (&a.layers[pf_mi >> 2].qkv)[pf_mi & 3] : a.out;
          const RowPart pr = row_part(pd.g_total, gridDim.x, blockIdx.x);
          const int pnbq = (pd.nb + 3) >> 2, pcq = pd.cb >> 2;
          const int pn = pr.R == 0 ? 0 : (pnbq + pcq - 1) / pcq;
          if (pf_k >= pn) { pf_mi++; pf_k = 0; continue; }
          if (pf_g >= lo) {
            const int pc = min(pcq, pnbq - pf_k * pcq);
            const uint8_t *pp = pd.w + (size_t) pr.row0 * pnbq * 80 + (size_t) pf_k * pcq * pr.R * 80;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pp), "r"((uint32_t) (pc * pr.R * 80)) : "memory");
          }
          pf_k++; pf_g++;
        }
      };
#endif
      for (int mi = 0; mi < n_mats; mi++) {
        const MatDesc &md = mi < 4 * a.n_layer ? (&a.layers[mi >> 2].qkv)[mi & 3] : a.out;
        const RowPart rp = row_part(md.g_total, gridDim.x, blockIdx.x);
        if (rp.R == 0) continue;
        const int nbq = (md.nb + 3) >> 2, cq = md.cb >> 2;
        const int nchunks = (nbq + cq - 1) / cq;
        const uint8_t *wbase = md.w + (size_t) rp.row0 * nbq * 80;
        for (int k = 0; k < nchunks; k++, g.next(S)) {
          const int s = g.s;
          // g.par is the parity of the fill about to start; the slot is free once the consumers released the
          // previous fill (parity par ^ 1).  On a fresh barrier that wait returns at once (first lap).
#if B200_IDLE_PREFETCH
          if (!mbar_try_wait(&sm.empty[s], g.par ^ 1u)) prefetch_while_blocked(g.g);
#endif
          if (!mbar_wait(&sm.empty[s], g.par ^ 1u, a.spin_limit)) return;   // (the consumers may be waiting for another GPU); abandoned: stop streaming
          const int cqk = min(cq, nbq - k * cq);
          const uint32_t bytes = (uint32_t) cqk * rp.R * 80;
          if (a.pace_ps_per_byte) {
            // paced streaming: a steady rate instead of bursts at full HBM speed, so that the latency-critical exchange
            // traffic of the serial phases meets shorter queues in L2
            while (globaltimer_ns() < pace_next) __nanosleep(100);
            pace_next = max(pace_next, globaltimer_ns() - 2000) + ((long long) bytes * a.pace_ps_per_byte) / 1000;
          }
          mbar_arrive_expect_tx(&sm.full[s], bytes);
#if B200_EVICT_FIRST
          tma_bulk_g2s_hint(sm.stages + (size_t) s * stage_bytes, wbase + (size_t) k * cq * rp.R * 80, bytes, &sm.full[s], l2pol);
#else
          tma_bulk_g2s(sm.stages + (size_t) s * stage_bytes, wbase + (size_t) k * cq * rp.R * 80, bytes, &sm.full[s]);
#endif
        }
      }
    }
    return;
  }

  // ===== compute warps =====
  // ONE loop over the token's phases with a single instance of the row loops: the kind of phase only selects the
  // prologue (how the activation vector is obtained and quantized) and the epilogue (the graph nodes that consume the
  // mat-vec).  Descriptors come from the kernel parameter block (constant bank).
  const int E = a.n_embd, HD = E / a.n_head, F = a.n_ff;
  const int T = a.tp.size, rank = a.tp.rank;
  const int e_loc = a.tp.e_loc, f_loc = a.tp.f_loc, v_loc = a.tp.v_loc;
  const int c0 = rank * e_loc;                        // first n_embd index (q/k/v column, wo/w2 row) owned by this rank
  const int nh_loc = a.n_head / T;
  const long long limit = a.spin_limit;
  const int pos = a.sp->pos;
  // flag values: the input of layer l of THIS launch carries seq0 + l (att / inpFF / h of layer l too, in their own
  // buffers); epoch >= 1 and never repeats, so a stale word of an earlier token can never match
  const uint32_t epoch = ld_vol_u32(a.epoch);
  const uint32_t seq0 = epoch * (uint32_t) (a.n_layer + 2) + 1u;
  // flagged area (uint2 units): inpL[2][E] | inpFF[2][E] | att[2][E] | h[2][F] | qkv[2][3E]; the buffer alternates with the layer
  // parity so a fast producer of layer l+1 can never overwrite words a slow consumer of layer l still polls
  const uint32_t o_inpL = 0u, o_inpFF = 2u * E, o_att = 4u * E, o_h = 6u * E, o_qkv = 6u * E + 2u * F;   // qkv[2][3E]: this GPU only
  uint2 *const ll_me = a.tp.ll[rank];
  // arrival-hint counters never reset: before this launch every buffer kind saw (epoch-1)*n_layer rounds of arrivals
  const unsigned int *const hint_me = a.tp.hint[rank];
  const unsigned int rounds0 = (epoch - 1u) * (unsigned int) a.n_layer;
  // one GPU: plain f32 words in the first half of each region + authoritative counters (see plain_wait)
  const bool plain = B200_PLAIN_LOCAL && T == 1;     // att
  // h is the one vector where halving the consumer-side bytes (44 KB instead of 88 KB per CTA at 7B) also pays inside a
  // group: plain f32 stored to every GPU, system-scope fence + counter as the release, ld.acquire.sys on the consumer
  const bool plain_h = B200_PLAIN_LOCAL && (T == 1 || B200_PLAIN_H_GROUP);
  const bool plain_x = B200_PLAIN_X && T == 1;       // inpL, inpFF
  unsigned int *const cnt_me = a.tp.hint[rank];
  if (tid < HD / 2) sm.ropev[tid] = a.rope[(size_t) pos * (HD / 2) + tid];   // visible after the first prologue's barriers
  RingPos gchunk = {0, 0u, 0u};
  int pm = 0;
  uint32_t ll_par = 0u;     // phase parity of the staging mbarrier
  PROF_MARK();   // 0: kernel start
  const int n_steps = 5 * a.n_layer + 1;
  for (int step = 0; step < n_steps; step++) {
    const int il = step / 5;
    const int kind = il < a.n_layer ? step - 5 * il : PH_OUT;
    const LayerDesc &L = a.layers[il < a.n_layer ? il : 0];
#if B200_JITTER
    {   // de-synchronise the CTAs (and the GPUs of a group): the exchange protocol must not depend on arrival order
      if (tid == 0) {
        uint32_t hsh = (blockIdx.x * 2654435761u) ^ ((uint32_t) step * 40503u) ^ (epoch * 2246822519u) ^ ((uint32_t) rank * 3266489917u);
        hsh ^= hsh >> 15; hsh *= 2654435761u; hsh ^= hsh >> 13;
        __nanosleep(hsh & 4095u);
      }
      named_bar_sync(1, MEGA_COMPUTE_THREADS);
    }
#endif
    const uint32_t seq = seq0 + (uint32_t) il;
    const uint32_t par = (uint32_t) il & 1u;

    for (int unit = blockIdx.x; kind == PH_QKV && pos > 0 && unit < 4 * nh_loc; unit += gridDim.x) {
      // The K/V rows of earlier positions that this CTA's attention phase will read were written many tokens ago and have
      // left L2 (the cache is 1 MB per position at 7B): start pulling them from HBM now, a whole mat-vec phase early.
      // K rows of the head are shared by its 4 CTAs (each takes every 4th position, 4 lines per row); V: this CTA's
      // 32-dim quarter of every row (1 line).
      const int hh = rank * nh_loc + (unit >> 2), qq = unit & 3;
      const float *kb = L.k_layer + hh * HD, *vb = L.v_layer + hh * HD + qq * 32;
      const int nk = ((pos + 3 - qq) >> 2) * 4;          // (positions j = qq, qq+4, ... < pos) x 4 lines
      for (int i = tid; i < nk + pos; i += MEGA_COMPUTE_THREADS) {
        const float *pl = i < nk ? kb + (size_t) (qq + 4 * (i >> 2)) * E + (i & 3) * 32 : vb + (size_t) (i - nk) * E;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pl));
      }
    }
    if (kind == PH_QKV && tid < 2) {
      // next layer's LayerNorm weights (read once per token, so never in L2): start the HBM fetch a whole layer early
      const float *nwp = il + 1 < a.n_layer ? (tid == 0 ? a.layers[il + 1].attn_norm : a.layers[il + 1].ffn_norm)
                                            : (tid == 0 ? a.final_norm : nullptr);
      if (nwp)
        for (int line = blockIdx.x; line * 32 < E; line += gridDim.x) asm volatile("prefetch.global.L2 [%0];" ::"l"(nwp + line * 32));
    }
    if (kind == PH_ATTN) {
      // ---- attention (PO.mm:614-646): this rank's heads, one (head, 32-dim quarter) unit per CTA -- a second round for
      // models with more than gridDim/4 heads per GPU (13B: 160 units on 148 CTAs); the result goes to every GPU ----
      for (int unit = blockIdx.x; unit < 4 * nh_loc; unit += gridDim.x)
        attention_phase(a, L, sm, rank * nh_loc + (unit >> 2), unit & 3, pos, a.sp->p_part, o_att + par * E,
                        ll_me + o_qkv + par * 3u * E, seq, limit, tid);
      PROF_MARK();
      continue;      // no barrier: the consumers poll the flagged words
    }

    const MatDesc &md = kind == PH_QKV ? L.qkv : kind == PH_WO ? L.wo : kind == PH_W13 ? L.w13 : kind == PH_W2 ? L.w2 : a.out;
    const RowPart rp = row_part(md.g_total, gridDim.x, blockIdx.x);
    if (rp.R == 0) {
      // no rows of this matrix on this CTA (tensor-parallel shards can have fewer row granules than SMs)
      PROF_MARK(); PROF_MARK(); PROF_MARK(); PROF_MARK();
      if (kind == PH_QKV) PROF_MARK();
#if B200_PROF_LN
      if (kind == PH_QKV || kind == PH_W13 || kind == PH_OUT) { PROF_MARK(); PROF_MARK(); PROF_MARK(); }
#endif
      continue;
    }

    // ---- prologue ----
    if (kind == PH_WO) {
      if (plain) plain_wait(hint_me + HINT_ATT, (rounds0 + il + 1u) * a.tp.p_att, limit, tid);
      else if (B200_HINTS) hint_wait(hint_me + HINT_ATT, (rounds0 + il + 1u) * a.tp.p_att, limit, tid);
      PROF_MARK();
      prologue_plain(ll_me + o_att + par * E, plain, E / 32, seq, limit, sm, tid);     // PO.mm:649-651
    } else if (kind == PH_W2) {
      if (plain_h) plain_wait(hint_me + HINT_H, (rounds0 + il + 1u) * a.tp.p_f, limit, tid, T > 1);
      else if (B200_HINTS) hint_wait(hint_me + HINT_H, (rounds0 + il + 1u) * a.tp.p_f, limit, tid);
      PROF_MARK();
      prologue_plain(ll_me + o_h + par * F, plain_h, F / 32, seq, limit, sm, tid);     // PO.mm:682-684
    } else {
      double xd[MEGA_NORM_ROUNDS][8];
      if (step == 0) {
        PROF_MARK();
        // get_rows: dequantize_row_q4_0 of the token's embedding row (ggml.c:6760-6785, 651-684) straight into registers
        const uint8_t *row = a.tok_emb + (size_t) a.sp->token * (E / 32) * 20;
        const int items = E / 8;
#pragma unroll
        for (int rd = 0; rd < MEGA_NORM_ROUNDS; rd++) {
          const int it = tid + rd * MEGA_COMPUTE_THREADS;
          if (it < items) {
            const uint8_t *blk = row + (it >> 2) * 20;
            const float d = __ldg(reinterpret_cast<const float *>(blk));
            const uint32_t by = __ldg(reinterpret_cast<const uint32_t *>(blk + 4) + (it & 3));   // 4 bytes = 8 nibbles
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const int qn = (by >> (4 * i)) & 0xf;             // element 2j = low nibble of byte j, 2j+1 = high nibble
              const float v = __fmul_rn((float) (qn - 8), d);
              xd[rd][i] = v;
              if (blockIdx.x == 0) {                            // residual source of layer 0 (this GPU only)
                if (plain_x) __stcg(reinterpret_cast<float *>(ll_me + o_inpL) + it * 8 + i, v);   // ordered by the qkv -> attention barrier
                else ll_store(ll_me + o_inpL + it * 8 + i, v, seq0);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; i++) xd[rd][i] = 0.0;
          }
        }
      } else {
        // inpFF(l) is round rounds0+l+1 of its counter; inpL(l), l >= 1, is round rounds0+l of its own (layer 0's input
        // is the embedding row: no arrival)
        const unsigned int *cnt = hint_me + (kind == PH_W13 ? HINT_INPFF : HINT_INPL);
        const unsigned int expected = (rounds0 + il + (kind == PH_W13 ? 1u : 0u)) * a.tp.p_e;
        if (plain_x) plain_wait(cnt, expected, limit, tid);
        else if (B200_HINTS) hint_wait(cnt, expected, limit, tid);
        PROF_MARK();
        load_items(xd, ll_me + (kind == PH_W13 ? o_inpFF : o_inpL) + par * E, plain_x, E / 32, seq, limit, sm, ll_par, tid);
      }
      const float *nw = kind == PH_QKV ? L.attn_norm : kind == PH_W13 ? L.ffn_norm : a.final_norm;
      prologue_norm_regs(xd, nw, E / 32, sm, tid, a.prof ? a.prof + (size_t) blockIdx.x * a.prof_marks : nullptr, pm);   // PO.mm:570-575, 660-665, 694-701
#if B200_PROF_LN
      PROF_MARK();
#endif
    }
    PROF_MARK();

    // the residual this thread's row will need (PO.mm:654, 687) has been there since an earlier phase: fetch it now,
    // under the row loop, instead of paying an L2 round trip in the epilogue (R <= 512: one row per thread)
    float resid = 0.0f;
    if ((kind == PH_WO || kind == PH_W2) && tid < rp.R && rp.row0 + tid < md.M) {
      const uint2 *rsrc = ll_me + (kind == PH_WO ? o_inpL : o_inpFF) + par * E;
      resid = plain_x ? __ldcg(reinterpret_cast<const float *>(rsrc) + c0 + rp.row0 + tid) : ll_wait1(rsrc + c0 + rp.row0 + tid, seq, limit);
    }

    // ---- the mat-vec ----
    gemv_dispatch(md, rp, sm, gchunk, S, stage_bytes, tid);
    PROF_MARK();

    // ---- epilogue ----
    if (kind == PH_QKV) {
      // fused local rows [0,e_loc) = wq, [e_loc,2e_loc) = wk, [2e_loc,3e_loc) = wv of this rank's heads.
      // rope (ggml.c:7110-7127, double math, host-built angles) + KV store (PO.mm:585-611); all rank-local
      for (int i = tid; i < rp.R / 2; i += MEGA_COMPUTE_THREADS) {
        const int g = rp.row0 + 2 * i;
        if (g >= md.M) continue;
        const int which = g / e_loc, col = c0 + (g - which * e_loc);
        float y0 = sm.rowres[2 * i], y1 = sm.rowres[2 * i + 1];
        if (which < 2) {
          const double2 cs = sm.ropev[(col % HD) / 2];
          const double x0 = y0, x1 = y1;
          y0 = (float) __dsub_rn(__dmul_rn(x0, cs.x), __dmul_rn(x1, cs.y));
          y1 = (float) __dadd_rn(__dmul_rn(x0, cs.y), __dmul_rn(x1, cs.x));
        }
        if (which > 0) {                            // the cache row of this position: read by FUTURE tokens (later launches)
          float *dst = which == 1 ? L.k_layer + (size_t) pos * E + col : L.v_layer + (size_t) pos * E + col;
          dst[0] = y0;
          dst[1] = y1;
        }
        uint2 *now = ll_me + o_qkv + par * 3u * E + (uint32_t) which * E + col;   // this token's attention reads these
        ll_store(now, y0, seq);
        ll_store(now + 1, y1, seq);
      }
      PROF_MARK();
    } else if (kind == PH_W13) {
      // fused rows 2i = w1 row i, 2i+1 = w3 row i: silu(w1 x) * (w3 x), PO.mm:678-680; silu via the fp16 table (ggml.c:1955-1963)
      for (int i = tid; i < rp.R / 2; i += MEGA_COMPUTE_THREADS) {
        const int g = rp.row0 / 2 + i;
        if (2 * g < md.M) {
          const uint16_t hx = __half_as_ushort(__float2half_rn(sm.rowres[2 * i]));
          const float sv = __half2float(__ushort_as_half(__ldg(a.silu_table + hx)));
          const float hv = __fmul_rn(sv, sm.rowres[2 * i + 1]);
          if (plain_h) {
            for (int p = 0; p < T; p++) __stcg(reinterpret_cast<float *>(a.tp.ll[p] + o_h + par * F) + rank * f_loc + g, hv);
          } else {
            ll_bcast(a.tp.ll, T, o_h + par * F + rank * f_loc + g, hv, seq);
          }
        }
      }
      if (plain_h) {
        named_bar_sync(1, MEGA_COMPUTE_THREADS);
        if (tid == 0) {
          if (T == 1) plain_arrive(cnt_me + HINT_H);
          else { __threadfence_system(); hint_arrive(a.tp.hint, T, HINT_H); }   // fence + relaxed add = release at system scope
        }
      } else if (B200_HINTS) {
        named_bar_sync(1, MEGA_COMPUTE_THREADS);
        if (tid == 0) hint_arrive(a.tp.hint, T, HINT_H);
      }
    } else if (kind == PH_OUT) {
      // the plain logits (PO.mm:705); every GPU of the group receives the full vector
      for (int i = tid; i < rp.R; i += MEGA_COMPUTE_THREADS) {
        const int g = rp.row0 + i;
        if (g < md.M) {
          const float v = sm.rowres[i];
          for (int p = 0; p < T; p++) a.tp.logits[p][rank * v_loc + g] = v;
        }
      }
    } else {
      // ggml_add with the residual stream (PO.mm:654, 687).  wo: inpFF(l) = wo.att + inpL(l); w2: inpL(l+1) = w2.h + inpFF(l).
      const uint32_t o_dst = kind == PH_WO ? o_inpFF + par * E : o_inpL + (par ^ 1u) * E;
      const uint32_t seq_dst = kind == PH_WO ? seq : seq + 1u;
      if (tid < rp.R && rp.row0 + tid < md.M) {
        const int g = rp.row0 + tid;
        const float v = __fadd_rn(sm.rowres[tid], resid);
        if (plain_x) __stcg(reinterpret_cast<float *>(ll_me + o_dst) + c0 + g, v);
        else ll_bcast(a.tp.ll, T, o_dst + c0 + g, v, seq_dst);
      }
      if (plain_x) {
        named_bar_sync(1, MEGA_COMPUTE_THREADS);
        if (tid == 0) plain_arrive(cnt_me + (kind == PH_WO ? HINT_INPFF : HINT_INPL));
      } else if (B200_HINTS) {
        named_bar_sync(1, MEGA_COMPUTE_THREADS);
        if (tid == 0) hint_arrive(a.tp.hint, T, kind == PH_WO ? HINT_INPFF : HINT_INPL);
      }
    }
    PROF_MARK();
  }

  // ---- end of token ----
  if (T > 1) {
    // The next kernel on this stream (arg-max, the D2H copy of the logits) must see every peer's logits stores: each CTA
    // fences its own peer stores system-wide, the last CTA of this GPU then tells every peer "rank r finished token
    // `epoch`" and waits until all peers said the same.  The kernel cannot complete before that CTA does.
    named_bar_sync(1, MEGA_COMPUTE_THREADS);
    if (tid == 0) {
      __threadfence_system();
      const unsigned int old = atomicAdd(a.bar + 1, 1u);
      if (old == gridDim.x - 1) {
        __threadfence_system();
        for (int p = 0; p < T; p++)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.tp.done[p] + rank), "r"(epoch) : "memory");
        for (int q = 0; q < T; q++) {
          const long long t0 = clock64();
          for (;;) {
            unsigned int v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.tp.done[rank] + q) : "memory");
            if ((int) (v - epoch) >= 0) break;
            if (wait_give_up(t0, limit)) break;
          }
        }
        *a.epoch = epoch + 1u;
      }
    }
  } else if (blockIdx.x == 0 && tid == 0) {
    *a.epoch = epoch + 1u;       // every CTA read the old value before CTA 0 can get here (it needed their rows)
  }
  if (a.fold_argmax && T == 1) {
    // ---- greedy pick (first maximum, like numpy.argmax): this CTA's rows are still in sm.rowres ----
    const RowPart rp = row_part(a.out.g_total, gridDim.x, blockIdx.x);
    float best = -CUDART_INF_F;
    int idx = 0x7fffffff;
    for (int i = tid; i < rp.R; i += MEGA_COMPUTE_THREADS) {
      const int g = rp.row0 + i;
      if (g < a.out.M) {
        const float v = sm.rowres[i];
        if (v > best || (v == best && g < idx)) { best = v; idx = g; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    float *cv = sm.redf;                                  // [16] values | [16] indices (as int bits)
    int *ci = reinterpret_cast<int *>(sm.redf + 16);
    named_bar_sync(1, MEGA_COMPUTE_THREADS);             // everybody is done with redf / rowres of the phases
    if ((tid & 31) == 0) { cv[tid >> 5] = best; ci[tid >> 5] = idx; }
    named_bar_sync(1, MEGA_COMPUTE_THREADS);
    if (tid < 32) {
      best = tid < MEGA_COMPUTE_WARPS ? cv[tid] : -CUDART_INF_F;
      idx = tid < MEGA_COMPUTE_WARPS ? ci[tid] : 0x7fffffff;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
      }
      unsigned int old = 0;
      if (tid == 0) {
        a.am[blockIdx.x] = make_float2(best, __int_as_float(idx));
        __threadfence();
        old = atomicAdd(a.bar, 1u);
      }
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old == gridDim.x - 1) {                         // the last CTA: every candidate is visible
        __threadfence();
        best = -CUDART_INF_F; idx = 0x7fffffff;
        for (int c = tid; c < (int) gridDim.x; c += 32) {
          const float2 cand = __ldcg(a.am + c);
          const int oi = __float_as_int(cand.y);
          if (cand.x > best || (cand.x == best && oi < idx)) { best = cand.x; idx = oi; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
          if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
        }
        if (tid == 0) {
          if (idx < 0 || idx >= a.out.M) idx = 0;         // all-NaN logits: never index the embedding table out of bounds
          StepParams *sp = a.sp;
          const int step = sp->step;
          a.token_log[step] = idx;
          sp->token = sp->forced ? a.forced_tokens[step] : idx;
          sp->pos += 1;
          sp->p_part += 1;
          sp->step = step + 1;
        }
      }
    }
  }
}

}  // namespace b200
