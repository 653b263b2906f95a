// Host launchers of the non-GEMM kernels of the MFP engine.  Every launcher returns MFP_OK or an error code.
#pragma once
#include "common.cuh"

namespace mfp {

struct TaskSet {
  int n;
  int ids[16];
};

struct ModifiedPtrs {
  void* cols[kMaxFields];          // int32 [T,C] / float [T,C]
  unsigned char* masks[kMaxFields];  // [T] (train mode only)
};

// masking.cu
// doc0: index of the batch's first document in the global batch (data-parallel shards; 0 for a single process): every Philox counter is
// formed from global document / element indices, so a sharded step draws exactly what the single-process step draws
int launch_sample_tasks(const TaskSet& allowed, int B, uint32_t seed, uint32_t step, int* tasks, cudaStream_t st, uint32_t doc0 = 0);
int launch_mask_corrupt(const Schema& sc, const BatchPtrs& in, const int* tasks, const MaskPtrs* test_masks, int B, int S, uint32_t seed,
                        uint32_t step, const ModifiedPtrs& out, cudaStream_t st, unsigned char* flags /*[n_num][T], optional*/ = nullptr, uint32_t doc0 = 0);
// reorders every sequence column: random permutation of the valid elements (shuffled_set) or lexicographic sort (sorted_set)
int launch_shuffle_inputs(const Schema& sc, const BatchPtrs& in, int B, int S, uint32_t seed, uint32_t step, int* perm /*[B][S]*/, const ModifiedPtrs& out,
                          cudaStream_t st, bool sorted = false, uint32_t doc0 = 0);
int launch_row_flags(const Schema& sc, const BatchPtrs& mod, int T, unsigned char* flags /*[n_num][T]*/, cudaStream_t st);

// encoder.cu
// pos: optional PositionEmbedding table [>= S][D] added per position (input_dtype != "set"), under dropout when pos_rate > 0
struct PosEmbed {
  const float* table;
  int S;
  float rate;  // 0 = inference
  uint32_t seed, step;
  int shift;   // 1 when a context token takes position 0 (encoder.py:247-252): the element in row s then has position s + 1
  uint32_t row0;  // global index of the batch's first row (doc0 * S): offsets the dropout counters
};
int launch_embed_fwd(const Schema& sc, const BatchPtrs& mod, const unsigned char* flags, const float* params, int T, float* h0, cudaStream_t st,
                     const PosEmbed& pos = PosEmbed{nullptr, 0, 0.f, 0u, 0u, 0, 0u});
int launch_pos_embed_bwd(const float* dh0, int B, int S, float rate, uint32_t seed, uint32_t step, float* dtable /*[S][D], overwritten*/, cudaStream_t st,
                         uint32_t row0 = 0);
// ... with a context token (PosEmbed::shift = 1): table row 0 collects the token rows (ctx_row[b]), row p >= 1 the elements in row p - 1
int launch_pos_embed_bwd_ctx(const float* dh0, const int* ctx_row, int B, int S, float rate, uint32_t seed, uint32_t step,
                             float* dtable /*[S + 1][D], overwritten*/, cudaStream_t st, uint32_t row0 = 0);
int launch_embed_onehot(const Schema& sc, const BatchPtrs& mod, const unsigned char* flags, int T, float* onehot /*[T][Rp]*/, cudaStream_t st,
                        const int* ctx_row = nullptr /*[B]: row of each document that holds the context token (all-zero one-hot row)*/, int S = 0);
// --context id / length (encoder.py:96-110,231-249): the special token of document b sits in row ctx_row[b] = length[b] + 1 of its S rows
// (self-attention without positions is order-free, so "after the last element" equals the reference's "prepended"); ctx_row doubles as
// the attention kernels' length array (zero-based: covers the elements and the token).
int launch_context_token(const float* table, int rows, const int* ids, const int* length, int B, int S, float* h0, int* ctx_row, cudaStream_t st,
                         const PosEmbed& pos = PosEmbed{nullptr, 0, 0.f, 0u, 0u, 0, 0u} /*given: the token also gets position 0 of the table, under the PositionEmbedding's dropout*/);
// --context canvas / canvas_add (encoder.py:177-199,228-230): vec[b, :] = sum over the canvas columns c of table_c[ids_c[b]]
struct CanvasArgs {
  int n;
  const int* ids[8];    // device [B]
  long long off[8];     // table offsets in the flat parameter buffer
  int rows[8];          // input_dim + 2
};
int launch_canvas_vector(const CanvasArgs& a, const float* params, int B, float* vec /*[B][D]*/, cudaStream_t st);
int launch_add_doc_vector(float* x /*[B*S][D], += vec[b]*/, const float* vec, int B, int S, cudaStream_t st);
int launch_sum_doc_rows(const float* dx /*[B*S][D]*/, int B, int S, float* dvec /*[B][D], overwritten*/, cudaStream_t st);
int launch_iota(int* iota /*[n] = 0..n-1*/, int* zeros /*[n] = 0*/, int n, cudaStream_t st);
int launch_context_token_bwd(const float* dh0, const int* ids, const int* ctx_row, int rows, int B, int S, float* dtable /*[rows][D], overwritten*/,
                             cudaStream_t st);
int launch_embed_scatter(const Schema& sc, const float* scratch /*[R][D]*/, float* grads, cudaStream_t st);

// transformer.cu
int launch_layernorm_fwd(const float* x, const float* gamma, const float* beta, int T, float* y, float* mean, float* rstd, cudaStream_t st);
// rowflags / n_masked / dx_masked (optional): also write n_masked copies of dx with the rows whose flag != 0 zeroed (encoder Dense wgrads)
int launch_layernorm_bwd(const float* x, const float* dy, const float* gamma, const float* mean, const float* rstd, const float* dres, int T,
                         float* dx, float* dgamma, float* dbeta, cudaStream_t st, const unsigned char* rowflags = nullptr, int n_masked = 0,
                         float* dx_masked = nullptr, float* dx_drop = nullptr, float drop_rate = 0.f, uint32_t drop_seed = 0, uint32_t drop_step = 0,
                         uint32_t drop_site = 0, uint32_t drop_row0 = 0,
                         float* det_part = nullptr /*[kLnBwdMaxCtas][2][D]: deterministic gamma / beta gradients (per-CTA partials + ordered sum)*/);
constexpr int kLnBwdMaxCtas = 148 * 4;
int launch_masked_copies(const float* src, const unsigned char* flags /*[n_copies][T]*/, int n_copies, int T, float* dst /*[n_copies][T][D]*/, cudaStream_t st);
int launch_attention_fwd(const float* qkv, const int* length, int B, int S, float* out, float* lse, cudaStream_t st);
int launch_attention_bwd(const float* qkv, const float* out, const float* lse, const float* dout, const int* length, int B, int S, float* dqkv,
                         cudaStream_t st);
// attention_tc.cu: the same attention core on tcgen05 tensor cores (S <= 128)
class TensorMapCache;
int launch_attention_fwd_tc(TensorMapCache* maps, const float* qkv, const int* length, int B, int S, float* out, float* lse, cudaStream_t st);
int launch_attention_bwd_tc(TensorMapCache* maps, const float* qkv, const float* out, const float* lse, const float* dout, const int* length, int B, int S,
                            float* dqkv, cudaStream_t st);
int launch_dropout_bwd(const float* dx, int T, float rate, uint32_t seed, uint32_t step, uint32_t site, float* dy, cudaStream_t st, uint32_t row0 = 0);
int launch_colsum(const float* x, int rows, int cols, int ld, float* out /*atomic accumulate*/, cudaStream_t st, bool deterministic = false);

// loss.cu
struct LossBuffers {
  float* part;      // [3][F][T]: loss, score, den per (field, position)
  int* idx_true;    // [T] permutation within each document (identity unless sorted)
  int* idx_pred;    // [T]
};
int launch_sort_indices(const Schema& sc, const BatchPtrs& targets, const float* logits, const unsigned char* sort_flag, const int* sort_tasks,
                        int pos_task_id, int B, int S,
                        const LossBuffers& buf, cudaStream_t st);
int launch_loss(const Schema& sc, const BatchPtrs& targets, const MaskPtrs& masks, const float* logits, int use_sort, int B, int S, float inv_batch,
                float* dlogits /*nullable*/, const LossBuffers& buf, float* metrics_out, cudaStream_t st);
int launch_merge_prediction(const Schema& sc, int field, const void* input_col, const unsigned char* mask, const float* logits, int T, float* out,
                            cudaStream_t st);

// optimizer.cu
struct VarDev {
  long long off;
  int rows, cols, ld, l2;
};
int launch_regularization_loss(const VarDev* vars, int V, const float* params, float* norms /*[V][16][2] partial sums*/, float l2, float* out, cudaStream_t st);
int launch_optimizer(const VarDev* vars, int V, float* params, const float* grads, float* m, float* v, float* norms /*[V][16][2] partial sums*/, int t, float lr,
                     float clipnorm, float l2, float* l2_loss_out, cudaStream_t st);

// allreduce.cu: two-shot NVLS all-reduce (multimem.ld_reduce / multimem.st) of a buffer that lives at the same offset of every rank's
// symmetric memory; pads_dev = device array of the ranks' signal pads, slot0 = first uint32 slot this engine may use (2 * world slots),
// call = 1, 2, 3, ... since the pads were zeroed (monotonic flags), local_call = the same count since `sync` (two uint32 in local memory) was zeroed
int launch_nvls_allreduce(float* multicast, uint32_t* const* pads_dev, int slot0, int rank, int world, size_t n_floats, uint32_t call, uint32_t local_call,
                          uint32_t* sync, cudaStream_t st);

}  // namespace mfp
