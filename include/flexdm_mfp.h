/* flexdm_mfp.h -- C ABI of the B200-native MFP (masked field prediction) train/eval step.
 *
 * The reference (CyberAgentAILab/flex-dm) has no FFI: the hot path sits behind the Keras Model/Layer
 * protocol (src/mfp/mfp/models/mfp.py:298 MFP.call; train.py:67-97; eval.py:155-172).  Each entry point
 * below names the reference code it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds
 * (flex_dm_b200/engine.py is that binding).
 *
 * Conventions: every pointer is a DEVICE pointer owned by the caller unless marked "host"; every call
 * that launches work takes a cudaStream_t (passed as void*) and is asynchronous on it; calls return
 * MFP_OK (0) or a negative error code, and mfp_last_error() returns the message (thread-local).
 * One handle per GPU/rank; a handle is not re-entrant, distinct handles are independent.
 * No hidden device allocation: the caller supplies parameters, optimiser state and a workspace.
 */
#ifndef FLEXDM_MFP_H_
#define FLEXDM_MFP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFP_MAX_FIELDS 16
#define MFP_MAX_CANVAS 8
#define MFP_NAME_LEN 96

enum { MFP_OK = 0, MFP_ERR_ARG = -1, MFP_ERR_CUDA = -2, MFP_ERR_STATE = -3, MFP_ERR_UNSUPPORTED = -4 };

/* One sequence column of DataSpec.make_input_columns (data/spec.py:144-211), in
 * get_valid_input_columns order (data/spec.py:393-403). */
typedef struct {
  char name[48];
  int32_t kind;       /* 0 = categorical, 1 = numerical */
  int32_t C;          /* column["shape"][-1]: sub-targets (categorical) or vector width (numerical) */
  int32_t input_dim;  /* categorical vocabulary size (0 for numerical) */
  int32_t task_id;    /* index in get_task_names (masking.py:18-21) of the attribute group holding it */
  int32_t has_cond;   /* "loss_condition" present (crello-spec.yml:88-121) */
  int32_t reserved;
  uint64_t cond_mask; /* bit i set <=> loss_condition["mask"][i] */
} mfp_field_desc;

/* MFP.__init__ keyword arguments (models/mfp.py:215-228) that shape the network. */
typedef struct {
  int32_t num_fields;
  int32_t type_field;  /* index of the "type" column (loss_condition key; sort key) */
  int32_t latent_dim;  /* must be 256 in this build (heads = 8, FFN = 2*latent_dim: transformer.py:43,164) */
  int32_t num_blocks;
  int32_t sort_pos;    /* 1 for rico: LossLayer sort branch (mfp.py:293-296, metrics.py:180-211) */
  int32_t pos_task_id; /* task_names.index("pos") */
  int32_t total_columns; /* len(input_columns) incl. demo-only: total_score divisor (metrics.py:298) */
  int32_t sort_fields[5]; /* indices of type,left,top,width,height (tensor_utils.py:11) */
  float dropout;       /* rate of the two residual-branch Dropouts (transformer.py:174-175) */
  float l2;            /* make_dense_options / make_emb_options coefficient (architecture/utils.py:8-22); <0 = None */
  int32_t input_dtype; /* 0 = "set" (default); 1 = "shuffled_set" / 2 = "sorted_set": the elements of every document are shuffled
                        * (tensor_utils.py:47-76) / sorted by (type,left,top,width,height) (tensor_utils.py:14-44) before the
                        * corruption (mfp.py:104-107; mfp_shuffle_inputs) and the encoder adds a learned PositionEmbedding with
                        * dropout (encoder.py:48-55,251-252; transformer.py:5-30) */
  int32_t length_input_dim; /* input_columns["length"]["input_dim"]: the PositionEmbedding table has this + 1 rows (>= S needed) */
  int32_t block_type;  /* 0 = "deepsvg" (pre-LayerNorm block, transformer.py:208-229; the default), 1 = "transformer"
                        * (post-LayerNorm TransformerBlock, transformer.py:187-205): same variables, same kernels, other wiring */
  int32_t context;     /* --context (encoder.py:96-110,231-249; decoder.py:74-78): 0 = None, 1 = "id" (a learned embedding of the document's task
                        * id), 2 = "length" (of its zero-based length) joins the sequence as one more token that every element attends to and
                        * that the heads ignore.  The engine keeps the token in row length[b] + 1 of the document's S rows (attention without
                        * positions is order-free, so this equals the reference's prepended token up to summation order): every document
                        * needs length[b] + 1 < S -- the host mirror pads the batch by one row.  With input_dtype != 0 the reference adds the
                        * positions after the token was put in front (encoder.py:247-252): the token takes table row 0 and the element in row s
                        * table row s + 1 (S + 1 <= length_input_dim + 1 needed).  That combination has not run on a GPU yet;
                        * the host mirror refuses it by default. */
  int32_t context_rows; /* rows of that embedding table: len(get_task_names(...)) for "id", input_columns["length"]["input_dim"] for "length" */
  /* context = 3 ("canvas") / 4 ("canvas_add") (encoder.py:34-37,177-199,228-249; decoder.py:25-43): the canvas-level categorical columns
   * (valid_input_columns with use_canvas: every non-sequence column but "length") are embedded (tables of input_dim + 2 rows) and
   * summed into one vector per document, which becomes the context token (3; its row as for id / length) or is added to every
   * element of the document (4).  With 3 the decoder also owns a Dense head per canvas column (decoder.py:32-43); LossLayer skips
   * non-sequence columns (metrics.py:226), so those heads only see their L2 term. */
  int32_t n_canvas;
  int32_t canvas_input_dim[MFP_MAX_CANVAS];
  char canvas_names[MFP_MAX_CANVAS][48];
} mfp_config;

/* One trainable variable of the reference model as a strided view of the flat
 * parameter buffer: element (r, c) lives at params[offset + r*ld + c]. */
typedef struct {
  char name[MFP_NAME_LEN]; /* reference attribute path, e.g. model/blocks/seq2seq/seq2seq_0/attn/dense_query/kernel */
  int64_t offset;
  int32_t rows, cols, ld;
  int32_t l2;             /* regularised (Dense kernel+bias, Embedding table; not LayerNorm) */
} mfp_variable;

/* A batch as DataSpec.parse_fn emits it (data/spec.py:255-287), restricted to what the path reads. */
typedef struct {
  const int32_t* length;             /* [B] zero-based (mask.py:28-29) */
  const void* cols[MFP_MAX_FIELDS];  /* categorical: int32 [B,S,C]; numerical: float [B,S,C] */
} mfp_batch;

typedef struct mfp_engine mfp_engine;

const char* mfp_last_error(void);
int mfp_version(void);

/* MFP.__init__ (models/mfp.py:215-296): builds the schema, the parameter layout and the variable table. */
int mfp_create(const mfp_config* cfg, const mfp_field_desc* fields, mfp_engine** out);
void mfp_destroy(mfp_engine* h);

/* Parameter buffer geometry (floats).  Includes alignment padding that belongs to no variable. */
int64_t mfp_param_count(const mfp_engine* h);
int32_t mfp_num_variables(const mfp_engine* h);
int mfp_get_variable(const mfp_engine* h, int32_t index, mfp_variable* out);
int32_t mfp_logit_width(const mfp_engine* h);                      /* padded row width of the logits matrix */
int32_t mfp_field_logit_offset(const mfp_engine* h, int32_t field); /* first column of a field's head (decoder.py:96-110) */

/* Workspace for a (B, S) batch shape; mfp_bind fixes the shape and all buffers (TMA descriptors are built here).
 * Limits (MFP_ERR_UNSUPPORTED): S <= 384; B * S * max(logit_width, 3 * latent_dim) < 2^31 (row offsets are 32-bit in the kernels;
 * at S = 128 and L = 4 that is 11 915 crello documents per step and GPU, whose workspace (110 KB per element) would fill the GPU's
 * 180 GB anyway). */
int64_t mfp_workspace_bytes(const mfp_engine* h, int32_t B, int32_t S);
int mfp_bind(mfp_engine* h, int32_t B, int32_t S, void* workspace, int64_t workspace_bytes,
             float* params, float* grads, float* adam_m, float* adam_v);

/* get_task_cat_dist_sampler(...).sample(B) (models/mfp.py:34-43,301): uniform over `allowed` task ids. */
int mfp_sample_tasks(mfp_engine* h, const int32_t* allowed_host, int32_t n_allowed, uint32_t seed, uint32_t step,
                     int32_t* tasks_out, void* stream);

/* preprocess_for_train (models/mfp.py:95-138): filter_padding + random/elem/feat masking variants selected
 * per document by task id (masking.py:24-53,116-155,227-269).  Writes modified columns and per-field masks [B,S]. */
int mfp_mask_corrupt(mfp_engine* h, const mfp_batch* inputs, const int32_t* tasks, uint32_t seed, uint32_t step,
                     void* const* modified_cols, uint8_t* const* masks_out, void* stream);

/* preprocess_for_test (models/mfp.py:72-92): filter_padding + apply_token(masks, "masked"). */
int mfp_mask_for_test(mfp_engine* h, const mfp_batch* inputs, const uint8_t* const* masks,
                      void* const* modified_cols, void* stream);

/* shuffle_inputs (models/tensor_utils.py:47-76; mfp.py:104-105, --input_dtype shuffled_set): a random permutation of the valid
 * elements of every document, applied to every sequence column (padding stays in place).  The reference draws it with Python's
 * `random`; here element (b, s) gets a Philox key (counter = (b*S+s, 1002), key = (seed, step)) and the valid elements are
 * ordered by it.  With input_dtype = "sorted_set" the same call applies sort_inputs (tensor_utils.py:14-44) instead.  shuffled_cols: one output column per field, same shapes as the inputs; perm_out (optional, int32 [B,S]):
 * source position of every output position. */
int mfp_shuffle_inputs(mfp_engine* h, const mfp_batch* inputs, uint32_t seed, uint32_t step, void* const* shuffled_cols,
                       int32_t* perm_out, void* stream);

/* Model.call(modified_inputs, training) (models/model.py:26-30): encoder -> blocks -> decoder.
 * Logits land in the workspace; `logits_out` (optional, [B*S, logit_width]) receives a copy. */
/* --context id: the per-document ids the context token embeds (device int32 [B]; MFP.call's `tasks`, mfp.py:137,301; eval.py:100-101).
 * The pointer is kept and read by every following mfp_forward / mfp_backward*; "length" contexts read the batch's own length array. */
int mfp_set_context_ids(mfp_engine* h, const int32_t* task_ids);
/* --context canvas / canvas_add: the canvas columns of the current batch (n_canvas device int32 [B] arrays, config order); kept like the ids. */
int mfp_set_canvas_columns(mfp_engine* h, const int32_t* const* columns);

/* Packed numerical columns.  A numerical sequence column of DataSpec.parse_fn (data/spec.py:255-287: float [B,S,512]) is mostly rows
 * the model never reads: filter_padding (masking.py:24-53) overwrites every padded position and every element whose type does not
 * carry the field (loss_condition, data/crello-spec.yml:88-121) with <UNUSED>, and LossLayer gates the same elements out of the loss
 * (metrics.py:251-267).  rowmaps[f] != NULL (device int32 [B*S]) declares that cols[f] of the batches passed to mfp_mask_corrupt,
 * mfp_mask_for_test (inputs) and mfp_loss (targets) is float [n_rows, C] holding only the rows that ARE read, rowmaps[f][t] being
 * element t's row or -1.  Results are identical to the dense column; host->device traffic and the corruption pass's reads shrink to the
 * rows in use.  The pointers are kept until the next call; rowmaps = NULL returns to dense columns.  Not accepted by mfp_shuffle_inputs. */
int mfp_set_packed_rows(mfp_engine* h, const int32_t* const* rowmaps);

int mfp_forward(mfp_engine* h, const mfp_batch* modified, int32_t training, uint32_t seed, uint32_t step,
                float* logits_out, void* stream);

/* LossLayer.call (models/metrics.py:173-299), optionally with the rico sort branch: a document is sorted when
 * sort_flag[b] != 0 (eval.py:105-106) or, if sort_flag is NULL, when sort_tasks[b] == pos_task_id (mfp.py:336-338);
 * both NULL = no sorting.
 * inv_batch = 1/B_global (reduce_mean over the batch, metrics.py:277).  With compute_grad != 0 also writes
 * d(loss)/d(logits) into the workspace for mfp_backward.  logits_in: NULL = the logits mfp_forward left in the
 * workspace (train path, mfp.py:335-340), else a [B*S, logit_width] matrix (eval.py:104-108 passes merged predictions).
 * metrics_out (device, 3*F+1 floats): per field loss, score_num, score_den; then the data loss total. */
int mfp_loss(mfp_engine* h, const mfp_batch* targets, const uint8_t* const* masks, const uint8_t* sort_flag,
             const int32_t* sort_tasks, const float* logits_in, float inv_batch, int32_t compute_grad,
             float* metrics_out, void* stream);

/* GradientTape.gradient of the data loss w.r.t. every variable (Keras default train_step).
 * Writes the flat gradient buffer bound in mfp_bind (overwrites; the L2 term is added in the optimiser). */
int mfp_backward(mfp_engine* h, const mfp_batch* modified, int32_t training, uint32_t seed, uint32_t step, void* stream);

/* The same backward pass in stages, so that a data-parallel host can start the all-reduce of a layer's gradients while the
 * layers below are still being differentiated (the reference has no distributed code).
 * Stage 0 = decoder heads (also clears the gradient buffer), stages 1..L = blocks L-1..0, stage L+1 = encoder.
 * Stages must run in ascending order, each exactly once per step; mfp_backward == stages 0..L+1.
 * mfp_backward_stage_range: the half-open range [lo, hi) of the flat gradient buffer that is final once `stage` has run. */
int32_t mfp_backward_num_stages(const mfp_engine* h);
int mfp_backward_stage_range(const mfp_engine* h, int32_t stage, int64_t* lo, int64_t* hi);
int mfp_backward_stages(mfp_engine* h, const mfp_batch* modified, int32_t training, uint32_t seed, uint32_t step,
                        int32_t first_stage, int32_t last_stage, void* stream);

/* Adam(learning_rate, clipnorm) apply_gradients (train.py:71-77) + the L2 regularisers' gradient and loss term
 * (architecture/utils.py:8-22): g += 2*l2*w; per-variable clip_by_norm; TF-form Adam.  t = 1-based step.
 * l2_loss_out (device float, optional) receives l2 * sum w^2 evaluated BEFORE the update. */
int mfp_optimizer_step(mfp_engine* h, int32_t t, float learning_rate, float clipnorm, float* l2_loss_out, void* stream);

/* Sum of the Keras regularisation losses, l2 * sum w^2 (architecture/utils.py:8-22), without an update:
 * Keras' test_step (model.evaluate, train.py:90) reports loss = add_loss + regularisers too. */
int mfp_regularization_loss(mfp_engine* h, float* l2_loss_out, void* stream);

/* merge_inputs_and_prediction (models/mfp.py:46-69) for one field: logits where masked, one-hot ground
 * truth (categorical) / input (numerical) elsewhere.  out: float [B,S,C,input_dim] or [B,S,C]. */
int mfp_merge_prediction(mfp_engine* h, int32_t field, const void* input_col, const uint8_t* mask, const float* logits_in,
                         float* out, void* stream);

/* Number of kernels launched by this handle since creation (bench.py's gpu_launches). */
int64_t mfp_launch_count(const mfp_engine* h);

/* Arithmetic of the dense contractions (the reference computes them in fp32, transformer.py:61-75; TensorFlow >= 2.4 itself uses TF32 on
 * Ampere and later):
 *   0 = tcgen05 kind::tf32 GEMMs and attention, operands rounded to TF32 (nearest even), fp32 accumulation: the product path (default);
 *   1 = fp32 SIMT GEMM with the same epilogue and fp32 SIMT attention: bring-up / test hook that pins every other kernel at fp32
 *       accuracy, independent of TF32 rounding;
 *   2 = fp32-accurate on the tensor cores: compensated "3xTF32" GEMMs (a_hi b_hi + a_lo b_hi + a_hi b_lo with x_lo = x - tf32(x), one
 *       accumulator, same fused epilogues) and the fp32 SIMT attention core.  About 2-3x the GEMM time of mode 0. */
int mfp_set_gemm_impl(mfp_engine* h, int32_t impl);

/* Deterministic gradient reductions (train.py:18-23 seeds everything "for reproducibility"; Keras on one device has a fixed reduction
 * order).  on != 0: split-K weight gradients, the fused bias-gradient column sums and the LayerNorm gamma / beta gradients are written
 * as per-CTA partials and summed in a fixed order by a second kernel, instead of TMA reduce-add / atomicAdd in arrival order -- two
 * runs of the same step are then bit-identical.  Off (default) is the faster arrival-order accumulation. */
int mfp_set_deterministic(mfp_engine* h, int32_t on);

/* Data-parallel shards (train.py:25 is the reference's single-process stub): index of the bound batch's first document in the GLOBAL
 * batch.  Every Philox counter of the step (task ids mfp.py:34-43, random_masking / elem_masking draws masking.py:98-113,227-269,
 * shuffle keys, dropout keep-masks transformer.py:218,224) is formed from global document / element indices, so rank r with
 * first_document = r * B_local draws exactly what a single process draws for those documents.  Default 0. */
int mfp_set_doc_offset(mfp_engine* h, int64_t first_document);

/* Data-parallel gradient exchange without NCCL on the critical path (the reference has no distributed code,
 * train.py:25): two-shot all-reduce over NVLink SHARP in ONE kernel on `stream`.  Requirements: the gradient buffer bound by mfp_bind
 * lives at the same offset of every rank's symmetric memory, `multicast_grads` is the multicast address of that buffer (every rank's
 * copy behind one address), `signal_pads_dev` a device array of the `world` ranks' signal pads (uint32 slots, zero-initialised;
 * slots [first_slot, first_slot + world) are used), `call` = 1, 2, 3, ... counts the calls since the pads were zeroed (same on every
 * rank).  On return (in stream order) every rank's buffer holds the sum over ranks.  All waits are bounded and trap. */
int mfp_allreduce_gradients_nvls(mfp_engine* h, float* multicast_grads, void* const* signal_pads_dev, int32_t first_slot, int32_t rank,
                                 int32_t world, uint32_t call, void* stream);

/* Optional device timing of kernel classes (bench.py's roofline): between begin and end every launch of the class is
 * bracketed by CUDA events on the launching stream; end synchronises and returns the summed milliseconds, the launch
 * count and the algorithmic HBM bytes (every operand and output once) per class (host arrays of MFP_PROFILE_CLASSES
 * entries).  Not meant for the throughput-timed region. */
#define MFP_PROFILE_CLASSES 2
#define MFP_PROFILE_GEMM 0
#define MFP_PROFILE_ATTENTION 1
int mfp_profile_begin(mfp_engine* h);
int mfp_profile_end(mfp_engine* h, float* ms_per_class_host, int32_t* launches_per_class_host, double* bytes_per_class_host);

/* Device-resident dataset cache (the B200 answer to dataset.cache() of DataSpec.make_dataset, data/spec.py:238-239): a parsed split is
 * kept ragged in HBM -- per sequence column one [total_elements, C] array of 32-bit words, documents back to back -- and a batch is
 * cut out of it on the device.  For every column c and every b < B, s < S:
 *     dst[c][b, s, :] = s < doc_len[idx[b]] ? src[c][doc_start[idx[b]] + s, :] : pad_word[c]
 * (pad_word = the parse default put through the column's preprocessor, like parse_sequence_example's padding).  Engine-independent. */
#define MFP_GATHER_MAX_COLUMNS 24
typedef struct {
  int32_t n_columns;
  int32_t words[MFP_GATHER_MAX_COLUMNS];       /* 32-bit words per element of the column (C for int32 / float32 columns) */
  uint32_t pad_word[MFP_GATHER_MAX_COLUMNS];
  const void* src[MFP_GATHER_MAX_COLUMNS];     /* device, [total_elements, words] */
  void* dst[MFP_GATHER_MAX_COLUMNS];           /* device, [B, S, words] */
} mfp_gather_desc;
int mfp_gather_documents(const mfp_gather_desc* desc, const int64_t* doc_start, const int32_t* doc_len, const int32_t* idx, int32_t B,
                         int32_t S, void* stream);

/* Bring-up hook: the attention core of MultiHeadSelfAttention (architecture/transformer.py:60-76) alone.
 * qkv [B*S, 768] (q | k | v, head h = columns 32h..32h+31 of each third), length [B] zero-based;
 * out [B*S, 256] heads merged, lse [B, 8, S].  impl: 0 = tcgen05 (S <= 128), 1 = SIMT. */
int mfp_debug_attention(const float* qkv, const int32_t* length, int32_t B, int32_t S, float* out, float* lse,
                        int32_t impl, void* stream);

/* Bring-up hook: backward of the attention core: dqkv [B*S, 768] from qkv, the forward's out / lse and dout [B*S, 256]. */
int mfp_debug_attention_bwd(const float* qkv, const int32_t* length, int32_t B, int32_t S, const float* out, const float* lse,
                            const float* dout, float* dqkv, int32_t impl, void* stream);

/* Bring-up hook: D[M,N] = epilogue(A . B^T) through the same tcgen05/TMA GEMM the engine uses.
 * a_mn / b_mn: 0 = operand is K-major ([rows=M|N][K] row-major, pitch ld), 1 = MN-major ([K][M|N] row-major).
 * Optional epilogue operands: bias [N]; relu != 0; residual [M,N] (pitch ldd) added last; relu_src [M,N] (pitch ldd):
 * result zeroed where relu_src <= 0; colsum [N] += column sums of the MN-major B operand (bias gradient of a wgrad).
 * splits > 1 accumulates into D (which must be pre-filled).  impl: 0 = tcgen05, 1 = SIMT bring-up kernel. */
int mfp_debug_gemm(const float* A, int32_t a_mn, int32_t lda, const float* B, int32_t b_mn, int32_t ldb,
                   float* D, int32_t ldd, int32_t M, int32_t N, int32_t K, const float* bias, int32_t relu,
                   const float* residual, const float* relu_src, float* colsum, int32_t splits, int32_t impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLEXDM_MFP_H_ */
