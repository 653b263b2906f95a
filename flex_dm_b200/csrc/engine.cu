// C ABI of the MFP engine (include/flexdm_mfp.h): schema + parameter layout, workspace plan, and the host-side
// orchestration of the forward / loss / backward / optimiser kernels on a caller-supplied stream.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "gemm.cuh"
#include "kernels.cuh"

namespace mfp {

static thread_local char g_error[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

struct BlockLayout {
  long long wqkv, bqkv, wo, bo, w1, b1, w2, b2, g1, be1, g2, be2;
};

struct Workspace {
  // byte offsets into the caller's workspace
  size_t vars, flags, perm, x, ln1, qkv, attn, xmid, ln2, hid, gates, stats, lse, logits, dlogits, dx, dtmp, dy, dqkv, dhid, dattn, dh0m, onehot, rowgrad, part, idx_true, idx_pred,
      norms, ctx_row, canvas_vec, dcanvas, iota, zeros, det, ln_part, lo_a, lo_b, ar_sync, total;
};

}  // namespace mfp

using namespace mfp;

struct mfp_engine {
  mfp_config cfg;
  Schema sc;
  std::vector<std::string> field_names;
  std::vector<mfp_variable> vars;
  std::vector<BlockLayout> blocks;
  long long wh = 0, bh = 0, param_count = 0;
  long long pos_off = -1;  // PositionEmbedding table (input_dtype != "set"), rows = length_input_dim + 1
  long long ctx_off = -1;  // --context id / length: embedding table of the context token, rows = cfg.context_rows
  const int32_t* ctx_ids = nullptr;  // --context id: device task ids of the current batch (mfp_set_context_ids)
  long long canvas_off[MFP_MAX_CANVAS] = {0};  // --context canvas / canvas_add: embedding tables of the canvas columns
  const int32_t* canvas_ids[MFP_MAX_CANVAS] = {nullptr};
  const int32_t* rowmaps[MFP_MAX_FIELDS] = {nullptr};  // packed numerical input / target columns (mfp_set_packed_rows)  // device columns of the current batch (mfp_set_canvas_columns)
  // bound state
  int B = 0, S = 0, T = 0;
  uint8_t* ws = nullptr;
  Workspace off{};
  float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
  TensorMapCache* maps = nullptr;
  std::vector<long long> stage_lo, stage_hi;  // flat-buffer range whose gradients are final after backward stage s (see mfp_backward_stages)
  const void* flags_for = nullptr;  // modified column (first numerical field) the workspace row flags were just derived from
  int gemm_impl = 0;
  bool gates_valid = false;  // the last forward wrote the FFN's ReLU gates as bits (workspace `gates`): the backward reads those instead of hid
  uint32_t ar_calls = 0;   // NVLS all-reduce calls since the last bind (monotonic barrier flags)
  int deterministic = 0;   // mfp_set_deterministic: fixed-order gradient reductions (bit-identical steps from run to run)
  uint32_t doc0 = 0;       // mfp_set_doc_offset: global index of the bound batch's first document (data-parallel shard)
  int64_t launches = 0;
  // optional per-kernel-class device timing (bench.py roofline): CUDA event pairs around each launch
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events[MFP_PROFILE_CLASSES];
  double prof_bytes[MFP_PROFILE_CLASSES] = {0, 0};  // algorithmic HBM bytes of the profiled launches
};

namespace mfp {

static long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

static void add_var(mfp_engine* h, const std::string& name, long long off, int rows, int cols, int ld, int l2) {
  mfp_variable v;
  memset(&v, 0, sizeof(v));
  snprintf(v.name, sizeof(v.name), "%s", name.c_str());
  v.offset = off;
  v.rows = rows;
  v.cols = cols;
  v.ld = ld;
  v.l2 = l2;
  h->vars.push_back(v);
}

static void build_layout(mfp_engine* h) {
  const int D = kD, L = h->cfg.num_blocks;
  long long cur = 0;
  auto alloc = [&](long long n) {
    const long long o = cur;
    cur = align_up(cur + n, 64);
    return o;
  };
  Schema& sc = h->sc;
  // logits geometry: each head starts at a multiple of 4 columns
  int lw = 0;
  for (int f = 0; f < sc.F; ++f) {
    FieldDev& fd = sc.f[f];
    fd.logit_w = (fd.kind == 0) ? fd.C * fd.input_dim : fd.C;
    fd.logit_off = lw;
    lw += (fd.logit_w + 3) & ~3;
  }
  sc.LWu = lw;
  sc.LW = (lw + 31) & ~31;
  // encoder (encoder.py:72-92)
  for (int f = 0; f < sc.F; ++f) {
    FieldDev& fd = sc.f[f];
    const std::string base = "model/encoder/input_layer/" + h->field_names[f];
    if (fd.kind == 0) {
      fd.table_off = alloc((long long)(fd.input_dim + 2) * D);
      fd.kernel_off = fd.bias_off = -1;
      add_var(h, base + "/embeddings", fd.table_off, fd.input_dim + 2, D, D, 1);
    } else {
      fd.table_off = alloc(2LL * D);
      fd.kernel_off = alloc((long long)fd.C * D);
      fd.bias_off = alloc(D);
      add_var(h, base + "_special/embeddings", fd.table_off, 2, D, D, 1);
      add_var(h, base + "/kernel", fd.kernel_off, fd.C, D, D, 1);
      add_var(h, base + "/bias", fd.bias_off, 1, D, D, 1);
    }
  }
  if (h->cfg.input_dtype != 0) {  // PositionEmbedding(latent_dim, maxlen = length input_dim) -> Embedding(maxlen + 1, D): encoder.py:48-55, transformer.py:17-21
    h->pos_off = alloc((long long)(h->cfg.length_input_dim + 1) * D);
    add_var(h, "model/encoder/input_layer/const/embeddings/embeddings", h->pos_off, h->cfg.length_input_dim + 1, D, D, 1);
  }
  if (h->cfg.context == 1 || h->cfg.context == 2) {  // encoder.py:96-110: input_layer["task"] / input_layer["length"]
    h->ctx_off = alloc((long long)h->cfg.context_rows * D);
    add_var(h, std::string("model/encoder/input_layer/") + (h->cfg.context == 1 ? "task" : "length") + "/embeddings", h->ctx_off, h->cfg.context_rows, D, D, 1);
  }
  if (h->cfg.context >= 3) {  // canvas columns are embedded like categorical sequence columns (encoder.py:72-79 over valid_input_columns with use_canvas)
    for (int c = 0; c < h->cfg.n_canvas; ++c) {
      const int rows = h->cfg.canvas_input_dim[c] + 2;
      h->canvas_off[c] = alloc((long long)rows * D);
      const std::string name(h->cfg.canvas_names[c], strnlen(h->cfg.canvas_names[c], sizeof(h->cfg.canvas_names[c])));
      add_var(h, "model/encoder/input_layer/" + name + "/embeddings", h->canvas_off[c], rows, D, D, 1);
    }
  }
  // backward stages: 0 = heads, 1..L = blocks L-1..0, L+1 = encoder; the layout is encoder | blocks | heads, each contiguous
  h->stage_lo.assign(L + 2, 0);
  h->stage_hi.assign(L + 2, 0);
  h->stage_lo[L + 1] = 0;
  h->stage_hi[L + 1] = cur;
  // blocks (transformer.py:54-57,161-173); Q|K|V kernels share one [D, 3D] matrix so the projection is one GEMM
  h->blocks.resize(L);
  for (int i = 0; i < L; ++i) {
    BlockLayout& b = h->blocks[i];
    h->stage_lo[L - i] = cur;
    b.wqkv = alloc(3LL * D * D); b.bqkv = alloc(3 * D);
    b.wo = alloc((long long)D * D); b.bo = alloc(D);
    b.w1 = alloc((long long)D * kF); b.b1 = alloc(kF);
    b.w2 = alloc((long long)kF * D); b.b2 = alloc(D);
    b.g1 = alloc(D); b.be1 = alloc(D); b.g2 = alloc(D); b.be2 = alloc(D);
    char p[64];
    snprintf(p, sizeof(p), "model/blocks/seq2seq/seq2seq_%d", i);
    const std::string s(p);
    const char* qkv_names[3] = {"dense_query", "dense_key", "dense_value"};
    for (int j = 0; j < 3; ++j) {
      add_var(h, s + "/attn/" + qkv_names[j] + "/kernel", b.wqkv + (long long)j * D, D, D, 3 * D, 1);
      add_var(h, s + "/attn/" + qkv_names[j] + "/bias", b.bqkv + (long long)j * D, 1, D, D, 1);
    }
    add_var(h, s + "/attn/combine_heads/kernel", b.wo, D, D, D, 1);
    add_var(h, s + "/attn/combine_heads/bias", b.bo, 1, D, D, 1);
    add_var(h, s + "/mlp/layer_with_weights-0/kernel", b.w1, D, kF, kF, 1);
    add_var(h, s + "/mlp/layer_with_weights-0/bias", b.b1, 1, kF, kF, 1);
    add_var(h, s + "/mlp/layer_with_weights-1/kernel", b.w2, kF, D, D, 1);
    add_var(h, s + "/mlp/layer_with_weights-1/bias", b.b2, 1, D, D, 1);
    add_var(h, s + "/norm1/gamma", b.g1, 1, D, D, 0);
    add_var(h, s + "/norm1/beta", b.be1, 1, D, D, 0);
    add_var(h, s + "/norm2/gamma", b.g2, 1, D, D, 0);
    add_var(h, s + "/norm2/beta", b.be2, 1, D, D, 0);
    h->stage_hi[L - i] = cur;
  }
  h->stage_lo[0] = cur;
  // decoder heads (decoder.py:32-43), concatenated into one [D, LW] matrix
  h->wh = alloc((long long)D * sc.LW);
  h->bh = alloc(sc.LW);
  for (int f = 0; f < sc.F; ++f) {
    const FieldDev& fd = sc.f[f];
    const std::string base = "model/decoder/decoders/" + h->field_names[f];
    add_var(h, base + "/kernel", h->wh + fd.logit_off, D, fd.logit_w, sc.LW, 1);
    add_var(h, base + "/bias", h->bh + fd.logit_off, 1, fd.logit_w, sc.LW, 1);
  }
  if (h->cfg.context == 3) {  // decoder.py:25-43: with use_canvas the decoder owns a head per canvas column; nothing reads them on this path
    for (int c = 0; c < h->cfg.n_canvas; ++c) {  // (LossLayer skips non-sequence columns, metrics.py:226): variables with an L2 term only
      const int w = h->cfg.canvas_input_dim[c];
      const long long wk = alloc((long long)D * w), wb = alloc(w);
      const std::string name(h->cfg.canvas_names[c], strnlen(h->cfg.canvas_names[c], sizeof(h->cfg.canvas_names[c])));
      add_var(h, "model/decoder/decoders/" + name + "/kernel", wk, D, w, w, 1);
      add_var(h, "model/decoder/decoders/" + name + "/bias", wb, 1, w, w, 1);
    }
  }
  h->param_count = cur;
  h->stage_hi[0] = cur;
}

// rows per chunk plane of the FFN's ReLU gate words ([block][chunk of 32 hidden units][row]): whole 128-byte lines per warp of 32 rows
static size_t gate_rows(size_t T) { return (T + 31) / 32 * 32; }
// FLEXDM_RELU_BITS=0 keeps the hidden activation itself as the ReLU-mask operand of the FFN's input gradient (A/B switch, read per call)
static bool relu_bits_enabled() { const char* e = getenv("FLEXDM_RELU_BITS"); return !(e && e[0] == '0'); }
// algorithmic bytes of a GEMM: each operand and the output once, plus the residual / ReLU-mask operand (or the gate words: one bit per output)
static double gemm_bytes(int M, int N, int K, const GemmEpilogue& ep) {
  return 4.0 * ((double)M * K + (double)N * K + (double)M * N * ((ep.residual || ep.relu_src) ? 2.0 : 1.0)) +
         ((ep.relu_bits || ep.relu_bits_out) ? (double)M * N / 8.0 : 0.0);
}

static Workspace plan_workspace(const mfp_engine* h, int B, int S) {
  Workspace w{};
  const size_t T = (size_t)B * S, L = h->cfg.num_blocks, D = kD, F = h->sc.F;
  size_t cur = 0;
  auto take = [&](size_t bytes) {
    const size_t o = cur;
    cur = (cur + bytes + 255) / 256 * 256;
    return o;
  };
  const size_t fl = sizeof(float);
  w.vars = take(h->vars.size() * sizeof(VarDev));
  w.flags = take((size_t)(h->sc.n_num > 0 ? h->sc.n_num : 1) * T);
  w.perm = take(T * sizeof(int));
  w.x = take((L + 1) * T * D * fl);
  w.ln1 = take(L * T * D * fl);
  w.qkv = take(L * T * 3 * D * fl);
  w.attn = take(L * T * D * fl);
  w.xmid = take(L * T * D * fl);
  w.ln2 = take(L * T * D * fl);
  w.hid = take(L * T * kF * fl);
  w.gates = take(L * (size_t)(kF / 32) * gate_rows(T) * sizeof(uint32_t));  // ReLU gates of the FFN hidden layer as bits: [block][chunk][row]
  w.stats = take(L * 4 * T * fl);
  w.lse = take(L * (size_t)B * kH * S * fl);
  w.logits = take(T * h->sc.LW * fl);
  w.dlogits = take(T * h->sc.LW * fl);
  w.dx = take(T * D * fl);
  w.dtmp = take(T * D * fl);
  w.dy = take(T * D * fl);
  w.dqkv = take(T * 3 * D * fl);
  w.dhid = take(T * kF * fl);
  w.dattn = take(T * D * fl);
  w.dh0m = take((size_t)(h->sc.n_num > 0 ? h->sc.n_num : 1) * T * D * fl);
  w.onehot = take(T * (size_t)h->sc.Rp * fl);
  w.rowgrad = take((size_t)h->sc.Rp * D * fl);
  w.part = take(3 * F * T * fl);
  w.idx_true = take(T * sizeof(int));
  w.idx_pred = take(T * sizeof(int));
  w.norms = take(2 * 16 * h->vars.size() * fl);
  w.ctx_row = take((size_t)B * sizeof(int));
  w.canvas_vec = take((size_t)B * D * fl);
  w.dcanvas = take((size_t)B * D * fl);
  w.iota = take((size_t)B * sizeof(int));
  w.zeros = take((size_t)B * sizeof(int));
  w.det = take(3 * kDetWsFloats * fl);  // one third per problem of a group launch
  w.ln_part = take((size_t)kLnBwdMaxCtas * 2 * D * fl);
  // 3xTF32 mode: low parts of the two operands of one GEMM (the widest operands are [T, LW] logits gradients and [T, 768] qkv)
  const size_t widest = std::max<size_t>(std::max<size_t>(h->sc.LW, 3 * D), std::max<size_t>((size_t)h->sc.Rp, 512));
  const size_t lo_rows = std::max<size_t>(T, 512);  // weight operands have up to 512 rows (FFN2, numerical Dense) whatever T is
  w.lo_a = take(lo_rows * widest * fl);
  w.lo_b = take(lo_rows * widest * fl);
  w.ar_sync = take(256);
  w.total = cur;
  return w;
}

template <typename Tp>
static Tp* wsp(const mfp_engine* h, size_t off) { return reinterpret_cast<Tp*>(h->ws + off); }

// packed: the batch is the caller's input / target batch, whose numerical columns may be packed (mfp_set_packed_rows); the engine's own
// modified columns are always dense
static BatchPtrs to_batch(const mfp_engine* h, const mfp_batch* b, bool packed = false) {
  BatchPtrs p{};
  p.length = b->length;
  for (int f = 0; f < h->sc.F; ++f) {
    p.cols[f] = b->cols[f];
    p.rowmap[f] = (packed && h->sc.f[f].kind == 1) ? h->rowmaps[f] : nullptr;
  }
  return p;
}

static bool any_packed(const mfp_engine* h) {
  for (int f = 0; f < h->sc.F; ++f)
    if (h->rowmaps[f]) return true;
  return false;
}

static int canvas_args(const mfp_engine* h, CanvasArgs* ca) {
  ca->n = h->cfg.n_canvas;
  for (int c = 0; c < ca->n; ++c) {
    if (!h->canvas_ids[c]) { set_error("context = canvas / canvas_add needs mfp_set_canvas_columns first"); return MFP_ERR_STATE; }
    ca->ids[c] = h->canvas_ids[c];
    ca->off[c] = h->canvas_off[c];
    ca->rows[c] = h->cfg.canvas_input_dim[c] + 2;
  }
  return MFP_OK;
}

static int first_numerical(const Schema& sc) {
  for (int f = 0; f < sc.F; ++f)
    if (sc.f[f].kind == 1) return f;
  return 0;
}

static int check_bound(const mfp_engine* h) {
  if (!h) { set_error("null engine"); return MFP_ERR_ARG; }
  if (!h->ws) { set_error("engine is not bound: call mfp_bind first"); return MFP_ERR_STATE; }
  return MFP_OK;
}

struct ProfScope {
  mfp_engine* h; int cls; cudaStream_t st; cudaEvent_t stop = nullptr;
  ProfScope(mfp_engine* h_, int cls_, cudaStream_t st_, double bytes = 0.0) : h(h_), cls(cls_), st(st_) {
    if (!h->profiling) return;
    h->prof_bytes[cls] += bytes;
    cudaEvent_t start;
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    cudaEventRecord(start, st);
    h->prof_events[cls].push_back(start);
    h->prof_events[cls].push_back(stop);
  }
  ~ProfScope() { if (stop) cudaEventRecord(stop, st); }
};

// colsum (optional, [N]): += sum over K of the B operand's columns -- the bias gradient of a wgrad GEMM, taken from the
// B tiles while they sit in shared memory (tcgen05 path) or by a separate kernel (SIMT bring-up path).
static int gemm(mfp_engine* h, const float* A, int a_mn, int lda, const float* Bp, int b_mn, int ldb, int M, int N, int K, const GemmEpilogue& ep,
                int splits, cudaStream_t st, float* colsum = nullptr) {
  GemmCall c{};
  c.a = GemmOperand{A, a_mn, lda};
  c.b = GemmOperand{Bp, b_mn, ldb};
  c.M = M; c.N = N; c.K = K;
  c.splits = splits;
  c.ep = ep;
  c.colsum = (h->gemm_impl != 1) ? colsum : nullptr;
  if (h->deterministic) { c.det_ws = wsp<float>(h, h->off.det); c.det_ws_floats = kDetWsFloats; }
  if (h->gemm_impl == 2) {
    // 3xTF32: x_lo = x - tf32(x) of both operands (same layout and pitch), then a_hi b_hi + a_lo b_hi + a_hi b_lo in one accumulator.
    // An operand is [rows = M|N][K] with pitch ld (K-major) or [K][M|N] (MN-major).
    float* alo = wsp<float>(h, h->off.lo_a);
    float* blo = wsp<float>(h, h->off.lo_b);
    // (whole pitch rows: widths like the one-hot matrix's R are not multiples of 4, pitches are)
    MFP_TRY(launch_split_tf32_lo(A, a_mn ? K : M, lda, lda, alo, st));
    MFP_TRY(launch_split_tf32_lo(Bp, b_mn ? K : N, ldb, ldb, blo, st));
    c.a_lo = alo; c.b_lo = blo;
    h->launches += 2;
  }
  h->launches++;
  if (h->deterministic && h->gemm_impl != 1 && (splits > 1 || colsum)) h->launches++;  // splitk_reduce_kernel
  if (colsum && h->gemm_impl == 1) {
    MFP_TRY(launch_colsum(Bp, K, N, ldb, colsum, st, h->deterministic != 0));
    h->launches++;
  }
  const double bytes = gemm_bytes(M, N, K, ep);
  ProfScope prof(h, MFP_PROFILE_GEMM, st, bytes);
  return launch_gemm(h->maps, c, h->gemm_impl == 2 ? 0 : h->gemm_impl, st);
}

// A GEMM of a group launch: the arguments of gemm() as a value.
struct GemmArgs {
  const float* A; int a_mn, lda;
  const float* B; int b_mn, ldb;
  int M, N, K;
  GemmEpilogue ep;
  int splits;
  float* colsum;
};

// Independent GEMMs (no output of one is an input of another) as ONE launch on the tcgen05 path: a weight gradient and the input
// gradient off the same dY, the encoder's weight gradients.  Their tiles form one tile space (problem 0 first), so the split-K reduce-add
// epilogue of a weight-gradient tile runs under the next tile's main loop and one drain / prologue / ramp per extra problem disappears.
// Other GEMM implementations (SIMT bring-up, 3xTF32 with its shared operand scratch) run them one after the other.
static int gemm_group(mfp_engine* h, const GemmArgs* g, int n, cudaStream_t st) {
  if (h->gemm_impl != 0 || n == 1) {
    for (int i = 0; i < n; ++i) MFP_TRY(gemm(h, g[i].A, g[i].a_mn, g[i].lda, g[i].B, g[i].b_mn, g[i].ldb, g[i].M, g[i].N, g[i].K, g[i].ep, g[i].splits, st, g[i].colsum));
    return MFP_OK;
  }
  GemmCall calls[3];
  double bytes = 0.0;
  if (n > 3) { set_error("gemm_group: at most 3 problems"); return MFP_ERR_ARG; }
  for (int i = 0; i < n; ++i) {
    GemmCall& c = calls[i];
    c = GemmCall{};
    c.a = GemmOperand{g[i].A, g[i].a_mn, g[i].lda};
    c.b = GemmOperand{g[i].B, g[i].b_mn, g[i].ldb};
    c.M = g[i].M; c.N = g[i].N; c.K = g[i].K;
    c.splits = g[i].splits;
    c.ep = g[i].ep;
    c.colsum = g[i].colsum;
    if (h->deterministic) {  // every problem of the launch sums its partials in its own third of the scratch block
      c.det_ws = wsp<float>(h, h->off.det) + (size_t)i * kDetWsFloats;
      c.det_ws_floats = kDetWsFloats;
      if (g[i].splits > 1 || g[i].colsum) h->launches++;  // splitk_reduce_kernel
    }
    bytes += gemm_bytes(c.M, c.N, c.K, c.ep);
  }
  h->launches++;
  ProfScope prof(h, MFP_PROFILE_GEMM, st, bytes);
  return launch_gemm_group(h->maps, calls, n, st);
}

// split-K factor of a weight-gradient GEMM (K = tokens): one wave of the persistent GEMM's tiles (gemm.cu knows the tile shape)
static int wgrad_splits(int M, int N, int K) { return gemm_wgrad_splits(M, N, K); }

}  // namespace mfp

// ===================================================================================================== C ABI
extern "C" {

const char* mfp_last_error(void) { return g_error; }
int mfp_version(void) { return 1; }

int mfp_create(const mfp_config* cfg, const mfp_field_desc* fields, mfp_engine** out) {
  if (!cfg || !fields || !out) { set_error("mfp_create: null argument"); return MFP_ERR_ARG; }
  if (cfg->latent_dim != kD) { set_error("mfp_create: latent_dim must be %d in this build (got %d)", kD, cfg->latent_dim); return MFP_ERR_UNSUPPORTED; }
  if (cfg->num_fields < 1 || cfg->num_fields > kMaxFields) { set_error("mfp_create: num_fields out of range"); return MFP_ERR_ARG; }
  if (cfg->num_blocks < 1 || cfg->num_blocks > 64) { set_error("mfp_create: num_blocks out of range"); return MFP_ERR_ARG; }
  if (cfg->input_dtype < 0 || cfg->input_dtype > 2) { set_error("mfp_create: input_dtype must be 0 (set), 1 (shuffled_set) or 2 (sorted_set)"); return MFP_ERR_ARG; }
  if (cfg->input_dtype != 0 && cfg->length_input_dim < 1) { set_error("mfp_create: shuffled_set / sorted_set need length_input_dim"); return MFP_ERR_ARG; }
  if (cfg->context < 0 || cfg->context > 4) { set_error("mfp_create: context must be 0 (None), 1 (id), 2 (length), 3 (canvas) or 4 (canvas_add)"); return MFP_ERR_ARG; }
  if ((cfg->context == 1 || cfg->context == 2) && cfg->context_rows < 1) { set_error("mfp_create: context needs context_rows >= 1"); return MFP_ERR_ARG; }
  if (cfg->context >= 3) {
    if (cfg->n_canvas < 1 || cfg->n_canvas > MFP_MAX_CANVAS) { set_error("mfp_create: canvas contexts need 1..%d canvas columns (encoder.py:205-206)", MFP_MAX_CANVAS); return MFP_ERR_ARG; }
    for (int c = 0; c < cfg->n_canvas; ++c)
      if (cfg->canvas_input_dim[c] < 1) { set_error("mfp_create: canvas column %d has no input_dim", c); return MFP_ERR_ARG; }
  }
  if (cfg->block_type != 0 && cfg->block_type != 1) { set_error("mfp_create: block_type must be 0 (deepsvg) or 1 (transformer)"); return MFP_ERR_ARG; }
  mfp_engine* h = new mfp_engine();
  h->cfg = *cfg;
  memset(&h->sc, 0, sizeof(h->sc));
  h->sc.F = cfg->num_fields;
  h->sc.type_field = cfg->type_field;
  for (int i = 0; i < 5; ++i) h->sc.sort_field[i] = cfg->sort_fields[i];
  int n_num = 0;
  for (int f = 0; f < cfg->num_fields; ++f) {
    const mfp_field_desc& d = fields[f];
    FieldDev& fd = h->sc.f[f];
    fd.kind = d.kind;
    fd.C = d.C;
    fd.input_dim = d.input_dim;
    fd.task_id = d.task_id;
    fd.has_cond = d.has_cond;
    fd.cond_mask = d.cond_mask;
    fd.num_slot = (d.kind == 1) ? n_num++ : -1;
    if (d.kind == 0 && (d.input_dim < 1 || d.input_dim > 256 || d.C < 1 || d.C > 32)) {
      set_error("mfp_create: categorical field %s needs 1 <= input_dim <= 256 and 1 <= C <= 32", d.name);
      delete h;
      return MFP_ERR_UNSUPPORTED;
    }
    if (d.kind == 1 && (d.C % 32 != 0 || d.C < 32)) {
      set_error("mfp_create: numerical field %s needs a width that is a multiple of 32", d.name);
      delete h;
      return MFP_ERR_UNSUPPORTED;
    }
    h->field_names.push_back(std::string(d.name, strnlen(d.name, sizeof(d.name))));
  }
  h->sc.n_num = n_num;
  int n_lk = 0;
  for (int f = 0; f < cfg->num_fields; ++f) {
    const int subs = (h->sc.f[f].kind == 0) ? h->sc.f[f].C : 1;
    if (n_lk + subs > kMaxLookups) { set_error("mfp_create: more than %d embedding lookups per element", kMaxLookups); delete h; return MFP_ERR_UNSUPPORTED; }
    for (int c = 0; c < subs; ++c) { h->sc.lk_field[n_lk] = (unsigned char)f; h->sc.lk_sub[n_lk] = (unsigned char)c; ++n_lk; }
  }
  h->sc.n_lookups = n_lk;
  int grow = 0;
  for (int f = 0; f < cfg->num_fields; ++f) {
    h->sc.f[f].grow_off = grow;
    grow += (h->sc.f[f].kind == 0) ? h->sc.f[f].input_dim + 2 : 3;
  }
  h->sc.R = grow;
  h->sc.Rp = (grow + 3) & ~3;
  if (cfg->type_field < 0 || cfg->type_field >= cfg->num_fields || h->sc.f[cfg->type_field].kind != 0) {
    set_error("mfp_create: type_field must index a categorical field");
    delete h;
    return MFP_ERR_ARG;
  }
  build_layout(h);
  const char* impl = getenv("FLEXDM_GEMM");
  h->gemm_impl = (impl && !strcmp(impl, "simt")) ? 1 : 0;
  h->maps = tensor_map_cache_create();
  *out = h;
  return MFP_OK;
}

void mfp_destroy(mfp_engine* h) {
  if (!h) return;
  tensor_map_cache_destroy(h->maps);
  delete h;
}

int64_t mfp_param_count(const mfp_engine* h) { return h ? h->param_count : 0; }
int32_t mfp_num_variables(const mfp_engine* h) { return h ? (int32_t)h->vars.size() : 0; }
int mfp_get_variable(const mfp_engine* h, int32_t index, mfp_variable* out) {
  if (!h || !out || index < 0 || index >= (int32_t)h->vars.size()) { set_error("mfp_get_variable: bad index"); return MFP_ERR_ARG; }
  *out = h->vars[index];
  return MFP_OK;
}
int32_t mfp_logit_width(const mfp_engine* h) { return h ? h->sc.LW : 0; }
int32_t mfp_field_logit_offset(const mfp_engine* h, int32_t field) {
  if (!h || field < 0 || field >= h->sc.F) return -1;
  return h->sc.f[field].logit_off;
}

int64_t mfp_workspace_bytes(const mfp_engine* h, int32_t B, int32_t S) {
  if (!h || B < 1 || S < 1) return 0;
  return (int64_t)plan_workspace(h, B, S).total;
}

int mfp_bind(mfp_engine* h, int32_t B, int32_t S, void* workspace, int64_t workspace_bytes, float* params, float* grads, float* adam_m,
             float* adam_v) {
  if (!h || !workspace || !params) { set_error("mfp_bind: null argument"); return MFP_ERR_ARG; }
  if (B < 1 || S < 1) { set_error("mfp_bind: bad shape"); return MFP_ERR_ARG; }
  if (S > 384) { set_error("mfp_bind: S = %d exceeds the attention kernels' shared-memory plan (max 384)", S); return MFP_ERR_UNSUPPORTED; }
  {  // element counts and row offsets are 32-bit inside the kernels: the widest row-major array ([T, LW] logits, [T, 768] qkv) must stay below 2^31 entries
    const long long widest = std::max<long long>(h->sc.LW, 3LL * h->cfg.latent_dim);
    if ((long long)B * S * widest >= (1LL << 31)) {
      set_error("mfp_bind: B * S = %lld elements exceed 32-bit indexing of a %lld-column array; split the batch", (long long)B * S, widest);
      return MFP_ERR_UNSUPPORTED;
    }
  }
  {
    const int positions = S + ((h->cfg.context >= 1 && h->cfg.context <= 3) ? 1 : 0);  // a context token takes position 0 (encoder.py:247-252)
    if (h->cfg.input_dtype != 0 && positions > h->cfg.length_input_dim + 1) {
      set_error("mfp_bind: %d positions exceed the PositionEmbedding table (%d rows)", positions, h->cfg.length_input_dim + 1);
      return MFP_ERR_ARG;
    }
  }
  if (h->cfg.context >= 1 && h->cfg.context <= 3 && S < 2) { set_error("mfp_bind: a context token needs S >= 2 (one row beyond the longest document)"); return MFP_ERR_ARG; }
  const Workspace w = plan_workspace(h, B, S);
  if ((size_t)workspace_bytes < w.total) { set_error("mfp_bind: workspace too small (%lld < %zu)", (long long)workspace_bytes, w.total); return MFP_ERR_ARG; }
  if (reinterpret_cast<uintptr_t>(workspace) & 255) { set_error("mfp_bind: workspace must be 256-byte aligned"); return MFP_ERR_ARG; }
  h->B = B; h->S = S; h->T = B * S;
  h->flags_for = nullptr;
  h->ws = reinterpret_cast<uint8_t*>(workspace);
  h->gates_valid = false;
  h->off = w;
  h->params = params; h->grads = grads; h->adam_m = adam_m; h->adam_v = adam_v;
  std::vector<VarDev> vd(h->vars.size());
  for (size_t i = 0; i < vd.size(); ++i) vd[i] = VarDev{h->vars[i].offset, h->vars[i].rows, h->vars[i].cols, h->vars[i].ld, h->vars[i].l2};
  MFP_CUDA_OK(cudaMemcpy(h->ws + w.vars, vd.data(), vd.size() * sizeof(VarDev), cudaMemcpyHostToDevice));
  MFP_CUDA_OK(cudaMemset(h->ws + w.ar_sync, 0, 256));
  h->ar_calls = 0;
  return MFP_OK;
}

int mfp_sample_tasks(mfp_engine* h, const int32_t* allowed_host, int32_t n_allowed, uint32_t seed, uint32_t step, int32_t* tasks_out, void* stream) {
  MFP_TRY(check_bound(h));
  if (n_allowed < 1 || n_allowed > 16) { set_error("mfp_sample_tasks: 1..16 task ids"); return MFP_ERR_ARG; }
  TaskSet ts{};
  ts.n = n_allowed;
  for (int i = 0; i < n_allowed; ++i) ts.ids[i] = allowed_host[i];
  h->launches++;
  return launch_sample_tasks(ts, h->B, seed, step, tasks_out, (cudaStream_t)stream, h->doc0);
}

int mfp_mask_corrupt(mfp_engine* h, const mfp_batch* inputs, const int32_t* tasks, uint32_t seed, uint32_t step, void* const* modified_cols,
                     uint8_t* const* masks_out, void* stream) {
  MFP_TRY(check_bound(h));
  ModifiedPtrs out{};
  for (int f = 0; f < h->sc.F; ++f) { out.cols[f] = modified_cols[f]; out.masks[f] = masks_out[f]; }
  h->launches++;
  h->flags_for = h->sc.n_num > 0 ? modified_cols[first_numerical(h->sc)] : nullptr;  // the encoder's row flags come out of the same pass
  return launch_mask_corrupt(h->sc, to_batch(h, inputs, true), tasks, nullptr, h->B, h->S, seed, step, out, (cudaStream_t)stream,
                             wsp<unsigned char>(h, h->off.flags), h->doc0);
}

int mfp_shuffle_inputs(mfp_engine* h, const mfp_batch* inputs, uint32_t seed, uint32_t step, void* const* shuffled_cols, int32_t* perm_out, void* stream) {
  MFP_TRY(check_bound(h));
  if (!inputs || !shuffled_cols) { set_error("mfp_shuffle_inputs: null argument"); return MFP_ERR_ARG; }
  if (any_packed(h)) { set_error("mfp_shuffle_inputs: packed columns are not supported here (expand them first)"); return MFP_ERR_UNSUPPORTED; }
  ModifiedPtrs out{};
  for (int f = 0; f < h->sc.F; ++f) {
    if (!shuffled_cols[f] || shuffled_cols[f] == inputs->cols[f]) { set_error("mfp_shuffle_inputs: output columns must be distinct buffers"); return MFP_ERR_ARG; }
    out.cols[f] = shuffled_cols[f];
  }
  int* perm = wsp<int>(h, h->off.perm);
  h->launches += 2;
  MFP_TRY(launch_shuffle_inputs(h->sc, to_batch(h, inputs), h->B, h->S, seed, step, perm, out, (cudaStream_t)stream, h->cfg.input_dtype == 2, h->doc0));
  if (perm_out) MFP_CUDA_OK(cudaMemcpyAsync(perm_out, perm, (size_t)h->T * sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return MFP_OK;
}

int mfp_mask_for_test(mfp_engine* h, const mfp_batch* inputs, const uint8_t* const* masks, void* const* modified_cols, void* stream) {
  MFP_TRY(check_bound(h));
  ModifiedPtrs out{};
  MaskPtrs tm{};
  for (int f = 0; f < h->sc.F; ++f) { out.cols[f] = modified_cols[f]; tm.m[f] = masks[f]; }
  h->launches++;
  h->flags_for = h->sc.n_num > 0 ? modified_cols[first_numerical(h->sc)] : nullptr;
  return launch_mask_corrupt(h->sc, to_batch(h, inputs, true), nullptr, &tm, h->B, h->S, 0, 0, out, (cudaStream_t)stream, wsp<unsigned char>(h, h->off.flags));
}

int mfp_set_context_ids(mfp_engine* h, const int32_t* task_ids) {
  if (!h) { set_error("mfp_set_context_ids: null engine"); return MFP_ERR_ARG; }
  if (h->cfg.context != 1) { set_error("mfp_set_context_ids: the engine was not created with context = id"); return MFP_ERR_STATE; }
  h->ctx_ids = task_ids;
  return MFP_OK;
}

int mfp_set_packed_rows(mfp_engine* h, const int32_t* const* rowmaps) {
  if (!h) { set_error("mfp_set_packed_rows: null engine"); return MFP_ERR_ARG; }
  for (int f = 0; f < h->sc.F; ++f) {
    const int32_t* m = rowmaps ? rowmaps[f] : nullptr;
    if (m && h->sc.f[f].kind != 1) { set_error("mfp_set_packed_rows: field %d is categorical (only numerical columns pack)", f); return MFP_ERR_ARG; }
    h->rowmaps[f] = m;
  }
  return MFP_OK;
}

int mfp_set_canvas_columns(mfp_engine* h, const int32_t* const* columns) {
  if (!h || !columns) { set_error("mfp_set_canvas_columns: null argument"); return MFP_ERR_ARG; }
  if (h->cfg.context < 3) { set_error("mfp_set_canvas_columns: the engine was not created with context = canvas / canvas_add"); return MFP_ERR_STATE; }
  for (int c = 0; c < h->cfg.n_canvas; ++c) {
    if (!columns[c]) { set_error("mfp_set_canvas_columns: column %d is null", c); return MFP_ERR_ARG; }
    h->canvas_ids[c] = columns[c];
  }
  return MFP_OK;
}

int mfp_forward(mfp_engine* h, const mfp_batch* modified, int32_t training, uint32_t seed, uint32_t step, float* logits_out, void* stream) {
  MFP_TRY(check_bound(h));
  cudaStream_t st = (cudaStream_t)stream;
  const Schema& sc = h->sc;
  const int T = h->T, D = kD, L = h->cfg.num_blocks;
  const BatchPtrs mod = to_batch(h, modified);
  const float* P = h->params;
  unsigned char* flags = wsp<unsigned char>(h, h->off.flags);
  float* x = wsp<float>(h, h->off.x);
  const size_t TD = (size_t)T * D;
  const bool drop = training && h->cfg.dropout > 0.f;
  const uint32_t row0 = h->doc0 * (uint32_t)h->S;  // global index of this batch's first row (dropout counters)
  // FFN 1 also writes its ReLU gates as bits (64 bytes per row instead of the 2 KB of hid the input gradient would re-read as a mask)
  const bool write_gates = h->gemm_impl != 1 && relu_bits_enabled();
  h->gates_valid = write_gates;
  const int gate_ld = (int)gate_rows((size_t)T);

  // ---- encoder (encoder.py:147-199)
  // special-token flags of the numerical fields: already written by the mask/corrupt pass when it produced exactly these
  // columns just before (consumed once: any later forward on the same buffers re-derives them by value)
  const bool have_flags = sc.n_num > 0 && h->flags_for != nullptr && h->flags_for == modified->cols[first_numerical(sc)];
  h->flags_for = nullptr;
  if (!have_flags) { MFP_TRY(launch_row_flags(sc, mod, T, flags, st)); h->launches++; }
  PosEmbed pos{nullptr, 0, 0.f, 0u, 0u, 0, 0u};
  const bool token_ctx = h->cfg.context >= 1 && h->cfg.context <= 3;
  if (h->pos_off >= 0) pos = PosEmbed{P + h->pos_off, h->S, drop ? h->cfg.dropout : 0.f, seed, step, token_ctx ? 1 : 0, h->doc0 * (uint32_t)h->S};
  MFP_TRY(launch_embed_fwd(sc, mod, flags, P, T, x, st, pos));
  h->launches += 1;
  // The first block's LayerNorm 1 rides the epilogue of the LAST numerical field's Dense GEMM (its output rows are the finished encoder
  // sum) when nothing else touches h0 afterwards: pre-LN blocks, no context token / canvas vector, a spec with numerical fields.
  static const bool fuse_ln_env0 = [] { const char* e = getenv("FLEXDM_FUSE_LN"); return !(e && e[0] == '0'); }();
  int last_num = -1;
  for (int f = 0; f < sc.F; ++f)
    if (sc.f[f].kind == 1) last_num = f;
  const bool ln1_in_encoder = fuse_ln_env0 && h->gemm_impl != 1 && h->cfg.block_type == 0 && h->cfg.context == 0 && L > 0 && last_num >= 0;
  for (int f = 0; f < sc.F; ++f) {
    const FieldDev& fd = sc.f[f];
    if (fd.kind != 1) continue;
    GemmEpilogue ep = make_epilogue(x, D);
    ep.residual = x; ep.ldr = D;
    ep.rowflag = flags + (size_t)fd.num_slot * T;
    if (ln1_in_encoder && f == last_num) {
      float* stats0 = wsp<float>(h, h->off.stats);
      ep.ln_out = wsp<float>(h, h->off.ln1); ep.ln_ldo = D; ep.ln_gamma = P + h->blocks[0].g1; ep.ln_beta = P + h->blocks[0].be1;
      ep.ln_mean = stats0; ep.ln_rstd = stats0 + T;
    }
    MFP_TRY(gemm(h, reinterpret_cast<const float*>(mod.cols[f]), 0, fd.C, P + fd.kernel_off, 1, D, T, D, fd.C, ep, 1, st));
  }
  // ---- context token (encoder.py:231-249): one more row per document, attended to by every element
  const int* attn_len = modified->length;
  if (h->cfg.context == 1 || h->cfg.context == 2) {
    const int* ids = h->cfg.context == 1 ? h->ctx_ids : modified->length;
    if (!ids) { set_error("mfp_forward: context = id needs mfp_set_context_ids first"); return MFP_ERR_STATE; }
    int* ctx_row = wsp<int>(h, h->off.ctx_row);
    MFP_TRY(launch_context_token(P + h->ctx_off, h->cfg.context_rows, ids, modified->length, h->B, h->S, x, ctx_row, st, pos));
    h->launches++;
    attn_len = ctx_row;
  } else if (h->cfg.context >= 3) {  // canvas (the token is the sum of the canvas columns' embeddings) / canvas_add (added to every element)
    CanvasArgs ca{};
    MFP_TRY(canvas_args(h, &ca));
    float* vec = wsp<float>(h, h->off.canvas_vec);
    MFP_TRY(launch_canvas_vector(ca, P, h->B, vec, st));
    h->launches++;
    if (h->cfg.context == 3) {
      int* ctx_row = wsp<int>(h, h->off.ctx_row);
      MFP_TRY(launch_iota(wsp<int>(h, h->off.iota), wsp<int>(h, h->off.zeros), h->B, st));
      MFP_TRY(launch_context_token(vec, h->B, wsp<int>(h, h->off.iota), modified->length, h->B, h->S, x, ctx_row, st, pos));  // "table" row b = document b's vector
      h->launches += 2;
      attn_len = ctx_row;
    } else {
      MFP_TRY(launch_add_doc_vector(x, vec, h->B, h->S, st));
      h->launches++;
    }
  }
  // ---- blocks (transformer.py:208-229)
  for (int i = 0; i < L; ++i) {
    const BlockLayout& b = h->blocks[i];
    float* xi = x + i * TD;
    float* xo = x + (i + 1) * TD;
    float* ln1 = wsp<float>(h, h->off.ln1) + i * TD;
    float* qkv = wsp<float>(h, h->off.qkv) + i * 3 * TD;
    float* attn = wsp<float>(h, h->off.attn) + i * TD;
    float* xmid = wsp<float>(h, h->off.xmid) + i * TD;
    float* ln2 = wsp<float>(h, h->off.ln2) + i * TD;
    float* hid = wsp<float>(h, h->off.hid) + (size_t)i * T * kF;
    float* stats = wsp<float>(h, h->off.stats) + (size_t)i * 4 * T;
    float* lse = wsp<float>(h, h->off.lse) + (size_t)i * h->B * kH * h->S;

    if (h->cfg.block_type == 1) {
      // post-LayerNorm TransformerBlock (transformer.py:187-205): z1 = x + drop(attn(x)); x1 = LN1(z1); z2 = x1 + drop(mlp(x1));
      // out = LN2(z2).  Same buffers under other roles: xmid holds z1, ln2 holds x1, ln1 holds z2.
      GemmEpilogue q1 = make_epilogue(qkv, 3 * D);
      q1.bias = P + b.bqkv;
      MFP_TRY(gemm(h, xi, 0, D, P + b.wqkv, 1, 3 * D, T, 3 * D, D, q1, 1, st));
      {
        ProfScope prof(h, MFP_PROFILE_ATTENTION, st, 4.0 * T * (3.0 * D + D) + 4.0 * h->B * kH * h->S);
        if (h->gemm_impl == 0 && h->S <= 128) MFP_TRY(launch_attention_fwd_tc(h->maps, qkv, attn_len, h->B, h->S, attn, lse, st));
        else MFP_TRY(launch_attention_fwd(qkv, attn_len, h->B, h->S, attn, lse, st));
      }
      GemmEpilogue q2 = make_epilogue(xmid, D);
      q2.bias = P + b.bo;
      q2.residual = xi; q2.ldr = D;
      if (drop) { q2.drop_enabled = 1; q2.drop_rate = h->cfg.dropout; q2.drop_seed = seed; q2.drop_step = step; q2.drop_site = kSiteDropout + 2 * i; q2.drop_row0 = row0; }
      MFP_TRY(gemm(h, attn, 0, D, P + b.wo, 1, D, T, D, D, q2, 1, st));
      MFP_TRY(launch_layernorm_fwd(xmid, P + b.g1, P + b.be1, T, ln2, stats, stats + T, st));
      GemmEpilogue q3 = make_epilogue(hid, kF);
      q3.bias = P + b.b1;
      q3.relu = 1;
      if (write_gates) { q3.relu_bits_out = wsp<uint32_t>(h, h->off.gates) + (size_t)i * (kF / 32) * gate_ld; q3.relu_bits_ld = gate_ld; }
      MFP_TRY(gemm(h, ln2, 0, D, P + b.w1, 1, kF, T, kF, D, q3, 1, st));
      GemmEpilogue q4 = make_epilogue(ln1, D);
      q4.bias = P + b.b2;
      q4.residual = ln2; q4.ldr = D;
      if (drop) { q4.drop_enabled = 1; q4.drop_rate = h->cfg.dropout; q4.drop_seed = seed; q4.drop_step = step; q4.drop_site = kSiteDropout + 2 * i + 1; q4.drop_row0 = row0; }
      MFP_TRY(gemm(h, hid, 0, kF, P + b.w2, 1, D, T, D, kF, q4, 1, st));
      MFP_TRY(launch_layernorm_fwd(ln1, P + b.g2, P + b.be2, T, xo, stats + 2 * T, stats + 3 * T, st));
      h->launches += 3;
      continue;
    }
    // LayerNorms of the pre-LN block (transformer.py:216,222) ride the epilogue of the GEMM that produces their input rows (N = 256: an
    // output tile holds whole rows): LN2 in the attention output projection, the next block's LN1 in FFN 2.  Only the first block's LN1
    // (input = the encoder's sum of embeddings) and the SIMT bring-up path run the standalone kernel.
    static const bool fuse_ln_env = [] { const char* e = getenv("FLEXDM_FUSE_LN"); return !(e && e[0] == '0'); }();  // A/B switch
    const bool fuse_ln = h->gemm_impl != 1 && fuse_ln_env;
    if ((i == 0 && !ln1_in_encoder) || !fuse_ln) { MFP_TRY(launch_layernorm_fwd(xi, P + b.g1, P + b.be1, T, ln1, stats, stats + T, st)); h->launches++; }
    GemmEpilogue e1 = make_epilogue(qkv, 3 * D);
    e1.bias = P + b.bqkv;
    MFP_TRY(gemm(h, ln1, 0, D, P + b.wqkv, 1, 3 * D, T, 3 * D, D, e1, 1, st));
    {
      ProfScope prof(h, MFP_PROFILE_ATTENTION, st, 4.0 * T * (3.0 * D + D) + 4.0 * h->B * kH * h->S);  // qkv in, out + lse
      if (h->gemm_impl == 0 && h->S <= 128) MFP_TRY(launch_attention_fwd_tc(h->maps, qkv, attn_len, h->B, h->S, attn, lse, st));
      else MFP_TRY(launch_attention_fwd(qkv, attn_len, h->B, h->S, attn, lse, st));
    }
    GemmEpilogue e2 = make_epilogue(xmid, D);
    e2.bias = P + b.bo;
    e2.residual = xi; e2.ldr = D;
    if (drop) { e2.drop_enabled = 1; e2.drop_rate = h->cfg.dropout; e2.drop_seed = seed; e2.drop_step = step; e2.drop_site = kSiteDropout + 2 * i; e2.drop_row0 = row0; }
    if (fuse_ln) { e2.ln_out = ln2; e2.ln_ldo = D; e2.ln_gamma = P + b.g2; e2.ln_beta = P + b.be2; e2.ln_mean = stats + 2 * T; e2.ln_rstd = stats + 3 * T; }
    MFP_TRY(gemm(h, attn, 0, D, P + b.wo, 1, D, T, D, D, e2, 1, st));
    if (!fuse_ln) { MFP_TRY(launch_layernorm_fwd(xmid, P + b.g2, P + b.be2, T, ln2, stats + 2 * T, stats + 3 * T, st)); h->launches++; }
    GemmEpilogue e3 = make_epilogue(hid, kF);
    e3.bias = P + b.b1;
    e3.relu = 1;
    if (write_gates) { e3.relu_bits_out = wsp<uint32_t>(h, h->off.gates) + (size_t)i * (kF / 32) * gate_ld; e3.relu_bits_ld = gate_ld; }
    MFP_TRY(gemm(h, ln2, 0, D, P + b.w1, 1, kF, T, kF, D, e3, 1, st));
    GemmEpilogue e4 = make_epilogue(xo, D);
    e4.bias = P + b.b2;
    e4.residual = xmid; e4.ldr = D;
    if (drop) { e4.drop_enabled = 1; e4.drop_rate = h->cfg.dropout; e4.drop_seed = seed; e4.drop_step = step; e4.drop_site = kSiteDropout + 2 * i + 1; e4.drop_row0 = row0; }
    if (fuse_ln && i + 1 < L) {  // the next block's LN1
      const BlockLayout& nb = h->blocks[i + 1];
      float* nstats = wsp<float>(h, h->off.stats) + (size_t)(i + 1) * 4 * T;
      e4.ln_out = wsp<float>(h, h->off.ln1) + (i + 1) * TD; e4.ln_ldo = D; e4.ln_gamma = P + nb.g1; e4.ln_beta = P + nb.be1; e4.ln_mean = nstats; e4.ln_rstd = nstats + T;
    }
    MFP_TRY(gemm(h, hid, 0, kF, P + b.w2, 1, D, T, D, kF, e4, 1, st));
    h->launches += 1;  // attention
  }
  // ---- decoder heads (decoder.py:72-111)
  float* logits = wsp<float>(h, h->off.logits);
  GemmEpilogue eh = make_epilogue(logits, sc.LW);
  eh.bias = P + h->bh;
  MFP_TRY(gemm(h, x + L * TD, 0, D, P + h->wh, 1, sc.LW, T, sc.LW, D, eh, 1, st));
  if (logits_out) MFP_CUDA_OK(cudaMemcpyAsync(logits_out, logits, (size_t)T * sc.LW * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return MFP_OK;
}

int mfp_loss(mfp_engine* h, const mfp_batch* targets, const uint8_t* const* masks, const uint8_t* sort_flag, const int32_t* sort_tasks,
             const float* logits_in, float inv_batch, int32_t compute_grad, float* metrics_out, void* stream) {
  MFP_TRY(check_bound(h));
  if (!metrics_out) { set_error("mfp_loss: metrics_out is required"); return MFP_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const BatchPtrs tg = to_batch(h, targets, true);
  MaskPtrs mp{};
  for (int f = 0; f < h->sc.F; ++f) mp.m[f] = masks[f];
  LossBuffers buf{wsp<float>(h, h->off.part), wsp<int>(h, h->off.idx_true), wsp<int>(h, h->off.idx_pred)};
  const float* logits = logits_in ? logits_in : wsp<float>(h, h->off.logits);
  const int use_sort = (h->cfg.sort_pos && (sort_flag || sort_tasks)) ? 1 : 0;
  if (use_sort) {
    MFP_TRY(launch_sort_indices(h->sc, tg, logits, sort_flag, sort_tasks, h->cfg.pos_task_id, h->B, h->S, buf, st));
    h->launches++;
  }
  h->launches += 3;
  return launch_loss(h->sc, tg, mp, logits, use_sort, h->B, h->S, inv_batch, compute_grad ? wsp<float>(h, h->off.dlogits) : nullptr, buf, metrics_out, st);
}

int32_t mfp_backward_num_stages(const mfp_engine* h) { return h ? h->cfg.num_blocks + 2 : 0; }

int mfp_backward_stage_range(const mfp_engine* h, int32_t stage, int64_t* lo, int64_t* hi) {
  if (!h || !lo || !hi || stage < 0 || stage >= h->cfg.num_blocks + 2) { set_error("mfp_backward_stage_range: bad argument"); return MFP_ERR_ARG; }
  *lo = h->stage_lo[stage];
  *hi = h->stage_hi[stage];
  return MFP_OK;
}

int mfp_backward(mfp_engine* h, const mfp_batch* modified, int32_t training, uint32_t seed, uint32_t step, void* stream) {
  return mfp_backward_stages(h, modified, training, seed, step, 0, h ? h->cfg.num_blocks + 1 : 0, stream);
}

int mfp_backward_stages(mfp_engine* h, const mfp_batch* modified, int32_t training, uint32_t seed, uint32_t step, int32_t first_stage,
                        int32_t last_stage, void* stream) {
  MFP_TRY(check_bound(h));
  if (!h->grads) { set_error("mfp_backward: no gradient buffer bound"); return MFP_ERR_STATE; }
  if (first_stage < 0 || last_stage > h->cfg.num_blocks + 1 || first_stage > last_stage) { set_error("mfp_backward_stages: bad stage range"); return MFP_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const Schema& sc = h->sc;
  const int T = h->T, D = kD, L = h->cfg.num_blocks;
  const BatchPtrs mod = to_batch(h, modified);
  const float* P = h->params;
  float* G = h->grads;
  const size_t TD = (size_t)T * D;
  const bool drop = training && h->cfg.dropout > 0.f;
  const uint32_t row0 = h->doc0 * (uint32_t)h->S;
  float* ln_det = h->deterministic ? wsp<float>(h, h->off.ln_part) : nullptr;
  float* x = wsp<float>(h, h->off.x);
  // --context: the forward pass left every document's context-token row in the workspace; it is also the attention length array
  const int* ctx_row = (h->cfg.context >= 1 && h->cfg.context <= 3) ? wsp<int>(h, h->off.ctx_row) : nullptr;
  const int* attn_len = ctx_row ? ctx_row : modified->length;
  float* dlogits = wsp<float>(h, h->off.dlogits);
  float* dx = wsp<float>(h, h->off.dx);
  float* dtmp = wsp<float>(h, h->off.dtmp);
  float* dyb = wsp<float>(h, h->off.dy);
  float* dqkv = wsp<float>(h, h->off.dqkv);
  float* dhid = wsp<float>(h, h->off.dhid);
  float* dattn = wsp<float>(h, h->off.dattn);
  // the FFN's ReLU gates as the forward's FFN 1 epilogue left them (bits), when it did and this path can read them
  const bool use_gates = h->gates_valid && h->gemm_impl != 1 && relu_bits_enabled();
  const int gate_ld = (int)gate_rows((size_t)T);

  if (first_stage == 0) {
    MFP_CUDA_OK(cudaMemsetAsync(G, 0, (size_t)h->param_count * sizeof(float), st));
    // ---- heads: dX = dlogits . Wh^T ; dWh = X^T . dlogits ; dbh = colsum(dlogits)
    const GemmArgs gh[2] = {
        {x + L * TD, 1, D, dlogits, 1, sc.LW, D, sc.LW, T, make_epilogue(G + h->wh, sc.LW), wgrad_splits(D, sc.LW, T), G + h->bh},
        {dlogits, 0, sc.LW, P + h->wh, 0, sc.LW, T, D, sc.LW, make_epilogue(dx, D), 1, nullptr}};
    MFP_TRY(gemm_group(h, gh, 2, st));
  }
  for (int i = L - 1; i >= 0; --i) {
    if (L - i < first_stage || L - i > last_stage) continue;  // stage L - i
    const BlockLayout& b = h->blocks[i];
    const float* xi = x + i * TD;
    const float* ln1 = wsp<float>(h, h->off.ln1) + i * TD;
    const float* qkv = wsp<float>(h, h->off.qkv) + i * 3 * TD;
    const float* attn = wsp<float>(h, h->off.attn) + i * TD;
    const float* xmid = wsp<float>(h, h->off.xmid) + i * TD;
    const float* ln2 = wsp<float>(h, h->off.ln2) + i * TD;
    const float* hid = wsp<float>(h, h->off.hid) + (size_t)i * T * kF;
    const uint32_t* gates = wsp<uint32_t>(h, h->off.gates) + (size_t)i * (kF / 32) * gate_ld;
    const float* stats = wsp<float>(h, h->off.stats) + (size_t)i * 4 * T;
    const float* lse = wsp<float>(h, h->off.lse) + (size_t)i * h->B * kH * h->S;

    if (h->cfg.block_type == 1) {
      // post-LayerNorm block, backward of the wiring above.  dx = d(out).
      //   dz2 = LN2'(dx);  FFN branch on drop2(dz2);  dx1 = dz2 + d(mlp input);  dz1 = LN1'(dx1);  attention branch on drop1(dz1);
      //   d(block input) = dz1 + d(QKV input)
      const float* z2 = ln1;
      const float* x1 = ln2;
      MFP_TRY(launch_layernorm_bwd(z2, dx, P + b.g2, stats + 2 * T, stats + 3 * T, nullptr, T, dx, G + b.g2, G + b.be2, st, nullptr, 0, nullptr,
                                   drop ? dyb : nullptr, h->cfg.dropout, seed, step, kSiteDropout + 2 * i + 1, row0, ln_det));
      const float* dy2 = drop ? dyb : dx;
      MFP_TRY(gemm(h, hid, 1, kF, dy2, 1, D, kF, D, T, make_epilogue(G + b.w2, D), wgrad_splits(kF, D, T), st, G + b.b2));
      GemmEpilogue ph = make_epilogue(dhid, kF);
      if (use_gates) { ph.relu_bits = gates; ph.relu_bits_ld = gate_ld; }
      else { ph.relu_src = hid; ph.ld_relu = kF; }
      MFP_TRY(gemm(h, dy2, 0, D, P + b.w2, 0, D, T, kF, D, ph, 1, st));
      MFP_TRY(gemm(h, x1, 1, D, dhid, 1, kF, D, kF, T, make_epilogue(G + b.w1, kF), wgrad_splits(D, kF, T), st, G + b.b1));
      GemmEpilogue p1 = make_epilogue(dx, D);  // dx1 = dz2 (in dx) + dhid . W1^T, in place
      p1.residual = dx; p1.ldr = D;
      MFP_TRY(gemm(h, dhid, 0, kF, P + b.w1, 0, kF, T, D, kF, p1, 1, st));
      MFP_TRY(launch_layernorm_bwd(xmid, dx, P + b.g1, stats, stats + T, nullptr, T, dx, G + b.g1, G + b.be1, st, nullptr, 0, nullptr,
                                   drop ? dyb : nullptr, h->cfg.dropout, seed, step, kSiteDropout + 2 * i, row0, ln_det));
      const float* dy1 = drop ? dyb : dx;
      MFP_TRY(gemm(h, attn, 1, D, dy1, 1, D, D, D, T, make_epilogue(G + b.wo, D), wgrad_splits(D, D, T), st, G + b.bo));
      MFP_TRY(gemm(h, dy1, 0, D, P + b.wo, 0, D, T, D, D, make_epilogue(dattn, D), 1, st));
      {
        ProfScope prof(h, MFP_PROFILE_ATTENTION, st, 4.0 * T * (3.0 * D + D + D + 3.0 * D) + 4.0 * h->B * kH * h->S);
        if (h->gemm_impl == 0 && h->S <= 128) MFP_TRY(launch_attention_bwd_tc(h->maps, qkv, attn, lse, dattn, attn_len, h->B, h->S, dqkv, st));
        else MFP_TRY(launch_attention_bwd(qkv, attn, lse, dattn, attn_len, h->B, h->S, dqkv, st));
      }
      MFP_TRY(gemm(h, xi, 1, D, dqkv, 1, 3 * D, D, 3 * D, T, make_epilogue(G + b.wqkv, 3 * D), wgrad_splits(D, 3 * D, T), st, G + b.bqkv));
      GemmEpilogue p2 = make_epilogue(dx, D);  // d(block input) = dz1 (in dx) + dqkv . Wqkv^T, in place
      p2.residual = dx; p2.ldr = D;
      MFP_TRY(gemm(h, dqkv, 0, 3 * D, P + b.wqkv, 0, 3 * D, T, D, 3 * D, p2, 1, st));
      if (i == 0 && sc.n_num > 0) {
        // the encoder's Dense wgrads want dh0 with the rows of special-token elements zeroed, one copy per numerical field
        MFP_TRY(launch_masked_copies(dx, wsp<unsigned char>(h, h->off.flags), sc.n_num, T, wsp<float>(h, h->off.dh0m), st));
        h->launches++;
      }
      h->launches += 3;
      continue;
    }
    // FFN branch: x_out = xmid + drop(relu(ln2.W1 + b1).W2 + b2)
    // dy = dx under the FFN branch's dropout mask: written by the LayerNorm backward of the block above (below: by the
    // separate kernel for the topmost block, whose dx comes out of the heads' dgrad GEMM)
    const float* dy = dx;
    if (drop) {
      if (i == L - 1) {
        MFP_TRY(launch_dropout_bwd(dx, T, h->cfg.dropout, seed, step, kSiteDropout + 2 * i + 1, dyb, st, row0));
        h->launches++;
      }
      dy = dyb;
    }
    // every weight gradient shares its launch with the input gradient that hangs off the same dY (gemm_group)
    GemmEpilogue eh = make_epilogue(dhid, kF);
    if (use_gates) { eh.relu_bits = gates; eh.relu_bits_ld = gate_ld; }
    else { eh.relu_src = hid; eh.ld_relu = kF; }
    const GemmArgs g1[2] = {{hid, 1, kF, dy, 1, D, kF, D, T, make_epilogue(G + b.w2, D), wgrad_splits(kF, D, T), G + b.b2},
                            {dy, 0, D, P + b.w2, 0, D, T, kF, D, eh, 1, nullptr}};
    MFP_TRY(gemm_group(h, g1, 2, st));
    const GemmArgs g2[2] = {{ln2, 1, D, dhid, 1, kF, D, kF, T, make_epilogue(G + b.w1, kF), wgrad_splits(D, kF, T), G + b.b1},
                            {dhid, 0, kF, P + b.w1, 0, kF, T, D, kF, make_epilogue(dtmp, D), 1, nullptr}};
    MFP_TRY(gemm_group(h, g2, 2, st));
    MFP_TRY(launch_layernorm_bwd(xmid, dtmp, P + b.g2, stats + 2 * T, stats + 3 * T, dx, T, dx, G + b.g2, G + b.be2, st, nullptr, 0, nullptr,
                                 drop ? dyb : nullptr, h->cfg.dropout, seed, step, kSiteDropout + 2 * i, row0, ln_det));
    // attention branch: xmid = x_in + drop(attn.Wo + bo)
    dy = drop ? dyb : dx;
    const GemmArgs g3[2] = {{attn, 1, D, dy, 1, D, D, D, T, make_epilogue(G + b.wo, D), wgrad_splits(D, D, T), G + b.bo},
                            {dy, 0, D, P + b.wo, 0, D, T, D, D, make_epilogue(dattn, D), 1, nullptr}};
    MFP_TRY(gemm_group(h, g3, 2, st));
    {
      ProfScope prof(h, MFP_PROFILE_ATTENTION, st, 4.0 * T * (3.0 * D + D + D + 3.0 * D) + 4.0 * h->B * kH * h->S);  // qkv, out, dout in; dqkv out
      if (h->gemm_impl == 0 && h->S <= 128) MFP_TRY(launch_attention_bwd_tc(h->maps, qkv, attn, lse, dattn, attn_len, h->B, h->S, dqkv, st));
      else MFP_TRY(launch_attention_bwd(qkv, attn, lse, dattn, attn_len, h->B, h->S, dqkv, st));
    }
    const GemmArgs g4[2] = {{ln1, 1, D, dqkv, 1, 3 * D, D, 3 * D, T, make_epilogue(G + b.wqkv, 3 * D), wgrad_splits(D, 3 * D, T), G + b.bqkv},
                            {dqkv, 0, 3 * D, P + b.wqkv, 0, 3 * D, T, D, 3 * D, make_epilogue(dtmp, D), 1, nullptr}};
    MFP_TRY(gemm_group(h, g4, 2, st));
    if (i == 0)
      MFP_TRY(launch_layernorm_bwd(xi, dtmp, P + b.g1, stats, stats + T, dx, T, dx, G + b.g1, G + b.be1, st, wsp<unsigned char>(h, h->off.flags), sc.n_num,
                                   sc.n_num > 0 ? wsp<float>(h, h->off.dh0m) : nullptr, nullptr, 0.f, 0u, 0u, 0u, 0u, ln_det));
    else
      MFP_TRY(launch_layernorm_bwd(xi, dtmp, P + b.g1, stats, stats + T, dx, T, dx, G + b.g1, G + b.be1, st, nullptr, 0, nullptr, drop ? dyb : nullptr,
                                   h->cfg.dropout, seed, step, kSiteDropout + 2 * (i - 1) + 1, row0, ln_det));
    h->launches += 3;
  }
  if (last_stage < L + 1) return MFP_OK;
  // ---- encoder: table / special / bias rows by a one-hot wgrad GEMM; Dense kernels by X^T . (dh0 with special-token rows zeroed) (encoder.cu)
  const unsigned char* flags = wsp<unsigned char>(h, h->off.flags);
  float* onehot = wsp<float>(h, h->off.onehot);
  float* rowgrad = wsp<float>(h, h->off.rowgrad);
  MFP_TRY(launch_embed_onehot(sc, mod, flags, T, onehot, st, ctx_row, h->S));
  MFP_CUDA_OK(cudaMemsetAsync(rowgrad, 0, (size_t)sc.Rp * D * sizeof(float), st));
  {
    GemmArgs ge[3];
    int ng = 0;
    ge[ng++] = GemmArgs{onehot, 1, sc.Rp, dx, 1, D, sc.R, D, T, make_epilogue(rowgrad, D), wgrad_splits(sc.R, D, T), nullptr};
    for (int f = 0; f < sc.F; ++f) {
      const FieldDev& fd = sc.f[f];
      if (fd.kind != 1) continue;
      if (ng == 3) { MFP_TRY(gemm_group(h, ge, ng, st)); ng = 0; }
      ge[ng++] = GemmArgs{reinterpret_cast<const float*>(mod.cols[f]), 1, fd.C, wsp<float>(h, h->off.dh0m) + (size_t)fd.num_slot * TD, 1, D, fd.C, D, T,
                          make_epilogue(G + fd.kernel_off, D), wgrad_splits(fd.C, D, T), nullptr};
    }
    MFP_TRY(gemm_group(h, ge, ng, st));
  }
  MFP_TRY(launch_embed_scatter(sc, rowgrad, G, st));
  h->launches += 2;
  if (h->cfg.context == 1 || h->cfg.context == 2) {  // d(context table): the token rows of dh0, summed per id (the one-hot rows of those positions are zero)
    const int* ids = h->cfg.context == 1 ? h->ctx_ids : modified->length;
    if (!ids) { set_error("mfp_backward: context = id needs mfp_set_context_ids first"); return MFP_ERR_STATE; }
    MFP_TRY(launch_context_token_bwd(dx, ids, ctx_row, h->cfg.context_rows, h->B, h->S, G + h->ctx_off, st));
    h->launches++;
  } else if (h->cfg.context >= 3) {
    // d(canvas vector)[b] = the token's row of dh0 (canvas) or the sum of the document's rows (canvas_add); then every canvas table gets
    // the vectors of the documents that picked each of its rows
    CanvasArgs ca{};
    MFP_TRY(canvas_args(h, &ca));
    float* dvec = wsp<float>(h, h->off.dcanvas);
    const int* iota = wsp<int>(h, h->off.iota);
    const int* zeros = wsp<int>(h, h->off.zeros);
    if (h->cfg.context == 3) {
      MFP_TRY(launch_context_token_bwd(dx, iota, ctx_row, h->B, h->B, h->S, dvec, st));
    } else {
      MFP_TRY(launch_iota(wsp<int>(h, h->off.iota), wsp<int>(h, h->off.zeros), h->B, st));
      MFP_TRY(launch_sum_doc_rows(dx, h->B, h->S, dvec, st));
      h->launches++;
    }
    for (int c = 0; c < ca.n; ++c) MFP_TRY(launch_context_token_bwd(dvec, ca.ids[c], zeros, ca.rows[c], h->B, 1, G + ca.off[c], st));
    h->launches += 1 + ca.n;
  }
  if (h->pos_off >= 0) {  // rows >= S (S + 1 with a context token) of the table get no gradient (G was cleared in stage 0)
    if (ctx_row) MFP_TRY(launch_pos_embed_bwd_ctx(dx, ctx_row, h->B, h->S, drop ? h->cfg.dropout : 0.f, seed, step, G + h->pos_off, st, row0));
    else MFP_TRY(launch_pos_embed_bwd(dx, h->B, h->S, drop ? h->cfg.dropout : 0.f, seed, step, G + h->pos_off, st, row0));
    h->launches++;
  }
  return MFP_OK;
}

int mfp_optimizer_step(mfp_engine* h, int32_t t, float learning_rate, float clipnorm, float* l2_loss_out, void* stream) {
  MFP_TRY(check_bound(h));
  if (!h->grads || !h->adam_m || !h->adam_v) { set_error("mfp_optimizer_step: gradient / Adam state buffers are not bound"); return MFP_ERR_STATE; }
  if (t < 1) { set_error("mfp_optimizer_step: t is 1-based"); return MFP_ERR_ARG; }
  h->launches += l2_loss_out ? 3 : 2;
  return launch_optimizer(wsp<VarDev>(h, h->off.vars), (int)h->vars.size(), h->params, h->grads, h->adam_m, h->adam_v, wsp<float>(h, h->off.norms), t,
                          learning_rate, clipnorm, h->cfg.l2, l2_loss_out, (cudaStream_t)stream);
}

int mfp_regularization_loss(mfp_engine* h, float* l2_loss_out, void* stream) {
  MFP_TRY(check_bound(h));
  if (!l2_loss_out) { set_error("mfp_regularization_loss: null output"); return MFP_ERR_ARG; }
  h->launches += 2;
  return launch_regularization_loss(wsp<VarDev>(h, h->off.vars), (int)h->vars.size(), h->params, wsp<float>(h, h->off.norms), h->cfg.l2, l2_loss_out,
                                    (cudaStream_t)stream);
}

int mfp_merge_prediction(mfp_engine* h, int32_t field, const void* input_col, const uint8_t* mask, const float* logits_in, float* out, void* stream) {
  MFP_TRY(check_bound(h));
  if (field < 0 || field >= h->sc.F) { set_error("mfp_merge_prediction: bad field"); return MFP_ERR_ARG; }
  h->launches++;
  return launch_merge_prediction(h->sc, field, input_col, mask, logits_in ? logits_in : wsp<float>(h, h->off.logits), h->T, out, (cudaStream_t)stream);
}

int64_t mfp_launch_count(const mfp_engine* h) { return h ? h->launches : 0; }

int mfp_set_gemm_impl(mfp_engine* h, int32_t impl) {
  if (!h || impl < 0 || impl > 2) { set_error("mfp_set_gemm_impl: bad argument"); return MFP_ERR_ARG; }
  h->gemm_impl = impl;
  return MFP_OK;
}

int mfp_set_deterministic(mfp_engine* h, int32_t on) {
  if (!h) { set_error("mfp_set_deterministic: null engine"); return MFP_ERR_ARG; }
  h->deterministic = on ? 1 : 0;
  return MFP_OK;
}

int mfp_set_doc_offset(mfp_engine* h, int64_t first_document) {
  if (!h || first_document < 0 || first_document > 0x7fffffffLL) { set_error("mfp_set_doc_offset: bad argument"); return MFP_ERR_ARG; }
  h->doc0 = (uint32_t)first_document;
  return MFP_OK;
}

int mfp_allreduce_gradients_nvls(mfp_engine* h, float* multicast_grads, void* const* signal_pads_dev, int32_t first_slot, int32_t rank, int32_t world,
                                 uint32_t call, void* stream) {
  MFP_TRY(check_bound(h));
  if (!h->grads || !multicast_grads || !signal_pads_dev) { set_error("mfp_allreduce_gradients_nvls: null argument"); return MFP_ERR_ARG; }
  if (call == 0) { set_error("mfp_allreduce_gradients_nvls: call numbers start at 1"); return MFP_ERR_ARG; }
  h->launches++;
  h->ar_calls++;
  return launch_nvls_allreduce(multicast_grads, reinterpret_cast<uint32_t* const*>(signal_pads_dev), first_slot, rank, world, (size_t)h->param_count, call,
                               h->ar_calls, wsp<uint32_t>(h, h->off.ar_sync), (cudaStream_t)stream);
}

int mfp_profile_begin(mfp_engine* h) {
  if (!h) { set_error("null engine"); return MFP_ERR_ARG; }
  for (auto& v : h->prof_events) { for (cudaEvent_t e : v) cudaEventDestroy(e); v.clear(); }
  for (double& b : h->prof_bytes) b = 0.0;
  h->profiling = true;
  return MFP_OK;
}

int mfp_profile_end(mfp_engine* h, float* ms_per_class_host, int32_t* launches_per_class_host, double* bytes_per_class_host) {
  if (!h || !ms_per_class_host || !launches_per_class_host || !bytes_per_class_host) { set_error("mfp_profile_end: null argument"); return MFP_ERR_ARG; }
  h->profiling = false;
  MFP_CUDA_OK(cudaDeviceSynchronize());
  for (int c = 0; c < MFP_PROFILE_CLASSES; ++c) {
    float total = 0.f;
    auto& v = h->prof_events[c];
    for (size_t i = 0; i + 1 < v.size(); i += 2) {
      float ms = 0.f;
      MFP_CUDA_OK(cudaEventElapsedTime(&ms, v[i], v[i + 1]));
      total += ms;
    }
    ms_per_class_host[c] = total;
    launches_per_class_host[c] = (int32_t)(v.size() / 2);
    bytes_per_class_host[c] = h->prof_bytes[c];
    for (cudaEvent_t e : v) cudaEventDestroy(e);
    v.clear();
  }
  return MFP_OK;
}

int mfp_debug_attention(const float* qkv, const int32_t* length, int32_t B, int32_t S, float* out, float* lse, int32_t impl, void* stream) {
  static TensorMapCache* cache = tensor_map_cache_create();
  if (!qkv || !length || !out || !lse || B < 1 || S < 1) { set_error("mfp_debug_attention: bad argument"); return MFP_ERR_ARG; }
  if (impl == 0) return launch_attention_fwd_tc(cache, qkv, length, B, S, out, lse, (cudaStream_t)stream);
  return launch_attention_fwd(qkv, length, B, S, out, lse, (cudaStream_t)stream);
}

int mfp_debug_attention_bwd(const float* qkv, const int32_t* length, int32_t B, int32_t S, const float* out, const float* lse, const float* dout,
                            float* dqkv, int32_t impl, void* stream) {
  static TensorMapCache* cache = tensor_map_cache_create();
  if (!qkv || !length || !out || !lse || !dout || !dqkv || B < 1 || S < 1) { set_error("mfp_debug_attention_bwd: bad argument"); return MFP_ERR_ARG; }
  if (impl == 0) return launch_attention_bwd_tc(cache, qkv, out, lse, dout, length, B, S, dqkv, (cudaStream_t)stream);
  return launch_attention_bwd(qkv, out, lse, dout, length, B, S, dqkv, (cudaStream_t)stream);
}

int mfp_debug_gemm(const float* A, int32_t a_mn, int32_t lda, const float* B, int32_t b_mn, int32_t ldb, float* D, int32_t ldd, int32_t M, int32_t N,
                   int32_t K, const float* bias, int32_t relu, const float* residual, const float* relu_src, float* colsum, int32_t splits, int32_t impl,
                   void* stream) {
  static TensorMapCache* cache = tensor_map_cache_create();
  GemmCall c{};
  c.a = GemmOperand{A, a_mn, lda};
  c.b = GemmOperand{B, b_mn, ldb};
  c.M = M; c.N = N; c.K = K;
  c.splits = splits;
  c.ep = make_epilogue(D, ldd);
  c.ep.bias = bias;
  c.ep.relu = relu;
  c.ep.residual = residual; c.ep.ldr = ldd;
  c.ep.relu_src = relu_src; c.ep.ld_relu = ldd;
  c.colsum = (impl == 0) ? colsum : nullptr;
  if (colsum && impl != 0) MFP_TRY(launch_colsum(B, K, N, ldb, colsum, (cudaStream_t)stream));
  return launch_gemm(cache, c, impl, (cudaStream_t)stream);
}

}  // extern "C"
