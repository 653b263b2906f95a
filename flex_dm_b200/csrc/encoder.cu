// Per-element multi-attribute embedding + additive fusion (reference: architecture/encoder.py:147-199) and its
// backward.  The numerical fields' Dense (512 -> D) runs on the tensor cores (gemm.cu); this file produces
// everything else of h0 in one pass: categorical gather-sum over sub-targets, <MASK>/<UNUSED> special rows
// and the Dense bias of unflagged rows.  The backward is a one-hot GEMM (see below).
#include "gemm.cuh"
#include "kernels.cuh"

namespace mfp {

constexpr int kTokPerCta = 8;

// One warp per element: lane l first resolves lookup l (a categorical sub-target's table row, or a numerical field's
// <MASK> / <UNUSED> / bias row) to a parameter offset, then the warp streams the rows -- independent 1 KB reads, two
// float4 per lane -- and accumulates them in the reference's order (fields in column order, sub-targets in order).
__global__ void __launch_bounds__(32 * kTokPerCta) embed_fwd_kernel(const __grid_constant__ Schema sc, const __grid_constant__ BatchPtrs mod,
                                                                    const unsigned char* __restrict__ flags, const float* __restrict__ params, int T,
                                                                    float* __restrict__ h0, const PosEmbed pos) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * kTokPerCta + (threadIdx.x >> 5);
  if (t >= T) return;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  for (int base = 0; base < sc.n_lookups; base += 32) {
    long long my_off = 0;
    const int l = base + lane;
    if (l < sc.n_lookups) {
      const int f = sc.lk_field[l], c = sc.lk_sub[l];
      const FieldDev& fd = sc.f[f];
      if (fd.kind == 0) {
        int i = __ldg(reinterpret_cast<const int*>(mod.cols[f]) + (size_t)t * fd.C + c);
        i = min(max(i, 0), fd.input_dim + 1);
        my_off = fd.table_off + (long long)i * kD;  // encoder.py:157-160
      } else {
        const int flag = flags[(size_t)fd.num_slot * T + t];
        my_off = flag ? fd.table_off + (long long)(flag - 1) * kD  // encoder.py:167-175
                      : fd.bias_off;                               // Dense bias; x.W is added by the GEMM
      }
    }
    const int cnt = min(32, sc.n_lookups - base);
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const long long off = __shfl_sync(0xffffffffu, my_off, j);
      const float4* row = reinterpret_cast<const float4*>(params + off);
      const float4 r0 = __ldg(row + lane), r1 = __ldg(row + 32 + lane);
      a0.x += r0.x; a0.y += r0.y; a0.z += r0.z; a0.w += r0.w;
      a1.x += r1.x; a1.y += r1.y; a1.z += r1.z; a1.w += r1.w;
    }
  }
  if (pos.table) {  // seq += PositionEmbedding(...): row of position s, under its own dropout site when training (transformer.py:24-30)
    const float4* row = reinterpret_cast<const float4*>(pos.table + (size_t)(t % pos.S + pos.shift) * kD);
    const float4 p0 = __ldg(row + lane), p1 = __ldg(row + 32 + lane);
    float va[4] = {p0.x, p0.y, p0.z, p0.w}, vb[4] = {p1.x, p1.y, p1.z, p1.w};
    if (pos.rate > 0.f) {
      dropout4(va, ((uint32_t)t + pos.row0) * kD + 4u * lane, pos.rate, pos.seed, pos.step, kSitePosDropout);
      dropout4(vb, ((uint32_t)t + pos.row0) * kD + 128u + 4u * lane, pos.rate, pos.seed, pos.step, kSitePosDropout);
    }
    a0.x += va[0]; a0.y += va[1]; a0.z += va[2]; a0.w += va[3];
    a1.x += vb[0]; a1.y += vb[1]; a1.z += vb[2]; a1.w += vb[3];
  }
  float4* out = reinterpret_cast<float4*>(h0 + (size_t)t * kD);
  out[lane] = a0;
  out[32 + lane] = a1;
}

// d(PositionEmbedding table)[s] = sum over documents of dh0[b, s] under the same dropout mask.  grid = S, block = D/4 threads.
__global__ void __launch_bounds__(kD / 4) pos_embed_bwd_kernel(const float* __restrict__ dh0, int B, int S, float rate, uint32_t seed, uint32_t step,
                                                               float* __restrict__ dtable, uint32_t row0) {
  pdl_wait();
  const int s = blockIdx.x, q = threadIdx.x;  // float4 index within the row
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const size_t t = (size_t)b * S + s;
    const float4 g = reinterpret_cast<const float4*>(dh0 + t * kD)[q];
    float v[4] = {g.x, g.y, g.z, g.w};
    if (rate > 0.f) dropout4(v, ((uint32_t)t + row0) * kD + 4u * q, rate, seed, step, kSitePosDropout);
    acc.x += v[0]; acc.y += v[1]; acc.z += v[2]; acc.w += v[3];
  }
  reinterpret_cast<float4*>(dtable + (size_t)s * kD)[q] = acc;
}

// The same with a context token in the sequence (encoder.py:247-252: the token takes position 0, the element in row s position s + 1):
// dtable[0] = sum of the token rows, dtable[p] = sum over documents of dh0[b, p - 1] unless that row holds the token.  grid = S + 1.
__global__ void __launch_bounds__(kD / 4) pos_embed_bwd_ctx_kernel(const float* __restrict__ dh0, const int* __restrict__ ctx_row, int B, int S, float rate,
                                                                   uint32_t seed, uint32_t step, float* __restrict__ dtable, uint32_t row0) {
  pdl_wait();
  const int p = blockIdx.x, q = threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const int tok = __ldg(ctx_row + b);
    int row;
    if (p == 0) {
      if (tok >= S) continue;  // a document without room for its token
      row = tok;
    } else {
      row = p - 1;
      if (row == tok) continue;
    }
    const size_t t = (size_t)b * S + row;
    const float4 g = reinterpret_cast<const float4*>(dh0 + t * kD)[q];
    float v[4] = {g.x, g.y, g.z, g.w};
    if (rate > 0.f) dropout4(v, ((uint32_t)t + row0) * kD + 4u * q, rate, seed, step, kSitePosDropout);
    acc.x += v[0]; acc.y += v[1]; acc.z += v[2]; acc.w += v[3];
  }
  reinterpret_cast<float4*>(dtable + (size_t)p * kD)[q] = acc;
}

// Backward of the above as a tensor-core contraction.  Every element selects a handful of "gradient rows" (one per
// categorical sub-target; for a numerical field the <MASK> row, the <UNUSED> row or the Dense bias), so
//     d(table rows)[R, D] = OneHot[T, R]^T . dh0[T, D]
// is a wgrad-shaped GEMM with an MN-major multi-hot A operand.  embed_onehot_kernel writes that operand (small integer
// counts, exact in TF32); the engine runs the GEMM into a scratch [R, D] block and embed_scatter_kernel moves the rows
// to their variables.  The Dense kernels' gradients are X^T . dh0m GEMMs, dh0m = dh0 with the rows of special-token
// elements zeroed (written by the last LayerNorm-backward launch, transformer.cu).
__global__ void __launch_bounds__(32 * kTokPerCta) embed_onehot_kernel(const __grid_constant__ Schema sc, const __grid_constant__ BatchPtrs mod,
                                                                       const unsigned char* __restrict__ flags, int T, float* __restrict__ onehot,
                                                                       const int* __restrict__ ctx_row, int S) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * kTokPerCta + (threadIdx.x >> 5);
  if (t >= T) return;
  float4* out = reinterpret_cast<float4*>(onehot + (size_t)t * sc.Rp);
  const int n4 = sc.Rp >> 2;
  for (int j = lane; j < n4; j += 32) out[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ctx_row && (t % S) == __ldg(ctx_row + t / S)) return;  // the context token's row: its gradient belongs to the context table only
  __syncwarp();
  for (int l = lane; l < sc.n_lookups; l += 32) {
    const int f = sc.lk_field[l], c = sc.lk_sub[l];
    const FieldDev& fd = sc.f[f];
    int r;
    if (fd.kind == 0) {
      const int i = __ldg(reinterpret_cast<const int*>(mod.cols[f]) + (size_t)t * fd.C + c);
      r = fd.grow_off + min(max(i, 0), fd.input_dim + 1);
    } else {
      const int flag = flags[(size_t)fd.num_slot * T + t];
      r = fd.grow_off + (flag ? flag - 1 : 2);
    }
    atomicAdd(onehot + (size_t)t * sc.Rp + r, 1.0f);  // sub-targets of one field may pick the same row (color): counts, not flags
  }
}

// grid = rows of the scratch block, block = D
__global__ void __launch_bounds__(kD) embed_scatter_kernel(const __grid_constant__ Schema sc, const float* __restrict__ scratch, float* __restrict__ grads) {
  pdl_wait();
  const int d = threadIdx.x;
  const int r = blockIdx.x;
  {
    for (int f = 0; f < sc.F; ++f) {
      const FieldDev& fd = sc.f[f];
      const int rows = (fd.kind == 0) ? fd.input_dim + 2 : 3;
      if (r >= fd.grow_off && r < fd.grow_off + rows) {
        const int k = r - fd.grow_off;
        const long long dst = (fd.kind == 0 || k < 2) ? fd.table_off + (long long)k * kD : fd.bias_off;
        grads[dst + d] = scratch[(size_t)r * kD + d];
        return;
      }
    }
  }
}

int launch_embed_fwd(const Schema& sc, const BatchPtrs& mod, const unsigned char* flags, const float* params, int T, float* h0, cudaStream_t st,
                     const PosEmbed& pos) {
  MFP_CUDA_OK(launch_pdl(embed_fwd_kernel, (T + kTokPerCta - 1) / kTokPerCta, 32 * kTokPerCta, 0, st, sc, mod, flags, params, T, h0, pos));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_pos_embed_bwd(const float* dh0, int B, int S, float rate, uint32_t seed, uint32_t step, float* dtable, cudaStream_t st, uint32_t row0) {
  MFP_CUDA_OK(launch_pdl(pos_embed_bwd_kernel, S, kD / 4, 0, st, dh0, B, S, rate, seed, step, dtable, row0));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_pos_embed_bwd_ctx(const float* dh0, const int* ctx_row, int B, int S, float rate, uint32_t seed, uint32_t step, float* dtable, cudaStream_t st,
                             uint32_t row0) {
  MFP_CUDA_OK(launch_pdl(pos_embed_bwd_ctx_kernel, S + 1, kD / 4, 0, st, dh0, ctx_row, B, S, rate, seed, step, dtable, row0));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

// h0[b, length[b] + 1, :] = table[ids[b]] and ctx_row[b] = length[b] + 1.  grid = B, block = D/4.  A document that fills all S rows has
// no room for the token (the host pads the batch by one row): it is left without one and ctx_row[b] = S (matches no row; the attention
// kernels clamp their length to S).
__global__ void __launch_bounds__(kD / 4) context_token_kernel(const float* __restrict__ table, int rows, const int* __restrict__ ids,
                                                               const int* __restrict__ length, int S, float* __restrict__ h0, int* __restrict__ ctx_row,
                                                               const PosEmbed pos) {
  pdl_wait();
  const int b = blockIdx.x, q = threadIdx.x;
  const int n = __ldg(length + b) + 1;
  if (n >= S) {
    if (q == 0) ctx_row[b] = S;
    return;
  }
  const int id = min(max(__ldg(ids + b), 0), rows - 1);
  float4 tok = __ldg(reinterpret_cast<const float4*>(table + (size_t)id * kD) + q);
  if (pos.table) {  // positions are added after the token was put in front (encoder.py:247-252): the token has position 0
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos.table) + q);
    float v[4] = {p0.x, p0.y, p0.z, p0.w};
    if (pos.rate > 0.f) dropout4(v, ((uint32_t)((size_t)b * S + n) + pos.row0) * kD + 4u * q, pos.rate, pos.seed, pos.step, kSitePosDropout);
    tok.x += v[0]; tok.y += v[1]; tok.z += v[2]; tok.w += v[3];
  }
  reinterpret_cast<float4*>(h0 + ((size_t)b * S + n) * kD)[q] = tok;
  if (q == 0) ctx_row[b] = n;
}

// d(context table)[r] = sum over the documents whose id is r of dh0[b, ctx_row[b]] (fixed order: deterministic).  grid = rows, block = D/4.
__global__ void __launch_bounds__(kD / 4) context_token_bwd_kernel(const float* __restrict__ dh0, const int* __restrict__ ids, const int* __restrict__ ctx_row,
                                                                   int rows, int B, int S, float* __restrict__ dtable) {
  pdl_wait();
  const int r = blockIdx.x, q = threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const int row = __ldg(ctx_row + b);
    if (row >= S || min(max(__ldg(ids + b), 0), rows - 1) != r) continue;
    const float4 g = reinterpret_cast<const float4*>(dh0 + ((size_t)b * S + row) * kD)[q];
    acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
  }
  reinterpret_cast<float4*>(dtable + (size_t)r * kD)[q] = acc;
}

// vec[b, :] = sum_c params[off_c + clamp(ids_c[b]) * D ...] in column order (encoder.py:156-160,194-199).  grid = B, block = D/4.
__global__ void __launch_bounds__(kD / 4) canvas_vector_kernel(const __grid_constant__ CanvasArgs a, const float* __restrict__ params, float* __restrict__ vec) {
  pdl_wait();
  const int b = blockIdx.x, q = threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < a.n; ++c) {
    const int id = min(max(__ldg(a.ids[c] + b), 0), a.rows[c] - 1);
    const float4 r = __ldg(reinterpret_cast<const float4*>(params + a.off[c] + (long long)id * kD) + q);
    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
  }
  reinterpret_cast<float4*>(vec + (size_t)b * kD)[q] = acc;
}

// x[b, s, :] += vec[b, :] for every row (seq += canvas, encoder.py:228-230).  grid = B * S, block = D/4.
__global__ void __launch_bounds__(kD / 4) add_doc_vector_kernel(float* __restrict__ x, const float* __restrict__ vec, int S) {
  pdl_wait();
  const int t = blockIdx.x, q = threadIdx.x;
  const float4 v = __ldg(reinterpret_cast<const float4*>(vec + (size_t)(t / S) * kD) + q);
  float4* row = reinterpret_cast<float4*>(x + (size_t)t * kD) + q;
  float4 r = *row;
  r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
  *row = r;
}

// dvec[b, :] = sum_s dx[b, s, :] (fixed order).  grid = B, block = D/4.
__global__ void __launch_bounds__(kD / 4) sum_doc_rows_kernel(const float* __restrict__ dx, int S, float* __restrict__ dvec) {
  pdl_wait();
  const int b = blockIdx.x, q = threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < S; ++s) {
    const float4 g = reinterpret_cast<const float4*>(dx + ((size_t)b * S + s) * kD)[q];
    acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
  }
  reinterpret_cast<float4*>(dvec + (size_t)b * kD)[q] = acc;
}

__global__ void iota_kernel(int* __restrict__ iota, int* __restrict__ zeros, int n) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { iota[i] = i; zeros[i] = 0; }
}

int launch_canvas_vector(const CanvasArgs& a, const float* params, int B, float* vec, cudaStream_t st) {
  MFP_CUDA_OK(launch_pdl(canvas_vector_kernel, B, kD / 4, 0, st, a, params, vec));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_add_doc_vector(float* x, const float* vec, int B, int S, cudaStream_t st) {
  MFP_CUDA_OK(launch_pdl(add_doc_vector_kernel, B * S, kD / 4, 0, st, x, vec, S));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_sum_doc_rows(const float* dx, int B, int S, float* dvec, cudaStream_t st) {
  MFP_CUDA_OK(launch_pdl(sum_doc_rows_kernel, B, kD / 4, 0, st, dx, S, dvec));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_iota(int* iota, int* zeros, int n, cudaStream_t st) {
  MFP_CUDA_OK(launch_pdl(iota_kernel, (n + 255) / 256, 256, 0, st, iota, zeros, n));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_context_token(const float* table, int rows, const int* ids, const int* length, int B, int S, float* h0, int* ctx_row, cudaStream_t st,
                         const PosEmbed& pos) {
  MFP_CUDA_OK(launch_pdl(context_token_kernel, B, kD / 4, 0, st, table, rows, ids, length, S, h0, ctx_row, pos));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_context_token_bwd(const float* dh0, const int* ids, const int* ctx_row, int rows, int B, int S, float* dtable, cudaStream_t st) {
  MFP_CUDA_OK(launch_pdl(context_token_bwd_kernel, rows, kD / 4, 0, st, dh0, ids, ctx_row, rows, B, S, dtable));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_embed_onehot(const Schema& sc, const BatchPtrs& mod, const unsigned char* flags, int T, float* onehot, cudaStream_t st, const int* ctx_row, int S) {
  MFP_CUDA_OK(launch_pdl(embed_onehot_kernel, (T + kTokPerCta - 1) / kTokPerCta, 32 * kTokPerCta, 0, st, sc, mod, flags, T, onehot, ctx_row, S));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_embed_scatter(const Schema& sc, const float* scratch, float* grads, cudaStream_t st) {
  MFP_CUDA_OK(launch_pdl(embed_scatter_kernel, sc.R, kD, 0, st, sc, scratch, grads));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

}  // namespace mfp
