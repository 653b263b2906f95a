// Non-GEMM kernels of the DeepSVG block (reference: architecture/transformer.py:33-99,208-229):
// LayerNorm fwd/bwd (warp-shuffle reductions, eps = 1e-3, biased variance), masked multi-head attention core
// fwd/bwd (flash-style per (document, head), keys limited to the document's length), dropout backward, column sums.
#include "kernels.cuh"
#include "gemm.cuh"

namespace mfp {

// ------------------------------------------------------------------------------------------------- LayerNorm
// one warp per row of D = 256: lane owns columns [4*lane, 4*lane+4) and [128 + 4*lane, ...)
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int T, float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= T) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)t * kD);
  const float4 a = xr[lane], b = xr[32 + lane];
  const float mu = warp_sum(a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w) * (1.0f / kD);
  const float4 ca = make_float4(a.x - mu, a.y - mu, a.z - mu, a.w - mu);
  const float4 cb = make_float4(b.x - mu, b.y - mu, b.z - mu, b.w - mu);
  const float var = warp_sum(ca.x * ca.x + ca.y * ca.y + ca.z * ca.z + ca.w * ca.w + cb.x * cb.x + cb.y * cb.y + cb.z * cb.z + cb.w * cb.w) * (1.0f / kD);
  const float rs = rsqrtf(var + kLnEps);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + lane), gb = __ldg(reinterpret_cast<const float4*>(gamma) + 32 + lane);
  const float4 ba = __ldg(reinterpret_cast<const float4*>(beta) + lane), bb = __ldg(reinterpret_cast<const float4*>(beta) + 32 + lane);
  float4* yr = reinterpret_cast<float4*>(y + (size_t)t * kD);
  yr[lane] = make_float4(ca.x * rs * ga.x + ba.x, ca.y * rs * ga.y + ba.y, ca.z * rs * ga.z + ba.z, ca.w * rs * ga.w + ba.w);
  yr[32 + lane] = make_float4(cb.x * rs * gb.x + bb.x, cb.y * rs * gb.y + bb.y, cb.z * rs * gb.z + bb.z, cb.w * rs * gb.w + bb.w);
  if (lane == 0) { mean[t] = mu; rstd[t] = rs; }
}

// dx = dres + rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)),  dxhat = dy * gamma
// dgamma += sum_t dy * xhat, dbeta += sum_t dy  (per-CTA register/shared reduction, one atomic per column per CTA)
// Four CTAs per SM (64 registers: at 80 the kernel lost a quarter of its loads in flight and a quarter of its bandwidth).
// One warp per row; lane l owns columns [4 l, 4 l + 4) and [128 + 4 l, ...): every warp-wide 16-byte access covers 512 contiguous bytes
// (eight consecutive columns per lane -- one Philox block of the dropout contract -- measured 8 % slower: half-used sectors per access).
__global__ void __launch_bounds__(256, 4) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ dres, int T, float* __restrict__ dx,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            const unsigned char* __restrict__ rowflags, int n_masked, float* __restrict__ dx_masked,
                                                            float* __restrict__ dx_drop, float drop_rate, uint32_t drop_seed, uint32_t drop_step,
                                                            uint32_t drop_site, uint32_t drop_row0, float* __restrict__ det_part) {
  pdl_wait();
  __shared__ float red[2][8][kD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + lane), gb = __ldg(reinterpret_cast<const float4*>(gamma) + 32 + lane);
  float dg[8] = {}, db[8] = {};
  for (int t = blockIdx.x * 8 + warp; t < T; t += gridDim.x * 8) {
    const float mu = mean[t], rs = rstd[t];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)t * kD);
    const float4* dr = reinterpret_cast<const float4*>(dy + (size_t)t * kD);
    const float4 xa = xr[lane], xb = xr[32 + lane], da = dr[lane], dbv = dr[32 + lane];
    const float xh[8] = {(xa.x - mu) * rs, (xa.y - mu) * rs, (xa.z - mu) * rs, (xa.w - mu) * rs, (xb.x - mu) * rs, (xb.y - mu) * rs, (xb.z - mu) * rs, (xb.w - mu) * rs};
    const float dyv[8] = {da.x, da.y, da.z, da.w, dbv.x, dbv.y, dbv.z, dbv.w};
    const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    float dxh[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dxh[i] = dyv[i] * g[i];
      s1 += dxh[i];
      s2 += dxh[i] * xh[i];
      dg[i] += dyv[i] * xh[i];
      db[i] += dyv[i];
    }
    s1 = warp_sum(s1) * (1.0f / kD);
    s2 = warp_sum(s2) * (1.0f / kD);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = rs * (dxh[i] - s1 - xh[i] * s2);
    if (dres) {
      const float4* rr = reinterpret_cast<const float4*>(dres + (size_t)t * kD);
      const float4 ra = rr[lane], rb = rr[32 + lane];
      o[0] += ra.x; o[1] += ra.y; o[2] += ra.z; o[3] += ra.w; o[4] += rb.x; o[5] += rb.y; o[6] += rb.z; o[7] += rb.w;
    }
    float4* out = reinterpret_cast<float4*>(dx + (size_t)t * kD);
    out[lane] = make_float4(o[0], o[1], o[2], o[3]);
    out[32 + lane] = make_float4(o[4], o[5], o[6], o[7]);
    // the gradient entering the next (earlier) sub-layer's branch is dx under that branch's dropout mask: written here so
    // that no separate dropout-backward pass re-reads dx
    if (dx_drop) {
      float va[4] = {o[0], o[1], o[2], o[3]}, vb[4] = {o[4], o[5], o[6], o[7]};
      dropout4(va, ((uint32_t)t + drop_row0) * kD + 4u * lane, drop_rate, drop_seed, drop_step, drop_site);
      dropout4(vb, ((uint32_t)t + drop_row0) * kD + 128u + 4u * lane, drop_rate, drop_seed, drop_step, drop_site);
      float4* po = reinterpret_cast<float4*>(dx_drop + (size_t)t * kD);
      po[lane] = make_float4(va[0], va[1], va[2], va[3]);
      po[32 + lane] = make_float4(vb[0], vb[1], vb[2], vb[3]);
    }
    // encoder: copies of dx with the rows of special-token elements zeroed, one per numerical field (B operand of its Dense wgrad)
    for (int s = 0; s < n_masked; ++s) {
      const bool keep = rowflags[(size_t)s * T + t] == 0;
      float4* mo = reinterpret_cast<float4*>(dx_masked + ((size_t)s * T + t) * kD);
      mo[lane] = keep ? make_float4(o[0], o[1], o[2], o[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
      mo[32 + lane] = keep ? make_float4(o[4], o[5], o[6], o[7]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[0][warp][4 * lane + i] = dg[i];
    red[0][warp][128 + 4 * lane + i] = dg[4 + i];
    red[1][warp][4 * lane + i] = db[i];
    red[1][warp][128 + 4 * lane + i] = db[4 + i];
  }
  __syncthreads();
  const int c = threadIdx.x;
  float sg = 0.f, sb = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { sg += red[0][w][c]; sb += red[1][w][c]; }
  if (det_part) {  // deterministic mode: per-CTA partials, summed in CTA order by ln_param_reduce_kernel
    det_part[(size_t)blockIdx.x * 2 * kD + c] = sg;
    det_part[(size_t)blockIdx.x * 2 * kD + kD + c] = sb;
  } else {
    atomicAdd(dgamma + c, sg);
    atomicAdd(dbeta + c, sb);
  }
}

__global__ void __launch_bounds__(2 * kD) ln_param_reduce_kernel(const float* __restrict__ part, int ctas, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_wait();
  const int c = threadIdx.x;  // [0, kD): gamma, [kD, 2 kD): beta
  float acc = 0.f;
  for (int b = 0; b < ctas; ++b) acc += part[(size_t)b * 2 * kD + c];
  if (c < kD) dgamma[c] += acc; else dbeta[c - kD] += acc;
}

// ------------------------------------------------------------------------------------------------- attention core
// qkv: [T, 3D] (q | k | v, head h = columns 32h..32h+31 of each third); out: [T, D] heads merged (transformer.py:92-97)
// grid = B*H, block = 128.  K and V of the (document, head) live in shared memory; one thread per query row keeps
// q, the running max / sum and the 32-wide output in registers (online softmax).  Keys j >= n are skipped: the
// reference adds -1e9 to them (transformer.py:73), which is exactly zero probability in fp32.
constexpr int kAttnThreads = 128;

__global__ void __launch_bounds__(kAttnThreads) attention_fwd_kernel(const float* __restrict__ qkv, const int* __restrict__ length, int S,
                                                                     float* __restrict__ out, float* __restrict__ lse) {
  pdl_wait();
  extern __shared__ float sm[];
  float* Ks = sm;
  float* Vs = sm + (size_t)S * kDh;
  const int b = blockIdx.x / kH, h = blockIdx.x % kH;
  const int n = min(S, length[b] + 1);
  const size_t row0 = (size_t)b * S;
  for (int i = threadIdx.x; i < n * (kDh / 4); i += kAttnThreads) {
    const int j = i / (kDh / 4), c = i % (kDh / 4);
    const float4* src = reinterpret_cast<const float4*>(qkv + (row0 + j) * (3 * kD) + h * kDh);
    reinterpret_cast<float4*>(Ks + j * kDh)[c] = src[kD / 4 + c];
    reinterpret_cast<float4*>(Vs + j * kDh)[c] = src[2 * kD / 4 + c];
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;  // 1/sqrt(32), transformer.py:62-63
  for (int i = threadIdx.x; i < S; i += kAttnThreads) {
    float q[kDh], o[kDh];
    const float4* qp = reinterpret_cast<const float4*>(qkv + (row0 + i) * (3 * kD) + h * kDh);
#pragma unroll
    for (int c = 0; c < kDh / 4; ++c) {
      const float4 v = qp[c];
      q[4 * c] = v.x * scale; q[4 * c + 1] = v.y * scale; q[4 * c + 2] = v.z * scale; q[4 * c + 3] = v.w * scale;
    }
#pragma unroll
    for (int c = 0; c < kDh; ++c) o[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n; ++j) {
      const float4* kp = reinterpret_cast<const float4*>(Ks + j * kDh);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < kDh / 4; ++c) {
        const float4 kv = kp[c];
        s = fmaf(q[4 * c], kv.x, s); s = fmaf(q[4 * c + 1], kv.y, s); s = fmaf(q[4 * c + 2], kv.z, s); s = fmaf(q[4 * c + 3], kv.w, s);
      }
      if (s > m) {
        const float corr = __expf(m - s);
        l *= corr;
#pragma unroll
        for (int c = 0; c < kDh; ++c) o[c] *= corr;
        m = s;
      }
      const float p = __expf(s - m);
      l += p;
      const float4* vp = reinterpret_cast<const float4*>(Vs + j * kDh);
#pragma unroll
      for (int c = 0; c < kDh / 4; ++c) {
        const float4 vv = vp[c];
        o[4 * c] = fmaf(p, vv.x, o[4 * c]); o[4 * c + 1] = fmaf(p, vv.y, o[4 * c + 1]);
        o[4 * c + 2] = fmaf(p, vv.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(p, vv.w, o[4 * c + 3]);
      }
    }
    const float inv = 1.0f / l;
    float4* op = reinterpret_cast<float4*>(out + (row0 + i) * kD + h * kDh);
#pragma unroll
    for (int c = 0; c < kDh / 4; ++c) op[c] = make_float4(o[4 * c] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv);
    lse[((size_t)b * kH + h) * S + i] = m + __logf(l);
  }
}

// Backward.  Rows i >= n carry zero upstream gradient (the loss weights them by seq_mask, metrics.py:263-267, and
// they are never keys), so both passes run over i, j < n and padded rows of dqkv are written as zeros.
__global__ void __launch_bounds__(kAttnThreads) attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ out,
                                                                     const float* __restrict__ lse, const float* __restrict__ dout,
                                                                     const int* __restrict__ length, int S, float* __restrict__ dqkv) {
  pdl_wait();
  extern __shared__ float sm[];
  float* Qs = sm;                         // pre-scaled by 1/sqrt(dh)
  float* Ks = Qs + (size_t)S * kDh;
  float* Vs = Ks + (size_t)S * kDh;
  float* dOs = Vs + (size_t)S * kDh;
  float* Ls = dOs + (size_t)S * kDh;      // lse
  float* Ds = Ls + S;                     // D_i = dO_i . O_i
  const int b = blockIdx.x / kH, h = blockIdx.x % kH;
  const int n = min(S, length[b] + 1);
  const size_t row0 = (size_t)b * S;
  const float scale = 0.17677669529663687f;
  for (int i = threadIdx.x; i < n * (kDh / 4); i += kAttnThreads) {
    const int j = i / (kDh / 4), c = i % (kDh / 4);
    const float4* src = reinterpret_cast<const float4*>(qkv + (row0 + j) * (3 * kD) + h * kDh);
    const float4 qv = src[c];
    reinterpret_cast<float4*>(Qs + j * kDh)[c] = make_float4(qv.x * scale, qv.y * scale, qv.z * scale, qv.w * scale);
    reinterpret_cast<float4*>(Ks + j * kDh)[c] = src[kD / 4 + c];
    reinterpret_cast<float4*>(Vs + j * kDh)[c] = src[2 * kD / 4 + c];
    reinterpret_cast<float4*>(dOs + j * kDh)[c] = reinterpret_cast<const float4*>(dout + (row0 + j) * kD + h * kDh)[c];
  }
  for (int i = threadIdx.x; i < n; i += kAttnThreads) {
    const float4* op = reinterpret_cast<const float4*>(out + (row0 + i) * kD + h * kDh);
    const float4* dp = reinterpret_cast<const float4*>(dout + (row0 + i) * kD + h * kDh);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < kDh / 4; ++c) {
      const float4 a = op[c], d = dp[c];
      acc += a.x * d.x + a.y * d.y + a.z * d.z + a.w * d.w;
    }
    Ds[i] = acc;
    Ls[i] = lse[((size_t)b * kH + h) * S + i];
  }
  __syncthreads();
  // pass A: thread per query row -> dQ_i = scale * sum_j ds_ij K_j,  ds_ij = p_ij (dO_i.V_j - D_i)
  for (int i = threadIdx.x; i < S; i += kAttnThreads) {
    float dq[kDh];
#pragma unroll
    for (int c = 0; c < kDh; ++c) dq[c] = 0.f;
    if (i < n) {
      float q[kDh], go[kDh];
#pragma unroll
      for (int c = 0; c < kDh; ++c) { q[c] = Qs[i * kDh + c]; go[c] = dOs[i * kDh + c]; }
      const float li = Ls[i], di = Ds[i];
      for (int j = 0; j < n; ++j) {
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int c = 0; c < kDh; ++c) { s = fmaf(q[c], Ks[j * kDh + c], s); dp = fmaf(go[c], Vs[j * kDh + c], dp); }
        const float ds = __expf(s - li) * (dp - di);
#pragma unroll
        for (int c = 0; c < kDh; ++c) dq[c] = fmaf(ds, Ks[j * kDh + c], dq[c]);
      }
    }
    float4* dst = reinterpret_cast<float4*>(dqkv + (row0 + i) * (3 * kD) + h * kDh);
#pragma unroll
    for (int c = 0; c < kDh / 4; ++c) dst[c] = make_float4(dq[4 * c] * scale, dq[4 * c + 1] * scale, dq[4 * c + 2] * scale, dq[4 * c + 3] * scale);
  }
  // pass B: thread per key row -> dV_j = sum_i p_ij dO_i,  dK_j = sum_i ds_ij Q_i(scaled)
  for (int j = threadIdx.x; j < S; j += kAttnThreads) {
    float dk[kDh], dv[kDh];
#pragma unroll
    for (int c = 0; c < kDh; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
    if (j < n) {
      float kj[kDh], vj[kDh];
#pragma unroll
      for (int c = 0; c < kDh; ++c) { kj[c] = Ks[j * kDh + c]; vj[c] = Vs[j * kDh + c]; }
      for (int i = 0; i < n; ++i) {
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int c = 0; c < kDh; ++c) { s = fmaf(Qs[i * kDh + c], kj[c], s); dp = fmaf(dOs[i * kDh + c], vj[c], dp); }
        const float p = __expf(s - Ls[i]);
        const float ds = p * (dp - Ds[i]);
#pragma unroll
        for (int c = 0; c < kDh; ++c) { dv[c] = fmaf(p, dOs[i * kDh + c], dv[c]); dk[c] = fmaf(ds, Qs[i * kDh + c], dk[c]); }
      }
    }
    float4* dkp = reinterpret_cast<float4*>(dqkv + (row0 + j) * (3 * kD) + kD + h * kDh);
    float4* dvp = reinterpret_cast<float4*>(dqkv + (row0 + j) * (3 * kD) + 2 * kD + h * kDh);
#pragma unroll
    for (int c = 0; c < kDh / 4; ++c) {
      dkp[c] = make_float4(dk[4 * c], dk[4 * c + 1], dk[4 * c + 2], dk[4 * c + 3]);
      dvp[c] = make_float4(dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------- small kernels
__global__ void __launch_bounds__(256) dropout_bwd_kernel(const float* __restrict__ dx, size_t n8, float rate, uint32_t seed, uint32_t step,
                                                          uint32_t site, float* __restrict__ dy, uint32_t row0) {
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // eight elements (one Philox block) per thread
  if (i >= n8) return;
  const float4 g0 = reinterpret_cast<const float4*>(dx)[2 * i], g1 = reinterpret_cast<const float4*>(dx)[2 * i + 1];
  float va[4] = {g0.x, g0.y, g0.z, g0.w}, vb[4] = {g1.x, g1.y, g1.z, g1.w};
  dropout8(va, vb, (uint32_t)(i * 8) + row0 * kD, rate, seed, step, site);
  reinterpret_cast<float4*>(dy)[2 * i] = make_float4(va[0], va[1], va[2], va[3]);
  reinterpret_cast<float4*>(dy)[2 * i + 1] = make_float4(vb[0], vb[1], vb[2], vb[3]);
}

// dst[s][t][:] = flags[s][t] ? 0 : src[t][:] -- the encoder's Dense weight gradients take dh0 with the rows of special-token elements
// zeroed, one copy per numerical field (the pre-LayerNorm path gets these copies out of its last LayerNorm backward)
__global__ void __launch_bounds__(256) masked_copies_kernel(const float* __restrict__ src, const unsigned char* __restrict__ flags, int n_copies, int T,
                                                            float* __restrict__ dst) {
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index into [T, kD]
  if (i >= (size_t)T * kD / 4) return;
  const int t = (int)(i / (kD / 4));
  const float4 v = reinterpret_cast<const float4*>(src)[i];
  for (int s = 0; s < n_copies; ++s)
    reinterpret_cast<float4*>(dst + (size_t)s * T * kD)[i] = flags[(size_t)s * T + t] ? make_float4(0.f, 0.f, 0.f, 0.f) : v;
}

// out[c] += sum_r x[r, c]; grid = (col blocks, row chunks)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int rows, int cols, int ld, int rows_per_chunk,
                                                     float* __restrict__ out) {
  pdl_wait();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += x[(size_t)r * ld + c];
  atomicAdd(out + c, acc);
}

// ------------------------------------------------------------------------------------------------- launchers
int launch_layernorm_fwd(const float* x, const float* gamma, const float* beta, int T, float* y, float* mean, float* rstd, cudaStream_t st) {
  MFP_CUDA_OK(launch_pdl(layernorm_fwd_kernel, (T + 7) / 8, 256, 0, st, x, gamma, beta, T, y, mean, rstd));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_layernorm_bwd(const float* x, const float* dy, const float* gamma, const float* mean, const float* rstd, const float* dres, int T,
                         float* dx, float* dgamma, float* dbeta, cudaStream_t st, const unsigned char* rowflags, int n_masked, float* dx_masked,
                         float* dx_drop, float drop_rate, uint32_t drop_seed, uint32_t drop_step, uint32_t drop_site, uint32_t drop_row0, float* det_part) {
  const int grid = min((T + 7) / 8, kLnBwdMaxCtas);
  MFP_CUDA_OK(launch_pdl(layernorm_bwd_kernel, grid, 256, 0, st, x, dy, gamma, mean, rstd, dres, T, dx, dgamma, dbeta, rowflags, n_masked, dx_masked, dx_drop, drop_rate,
                                             drop_seed, drop_step, drop_site, drop_row0, det_part));
  MFP_CUDA_OK(cudaGetLastError());
  if (det_part) {
    MFP_CUDA_OK(launch_pdl(ln_param_reduce_kernel, 1, 2 * kD, 0, st, (const float*)det_part, grid, dgamma, dbeta));
    MFP_CUDA_OK(cudaGetLastError());
  }
  return MFP_OK;
}

static int attn_smem_check(size_t bytes, const void* fn, size_t* configured) {
  if (bytes > 220 * 1024) { set_error("attention: sequence length needs %zu bytes of shared memory (limit 220 KB)", bytes); return MFP_ERR_UNSUPPORTED; }
  if (bytes > *configured) {
    MFP_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    *configured = bytes;
  }
  return MFP_OK;
}

int launch_attention_fwd(const float* qkv, const int* length, int B, int S, float* out, float* lse, cudaStream_t st) {
  static size_t configured = 0;
  const size_t smem = (size_t)2 * S * kDh * sizeof(float);
  MFP_TRY(attn_smem_check(smem, (const void*)attention_fwd_kernel, &configured));
  MFP_CUDA_OK(launch_pdl(attention_fwd_kernel, B * kH, kAttnThreads, smem, st, qkv, length, S, out, lse));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_attention_bwd(const float* qkv, const float* out, const float* lse, const float* dout, const int* length, int B, int S, float* dqkv,
                         cudaStream_t st) {
  static size_t configured = 0;
  const size_t smem = ((size_t)4 * S * kDh + 2 * S) * sizeof(float);
  MFP_TRY(attn_smem_check(smem, (const void*)attention_bwd_kernel, &configured));
  MFP_CUDA_OK(launch_pdl(attention_bwd_kernel, B * kH, kAttnThreads, smem, st, qkv, out, lse, dout, length, S, dqkv));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_dropout_bwd(const float* dx, int T, float rate, uint32_t seed, uint32_t step, uint32_t site, float* dy, cudaStream_t st, uint32_t row0) {
  const size_t n8 = (size_t)T * kD / 8;
  MFP_CUDA_OK(launch_pdl(dropout_bwd_kernel, (unsigned)((n8 + 255) / 256), 256, 0, st, dx, n8, rate, seed, step, site, dy, row0));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_masked_copies(const float* src, const unsigned char* flags, int n_copies, int T, float* dst, cudaStream_t st) {
  const size_t n4 = (size_t)T * kD / 4;
  MFP_CUDA_OK(launch_pdl(masked_copies_kernel, (unsigned)((n4 + 255) / 256), 256, 0, st, src, flags, n_copies, T, dst));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_colsum(const float* x, int rows, int cols, int ld, float* out, cudaStream_t st, bool deterministic) {
  const int chunks = deterministic ? 1 : min(128, (rows + 63) / 64);  // one chunk: no atomics between CTAs (bring-up path only)
  const int rows_per_chunk = (rows + chunks - 1) / chunks;
  MFP_CUDA_OK(launch_pdl(colsum_kernel, dim3((cols + 255) / 256, chunks), 256, 0, st, x, rows, cols, ld, rows_per_chunk, out));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

}  // namespace mfp
