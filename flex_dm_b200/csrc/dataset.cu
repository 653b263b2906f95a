// Device-resident dataset cache: cutting a padded batch out of a ragged split that lives in HBM (mfp_gather_documents,
// include/flexdm_mfp.h).  Reference: DataSpec.make_dataset(..., cache=True) (data/spec.py:213-253) keeps the *serialized* records in
// host memory and re-parses them every epoch; a B200 holds the parsed crello / rico splits (a few GB) many times over, so the
// steady-state input cost becomes one HBM -> HBM gather per step (read + write of the batch's columns, ~270 MB at cfg2 = ~45 us)
// and B indices over PCIe.
//
// Layout: per column a [total_elements, W] array of 32-bit words with the documents back to back, so the W * n words of one document
// are contiguous and a warp reads them coalesced.  blockIdx.y = column; the x-blocks of a column grid-stride over its B * S * W output
// words in 16-byte groups when W is a multiple of 4 (the 512-float embeddings), word by word otherwise (C = 1 or 3 categorical columns).
#include <cstdint>

#include "common.cuh"
#include "../../include/flexdm_mfp.h"

namespace mfp {


struct GatherArgs {
  int n_columns;
  int words[MFP_GATHER_MAX_COLUMNS];
  uint32_t pad[MFP_GATHER_MAX_COLUMNS];
  const uint32_t* src[MFP_GATHER_MAX_COLUMNS];
  uint32_t* dst[MFP_GATHER_MAX_COLUMNS];
};

__global__ void __launch_bounds__(256) gather_documents_kernel(const __grid_constant__ GatherArgs a, const long long* __restrict__ doc_start,
                                                               const int* __restrict__ doc_len, const int* __restrict__ idx, int B, int S) {
  const int c = blockIdx.y;
  const int W = a.words[c];
  const uint32_t pad = a.pad[c];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((W & 3) == 0) {
    const int W4 = W >> 2;
    const size_t per_doc = (size_t)S * W4, total = (size_t)B * per_doc;
    const uint4* src = reinterpret_cast<const uint4*>(a.src[c]);
    uint4* dst = reinterpret_cast<uint4*>(a.dst[c]);
    const uint4 padv = make_uint4(pad, pad, pad, pad);
    for (size_t i = first; i < total; i += stride) {
      const int b = (int)(i / per_doc);
      const size_t r = i - (size_t)b * per_doc;  // (s, word group) within the document
      const int d = __ldg(idx + b);
      const size_t valid = (size_t)__ldg(doc_len + d) * W4;
      dst[i] = r < valid ? __ldg(src + (size_t)__ldg(doc_start + d) * W4 + r) : padv;
    }
  } else {
    const size_t per_doc = (size_t)S * W, total = (size_t)B * per_doc;
    const uint32_t* src = a.src[c];
    uint32_t* dst = a.dst[c];
    for (size_t i = first; i < total; i += stride) {
      const int b = (int)(i / per_doc);
      const size_t r = i - (size_t)b * per_doc;
      const int d = __ldg(idx + b);
      const size_t valid = (size_t)__ldg(doc_len + d) * W;
      dst[i] = r < valid ? __ldg(src + (size_t)__ldg(doc_start + d) * W + r) : pad;
    }
  }
}

}  // namespace mfp

extern "C" int mfp_gather_documents(const mfp_gather_desc* desc, const int64_t* doc_start, const int32_t* doc_len, const int32_t* idx, int32_t B,
                                    int32_t S, void* stream) {
  using namespace mfp;
  if (!desc || !doc_start || !doc_len || !idx) { set_error("mfp_gather_documents: null argument"); return MFP_ERR_ARG; }
  if (desc->n_columns < 1 || desc->n_columns > MFP_GATHER_MAX_COLUMNS || B < 1 || S < 0) { set_error("mfp_gather_documents: bad shape"); return MFP_ERR_ARG; }
  if (S == 0) return MFP_OK;
  GatherArgs a{};
  a.n_columns = desc->n_columns;
  size_t most = 0;
  for (int c = 0; c < desc->n_columns; ++c) {
    if (desc->words[c] < 1 || !desc->src[c] || !desc->dst[c]) { set_error("mfp_gather_documents: column %d is incomplete", c); return MFP_ERR_ARG; }
    if ((desc->words[c] & 3) == 0 && ((reinterpret_cast<uintptr_t>(desc->src[c]) | reinterpret_cast<uintptr_t>(desc->dst[c])) & 15)) {
      set_error("mfp_gather_documents: column %d needs 16-byte aligned buffers", c);
      return MFP_ERR_ARG;
    }
    a.words[c] = desc->words[c];
    a.pad[c] = desc->pad_word[c];
    a.src[c] = static_cast<const uint32_t*>(desc->src[c]);
    a.dst[c] = static_cast<uint32_t*>(desc->dst[c]);
    const size_t units = (size_t)B * S * ((desc->words[c] & 3) == 0 ? desc->words[c] >> 2 : desc->words[c]);
    most = units > most ? units : most;
  }
  // enough x-blocks for the widest column to fill the machine (148 SMs x 8 resident 256-thread blocks), no more than its work
  size_t bx = (most + 255) / 256;
  if (bx > 148 * 8) bx = 148 * 8;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)desc->n_columns);
  gather_documents_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, reinterpret_cast<const long long*>(doc_start), doc_len, idx, B, S);
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}
