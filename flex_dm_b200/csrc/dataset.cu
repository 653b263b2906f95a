// Device-resident dataset cache: cutting a padded batch out of a ragged split that lives in HBM (mfp_gather_documents,
// include/flexdm_mfp.h).  Reference: DataSpec.make_dataset(..., cache=True) (data/spec.py:213-253) keeps the *serialized* records in
// host memory and re-parses them every epoch; a B200 holds the parsed crello / rico splits (a few GB) many times over, so the
// steady-state input cost becomes one HBM -> HBM gather per step (read + write of the batch's columns, ~270 MB at cfg2 = ~45 us)
// and B indices over PCIe.
//
// Layout: per column a [total_elements, W] array of 32-bit words with the documents back to back, so the W * n words of one document
// are contiguous and a warp reads them coalesced.  blockIdx.y = column; the x-blocks of a column grid-stride over (document, chunk)
// items of its B * S * W output words, in 16-byte units when W is a multiple of 4 (the 512-float embeddings), word by word otherwise
// (C = 1 or 3 categorical columns).
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

// Work item = (document b, chunk of kChunk units of its S * W output words); unit = 16 bytes when W % 4 == 0, one word otherwise.
// Items are grid-strided with 32-bit arithmetic (one division per item, none per load).
constexpr int kChunk = 2048;  // units per item: 8 per thread

template <typename Unit>
__device__ __forceinline__ void gather_column(const Unit* __restrict__ src, Unit* __restrict__ dst, Unit padv, int units_per_elem, const long long* __restrict__ doc_start,
                                              const int* __restrict__ doc_len, const int* __restrict__ idx, int B, int S) {
  const unsigned per_doc = (unsigned)S * (unsigned)units_per_elem;
  const unsigned chunks = (per_doc + kChunk - 1) / kChunk;
  const unsigned items = (unsigned)B * chunks;
  for (unsigned item = blockIdx.x; item < items; item += gridDim.x) {
    const unsigned b = item / chunks, ch = item - b * chunks;
    const int d = __ldg(idx + b);
    const unsigned valid = (unsigned)__ldg(doc_len + d) * (unsigned)units_per_elem;
    const Unit* from = src + (size_t)__ldg(doc_start + d) * units_per_elem;
    Unit* to = dst + (size_t)b * per_doc;
    const unsigned hi = min(per_doc, (ch + 1) * kChunk);
#pragma unroll 4
    for (unsigned r = ch * kChunk + threadIdx.x; r < hi; r += blockDim.x) to[r] = r < valid ? __ldg(from + r) : padv;
  }
}

__global__ void __launch_bounds__(256) gather_documents_kernel(const __grid_constant__ GatherArgs a, const long long* __restrict__ doc_start,
                                                               const int* __restrict__ doc_len, const int* __restrict__ idx, int B, int S) {
  const int c = blockIdx.y;
  const int W = a.words[c];
  const uint32_t pad = a.pad[c];
  if ((W & 3) == 0)
    gather_column<uint4>(reinterpret_cast<const uint4*>(a.src[c]), reinterpret_cast<uint4*>(a.dst[c]), make_uint4(pad, pad, pad, pad), W >> 2, doc_start, doc_len, idx, B, S);
  else
    gather_column<uint32_t>(a.src[c], a.dst[c], pad, W, doc_start, doc_len, idx, B, S);
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
    const size_t per_doc = (size_t)S * ((desc->words[c] & 3) == 0 ? desc->words[c] >> 2 : desc->words[c]);
    if (per_doc * (size_t)B >= (1ull << 31)) { set_error("mfp_gather_documents: column %d is too large for 32-bit indexing", c); return MFP_ERR_UNSUPPORTED; }
    const size_t items = (size_t)B * ((per_doc + kChunk - 1) / kChunk);
    most = items > most ? items : most;
  }
  // x-blocks: the items of the widest column, at most one full wave of 8 resident 256-thread blocks per SM
  size_t bx = most;
  if (bx > 148 * 8) bx = 148 * 8;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)desc->n_columns);
  gather_documents_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, reinterpret_cast<const long long*>(doc_start), doc_len, idx, B, S);
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}
