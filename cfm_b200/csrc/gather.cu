// Row gather  out[k, :] = x[idx[k], :]   (reference: x0[i], x1[j] at
// torchcfm/optimal_transport.py:145 and the label variant at :213-218).
// Pure HBM-bound byte movement: rows are moved as 16-byte words when size and alignment allow,
// one word per thread, consecutive threads on consecutive words of a row (coalesced).
#include "common.cuh"

namespace cfm {

template <class W>
__global__ void gather_rows_kernel(const W* __restrict__ x, int64_t words_per_row,
                                   const int64_t* __restrict__ idx, int64_t n_idx,
                                   W* __restrict__ out) {
  const int64_t total = n_idx * words_per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = t / words_per_row, c = t - k * words_per_row;
    out[t] = x[idx[k] * words_per_row + c];
  }
}

template <class W>
static cudaError_t launch_gather(const void* x, int64_t words, const int64_t* idx, int64_t n_idx,
                                 void* out, cudaStream_t s) {
  const int64_t total = n_idx * words;
  if (total == 0) return cudaSuccess;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  gather_rows_kernel<W><<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const W*>(x), words, idx,
                                                         n_idx, reinterpret_cast<W*>(out)); ::cfm::note_launches(1);
  return cudaGetLastError();
}

}  // namespace cfm

using namespace cfm;

extern "C" int cfm_gather_rows(const void* x, int64_t row_elems, int elem_bytes, const int64_t* idx,
                               int64_t n_idx, void* out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  CFM_REQUIRE(n_idx >= 0 && row_elems >= 0, "cfm_gather_rows: negative size");
  if (n_idx == 0 || row_elems == 0) return CFM_OK;
  CFM_REQUIRE(x && idx && out, "cfm_gather_rows: null pointer");
  CFM_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8,
              "cfm_gather_rows: elem_bytes must be 1, 2, 4 or 8 (got %d)", elem_bytes);
  const int64_t row_bytes = row_elems * elem_bytes;
  const uintptr_t a = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out);
  if ((row_bytes & 15) == 0 && (a & 15) == 0)
    CFM_CUDA_OK(launch_gather<uint4>(x, row_bytes / 16, idx, n_idx, out, s));
  else if ((row_bytes & 7) == 0 && (a & 7) == 0)
    CFM_CUDA_OK(launch_gather<uint2>(x, row_bytes / 8, idx, n_idx, out, s));
  else if ((row_bytes & 3) == 0 && (a & 3) == 0)
    CFM_CUDA_OK(launch_gather<uint32_t>(x, row_bytes / 4, idx, n_idx, out, s));
  else if ((row_bytes & 1) == 0 && (a & 1) == 0)
    CFM_CUDA_OK(launch_gather<uint16_t>(x, row_bytes / 2, idx, n_idx, out, s));
  else
    CFM_CUDA_OK(launch_gather<uint8_t>(x, row_bytes, idx, n_idx, out, s));
  return CFM_OK;
}
