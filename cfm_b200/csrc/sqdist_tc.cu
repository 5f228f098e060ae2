// tcgen05 cost-matrix path:  M = cdist(x0, x1)**2 with the x0.x1^T contraction on the 5th-gen
// tensor cores (gemm_tc.cuh, 3xTF32) and the |x0|^2 + |x1|^2 - 2 acc, clamp, sqrt, square, max
// epilogue fused on the TMEM read-out (reference: torchcfm/optimal_transport.py:84-86).
#include <stdlib.h>

#include <mutex>

#include "gemm_h3.cuh"

namespace cfm {

// ---- pre-pass: hi / lo split ------------------------------------------------------------------------
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                  float* __restrict__ lo, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float h, l;
    split_tf32(x[i], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

struct SqDistTcEpilogue {
  float* M;
  int64_t ldm;
  const float* nx;
  const float* ny;
  float* cost_max;
  int squared;
  float a, tmax;
  __device__ __forceinline__ void begin_row(int row, bool ok) { a = ok ? __ldg(nx + row) : 0.f; }
  __device__ __forceinline__ float one(float acc, float b) {
    const float d2 = (a + b) - 2.f * acc;
    float v = d2 < 0.f ? 0.f : d2;  // clamp_min(0) that lets NaN through, like ATen's
    const float s = __fsqrt_rn(v);
    return squared ? s * s : s;
  }
  __device__ __forceinline__ void store32(int row0, int lane, int col0, const uint32_t (&r)[32], int n0, int n1,
                                          float* tile) {
    const int row = row0 + lane;
    const bool ok = row < n0;
    if (col0 + 32 <= n1) {  // full chunk: float4 loads of |x1|^2 (16B aligned: col0 % 32 == 0), coalesced staged store
      float o[32];
      float m = 0.f;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ny + col0 + c));
        o[c] = one(__uint_as_float(r[c]), b.x); o[c + 1] = one(__uint_as_float(r[c + 1]), b.y);
        o[c + 2] = one(__uint_as_float(r[c + 2]), b.z); o[c + 3] = one(__uint_as_float(r[c + 3]), b.w);
        m = fmaxf(m, fmaxf(fmaxf(o[c], o[c + 1]), fmaxf(o[c + 2], o[c + 3])));
      }
      if (ok) tmax = fmaxf(tmax, m);
      tc_store_chunk32(tile, o, M + (int64_t)row0 * ldm + col0, ldm, n0 - row0, lane);
    } else if (ok) {
      float* dst = M + (int64_t)row * ldm + col0;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (col0 + c < n1) {
          const float v = one(__uint_as_float(r[c]), __ldg(ny + col0 + c));
          tmax = fmaxf(tmax, v);
          dst[c] = v;
        }
    }
  }
  __device__ __forceinline__ void finish(int lane) {
    if (cost_max == nullptr) return;
    const float m = warp_max(tmax);
    if (lane == 0) atomic_max_nonneg(cost_max, m);
  }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// ---- TMA descriptor cache ---------------------------------------------------------------------------
// A CUtensorMap is a pure function of (base pointer, shape, strides, box, element type); a training loop
// hands the same buffers to the GEMMs every step, so the encode call is made once per distinct key and the
// 128-byte descriptor is replayed from a small direct-mapped table afterwards (the only library state
// besides the error string; guarded by a mutex, safe from several host threads).
namespace {
struct MapKey {
  const void* base; int rows, d, box_rows, elem; int64_t ld;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && d == o.d && box_rows == o.box_rows && elem == o.elem && ld == o.ld;
  }
};
struct MapSlot { MapKey key; CUtensorMap map; bool used; };
constexpr int kMapSlots = 256;
MapSlot g_maps[kMapSlots];
std::mutex g_map_mu;
inline unsigned map_hash(const MapKey& k) {
  uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
  h ^= ((uint64_t)k.rows << 32) ^ (uint64_t)k.d ^ ((uint64_t)k.box_rows << 20) ^ ((uint64_t)k.elem << 12) ^ ((uint64_t)k.ld << 40);
  h *= 0xC2B2AE3D27D4EB4Full;
  return (unsigned)(h >> 40) % kMapSlots;
}
}  // namespace

static int make_map_cached(CUtensorMap* m, const void* base, int rows, int d, int64_t ld, int box_rows, int elem) {
  const MapKey key{base, rows, d, box_rows, elem, ld};
  const unsigned slot = map_hash(key);
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    if (g_maps[slot].used && g_maps[slot].key == key) { *m = g_maps[slot].map; return CFM_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return CFM_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)elem};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem), (cuuint32_t)box_rows};  // 128-byte swizzle atom wide
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d d=%d ld=%lld elem=%d", (int)r, rows, d, (long long)ld, elem);
    return CFM_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(g_map_mu);
  g_maps[slot].key = key; g_maps[slot].map = *m; g_maps[slot].used = true;
  return CFM_OK;
}

int tc_make_map(CUtensorMap* m, const float* base, int rows, int d, int64_t ld, int box_rows) {
  return make_map_cached(m, base, rows, d, ld, box_rows, 4);
}
int tc_make_map_f16(CUtensorMap* m, const __half* base, int rows, int d, int64_t ld, int box_rows) {
  return make_map_cached(m, base, rows, d, ld, box_rows, 2);
}

static unsigned long long* g_tc_dbg = nullptr;
unsigned long long* tc_debug_buffer() { return g_tc_dbg; }

int tc_split(const float* x, float* hi, float* lo, int64_t n, cudaStream_t s) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, hi, lo, n); ::cfm::note_launches(1);
  CFM_CUDA_OK(cudaGetLastError());
  return CFM_OK;
}

int sqdist_tc_supported(int n0, int n1, int d, const float* x0, const float* x1, const float* M, int64_t ldm) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(M);
  return (d % 4 == 0) && (a % 16 == 0) && (ldm % 4 == 0) && n0 > 0 && n1 > 0;
}

size_t sqdist_tc_workspace_bytes(int n0, int n1, int d) {
  return 2 * align_up((size_t)n0 * d * 4, 256) + 2 * align_up((size_t)n1 * d * 4, 256);
}

int sqdist_tc_launch(const float* x0, const float* x1, float* M, int n0, int n1, int d, int64_t ldm, int squared,
                     float* cost_max, const float* nx, const float* ny, void* ws, size_t ws_bytes,
                     cudaStream_t s) {
  CFM_REQUIRE(ws_bytes >= sqdist_tc_workspace_bytes(n0, n1, d), "sqdist tcgen05: workspace too small");
  char* w = reinterpret_cast<char*>(ws);
  const size_t a0 = align_up((size_t)n0 * d * 4, 256), a1 = align_up((size_t)n1 * d * 4, 256);
  float* ah = reinterpret_cast<float*>(w);
  float* al = reinterpret_cast<float*>(w + a0);
  float* bh = reinterpret_cast<float*>(w + 2 * a0);
  float* bl = reinterpret_cast<float*>(w + 2 * a0 + a1);
  int rc;
  if ((rc = tc_split(x0, ah, al, (int64_t)n0 * d, s)) != CFM_OK) return rc;
  if ((rc = tc_split(x1, bh, bl, (int64_t)n1 * d, s)) != CFM_OK) return rc;
  SqDistTcEpilogue epi{M, ldm, nx, ny, cost_max, squared, 0.f, 0.f};
  static int tn = -1;  // CFM_TC_TN=128 selects the 128-wide N tile (3 smem stages) for experiments
  if (tn < 0) { const char* e = getenv("CFM_TC_TN"); tn = e ? atoi(e) : 256; }
  if (tn == 128) return launch_gemm_tc<128>(ah, al, n0, (int64_t)d, bh, bl, n1, (int64_t)d, d, epi, s);
  return launch_gemm_tc<256>(ah, al, n0, (int64_t)d, bh, bl, n1, (int64_t)d, d, epi, s);
}

}  // namespace cfm

// debugging aid: per-CTA globaltimer checkpoints of the next tcgen05 GEMM launches (8 x u64 per CTA,
// buffer of at least 8 * sm_count entries; pass NULL to switch off)
extern "C" int cfm_tc_debug_buffer(unsigned long long* buf) {
  cfm::g_tc_dbg = buf;
  return CFM_OK;
}
