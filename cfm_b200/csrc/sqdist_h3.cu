// Cost matrix  M = cdist(x0, x1)**2  on the kind::f16 tensor pipe (gemm_h3.cuh, fp16x3 scheme)
// (reference: torchcfm/optimal_transport.py:84-86).
//
// One pre-pass per input (one warp per row): |x_r|^2 in fp32, the row's absolute maximum, a power-of-two
// row scale s_r that puts max|x_r| * s_r in [2^13, 2^14) (exact; fp16 overflows at 65504 and the data range
// is the caller's), and the (hi, lo) fp16 split of the scaled row.  The GEMM epilogue undoes the scales --
// dot_ij = (acc0 + acc1 2^-11) / (s_i t_j), again exact powers of two -- and applies
// |x0_i|^2 + |x1_j|^2 - 2 dot, clamp, sqrt, square, running max exactly like the other two paths.
#include "gemm_h3.cuh"

namespace cfm {

// one warp per row; d % 4 == 0, x 16-byte aligned, ldo % 8 == 0.  `uniform_absmax` (nullable, device): scale every
// row by the power of two derived from THIS value instead of the row's own maximum (the MLP weights: one scale per
// layer, so the epilogue multiplies by a kernel-uniform constant)
__global__ void prep_rows_h3_kernel(const float* __restrict__ X, int rows, int d, __half* __restrict__ hi,
                                    __half* __restrict__ lo, int64_t ldo, float* __restrict__ sqnorm,
                                    float* __restrict__ inv_scale, const float* __restrict__ uniform_absmax) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(X + (int64_t)warp * d);
  const int n4 = d >> 2;
  float s = 0.f, am = 0.f;
  for (int k = lane; k < n4; k += 32) {
    const float4 v = __ldg(p + k);
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    am = fmaxf(am, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  s = warp_sum(s);
  am = warp_max(am);
  if (uniform_absmax != nullptr) am = __ldg(uniform_absmax);
  // scale = 2^(13 - floor(log2(am))); non-finite or zero rows keep scale 1 (their products are what they are)
  float sc = 1.f;
  if (am > 0.f && am < 3.0e38f) {
    int e = ilogbf(am);
    e = max(-100, min(100, 13 - e));
    sc = scalbnf(1.f, e);
  }
  if (lane == 0) {
    if (sqnorm) sqnorm[warp] = s;
    inv_scale[warp] = 1.f / sc;  // exact
  }
  __half* h = hi + (int64_t)warp * ldo;
  __half* l = lo + (int64_t)warp * ldo;
  for (int k = lane; k < n4; k += 32) {
    const float4 v = __ldg(p + k);  // second pass: L1-resident (a row is d*4 bytes)
    __half hh[4], ll[4];
    split_h3(v.x * sc, hh[0], ll[0]); split_h3(v.y * sc, hh[1], ll[1]);
    split_h3(v.z * sc, hh[2], ll[2]); split_h3(v.w * sc, hh[3], ll[3]);
    *reinterpret_cast<uint2*>(h + 4 * k) = make_uint2(
        (uint32_t)__half_as_ushort(hh[0]) | ((uint32_t)__half_as_ushort(hh[1]) << 16),
        (uint32_t)__half_as_ushort(hh[2]) | ((uint32_t)__half_as_ushort(hh[3]) << 16));
    *reinterpret_cast<uint2*>(l + 4 * k) = make_uint2(
        (uint32_t)__half_as_ushort(ll[0]) | ((uint32_t)__half_as_ushort(ll[1]) << 16),
        (uint32_t)__half_as_ushort(ll[2]) | ((uint32_t)__half_as_ushort(ll[3]) << 16));
  }
}

// max |X| over a (rows x d) matrix into *out (zeroed by the caller)
__global__ void absmax_h3_kernel(const float* __restrict__ X, int64_t n, float* __restrict__ out) {
  float am = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = fabsf(X[i]);
    am = (v <= 3.0e38f) ? fmaxf(am, v) : am;  // non-finite entries do not set the scale
  }
  am = warp_max(am);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(out, am);
}

int prep_rows_h3(const float* X, int rows, int d, __half* hi, __half* lo, int64_t ldo, float* sqnorm,
                 float* inv_scale, cudaStream_t s, float* uniform_absmax) {
  if (uniform_absmax != nullptr) {
    CFM_CUDA_OK(cudaMemsetAsync(uniform_absmax, 0, sizeof(float), s));
    const int64_t n = (int64_t)rows * d;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 1024) blocks = 1024;
    absmax_h3_kernel<<<blocks, 256, 0, s>>>(X, n, uniform_absmax);
    ::cfm::note_launches(1);
  }
  prep_rows_h3_kernel<<<(rows + 7) / 8, 256, 0, s>>>(X, rows, d, hi, lo, ldo, sqnorm, inv_scale, uniform_absmax);
  ::cfm::note_launches(1);
  CFM_CUDA_OK(cudaGetLastError());
  return CFM_OK;
}

struct SqDistH3Epilogue {
  float* M;
  int64_t ldm;
  const float* nx;   // |x0_i|^2
  const float* ny;   // |x1_j|^2
  const float* isx;  // 1 / row scale of x0
  const float* isy;  // 1 / row scale of x1
  float* cost_max;
  int squared;
  float a, ia, tmax;
  __device__ __forceinline__ void begin_row(int row, bool ok) {
    a = ok ? __ldg(nx + row) : 0.f;
    ia = ok ? __ldg(isx + row) : 0.f;
  }
  __device__ __forceinline__ float one(float acc, float b, float ib) const {
    const float dot = (acc * ia) * ib;  // exact rescale
    const float d2 = (a + b) - 2.f * dot;
    const float v = d2 < 0.f ? 0.f : d2;  // clamp_min(0) that lets NaN through, like ATen's
    // sqrt-then-square mimics cdist(...)**2; on this path the contraction itself is good to ~1e-6 of the norms,
    // so the 2-ulp approximate square root (one MUFU, no IEEE fix-up subroutine) does not show
    float s;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(v));
    return squared ? s * s : s;
  }
  __device__ __forceinline__ void store32(int row0, int lane, int col0, const float (&r)[32], int n0, int n1,
                                          float* tile) {
    const int row = row0 + lane;
    const bool ok = row < n0;
    if (col0 + 32 <= n1) {
      float o[32];
      float m = 0.f;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ny + col0 + c));
        const float4 ib = __ldg(reinterpret_cast<const float4*>(isy + col0 + c));
        o[c] = one(r[c], b.x, ib.x); o[c + 1] = one(r[c + 1], b.y, ib.y);
        o[c + 2] = one(r[c + 2], b.z, ib.z); o[c + 3] = one(r[c + 3], b.w, ib.w);
        m = fmaxf(m, fmaxf(fmaxf(o[c], o[c + 1]), fmaxf(o[c + 2], o[c + 3])));
      }
      if (ok) tmax = fmaxf(tmax, m);
      tc_store_chunk32(tile, o, M + (int64_t)row0 * ldm + col0, ldm, n0 - row0, lane);
    } else if (ok) {
      float* dst = M + (int64_t)row * ldm + col0;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (col0 + c < n1) {
          const float v = one(r[c], __ldg(ny + col0 + c), __ldg(isy + col0 + c));
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

struct H3Ws { size_t ah, al, bh, bl, isx, isy, total; int64_t ld; };
static H3Ws h3_ws(int n0, int n1, int d) {
  H3Ws w;
  w.ld = (d + 7) / 8 * 8;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  w.ah = take((size_t)n0 * w.ld * 2); w.al = take((size_t)n0 * w.ld * 2);
  w.bh = take((size_t)n1 * w.ld * 2); w.bl = take((size_t)n1 * w.ld * 2);
  w.isx = take((size_t)n0 * 4); w.isy = take((size_t)n1 * 4);
  w.total = o;
  return w;
}

size_t sqdist_h3_workspace_bytes(int n0, int n1, int d) { return h3_ws(n0, n1, d).total; }

// nx, ny are written here (the fused pre-pass computes the norms as well)
int sqdist_h3_launch(const float* x0, const float* x1, float* M, int n0, int n1, int d, int64_t ldm, int squared,
                     float* cost_max, float* nx, float* ny, void* ws, size_t ws_bytes, cudaStream_t s) {
  const H3Ws W = h3_ws(n0, n1, d);
  CFM_REQUIRE(ws_bytes >= W.total, "sqdist fp16x3: workspace too small (%zu < %zu)", ws_bytes, W.total);
  char* w = reinterpret_cast<char*>(ws);
  __half* ah = reinterpret_cast<__half*>(w + W.ah);
  __half* al = reinterpret_cast<__half*>(w + W.al);
  __half* bh = reinterpret_cast<__half*>(w + W.bh);
  __half* bl = reinterpret_cast<__half*>(w + W.bl);
  float* isx = reinterpret_cast<float*>(w + W.isx);
  float* isy = reinterpret_cast<float*>(w + W.isy);
  int rc;
  if ((rc = prep_rows_h3(x0, n0, d, ah, al, W.ld, nx, isx, s, nullptr)) != CFM_OK) return rc;
  if ((rc = prep_rows_h3(x1, n1, d, bh, bl, W.ld, ny, isy, s, nullptr)) != CFM_OK) return rc;
  SqDistH3Epilogue epi{M, ldm, nx, ny, isx, isy, cost_max, squared, 0.f, 0.f, 0.f};
  // N tile: 128 (accumulator pairs double-buffered: the read-out of a tile overlaps the MMAs of the next one) or
  // 256 (half the B-operand traffic, but all 512 TMEM columns in use, read-out exposed).  Measured at C2 on the
  // same box: 128 -> 0.307 ms for the cost stage, 256 -> 0.337 ms.  CFM_H3_TN overrides.
  static int tn = -1;
  if (tn < 0) { const char* e = getenv("CFM_H3_TN"); tn = e ? atoi(e) : 128; }
  if (tn == 128) return launch_gemm_h3<128>(ah, al, n0, W.ld, bh, bl, n1, W.ld, d, epi, s);
  return launch_gemm_h3<256>(ah, al, n0, W.ld, bh, bl, n1, W.ld, d, epi, s);
}

}  // namespace cfm
