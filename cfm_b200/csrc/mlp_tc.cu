// tcgen05 path of the vector-field MLP  y = net(cat([x, t], 1))
// (reference: torchcfm/models/models.py:20-21 + torchcfm/utils.py:51-52).
//
// Four launches of the 3xTF32 tensor-core GEMM core (gemm_tc.cuh), one per Linear layer, with the
// bias (+ t * W0[:, -1] for the first layer), the SELU/SiLU activation AND the TF32 hi/lo split of
// the result fused into the TMEM read-out: hidden activations never exist as plain fp32 in memory,
// they are written once as the (hi, lo) operand pair the next layer's TMA loads consume (10 MB per
// layer at B = 10 000: they stay in the 126 MB L2).  Weights are split once per weight set by
// cfm_mlp_prepare into the blob's tensor-core section.
#include "gemm_tc.cuh"
#include "mlp_common.cuh"

namespace cfm {

constexpr int kMlpTN = 128;  // N tile of the MLP layers: 2x the tiles of the 256-wide one, 3 smem stages

struct MlpTcEpilogue {
  const float* bias;
  const float* tcol;   // nullable: + t * tcol[col]
  const float* t_dev;  // nullable device scalar
  float t_host;
  int act;             // -1: none
  float* out;          // nullable: plain fp32 result (last layer)
  float* out_hi;       // nullable: TF32 split of the result (hidden layers)
  float* out_lo;
  int64_t ldo;
  float t;
  __device__ __forceinline__ void begin_row(int, bool) { t = tcol ? (t_dev ? __ldg(t_dev) : t_host) : 0.f; }
  __device__ __forceinline__ float one(float acc, float b, float tc) const {
    float v = acc + b;
    if (tcol) v = fmaf(t, tc, v);
    return act >= 0 ? act_apply_fast(v, act) : v;
  }
  __device__ __forceinline__ void store32(int row0, int lane, int col0, const uint32_t (&r)[32], int n0, int n1,
                                          float* tile) {
    const int row = row0 + lane;
    if ((col0 + 32 <= n1) && ((ldo & 3) == 0)) {
      float o[32];
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + c));
        float4 tc4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tcol) tc4 = __ldg(reinterpret_cast<const float4*>(tcol + col0 + c));
        o[c] = one(__uint_as_float(r[c]), b.x, tc4.x); o[c + 1] = one(__uint_as_float(r[c + 1]), b.y, tc4.y);
        o[c + 2] = one(__uint_as_float(r[c + 2]), b.z, tc4.z); o[c + 3] = one(__uint_as_float(r[c + 3]), b.w, tc4.w);
      }
      const int64_t base = (int64_t)row0 * ldo + col0;
      const int rows_valid = n0 - row0;
      if (out) tc_store_chunk32(tile, o, out + base, ldo, rows_valid, lane);
      if (out_hi) {
        float l[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) { float h; split_tf32(o[c], h, l[c]); o[c] = h; }
        tc_store_chunk32(tile, o, out_hi + base, ldo, rows_valid, lane);
        tc_store_chunk32(tile, l, out_lo + base, ldo, rows_valid, lane);
      }
    } else if (row < n0) {
      const int64_t base = (int64_t)row * ldo + col0;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (col0 + c < n1) {
          const float v = one(__uint_as_float(r[c]), __ldg(bias + col0 + c), tcol ? __ldg(tcol + col0 + c) : 0.f);
          if (out) out[base + c] = v;
          if (out_hi) { float h, l; split_tf32(v, h, l); out_hi[base + c] = h; out_lo[base + c] = l; }
        }
    }
  }
  __device__ __forceinline__ void finish(int) {}
};

struct TcBlob {  // offsets (bytes) inside the blob's tensor-core section
  size_t w0h, w0l, w1h, w1l, w2h, w2l, w3h, w3l, total;
};
static TcBlob tc_blob(int dimp, int w, int out_dim) {
  TcBlob b;
  size_t o = 0;
  auto take = [&](size_t floats) { size_t r = o; o += align_up(floats * 4, 256); return r; };
  b.w0h = take((size_t)w * dimp); b.w0l = take((size_t)w * dimp);
  b.w1h = take((size_t)w * w); b.w1l = take((size_t)w * w);
  b.w2h = take((size_t)w * w); b.w2l = take((size_t)w * w);
  b.w3h = take((size_t)out_dim * w); b.w3l = take((size_t)out_dim * w);
  b.total = o;
  return b;
}

size_t mlp_tc_blob_bytes(int dim, int w, int out_dim) { return tc_blob((dim + 3) / 4 * 4, w, out_dim).total; }

int mlp_tc_prepare(const MlpBlobHeader& h, void* blob, cudaStream_t s) {
  char* B = reinterpret_cast<char*>(blob);
  char* T = B + h.off_tc;
  const TcBlob tb = tc_blob(h.dimp, h.w, h.out_dim);
  auto F = [&](int64_t off) { return reinterpret_cast<const float*>(B + off); };
  auto G = [&](size_t off) { return reinterpret_cast<float*>(T + off); };
  int rc;
  if ((rc = tc_split(F(h.off_w0x), G(tb.w0h), G(tb.w0l), (int64_t)h.w * h.dimp, s)) != CFM_OK) return rc;
  if ((rc = tc_split(F(h.off_w1), G(tb.w1h), G(tb.w1l), (int64_t)h.w * h.w, s)) != CFM_OK) return rc;
  if ((rc = tc_split(F(h.off_w2), G(tb.w2h), G(tb.w2l), (int64_t)h.w * h.w, s)) != CFM_OK) return rc;
  if ((rc = tc_split(F(h.off_w3), G(tb.w3h), G(tb.w3l), (int64_t)h.out_dim * h.w, s)) != CFM_OK) return rc;
  return CFM_OK;
}

int mlp_tc_supported(int batch, int dim, int w, int out_dim) {
  // 16-byte aligned rows for the TMA maps and float4 stores; tiny problems stay on the SIMT path
  return batch >= 128 && dim >= 32 && (dim % 4 == 0) && (w % 4 == 0) && w >= 32 && (out_dim % 4 == 0);
}

struct TcWs { size_t xh, xl, ah, al, bh, bl, total; };
static TcWs tc_ws(int batch, int dim, int w) {
  TcWs t;
  size_t o = 0;
  auto take = [&](size_t floats) { size_t r = o; o += align_up(floats * 4, 256); return r; };
  t.xh = take((size_t)batch * dim); t.xl = take((size_t)batch * dim);
  t.ah = take((size_t)batch * w); t.al = take((size_t)batch * w);
  t.bh = take((size_t)batch * w); t.bl = take((size_t)batch * w);
  t.total = o;
  return t;
}
size_t mlp_tc_workspace_bytes(int batch, int dim, int w, int) { return tc_ws(batch, dim, w).total; }

int mlp_tc_forward(const MlpBlobHeader& h, const void* blob, const float* x, const float* x_hi,
                   const float* x_lo, int batch, const float* t_dev, float t_host, int act, float* y, void* ws,
                   size_t ws_bytes, cudaStream_t s) {
  const TcWs W = tc_ws(batch, h.dim, h.w);
  CFM_REQUIRE(ws_bytes >= W.total, "mlp tcgen05: workspace too small (%zu < %zu)", ws_bytes, W.total);
  CFM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_hi) | reinterpret_cast<uintptr_t>(x_lo) |
                reinterpret_cast<uintptr_t>(y)) & 15) == 0, "mlp tcgen05: x and y must be 16-byte aligned");
  const char* B = reinterpret_cast<const char*>(blob);
  const char* T = B + h.off_tc;
  const TcBlob tb = tc_blob(h.dimp, h.w, h.out_dim);
  char* w = reinterpret_cast<char*>(ws);
  auto F = [&](int64_t off) { return reinterpret_cast<const float*>(B + off); };
  auto G = [&](size_t off) { return reinterpret_cast<const float*>(T + off); };
  auto Wp = [&](size_t off) { return reinterpret_cast<float*>(w + off); };
  int rc;
  const float* xh = x_hi;
  const float* xl = x_lo;
  if (x_hi == nullptr) {  // plain fp32 input: split it here; otherwise the caller already did
    if ((rc = tc_split(x, Wp(W.xh), Wp(W.xl), (int64_t)batch * h.dim, s)) != CFM_OK) return rc;
    xh = Wp(W.xh); xl = Wp(W.xl);
  }
  MlpTcEpilogue e0{F(h.off_b0), h.time_varying ? F(h.off_w0t) : nullptr, t_dev, t_host, act, nullptr,
                   Wp(W.ah), Wp(W.al), (int64_t)h.w, 0.f};
  if ((rc = launch_gemm_tc<kMlpTN>(xh, xl, batch, (int64_t)h.dim, G(tb.w0h), G(tb.w0l), h.w,
                           (int64_t)h.dimp, h.dim, e0, s)) != CFM_OK) return rc;
  MlpTcEpilogue e1{F(h.off_b1), nullptr, nullptr, 0.f, act, nullptr, Wp(W.bh), Wp(W.bl), (int64_t)h.w, 0.f};
  if ((rc = launch_gemm_tc<kMlpTN>(Wp(W.ah), Wp(W.al), batch, (int64_t)h.w, G(tb.w1h), G(tb.w1l), h.w, (int64_t)h.w,
                           h.w, e1, s)) != CFM_OK) return rc;
  MlpTcEpilogue e2{F(h.off_b2), nullptr, nullptr, 0.f, act, nullptr, Wp(W.ah), Wp(W.al), (int64_t)h.w, 0.f};
  if ((rc = launch_gemm_tc<kMlpTN>(Wp(W.bh), Wp(W.bl), batch, (int64_t)h.w, G(tb.w2h), G(tb.w2l), h.w, (int64_t)h.w,
                           h.w, e2, s)) != CFM_OK) return rc;
  MlpTcEpilogue e3{F(h.off_b3), nullptr, nullptr, 0.f, -1, y, nullptr, nullptr, (int64_t)h.out_dim, 0.f};
  return launch_gemm_tc<kMlpTN>(Wp(W.ah), Wp(W.al), batch, (int64_t)h.w, G(tb.w3h), G(tb.w3l), h.out_dim, (int64_t)h.w,
                        h.w, e3, s);
}

}  // namespace cfm
