// Fused pair gather + probability-path sample + conditional flow.
//
// Replaces, for one batch of pairs, the reference sequence
//     x0[i], x1[j]                                   torchcfm/optimal_transport.py:145
//     xt = mu_t + sigma_t * eps ; ut = u_t(x1 | x0)   torchcfm/conditional_flow_matching.py:104-154 and the
//                                                     overrides at :329-394 (target), :429-478 (SB), :569-618 (VP)
// -- two gathers and ~8 elementwise passes over (N, d) -- by ONE pass that reads x0[i], x1[j], eps once and
// writes xt, ut once (HBM-bound: 3 reads + 2 writes of N*d*4 B).  The gathered batches are never stored.
//
// Bit-exactness: the per-row scalars (t, 1-t, sigma_t, ...) are computed by the caller with the very torch
// expressions the reference uses; here every per-element operation is an explicitly rounded fp32 op
// (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn: no FMA contraction) in the reference's association order,
// so xt and ut equal the reference's tensors bit for bit.
#include "common.cuh"

namespace cfm {

struct FlowArgs {
  const float* x0;
  const float* x1;
  const int64_t* i_idx;  // nullable: identity pairing
  const int64_t* j_idx;
  const float* eps;
  const float* ra;       // per-row coefficient a (meaning depends on kind)
  const float* rb;       // per-row coefficient b
  const float* rs;       // per-row sigma_t (nullable: scalar `s` is used)
  const float* rc;       // per-row extra coefficient (SB: sigma_t'/sigma_t ; target: denominator)
  float s, k;            // scalar sigma / scalar constant (target: 1 - sigma ; VP: pi / 2)
  float* xt;
  float* ut;
  int64_t n, row;
  int kind;
};

__device__ __forceinline__ void flow_one(const FlowArgs& a, int64_t r, float v0, float v1, float e, float& xt,
                                         float& ut) {
  const float sg = a.rs ? __ldg(a.rs + r) : a.s;
  switch (a.kind) {
    case CFM_FLOW_ICFM: {  // mu = t*x1 + (1-t)*x0 ; xt = mu + sigma*eps ; ut = x1 - x0
      const float mu = __fadd_rn(__fmul_rn(__ldg(a.ra + r), v1), __fmul_rn(__ldg(a.rb + r), v0));
      xt = __fadd_rn(mu, __fmul_rn(sg, e));
      ut = __fsub_rn(v1, v0);
      break;
    }
    case CFM_FLOW_TARGET: {  // mu = t*x1 ; xt = mu + sigma_t*eps ; ut = (x1 - (1-sigma)*xt) / (1 - (1-sigma)*t)
      const float mu = __fmul_rn(__ldg(a.ra + r), v1);
      xt = __fadd_rn(mu, __fmul_rn(sg, e));
      ut = __fdiv_rn(__fsub_rn(v1, __fmul_rn(a.k, xt)), __ldg(a.rc + r));
      break;
    }
    case CFM_FLOW_SB: {  // ut = coef*(xt - mu) + x1 - x0
      const float mu = __fadd_rn(__fmul_rn(__ldg(a.ra + r), v1), __fmul_rn(__ldg(a.rb + r), v0));
      xt = __fadd_rn(mu, __fmul_rn(sg, e));
      ut = __fsub_rn(__fadd_rn(__fmul_rn(__ldg(a.rc + r), __fsub_rn(xt, mu)), v1), v0);
      break;
    }
    default: {  // CFM_FLOW_VP: mu = cos*x0 + sin*x1 ; ut = (pi/2) * (cos*x1 - sin*x0)
      const float c = __ldg(a.ra + r), s = __ldg(a.rb + r);
      const float mu = __fadd_rn(__fmul_rn(c, v0), __fmul_rn(s, v1));
      xt = __fadd_rn(mu, __fmul_rn(sg, e));
      ut = __fmul_rn(a.k, __fsub_rn(__fmul_rn(c, v1), __fmul_rn(s, v0)));
      break;
    }
  }
}

template <bool VEC>
__global__ void flow_pairs_kernel(const FlowArgs a) {
  const int64_t per_row = VEC ? a.row / 4 : a.row;
  const int64_t total = a.n * per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / per_row, c = t - r * per_row;
    const int64_t r0 = a.i_idx ? a.i_idx[r] : r, r1 = a.j_idx ? a.j_idx[r] : r;
    if (VEC) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(a.x0 + r0 * a.row) + c);
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(a.x1 + r1 * a.row) + c);
      const float4 e = __ldg(reinterpret_cast<const float4*>(a.eps + r * a.row) + c);
      float4 xt, ut;
      flow_one(a, r, v0.x, v1.x, e.x, xt.x, ut.x);
      flow_one(a, r, v0.y, v1.y, e.y, xt.y, ut.y);
      flow_one(a, r, v0.z, v1.z, e.z, xt.z, ut.z);
      flow_one(a, r, v0.w, v1.w, e.w, xt.w, ut.w);
      reinterpret_cast<float4*>(a.xt + r * a.row)[c] = xt;
      reinterpret_cast<float4*>(a.ut + r * a.row)[c] = ut;
    } else {
      float xt, ut;
      flow_one(a, r, __ldg(a.x0 + r0 * a.row + c), __ldg(a.x1 + r1 * a.row + c), __ldg(a.eps + r * a.row + c), xt, ut);
      a.xt[r * a.row + c] = xt;
      a.ut[r * a.row + c] = ut;
    }
  }
}

}  // namespace cfm

using namespace cfm;

extern "C" int cfm_flow_pairs_f32(int kind, const float* x0, const float* x1, const int64_t* i_idx,
                                  const int64_t* j_idx, const float* eps, const float* row_a, const float* row_b,
                                  const float* row_sigma, const float* row_c, float sigma, float konst, float* xt,
                                  float* ut, int64_t n, int64_t row_elems, void* stream) {
  CFM_REQUIRE(kind >= CFM_FLOW_ICFM && kind <= CFM_FLOW_VP, "cfm_flow_pairs_f32: unknown kind %d", kind);
  CFM_REQUIRE(n >= 0 && row_elems >= 0, "cfm_flow_pairs_f32: negative size");
  if (n == 0 || row_elems == 0) return CFM_OK;
  CFM_REQUIRE(x0 && x1 && eps && row_a && xt && ut, "cfm_flow_pairs_f32: null pointer");
  CFM_REQUIRE(kind == CFM_FLOW_TARGET || row_b, "cfm_flow_pairs_f32: row_b required for this kind");
  CFM_REQUIRE((kind != CFM_FLOW_TARGET && kind != CFM_FLOW_SB) || row_c, "cfm_flow_pairs_f32: row_c required");
  FlowArgs a{x0, x1, i_idx, j_idx, eps, row_a, row_b, row_sigma, row_c, sigma, konst, xt, ut, n, row_elems, kind};
  const uintptr_t al = reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(x1) |
                       reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(xt) |
                       reinterpret_cast<uintptr_t>(ut);
  const bool vec = (row_elems % 4 == 0) && (al % 16 == 0);
  const int64_t total = n * (vec ? row_elems / 4 : row_elems);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (vec) flow_pairs_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  else flow_pairs_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  note_launches(1);
  CFM_CUDA_OK(cudaGetLastError());
  return CFM_OK;
}
