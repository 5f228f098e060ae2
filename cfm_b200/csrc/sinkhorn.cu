// Persistent log-domain Sinkhorn on uniform marginals.
//
// Replaces pot.sinkhorn(a, b, M, reg) as called at torchcfm/optimal_transport.py:87 (bound at
// :51) with POT's sinkhorn_log algorithm (ot/bregman/_sinkhorn.py): per iteration
//     v = logb - LSE_i(Mr_ij + u_i)      (column pass)
//     u = loga - LSE_j(Mr_ij + v_j)      (row pass),        Mr = -M/reg
// with the column-marginal L2 error tested every `check_every` iterations.
//
// B200 design (one cooperative launch for the whole solve, grid = resident CTAs):
//   * CTA b owns a contiguous slab of rows.  One *fused sweep* per iteration: for each chunk of
//     16 rows the CTA first computes the row LSEs (u update, one warp per row, streaming M from
//     HBM), then immediately re-reads the same 16 rows -- now L2 hits -- to add them, with the
//     fresh u, to per-thread running (max, sum) accumulators of the columns it owns.  M therefore
//     crosses HBM ONCE per iteration, not twice (algorithmic bytes 2*N^2*4 per iteration; DRAM
//     traffic ~ half of that).
//   * per-CTA column partials (max, sum) go to a [cta][n1] workspace; after a grid barrier each
//     CTA combines a slice of columns (8 lanes per column), writes the new v and the marginal
//     error, and a second barrier publishes them.
//   * the opposite potential v is staged in shared memory once per sweep (the row phase reads
//     it 16x per chunk); potentials written during the kernel are only ever read through
//     ld.global.cg or smem, never through the non-coherent path.
//   * two arithmetic modes chosen on the device: fast (fp32 log2-domain, one FFMA + one ex2 per
//     element) and precise (float64 potentials, IEEE fp32 division forming -M/reg exactly as
//     NumPy does, for |M/reg| >> 64 where fp32 exponents lose the answer).
#include <stdlib.h>

#include "sinkhorn_common.cuh"

namespace cfm {

// ---- row phase: one warp computes LSE_j(x(M_rj, v_j)) for one row ---------------------------
template <bool P, bool VEC, bool SM, bool MIX>
__device__ __forceinline__ typename Tr<P>::pot_t row_lse(const float* __restrict__ row,
                                                         const typename Tr<P>::pot_t* vsrc,
                                                         int n1, int ng,
                                                         const Xf<P>& xf, int lane) {
  using pot_t = typename Tr<P>::pot_t;
  using sum_t = typename Tr<P>::sum_t;
  pot_t m = Tr<P>::init();
  sum_t s = 0;
  constexpr int U = P ? 4 : 8;
  for (int g0 = 0; g0 < ng; g0 += 32 * U) {
    float4 c[U];
    Vec4<pot_t> pv[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int g = g0 + q * 32 + lane;
      if (g < ng) {
        c[q] = load_cost4<VEC>(row, g * 4, n1);
        pv[q] = load_pot4<SM, pot_t>(vsrc, g * 4);
      } else {
        const float inf = __int_as_float(0x7f800000);
        c[q] = make_float4(inf, inf, inf, inf);
        pv[q].x = pv[q].y = pv[q].z = pv[q].w = (pot_t)0;
      }
    }
    pot_t x[U][4];
    pot_t bm = m;
#pragma unroll
    for (int q = 0; q < U; ++q) {
      x[q][0] = xf(c[q].x, pv[q].x); x[q][1] = xf(c[q].y, pv[q].y);
      x[q][2] = xf(c[q].z, pv[q].z); x[q][3] = xf(c[q].w, pv[q].w);
      bm = vmax(bm, vmax(vmax(x[q][0], x[q][1]), vmax(x[q][2], x[q][3])));
    }
    sum_t acc = s * expd<MIX>(m, bm);
#pragma unroll
    for (int q = 0; q < U; ++q)
      acc += (expd<MIX>(x[q][0], bm) + expd<MIX>(x[q][1], bm)) +
             (expd<MIX>(x[q][2], bm) + expd<MIX>(x[q][3], bm));
    s = acc;
    m = bm;
  }
  // warp combine of (m, s)
  pot_t wm = warp_max(m);
  sum_t ws = warp_sum(s * expd<MIX>(m, wm));
  return lse_fin(wm, ws);
}

// ---- the solver body ---------------------------------------------------------------------------
template <bool P, bool VEC, int KG, bool MIX>
__device__ void sinkhorn_run(const SkParams& p, unsigned char* smem_raw) {
  using pot_t = typename Tr<P>::pot_t;
  using sum_t = typename Tr<P>::sum_t;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = gridDim.x, b = blockIdx.x;
  const int n0 = p.n0, n1 = p.n1, n1p = p.n1p, ng = n1p / 4;

  __shared__ pot_t u_chunk[kSkChunk];
  __shared__ double red[kSkWarps];
  pot_t* v_s = reinterpret_cast<pot_t*>(smem_raw);  // n1p entries when p.v_in_smem
  if (tid < kSkChunk) u_chunk[tid] = (pot_t)0;      // padding rows must never see NaN bits
  __syncthreads();

  pot_t* u_work = reinterpret_cast<pot_t*>(p.u_work);
  pot_t* v_work[2] = {reinterpret_cast<pot_t*>(p.v_work[0]), reinterpret_cast<pot_t*>(p.v_work[1])};
  pot_t* part_m = reinterpret_cast<pot_t*>(p.part_m);
  sum_t* part_s = reinterpret_cast<sum_t*>(p.part_s);

  // element transform and the log-marginals in this mode's units
  Xf<P> xf;
  pot_t loga, logb;
  const float cmax = p.cost_max ? __ldg(p.cost_max) : 1.f;
  if constexpr (P) {
    xf.reg = p.reg; xf.cmax = cmax; xf.norm = p.normalize;
    loga = -log((double)n0); logb = -log((double)n1);
  } else {
    const float scale = p.normalize ? cmax : 1.f;
    xf.c2 = -kLog2e / (p.reg * scale);
    loga = -log2f((float)n0); logb = -log2f((float)n1);
  }
  const double to_ln = P ? 1.0 : kLn2d;  // working units -> natural log

  // slab of rows owned by this CTA (first `rem` CTAs get one extra row)
  const int base = n0 / nblk, rem = n0 % nblk;
  const int r_begin = b * base + min(b, rem);
  const int r_end = r_begin + base + (b < rem ? 1 : 0);
  const int npanel = (n1p + kPanelCols - 1) / kPanelCols;

  // one sweep: (optional) row phase -> u ; (optional) column partials with that u
  auto sweep = [&](bool do_row, bool do_col, const pot_t* v_cur) {
    if (do_row && p.v_in_smem) {
      for (int j = tid; j < n1p; j += kSkThreads) v_s[j] = (j < n1) ? __ldcg(v_cur + j) : (pot_t)0;
      __syncthreads();
    }
    const pot_t* vsrc = p.v_in_smem ? v_s : v_cur;
    for (int panel = 0; panel < npanel; ++panel) {
      pot_t cm[KG][4];
      sum_t cs[KG][4];
#pragma unroll
      for (int k = 0; k < KG; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) { cm[k][c] = Tr<P>::init(); cs[k][c] = 0; }
      const int gpan = panel * (kPanelCols / 4);

      for (int r0 = r_begin; r0 < r_end; r0 += kSkChunk) {
        const int R = min(kSkChunk, r_end - r0);
        // -- row phase (first panel only; later panels re-use u from global) --
        if (panel == 0) {
          if (warp < R) {
            pot_t uval = (pot_t)0;
            if (do_row) {
              const float* row = p.M + (int64_t)(r0 + warp) * p.ldm;
              pot_t lse;
              if (p.v_in_smem) lse = row_lse<P, VEC, true, MIX>(row, vsrc, n1, ng, xf, lane);
              else lse = row_lse<P, VEC, false, MIX>(row, vsrc, n1, ng, xf, lane);
              uval = loga - lse;
              if (lane == 0) {
                u_work[r0 + warp] = uval;
                p.log_u[r0 + warp] = (double)uval * to_ln;
              }
            }
            if (lane == 0) u_chunk[warp] = uval;
          }
        } else if (warp < R && lane == 0) {
          u_chunk[warp] = do_row ? u_work[r0 + warp] : (pot_t)0;
        }
        if (!do_col) continue;
        __syncthreads();
        // -- column phase: thread owns float4 groups g = gpan + tid + 512*k --
#pragma unroll
        for (int k = 0; k < KG; ++k) {
          const int g = gpan + tid + kSkThreads * k;
          if (g >= ng) continue;
          for (int rb = 0; rb < R; rb += 8) {
            float4 c[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (rb + q < R) {
                c[q] = load_cost4<VEC>(p.M + (int64_t)(r0 + rb + q) * p.ldm, g * 4, n1);
              } else {
                const float inf = __int_as_float(0x7f800000);
                c[q] = make_float4(inf, inf, inf, inf);
              }
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              pot_t x[8];
              pot_t bm = cm[k][cc];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float cv = cc == 0 ? c[q].x : cc == 1 ? c[q].y : cc == 2 ? c[q].z : c[q].w;
                x[q] = xf(cv, u_chunk[(rb + q) & (kSkChunk - 1)]);
                bm = vmax(bm, x[q]);
              }
              sum_t acc = cs[k][cc] * expd<MIX>(cm[k][cc], bm);
#pragma unroll
              for (int q = 0; q < 8; ++q) acc += expd<MIX>(x[q], bm);
              cs[k][cc] = acc;
              cm[k][cc] = bm;
            }
          }
        }
        __syncthreads();  // u_chunk is rewritten by the next chunk's row phase
      }
      if (do_col) {
#pragma unroll
        for (int k = 0; k < KG; ++k) {
          const int g = gpan + tid + kSkThreads * k;
          if (g >= ng) continue;
          pot_t* pm = part_m + (int64_t)b * n1p + g * 4;
          sum_t* ps = part_s + (int64_t)b * n1p + g * 4;
#pragma unroll
          for (int c = 0; c < 4; ++c) { pm[c] = cm[k][c]; ps[c] = cs[k][c]; }
        }
      }
    }
  };

  // combine the per-CTA column partials of a slice of columns -> v_new, marginal error
  auto combine = [&](const pot_t* v_cur, pot_t* v_new, bool have_cur, double* err_slot) {
    const int cpc = (n1 + nblk - 1) / nblk;
    const int c_begin = b * cpc, c_end = min(n1, c_begin + cpc);
    double err_local = 0.0;
    const int sub = tid & 7;
    for (int j0 = c_begin; j0 < c_end; j0 += (kSkThreads >> 3)) {
      // every thread takes the same trip count (shuffles below); idle columns contribute nothing
      const int j = j0 + (tid >> 3);
      const bool act = j < c_end;
      pot_t m = Tr<P>::init();
      sum_t s = 0;
      if (act) {
        for (int c0 = sub; c0 < nblk; c0 += 8 * 4) {
          pot_t mm[4];
          sum_t ss[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = c0 + q * 8;
            if (c < nblk) {
              mm[q] = __ldcg(part_m + (int64_t)c * n1p + j);
              ss[q] = __ldcg(part_s + (int64_t)c * n1p + j);
            } else { mm[q] = Tr<P>::init(); ss[q] = 0; }
          }
          pot_t bm = vmax(vmax(mm[0], mm[1]), vmax(mm[2], mm[3]));
          bm = vmax(bm, m);
          sum_t acc = s * expd<MIX>(m, bm);
#pragma unroll
          for (int q = 0; q < 4; ++q) acc += ss[q] * expd<MIX>(mm[q], bm);
          s = acc; m = bm;
        }
      }
      // reduce over the 8 lanes of this column
      pot_t gm = m;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) gm = vmax(gm, __shfl_xor_sync(0xffffffffu, gm, o));
      sum_t gs = s * expd<MIX>(m, gm);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
      if (act && sub == 0) {
        const pot_t lse = lse_fin(gm, gs);
        const pot_t vn = logb - lse;
        v_new[j] = vn;
        if (have_cur) {
          // column marginal of (u, v_cur): exp(v_cur + lse); POT: err = ||colsum - b||_2
          const double d = ((double)__ldcg(v_cur + j) - (double)vn) * to_ln;
          const double e = expm1(d) / (double)n1;
          err_local += e * e;
        }
      }
    }
    if (err_slot != nullptr) {
      err_local = warp_sum(err_local);
      if (lane == 0) red[warp] = err_local;
      __syncthreads();
      if (warp == 0) {
        double t = lane < kSkWarps ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) atomicAdd(err_slot, t);
      }
      __syncthreads();
    }
  };

  // ---- prologue: v^0 = logb - LSE_i(Mr_ij + 0) ----
  if (b == 0 && tid < 4) p.err_ring[tid] = 0.0;
  sweep(false, true, nullptr);
  grid.sync();
  combine(nullptr, v_work[0], false, nullptr);
  grid.sync();

  int cur = 0, iters = 0;
  bool converged = false;
  double err = 1.0, prev_check_err = -1.0;
  for (int it = 0; it < p.max_iters; ++it) {
    const bool last = (it == p.max_iters - 1);
    const bool check = (it % p.check_every) == 0;
    const bool do_col = !last || check;
    sweep(true, do_col, v_work[cur]);
    iters = it + 1;
    if (!do_col) break;
    grid.sync();
    if (b == 0 && tid == 0) p.err_ring[(it + 2) & 3] = 0.0;
    combine(v_work[cur], v_work[cur ^ 1], true, &p.err_ring[it & 3]);
    grid.sync();
    if (check) {
      err = sqrt(__ldcg(&p.err_ring[it & 3]));
      if (err < p.stop_thr) { converged = true; break; }
      if (p.stall_tol > 0.0 && prev_check_err >= 0.0 &&
          err > (1.0 - p.stall_tol) * prev_check_err &&
          err * sqrt((double)n1) < 1e-5) { converged = true; break; }
      prev_check_err = err;
    }
    if (last) break;
    cur ^= 1;
  }

  // ---- outputs ----
  for (int j = b * kSkThreads + tid; j < n1; j += nblk * kSkThreads)
    p.log_v[j] = (double)__ldcg(v_work[cur] + j) * to_ln;
  if (b == 0 && tid == 0) {
    int flags = 0;
    if (!converged) flags |= CFM_FLAG_NOT_CONVERGED;
    if (!(err == err)) flags |= CFM_FLAG_NONFINITE;
    p.status[0] = flags;
    p.status[1] = iters;
    p.status[2] = P ? (MIX ? 2 : 1) : 0;  // 0 fp32, 1 float64, 2 float64 arguments + fp32 exponentials
    p.status[3] = 0;
    *p.err_out = err;
  }
}

template <bool VEC, int KG>
__global__ void __launch_bounds__(kSkThreads, 1) sinkhorn_kernel(const SkParams p) {
  extern __shared__ __align__(16) unsigned char sk_smem[];
  bool precise = p.precise == 1;
  if (p.precise < 0) {
    const float cmax = p.cost_max ? __ldg(p.cost_max) : 0.f;
    const float span = (p.normalize ? 1.f : cmax) / p.reg;
    precise = !(span <= 64.f);
  }
  if (p.run_if == 2 && !precise) return;  // the fast case was taken by sinkhorn_v2_kernel
  if (precise) {
    if (p.mixed) sinkhorn_run<true, VEC, KG, true>(p, sk_smem);
    else sinkhorn_run<true, VEC, KG, false>(p, sk_smem);
  } else {
    sinkhorn_run<false, VEC, KG, false>(p, sk_smem);
  }
}

// ---- plan materialisation: plan_ij = exp(-M_ij/reg + log_u_i + log_v_j) in float64 ----------
__global__ void plan_materialize_kernel(const float* __restrict__ M, int n0, int n1, int64_t ldm,
                                        float reg, const float* cost_max, int normalize,
                                        const double* __restrict__ lu, const double* __restrict__ lv,
                                        double* __restrict__ plan, double* mass, int32_t* status) {
  const float cmax = (normalize && cost_max) ? __ldg(cost_max) : 1.f;
  const int i = blockIdx.y;
  const double ui = lu[i];
  double local = 0.0;
  bool bad = false;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n1; j += gridDim.x * blockDim.x) {
    float m = M[(int64_t)i * ldm + j];
    if (normalize) m = __fdiv_rn(m, cmax);
    const double v = exp((double)(-__fdiv_rn(m, reg)) + ui + lv[j]);
    plan[(int64_t)i * n1 + j] = v;
    local += v;
    bad |= !isfinite(v);
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(mass, local);
  if (bad) atomicOr(status, CFM_FLAG_NONFINITE);
}

struct SkLayout {
  size_t u, v0, v1, pm, ps, ring, total;
};
static SkLayout sk_layout(int n0, int n1, int grid) {
  const size_t n1p = (size_t)(n1 + 3) / 4 * 4;
  SkLayout L;
  size_t o = 0;
  L.u = o; o += align_up((size_t)n0 * 8, 256);
  L.v0 = o; o += align_up(n1p * 8, 256);
  L.v1 = o; o += align_up(n1p * 8, 256);
  L.pm = o; o += align_up((size_t)grid * n1p * 8, 256);
  L.ps = o; o += align_up((size_t)grid * n1p * 8, 256);
  L.ring = o; o += 256;
  L.total = o;
  return L;
}
static int sk_grid_upper() { return sm_count() * 2; }
int sinkhorn_v2_launch(SkParams& p, cudaStream_t s);  // sinkhorn_v2.cu

}  // namespace cfm

using namespace cfm;

extern "C" size_t cfm_sinkhorn_workspace_bytes(int n0, int n1) {
  return sk_layout(n0, n1, sk_grid_upper()).total;
}

template <bool VEC, int KG>
static int sk_launch(SkParams& p, size_t smem, void* workspace, cudaStream_t s) {
  auto kern = sinkhorn_kernel<VEC, KG>;
  CFM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CFM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSkThreads, smem));
  CFM_REQUIRE(per_sm >= 1, "cfm_sinkhorn_log_f32: kernel does not fit on an SM (smem %zu)", smem);
  int grid = sm_count() * (per_sm > 2 ? 2 : per_sm);
  if (grid > p.n0) grid = p.n0;  // at least one row per CTA
  if (grid > sk_grid_upper()) grid = sk_grid_upper();
  const SkLayout L = sk_layout(p.n0, p.n1, sk_grid_upper());
  char* w = reinterpret_cast<char*>(workspace);
  p.u_work = w + L.u;
  p.v_work[0] = w + L.v0;
  p.v_work[1] = w + L.v1;
  p.part_m = w + L.pm;
  p.part_s = w + L.ps;
  p.err_ring = reinterpret_cast<double*>(w + L.ring);
  void* args[] = {(void*)&p};
  CFM_CUDA_OK(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kSkThreads), args, smem, s));
  note_launches(1);
  return CFM_OK;
}

extern "C" int cfm_sinkhorn_log_f32(const float* M, int n0, int n1, int64_t ldm, float reg,
                                       const float* cost_max, int normalize, int max_iters,
                                       double stop_thr, int check_every, int precise,
                                       double stall_tol, double* log_u, double* log_v,
                                       int32_t* status, double* err, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  CFM_REQUIRE(M && log_u && log_v && status && err && workspace, "cfm_sinkhorn_log_f32: null pointer");
  CFM_REQUIRE(n0 > 0 && n1 > 0 && ldm >= n1, "cfm_sinkhorn_log_f32: bad shape n0=%d n1=%d ldm=%lld", n0,
              n1, (long long)ldm);
  CFM_REQUIRE(reg > 0.f, "cfm_sinkhorn_log_f32: reg must be > 0 (got %g)", (double)reg);
  CFM_REQUIRE(max_iters >= 1 && check_every >= 1, "cfm_sinkhorn_log_f32: max_iters/check_every must be >= 1");
  CFM_REQUIRE(!(normalize && !cost_max), "cfm_sinkhorn_log_f32: normalize needs cost_max");
  CFM_REQUIRE(!(precise < 0 && !cost_max), "cfm_sinkhorn_log_f32: precise=-1 (auto) needs cost_max");
  CFM_REQUIRE(workspace_bytes >= cfm_sinkhorn_workspace_bytes(n0, n1),
              "cfm_sinkhorn_log_f32: workspace too small (%zu < %zu)", workspace_bytes,
              cfm_sinkhorn_workspace_bytes(n0, n1));
  SkParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.n0 = n0; p.n1 = n1; p.ldm = ldm; p.reg = reg; p.cost_max = cost_max;
  p.normalize = normalize; p.max_iters = max_iters; p.stop_thr = stop_thr;
  p.check_every = check_every; p.precise = precise; p.stall_tol = stall_tol;
  p.log_u = log_u; p.log_v = log_v; p.status = status; p.err_out = err;
  p.n1p = (n1 + 3) / 4 * 4;
  p.vec = ((n1 & 3) == 0) && ((ldm & 3) == 0) && ((reinterpret_cast<uintptr_t>(M) & 15) == 0);
  const SkLayout L = sk_layout(n0, n1, sk_grid_upper());
  {
    char* w = reinterpret_cast<char*>(workspace);
    p.u_work = w + L.u; p.v_work[0] = w + L.v0; p.v_work[1] = w + L.v1;
    p.part_m = w + L.pm; p.part_s = w + L.ps; p.err_ring = reinterpret_cast<double*>(w + L.ring);
  }
  // precise: -1 auto, 0 fp32, 1 float64, 2 fp32 on the generic kernel, 3 float64 with fp32 exponentials ("mixed")
  static int auto_mixed = -1;  // CFM_SK_MIXED: what auto mode uses when it needs float64 potentials
  if (auto_mixed < 0) { const char* e = getenv("CFM_SK_MIXED"); auto_mixed = e ? atoi(e) : 1; }  // measured at C4 (N=4096, 100 it): 16.5 ms float64, 9.5 ms mixed, same err to 1e-9
  if (precise == 3) { precise = 1; p.precise = 1; p.mixed = 1; }
  else if (precise < 0) p.mixed = auto_mixed;
  // fast fp32 mode on aligned n1 <= 8192: smem-staged kernel (sinkhorn_v2.cu)
  if (precise == 2) { precise = 0; p.precise = 0; }  // fast arithmetic, generic kernel (cross-checks)
  else if (precise != 1) {
    p.run_if = precise < 0 ? 1 : 0;
    const int rc = sinkhorn_v2_launch(p, s);
    if (rc < 0) return rc;
    if (rc == CFM_OK) {
      if (precise == 0) return CFM_OK;
      p.run_if = 2;  // auto: the generic kernel below only runs when float64 potentials are needed
    } else {
      p.run_if = 0;
    }
  }
  const size_t vbytes = (size_t)p.n1p * 8;
  p.v_in_smem = vbytes <= 160 * 1024;
  const size_t smem = p.v_in_smem ? vbytes : 0;
  const int ng = p.n1p / 4;
  const int kg_need = (ng + kSkThreads - 1) / kSkThreads;
  if (p.vec) {
    if (kg_need <= 1) return sk_launch<true, 1>(p, smem, workspace, s);
    if (kg_need <= 2) return sk_launch<true, 2>(p, smem, workspace, s);
    return sk_launch<true, 4>(p, smem, workspace, s);  // n1 > 8192: several column panels
  }
  if (kg_need <= 1) return sk_launch<false, 1>(p, smem, workspace, s);
  if (kg_need <= 2) return sk_launch<false, 2>(p, smem, workspace, s);
  return sk_launch<false, 4>(p, smem, workspace, s);
}

extern "C" int cfm_plan_materialize_f64(const float* M, int n0, int n1, int64_t ldm, float reg,
                                        const float* cost_max, int normalize, const double* log_u,
                                        const double* log_v, double* plan, double* mass,
                                        int32_t* status, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  CFM_REQUIRE(M && log_u && log_v && plan && mass && status, "cfm_plan_materialize_f64: null pointer");
  CFM_REQUIRE(n0 > 0 && n1 > 0 && ldm >= n1, "cfm_plan_materialize_f64: bad shape");
  CFM_REQUIRE(!(normalize && !cost_max), "cfm_plan_materialize_f64: normalize needs cost_max");
  CFM_CUDA_OK(cudaMemsetAsync(mass, 0, sizeof(double), s));
  dim3 grid((n1 + 1023) / 1024 > 8 ? 8 : (n1 + 1023) / 1024, n0);
  plan_materialize_kernel<<<grid, 256, 0, s>>>(M, n0, n1, ldm, reg, cost_max, normalize, log_u, log_v,
                                              plan, mass, status); ::cfm::note_launches(1);
  CFM_CUDA_OK(cudaGetLastError());
  return CFM_OK;
}
