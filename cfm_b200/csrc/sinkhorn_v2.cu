// Sinkhorn V2: smem-staged fused sweep fed by bulk async copies (fast fp32 mode, n1 <= 8192,
// 16-byte aligned rows).  Same algorithm and outputs as sinkhorn.cu (POT sinkhorn_log, reference
// call site torchcfm/optimal_transport.py:87); different data movement:
//
//   * the CTA's row slab is streamed, R rows per stage, into an S-stage shared-memory ring with
//     cp.async.bulk (one contiguous 4*n1-byte copy per row, mbarrier complete_tx).  Thread 0 issues
//     the refill of a stage right after the per-chunk block barrier that proves every warp has
//     finished reading it, so the ring stays S chunks ahead -- across the grid barriers too (the
//     copies for the next sweep are already in flight while the column partials are combined).
//   * all 16 warps are consumers.  Thread t owns the same 4*KG columns in both phases, so v_j and the
//     running column accumulators live in registers for the whole sweep.  Per stage:
//       row phase   x = M*c2 + v_j for its columns -> warp max -> sum ex2 -> 16 per-warp partials
//                   -> one named barrier -> every warp folds the 16 partials -> u_r
//       column phase  re-reads the same smem rows, x = M*c2 + u_r, accumulates into (cm, cs)
//                   with a lazily updated reference maximum (rescale only when x > cm + 40)
//     so each element of M costs one HBM read, two smem reads and two ex2 per iteration.
//   * FACTORED path (taken on the device when span = max|M/reg|*log2(e) <= 40, the regime where
//     kernel-space Sinkhorn is finite, e.g. BASELINE config 2): E_ij = ex2(M_ij*c2 - kappa) is
//     formed ONCE per element into registers; the row phase is s_r += E*V_j with V_j = ex2(v_j - vref)
//     and the column phase c_j += E*U_r with U_r = ex2(u_r - uref): 4 instructions and ONE ex2 per
//     element per iteration instead of ~11 and two, and the smem stage is released as soon as it
//     has been read.  Ranges are safe by construction: LSE is 1-Lipschitz, so the spreads of u and
//     v are <= span and every product stays within 2^(+-3*span) << 2^127.
//   * per-CTA column partials -> [cta][n1] workspace -> grid barrier -> sliced combine -> v_new,
//     marginal error -> grid barrier (identical to sinkhorn.cu).
#include <stdlib.h>

#include "sinkhorn_common.cuh"

namespace cfm {

// NT = threads per CTA (512: one CTA per SM; 256: two CTAs per SM whose barrier domains are
// independent, so one CTA's per-chunk latency chain is hidden behind the other's work)

__device__ __forceinline__ uint32_t v2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void v2_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(v2_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void v2_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(v2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void v2_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(v2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void v2_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(v2_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void v2_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(v2_smem_u32(dst)), "l"(src), "r"(bytes), "r"(v2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void v2_bulk_load_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(v2_smem_u32(dst)), "l"(src), "r"(bytes), "r"(v2_smem_u32(bar)), "l"(policy) : "memory");
}
// fire-and-forget L2 prefetch of a contiguous global range (no shared memory, no completion tracking)
__device__ __forceinline__ void v2_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t v2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t v2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// Minimal grid barrier for the co-resident (cooperatively launched) grid: thread 0 of every CTA
// releases its CTA's global writes, arrives with one atomic and spins with acquire loads until the
// running counter reaches `target` (= barrier ordinal * gridDim.x).  Cheaper than
// cooperative_groups::grid.sync(), which costs ~3 us per call at 148 x 512 threads.
__device__ __forceinline__ void v2_grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}
template <int NT>
__device__ __forceinline__ void v2_consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// Reduce R per-lane values over the warp with a shared shuffle tree: at the stage with offset
// `off` lanes whose bit `off` is set keep the odd member of each pair, the others the even one, so
// R values cost R-1 + log2(32/R) shuffles instead of 5R.  On return v[0] holds, in EVERY lane, the
// warp total of row  multi_row(lane)  (the row's index bits are lane bits 4, 3, 2 for R = 2, 4, 8).
template <int R>
__device__ __forceinline__ void multi_reduce(float (&v)[R], int lane) {
  static_assert(R == 1 || R == 2 || R == 4 || R == 8, "R must be a power of two <= 8");
  int off = 16;
#pragma unroll
  for (int n = R; n > 1; n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float a = v[2 * i], b = v[2 * i + 1];
      const float send = upper ? a : b, keep = upper ? b : a;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}
template <int R>
__device__ __forceinline__ int multi_row(int lane) {
  int r = 0, bit = 4;
#pragma unroll
  for (int n = R, sh = 0; n > 1; n >>= 1, ++sh, --bit) r |= ((lane >> bit) & 1) << sh;
  return r;
}

template <int NT, int KG, int R>
__global__ void __launch_bounds__(NT, 512 / NT) sinkhorn_v2_kernel(const SkParams p, const int S) {
  constexpr int kV2Consumers = NT, kV2Threads = NT, kV2Warps = NT / 32;
  constexpr bool kVSmem = KG > 4;  // 32 columns per thread: keep V_j in thread-private smem slots
  extern __shared__ __align__(128) unsigned char v2_smem[];
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = gridDim.x, b = blockIdx.x;
  const int n0 = p.n0, n1 = p.n1, n1p = p.n1p, ng = n1p / 4;

  // device-side mode selection shared with sinkhorn.cu (exactly one of the two kernels works)
  if (p.run_if != 0) {
    const float cmax = p.cost_max ? __ldg(p.cost_max) : 0.f;
    const bool precise = !((p.normalize ? 1.f : cmax) / p.reg <= 64.f);
    if (precise) return;  // uniform over the whole grid: nobody reaches a grid barrier
  }

  const uint32_t stage_floats = (uint32_t)R * (uint32_t)n1p;
  float* stages = reinterpret_cast<float*>(v2_smem);
  float2* rowpart = reinterpret_cast<float2*>(stages + (size_t)S * stage_floats);  // [2][R][16]
  uint64_t* full = reinterpret_cast<uint64_t*>(rowpart + 2 * R * kV2Warps);
  float* V_s = reinterpret_cast<float*>(full + 8);  // n1p floats when kVSmem
  __shared__ double red[kV2Warps];

  if (tid == 0) {
    for (int s = 0; s < S; ++s) v2_mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  float* u_work = reinterpret_cast<float*>(p.u_work);
  float* v_work[2] = {reinterpret_cast<float*>(p.v_work[0]), reinterpret_cast<float*>(p.v_work[1])};
  float* part_m = reinterpret_cast<float*>(p.part_m);
  float* part_s = reinterpret_cast<float*>(p.part_s);

  const float cmaxv = p.cost_max ? __ldg(p.cost_max) : 1.f;
  const float c2 = -kLog2e / (p.reg * (p.normalize ? cmaxv : 1.f));
  const float loga = -log2f((float)n0), logb = -log2f((float)n1);
  const float kappa = c2 * cmaxv;  // = min_ij M_ij*c2 (most negative exponent), so E = ex2(M*c2 - kappa) >= 1
  const bool fact = (p.cost_max != nullptr) && (-kappa <= 40.f);  // span in log2 units

  const int base = n0 / nblk, rem = n0 % nblk;
  const int r_begin = b * base + min(b, rem);
  const int nrows = base + (b < rem ? 1 : 0);
  const int nchunks = (nrows + R - 1) / R;

  // ---- ring bookkeeping (incremental: no 64-bit div/mod in the hot loop) ----
  // chunk sequence q = 0,1,2,... over ALL sweeps; chunk q sits in stage q % S, slab chunk q % nchunks
  int c_st = 0;             // consumers: stage of the next chunk to consume
  uint32_t c_par = 0;       //            parity to wait for on full[c_st]
  int p_st = 0, p_chunk = 0;  // producer (thread 0): stage / slab chunk of the next chunk to issue
  int p_dir = 1;              // sweeps alternate direction (boustrophedon): the rows a sweep ends on are
                              // the rows the next one starts on, and those are still resident in L2
  int sweep_no = 0;
  unsigned int* gbar = reinterpret_cast<unsigned int*>(p.err_ring + 4);  // zeroed by the launcher
  unsigned int gbar_n = 0;
  // measured: cg::grid.sync() gives an 11.3 us per-iteration floor, the hand-rolled atomic-counter v2_grid_barrier
  // 12.3 us; a flag-array barrier (one word per CTA, warp 0 polling all words) +4.2 us with acquire polls and
  // +11 us with relaxed polls (148 pollers x 148 writers on five cache lines) -- grid.sync() stays
  auto grid_sync = [&]() { (void)gbar; (void)gbar_n; grid.sync(); };
  long long issued = 0, consumed_total = 0;
  bool gvalid[KG];
#pragma unroll
  for (int k = 0; k < KG; ++k) gvalid[k] = (tid + kV2Consumers * k) < ng;
  const int tcol = tid * 4;  // first owned column; group k adds 2048*k
  const bool full_cols = ng == kV2Consumers * KG;  // every thread owns KG valid float4 groups

  // L2 residency: the first `resident` chunks of every slab are loaded evict_last (they stay in the
  // 126 MB L2 from sweep to sweep), the rest evict_first (pure streaming): only the streaming part
  // crosses HBM every iteration.  p.l2_resident_frac <= 0 leaves the default policy.
  const int resident = p.l2_resident_frac > 0.f ? (int)(p.l2_resident_frac * (float)nchunks) : -1;
  const uint64_t pol_last = v2_policy_evict_last(), pol_first = v2_policy_evict_first();
  // thread 0: issue the bulk copies of the next chunk into its (free) stage
  auto issue_next = [&]() {
    const int r0 = r_begin + p_chunk * R;
    const int rv = min(R, r_begin + nrows - r0);
    v2_mbar_expect_tx(&full[p_st], (uint32_t)rv * (uint32_t)n1 * 4u);
    for (int r = 0; r < rv; ++r) {
      float* dst = stages + (size_t)p_st * stage_floats + (size_t)r * n1p;
      const float* src = p.M + (int64_t)(r0 + r) * p.ldm;
      if (resident < 0) v2_bulk_load(dst, src, (uint32_t)n1 * 4u, &full[p_st]);
      else v2_bulk_load_hint(dst, src, (uint32_t)n1 * 4u, &full[p_st], p_chunk < resident ? pol_last : pol_first);
    }
    ++issued;
    p_chunk += p_dir;
    if (p_chunk == nchunks) { p_chunk = nchunks - 1; p_dir = -1; }   // next sweep runs backwards
    else if (p_chunk < 0) { p_chunk = 0; p_dir = 1; }
    if (++p_st == S) p_st = 0;
  };
  if (tid == 0)
    for (int q = 0; q < S; ++q) issue_next();  // initial fill (wraps into the next sweep if S > nchunks)
  // While the grid sits in the two barriers and the combine between two sweeps (~10 us per iteration) the HBM
  // is idle except for the refill of the shared-memory ring (S chunks).  Thread 0 uses that window: at the end
  // of a sweep it prefetches into L2 the p.prefetch_chunks chunks the next sweep will ask for right after the
  // ones already in the ring, so the sweep starts on L2 hits and the DRAM transfer of those rows overlaps the
  // barrier latency instead of the streaming phase.  A hint only: results do not depend on it.
  auto prefetch_ahead = [&]() {
    int c = p_chunk, dir = p_dir;
    for (int q = 0; q < p.prefetch_chunks; ++q) {
      if (c < 0 || c >= nchunks) break;  // the sweep after next is not worth fetching yet
      if (!(resident >= 0 && c < resident)) {  // resident chunks are L2 hits already
        const int r0 = r_begin + c * R;
        const int rv = min(R, r_begin + nrows - r0);
        for (int r = 0; r < rv; ++r) v2_prefetch_l2(p.M + (int64_t)(r0 + r) * p.ldm, (uint32_t)n1 * 4u);
      }
      c += dir;
    }
  };

  // ---- one fused sweep ----
  // v_in (single-barrier mode): this thread's columns of v in registers and the common reference vref_in;
  // acc_out (single-barrier mode): fixed-point column accumulator the partials are added to
  // v_reload (single-barrier mode): where this thread parked its v before the sweep (re-read for the partial scaling
  // at the end, so that 16 registers do not stay live across the streaming loop); null = v is zero (prologue)
  auto sweep = [&](bool do_row, bool do_col, const float* v_cur, const float4* v_in, float vref_in,
                   unsigned long long* acc_out, const float* v_reload) {
    if (fact) {
      // ---------------- factored (kernel-space in registers) sweep ----------------
      const float vref = do_row ? (v_in ? vref_in : __ldcg(v_cur)) : 0.f;
      const float nkap = -kappa;
      float4 V[kVSmem ? 1 : KG];
      float cs[KG][4];
#pragma unroll
      for (int k = 0; k < KG; ++k) {
        float4 vk = make_float4(0.f, 0.f, 0.f, 0.f);  // idle columns: weight 0
        if (gvalid[k]) {
          if (do_row) {
            const float4 vv = v_in ? v_in[k]
                                   : __ldcg(reinterpret_cast<const float4*>(v_cur + tcol + kV2Consumers * 4 * k));
            vk = make_float4(ex2f(vv.x - vref), ex2f(vv.y - vref), ex2f(vv.z - vref), ex2f(vv.w - vref));
          } else {
            vk = make_float4(1.f, 1.f, 1.f, 1.f);
          }
        }
        if (kVSmem) *reinterpret_cast<float4*>(V_s + tcol + kV2Consumers * 4 * k) = vk;  // private slot
        else V[k] = vk;
#pragma unroll
        for (int c = 0; c < 4; ++c) cs[k][c] = 0.f;
      }
      float uref = 0.f;
      bool have_uref = !do_row;  // prologue: u = 0 everywhere, U = 1
      for (int cc = 0; cc < nchunks; ++cc) {
        const int chunk = (sweep_no & 1) ? nchunks - 1 - cc : cc;
        const int r0 = r_begin + chunk * R;
        const int rv = min(R, r_begin + nrows - r0);
        const float* sbase = stages + (uint32_t)c_st * stage_floats + tcol;
        float2* rp = rowpart + (cc & 1) * R * kV2Warps;
        v2_mbar_wait(&full[c_st], c_par);
        float4 E[R][KG];
        if (full_cols && rv == R) {  // common case: no predicates in the hot loop
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < KG; ++k) {
              const float4 mv = *reinterpret_cast<const float4*>(sbase + r * n1p + kV2Consumers * 4 * k);
              E[r][k] = make_float4(ex2f(fmaf(mv.x, c2, nkap)), ex2f(fmaf(mv.y, c2, nkap)),
                                    ex2f(fmaf(mv.z, c2, nkap)), ex2f(fmaf(mv.w, c2, nkap)));
            }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float* srow = sbase + r * n1p;
#pragma unroll
            for (int k = 0; k < KG; ++k) {
              E[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (gvalid[k] && r < rv) {
                const float4 mv = *reinterpret_cast<const float4*>(srow + kV2Consumers * 4 * k);
                E[r][k] = make_float4(ex2f(fmaf(mv.x, c2, nkap)), ex2f(fmaf(mv.y, c2, nkap)),
                                      ex2f(fmaf(mv.z, c2, nkap)), ex2f(fmaf(mv.w, c2, nkap)));
              }
            }
          }
        }
        if (++c_st == S) { c_st = 0; c_par ^= 1u; }
        if (do_row) {
          float sr[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < KG; ++k) {
              const float4 vk = kVSmem ? *reinterpret_cast<const float4*>(V_s + tcol + kV2Consumers * 4 * k)
                                       : V[kVSmem ? 0 : k];
              s = fmaf(E[r][k].x, vk.x, s); s = fmaf(E[r][k].y, vk.y, s);
              s = fmaf(E[r][k].z, vk.z, s); s = fmaf(E[r][k].w, vk.w, s);
            }
            sr[r] = s;
          }
          multi_reduce<R>(sr, lane);  // lane L now holds the warp total of row multi_row<R>(L)
          if ((lane & (32 / R - 1)) == 0) rp[multi_row<R>(lane) * kV2Warps + warp] = make_float2(sr[0], 0.f);
        }
        // every warp has now copied its part of the stage into registers: after this barrier the
        // stage is free, so thread 0 refills it with the chunk S positions ahead
        v2_consumer_barrier<NT>();
        if (tid == 0) issue_next();
        float U[R];
        if (do_row) {
          // fold the 16 per-warp partials: lanes 0-15 take row r, lanes 16-31 row r+1 (one tree for two rows)
#pragma unroll
          for (int rr = 0; rr < R; rr += 2) {
            const int myrow = rr + (lane >> 4);
            float pr = ((R == 1 && lane >= 16) || (lane & 15) >= kV2Warps)
                           ? 0.f : rp[(R == 1 ? 0 : myrow) * kV2Warps + (lane & 15)].x;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, o);
            float Srow[2];
            Srow[0] = __shfl_sync(0xffffffffu, pr, 0);
            Srow[1] = __shfl_sync(0xffffffffu, pr, 16);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = rr + h;
              if (r < R) {
                const float u2 = loga - (kappa + vref + lg2_abs(Srow[h]));  // lse_r = kappa + vref + log2(S_r)
                if (!have_uref) { uref = u2; have_uref = true; }         // first row of the slab fixes uref
                U[r] = r < rv ? ex2f(u2 - uref) : 0.f;  // padding rows: E = 0 and U = 0 (never 0*inf)
                if (warp == 0 && lane == 0 && r < rv) {
                  u_work[r0 + r] = u2;
                  p.log_u[r0 + r] = (double)u2 * kLn2d;
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) U[r] = 1.f;
        }
        if (do_col) {
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < KG; ++k) {
              cs[k][0] = fmaf(E[r][k].x, U[r], cs[k][0]); cs[k][1] = fmaf(E[r][k].y, U[r], cs[k][1]);
              cs[k][2] = fmaf(E[r][k].z, U[r], cs[k][2]); cs[k][3] = fmaf(E[r][k].w, U[r], cs[k][3]);
            }
        }
      }
      consumed_total += nchunks;
      ++sweep_no;
      if (tid == 0) prefetch_ahead();
      if (do_col && acc_out != nullptr) {
        // Single-barrier mode.  Column sum C_j = sum_b cs_bj 2^(kappa + uref_b) and v_new_j = logb - log2 C_j, so
        // with q_bj = cs_bj 2^(kappa + uref_b + v_j - logb):  sum_b q_bj = 2^(v_j - v_new_j)  (= 1 at the fixed
        // point).  The q's are added as 32.32 fixed-point integers: integer addition is associative, so the
        // total -- unlike a floating-point atomic sum -- does not depend on the arrival order of the CTAs.
        const float sh = kappa + uref - logb;
#pragma unroll
        for (int k = 0; k < KG; ++k)
          if (gvalid[k]) {
            const float4 vv = v_reload ? __ldcg(reinterpret_cast<const float4*>(v_reload + tcol + kV2Consumers * 4 * k))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            const float vq[4] = {vv.x, vv.y, vv.z, vv.w};
            unsigned long long* dst = acc_out + tcol + kV2Consumers * 4 * k;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float q = cs[k][c] * ex2f(sh + vq[c]) * 4294967296.f;
              atomicAdd(dst + c, (unsigned long long)__float2ll_rn(fminf(fmaxf(q, 0.f), 4.0e18f)));
            }
          }
      } else if (do_col) {
        // sum_i ex2(M c2 + u_i) over this slab = cs * 2^(kappa + uref): partial (max, sum) form
        const float pm = kappa + uref;
#pragma unroll
        for (int k = 0; k < KG; ++k)
          if (gvalid[k]) {
            *reinterpret_cast<float4*>(part_m + (int64_t)b * n1p + tcol + kV2Consumers * 4 * k) =
                make_float4(pm, pm, pm, pm);
            *reinterpret_cast<float4*>(part_s + (int64_t)b * n1p + tcol + kV2Consumers * 4 * k) =
                make_float4(cs[k][0], cs[k][1], cs[k][2], cs[k][3]);
          }
      }
      return;
    }
    // ---------------- log-domain sweep (span > 40): two smem reads per element ----------------
    float4 vreg[KG];
    float cm[KG][4], cs[KG][4];
#pragma unroll
    for (int k = 0; k < KG; ++k) {
      vreg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (do_row && gvalid[k]) {
        const float* vp = v_cur + tcol + kV2Consumers * 4 * k;  // written before the last grid barrier
        vreg[k] = make_float4(__ldcg(vp), __ldcg(vp + 1), __ldcg(vp + 2), __ldcg(vp + 3));
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) { cm[k][c] = -1.0e30f; cs[k][c] = 0.f; }
    }
    const float inf = __int_as_float(0x7f800000);
    for (int cc = 0; cc < nchunks; ++cc) {
      const int chunk = (sweep_no & 1) ? nchunks - 1 - cc : cc;
      const int r0 = r_begin + chunk * R;
      const int rv = min(R, r_begin + nrows - r0);
      const float* sbase = stages + (uint32_t)c_st * stage_floats + tcol;
      float2* rp = rowpart + (cc & 1) * R * kV2Warps;
      v2_mbar_wait(&full[c_st], c_par);
      if (++c_st == S) { c_st = 0; c_par ^= 1u; }
      float u2[R];
      if (do_row) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float x[KG][4];
          float tmax = -1.0e30f;
#pragma unroll
          for (int k = 0; k < KG; ++k) {
            float4 mv = make_float4(inf, inf, inf, inf);
            if (gvalid[k] && r < rv) mv = *reinterpret_cast<const float4*>(sbase + r * n1p + kV2Consumers * 4 * k);
            x[k][0] = fmaf(mv.x, c2, vreg[k].x); x[k][1] = fmaf(mv.y, c2, vreg[k].y);
            x[k][2] = fmaf(mv.z, c2, vreg[k].z); x[k][3] = fmaf(mv.w, c2, vreg[k].w);
            tmax = fmaxf(tmax, fmaxf(fmaxf(x[k][0], x[k][1]), fmaxf(x[k][2], x[k][3])));
          }
          const float wm = warp_max(tmax);
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < KG; ++k)
            s += (ex2f(x[k][0] - wm) + ex2f(x[k][1] - wm)) + (ex2f(x[k][2] - wm) + ex2f(x[k][3] - wm));
          s = warp_sum(s);
          if (lane == 0) rp[r * kV2Warps + warp] = make_float2(wm, s);
        }
      }
      v2_consumer_barrier<NT>();
#pragma unroll
      for (int r = 0; r < R; ++r) {
        u2[r] = 0.f;
        if (do_row) {
          const float2 pr = lane < kV2Warps ? rp[r * kV2Warps + lane] : make_float2(-1.0e30f, 0.f);
          const float gm = warp_max(pr.x);
          const float gs = warp_sum(pr.y * ex2f(pr.x - gm));
          u2[r] = loga - (gm + log2f(gs));
          if (warp == 0 && lane == 0 && r < rv) {
            u_work[r0 + r] = u2[r];
            p.log_u[r0 + r] = (double)u2[r] * kLn2d;
          }
        }
      }
      if (do_col) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r < rv) {
#pragma unroll
            for (int k = 0; k < KG; ++k) {
              if (gvalid[k]) {
                const float4 mv = *reinterpret_cast<const float4*>(sbase + r * n1p + kV2Consumers * 4 * k);
                const float xv[4] = {fmaf(mv.x, c2, u2[r]), fmaf(mv.y, c2, u2[r]), fmaf(mv.z, c2, u2[r]),
                                     fmaf(mv.w, c2, u2[r])};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  if (xv[c] > cm[k][c] + 40.f) {  // rare: reference maximum lags by more than 2^40
                    cs[k][c] *= ex2f(cm[k][c] - xv[c]);
                    cm[k][c] = xv[c];
                  }
                  cs[k][c] += ex2f(xv[c] - cm[k][c]);
                }
              }
            }
          }
        }
      }
      // the stage is read again by the column phase: it is free only once every warp is past this
      // point, which the NEXT chunk's barrier (or the one below for the last chunk) certifies
      v2_consumer_barrier<NT>();
      if (tid == 0) issue_next();
    }
    consumed_total += nchunks;
    ++sweep_no;
    if (tid == 0) prefetch_ahead();
    if (do_col) {
#pragma unroll
      for (int k = 0; k < KG; ++k)
        if (gvalid[k]) {
          *reinterpret_cast<float4*>(part_m + (int64_t)b * n1p + tcol + kV2Consumers * 4 * k) =
              make_float4(cm[k][0], cm[k][1], cm[k][2], cm[k][3]);
          *reinterpret_cast<float4*>(part_s + (int64_t)b * n1p + tcol + kV2Consumers * 4 * k) =
              make_float4(cs[k][0], cs[k][1], cs[k][2], cs[k][3]);
        }
    }
  };

  // ---- combine a slice of columns over all CTAs' partials ----
  auto combine = [&](const float* v_cur, float* v_new, bool have_cur, double* err_slot) {
    double err_local = 0.0;
    const int cpc = (n1 + nblk - 1) / nblk;
    const int c_begin = b * cpc, c_end = min(n1, c_begin + cpc);
    const int sub = tid & 7;
    for (int j0 = c_begin; j0 < c_end; j0 += (kV2Consumers >> 3)) {  // NT/8 columns per pass
      const int j = j0 + (tid >> 3);
      const bool act = j < c_end;
      float m = -1.0e30f, s = 0.f;
      if (act) {
        if (p.dbg_flags & 1) {  // A/B: previous 4-wide variant
          for (int c0 = sub; c0 < nblk; c0 += 32) {
            float mm[4], ss[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int c = c0 + q * 8;
              if (c < nblk) {
                mm[q] = __ldcg(part_m + (int64_t)c * n1p + j);
                ss[q] = __ldcg(part_s + (int64_t)c * n1p + j);
              } else { mm[q] = -1.0e30f; ss[q] = 0.f; }
            }
            const float bm = fmaxf(fmaxf(fmaxf(mm[0], mm[1]), fmaxf(mm[2], mm[3])), m);
            float acc = s * ex2f(m - bm);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += ss[q] * ex2f(mm[q] - bm);
            s = acc; m = bm;
          }
        } else if (!(p.dbg_flags & 4) && nblk <= 152) {
          // every lane loads ALL its <= 19 (max, sum) pairs before it reduces: ONE L2 round trip per column instead of
          // three dependent ones (the combine sits between the two grid barriers: its latency is fixed cost per iteration)
          float mm[19], ss[19];
#pragma unroll
          for (int q = 0; q < 19; ++q) {
            const int c = sub + q * 8;
            if (c < nblk) {
              mm[q] = __ldcg(part_m + (int64_t)c * n1p + j);
              ss[q] = __ldcg(part_s + (int64_t)c * n1p + j);
            } else { mm[q] = -1.0e30f; ss[q] = 0.f; }
          }
          float bm = m;
#pragma unroll
          for (int q = 0; q < 19; ++q) bm = fmaxf(bm, mm[q]);
          float acc = 0.f;
#pragma unroll
          for (int q = 0; q < 19; ++q) acc += ss[q] * ex2f(mm[q] - bm);
          s = acc; m = bm;
        } else
        for (int c0 = sub; c0 < nblk; c0 += 64) {  // 8 independent (max, sum) pairs in flight per lane
          float mm[8], ss[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = c0 + q * 8;
            if (c < nblk) {
              mm[q] = __ldcg(part_m + (int64_t)c * n1p + j);
              ss[q] = __ldcg(part_s + (int64_t)c * n1p + j);
            } else { mm[q] = -1.0e30f; ss[q] = 0.f; }
          }
          float bm = m;
#pragma unroll
          for (int q = 0; q < 8; ++q) bm = fmaxf(bm, mm[q]);
          float acc = s * ex2f(m - bm);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc += ss[q] * ex2f(mm[q] - bm);
          s = acc; m = bm;
        }
      }
      float gm = m;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
      float gs = s * ex2f(m - gm);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
      if (act && sub == 0) {
        const float vn = logb - (gm + log2f(gs));
        v_new[j] = vn;
        if (have_cur) {
          const double d = ((double)__ldcg(v_cur + j) - (double)vn) * kLn2d;
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
        double t = lane < kV2Warps ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) atomicAdd(err_slot, t);
      }
      __syncthreads();
    }
  };

  // ---- prologue: v^0 from u = 0 ----
  if (b == 0 && tid < 4) p.err_ring[tid] = 0.0;

  int cur = 0, iters = 0;
  bool converged = false;
  double err = 1.0, prev_check_err = -1.0;
  const bool single = fact && p.atomic_cols != 0;
  if (single) {
    // ---- single-barrier iteration: column partials reduced by fixed-point atomics, ONE grid barrier per sweep ----
    // acc[3][n1p] (zeroed by the launcher) lives in the part_s workspace; sweep s adds into acc[s % 3], after the
    // barrier every thread reads the totals of ITS columns and forms v_new itself (no combine phase, no second
    // barrier); buffer (s + 2) % 3 -- last read one barrier ago -- is zeroed by slices for the sweep after next.
    // v is kept per CTA in the part_m workspace row b (each thread re-reads only what it wrote).
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(p.part_s);
    float* vmine = part_m + (int64_t)b * n1p;
    // v used by the sweep that produced the totals being absorbed lives in vmine (this thread's own columns)
#pragma unroll
    for (int k = 0; k < KG; ++k)
      if (gvalid[k]) *reinterpret_cast<float4*>(vmine + tcol + kV2Consumers * 4 * k) = make_float4(0.f, 0.f, 0.f, 0.f);
    float vref_cur = 0.f;
    const int cpc = ((n1 + nblk - 1) / nblk + 3) & ~3;
    sweep(false, true, nullptr, nullptr, 0.f, acc, nullptr);
    grid_sync();
    for (int it = 0;; ++it) {
      const unsigned long long* tot = acc + (size_t)(it % 3) * n1p;
      {  // zero this CTA's slice of the buffer of sweep it + 2
        unsigned long long* z = acc + (size_t)((it + 2) % 3) * n1p;
        for (int j = b * cpc + tid; j < min(n1, (b + 1) * cpc); j += kV2Threads) z[j] = 0ull;
      }
      // totals t_j = 2^(v_j - v_new_j) of this thread's columns (and of column 0, the common reference)
      float4 vnew[KG];
      double e2 = 0.0;
      const float t0 = (float)((double)__ldcg(tot) * 2.3283064365386963e-10);
      const float vref_new = vref_cur - lg2_abs(t0);
#pragma unroll
      for (int k = 0; k < KG; ++k) {
        float4 vcur_k = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gvalid[k]) vcur_k = __ldcg(reinterpret_cast<const float4*>(vmine + tcol + kV2Consumers * 4 * k));
        vnew[k] = vcur_k;
        if (gvalid[k]) {
          const ulonglong2 a = __ldcg(reinterpret_cast<const ulonglong2*>(tot + tcol + kV2Consumers * 4 * k));
          const ulonglong2 c = __ldcg(reinterpret_cast<const ulonglong2*>(tot + tcol + kV2Consumers * 4 * k + 2));
          const double td[4] = {(double)a.x * 2.3283064365386963e-10, (double)a.y * 2.3283064365386963e-10,
                                (double)c.x * 2.3283064365386963e-10, (double)c.y * 2.3283064365386963e-10};
          vnew[k] = make_float4(vcur_k.x - lg2_abs((float)td[0]), vcur_k.y - lg2_abs((float)td[1]),
                                vcur_k.z - lg2_abs((float)td[2]), vcur_k.w - lg2_abs((float)td[3]));
#pragma unroll
          for (int q = 0; q < 4; ++q) e2 += (td[q] - 1.0) * (td[q] - 1.0);
        }
      }
      // stopping rule of iteration it - 1 (POT: every check_every iterations, column marginal of (u, v) in L2 norm):
      // expm1(v_prev - v_new) = t - 1.  Every CTA evaluates it on the same numbers in the same order.
      if (it >= 1 && (((it - 1) % p.check_every) == 0 || (p.dbg_flags & 2))) {
        e2 = warp_sum(e2);
        if (lane == 0) red[warp] = e2;
        __syncthreads();
        double t = 0.0;
        for (int w = 0; w < kV2Warps; ++w) t += red[w];
        __syncthreads();
        err = sqrt(t) / (double)n1;
        if (((it - 1) % p.check_every) == 0) {
          if (err < p.stop_thr) { converged = true; break; }
          if (p.stall_tol > 0.0 && prev_check_err >= 0.0 && err > (1.0 - p.stall_tol) * prev_check_err &&
              err * sqrt((double)n1) < 1e-5) { converged = true; break; }
          prev_check_err = err;
        }
      }
      if (it == p.max_iters) break;  // (only reached when the last iteration ran its column phase for a check)
#pragma unroll
      for (int k = 0; k < KG; ++k)
        if (gvalid[k]) *reinterpret_cast<float4*>(vmine + tcol + kV2Consumers * 4 * k) = vnew[k];
      vref_cur = vref_new;
      const bool last = (it == p.max_iters - 1);
      const bool check = (it % p.check_every) == 0;
      const bool do_col = !last || check;
      sweep(true, do_col, nullptr, vnew, vref_cur, do_col ? acc + (size_t)((it + 1) % 3) * n1p : nullptr, vmine);
      iters = it + 1;
      if (!do_col) break;
      grid_sync();
    }
    // outputs: v of the last sweep (parked in vmine by this very thread), u written by the sweep
    if (b == 0) {
#pragma unroll
      for (int k = 0; k < KG; ++k)
        if (gvalid[k]) {
          const int j = tcol + kV2Consumers * 4 * k;
          const float4 vv = __ldcg(reinterpret_cast<const float4*>(vmine + j));
          p.log_v[j] = (double)vv.x * kLn2d; p.log_v[j + 1] = (double)vv.y * kLn2d;
          p.log_v[j + 2] = (double)vv.z * kLn2d; p.log_v[j + 3] = (double)vv.w * kLn2d;
        }
    }
  } else {
  sweep(false, true, nullptr, nullptr, 0.f, nullptr, nullptr);
  grid_sync();
  combine(nullptr, v_work[0], false, nullptr);
  grid_sync();

  for (int it = 0; it < p.max_iters; ++it) {
    const bool last = (it == p.max_iters - 1);
    const bool check = (it % p.check_every) == 0;
    const bool do_col = !last || check;
    sweep(true, do_col, v_work[cur], nullptr, 0.f, nullptr, nullptr);
    iters = it + 1;
    if (!do_col) break;
    grid_sync();
    if (b == 0 && tid == 0) p.err_ring[(it + 2) & 3] = 0.0;
    const bool want_err = check || (p.dbg_flags & 2);
    combine(v_work[cur], v_work[cur ^ 1], want_err, want_err ? &p.err_ring[it & 3] : nullptr);
    grid_sync();
    if (check) {
      err = sqrt(__ldcg(&p.err_ring[it & 3]));
      if (err < p.stop_thr) { converged = true; break; }
      if (p.stall_tol > 0.0 && prev_check_err >= 0.0 && err > (1.0 - p.stall_tol) * prev_check_err &&
          err * sqrt((double)n1) < 1e-5) { converged = true; break; }
      prev_check_err = err;
    }
    if (last) break;
    cur ^= 1;
  }
  }  // two-barrier path

  // ---- drain speculative prefetches so no bulk copy is in flight when the CTA exits ----
  if (tid == 0) {
    for (long long q = consumed_total; q < issued; ++q) {
      v2_mbar_wait(&full[c_st], c_par);
      if (++c_st == S) { c_st = 0; c_par ^= 1u; }
    }
  }

  if (!single)
    for (int j = b * kV2Threads + tid; j < n1; j += nblk * kV2Threads)
      p.log_v[j] = (double)__ldcg(v_work[cur] + j) * kLn2d;
  if (b == 0 && tid == 0) {
    int flags = 0;
    if (!converged) flags |= CFM_FLAG_NOT_CONVERGED;
    if (!(err == err)) flags |= CFM_FLAG_NONFINITE;
    p.status[0] = flags;
    p.status[1] = iters;
    p.status[2] = 0;
    p.status[3] = fact ? (single ? 4 : 3) : 2;  // kernel variant: 4 = factored + single barrier, 3 = factored, 2 = log-domain
    *p.err_out = err;
  }
}

template <int NT, int KG, int R>
static int v2_launch_t(SkParams& p, cudaStream_t s) {
  auto kern = sinkhorn_v2_kernel<NT, KG, R>;
  constexpr int kPerSm = 512 / NT;
  const size_t stage_bytes = (size_t)R * p.n1p * 4;
  const size_t fixed = (size_t)2 * R * (NT / 32) * sizeof(float2) + 8 * sizeof(uint64_t) + 128 +
                       (KG > 4 ? (size_t)p.n1p * 4 : 0);
  const size_t budget = (kPerSm == 1 ? 220 : 108) * 1024;
  if (budget < fixed + 2 * stage_bytes) return 1;
  int S = (int)((budget - fixed) / stage_bytes);
  if (S > 8) S = 8;
  const size_t smem = (size_t)S * stage_bytes + fixed;
  CFM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CFM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
  if (per_sm < 1) return 1;
  if (per_sm > kPerSm) per_sm = kPerSm;
  int grid = sm_count() * per_sm;
  if (grid > p.n0) grid = p.n0;
  void* args[] = {(void*)&p, (void*)&S};
  CFM_CUDA_OK(cudaMemsetAsync(p.err_ring + 4, 0, 16, s));  // grid-barrier counter
  if (p.atomic_cols)  // the three fixed-point column accumulators of the single-barrier mode
    CFM_CUDA_OK(cudaMemsetAsync(p.part_s, 0, (size_t)3 * p.n1p * sizeof(unsigned long long), s));
  CFM_CUDA_OK(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(NT), args, smem, s));
  note_launches(1);
  return CFM_OK;
}

// returns CFM_OK when launched, 1 when this shape is not covered (caller uses sinkhorn.cu), <0 on error
int sinkhorn_v2_launch(SkParams& p, cudaStream_t s) {
  if (!p.vec || p.n1p > 8192 || p.n1p != p.n1) return 1;
  static float l2frac = -2.f;  // CFM_SK_L2: fraction of each slab kept L2-resident (default below)
  if (l2frac < -1.f) { const char* e = getenv("CFM_SK_L2"); l2frac = e ? (float)atof(e) : 0.15f; }  // measured: 0.1-0.2 best (4.68 ms vs 4.92 ms at 0)
  // only worth it when M does not fit in L2 anyway (it is then fully resident by itself)
  p.l2_resident_frac = ((size_t)p.n0 * p.n1 * 4 > (size_t)100 << 20) ? l2frac : 0.f;
  static int dbg = -1;  // CFM_SK_DBG: A/B switches (bit0: 4-wide combine loop, bit2: 8-wide loop, bit1: marginal error every iteration)
  // same-box A/B at C2 (solve, ms): 4-wide loop (bit 0) 4.469, 8-wide loop (bit 2) 4.468, all <= 19 partials of a lane in one
  // round trip (0, default) 4.448 -- the combine's loads are ~0.2 us of the 10 us between two sweeps; error every iteration +0.9 us
  if (dbg < 0) { const char* e = getenv("CFM_SK_DBG"); dbg = e ? atoi(e) : 0; }
  p.dbg_flags = dbg;
  static int pf = -1;  // CFM_SK_PF: chunks (of R rows) per CTA prefetched into L2 at the end of every sweep
  // measured on B200 at C2 (same box, 10 steps each): 0 -> 4.53 ms, 3 -> 4.60, 6 -> 4.76, 10 -> 5.16, 14 -> 5.61:
  // the prefetched rows displace the evict_last resident part of the slabs from L2, so the default is off
  if (pf < 0) { const char* e = getenv("CFM_SK_PF"); pf = e ? atoi(e) : 0; }
  p.prefetch_chunks = ((size_t)p.n0 * p.n1 * 4 > (size_t)100 << 20) ? pf : 0;  // pointless when M lives in L2 anyway
  static int atom = -1;  // CFM_SK_ATOMIC: 1 = single-barrier iteration (fixed-point atomic column sums), 0 = two barriers + combine
  // measured at C2 on one box: two barriers + sliced combine 4.435 ms, single barrier + 1.2 M 64-bit L2 atomics per
  // iteration 5.62 ms -- the atomics cost three times what the second barrier and the combine do; default off
  if (atom < 0) { const char* e = getenv("CFM_SK_ATOMIC"); atom = e ? atoi(e) : 0; }
  p.atomic_cols = atom;
  static long persist = -2;  // CFM_SK_PERSIST_MB: raise cudaLimitPersistingL2CacheSize before the first solve (experiment)
  if (persist == -2) {
    const char* e = getenv("CFM_SK_PERSIST_MB");
    persist = e ? atol(e) : -1;
    if (persist >= 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)persist << 20);
  }
  static int cfg = -1;  // CFM_SK_CONFIG: 0 heuristic (default), 1 force 512-thread CTAs, 2 force 256-thread CTAs
  if (cfg < 0) { const char* e = getenv("CFM_SK_CONFIG"); cfg = e ? atoi(e) : 0; }
  const int ng = p.n1p / 4;
  // measured on B200 at N=8192: one 512-thread CTA per SM 5.16 ms / 100 it, two 256-thread CTAs 6.42 ms
#ifdef CFM_SK_BUILD_256  // experiment only (measured slower): two 256-thread CTAs per SM
  if (cfg == 2) {
    const int kg = (ng + 255) / 256;
    int rc = 1;
    if (kg <= 1) rc = v2_launch_t<256, 1, 8>(p, s);
    else if (kg <= 2) rc = v2_launch_t<256, 2, 4>(p, s);
    else if (kg <= 4) rc = v2_launch_t<256, 4, 2>(p, s);
    else rc = v2_launch_t<256, 8, 1>(p, s);
    if (rc != 1) return rc;
  }
#endif
  const int kg = (ng + 511) / 512;
  if (kg <= 1) return v2_launch_t<512, 1, 8>(p, s);
  if (kg <= 2) return v2_launch_t<512, 2, 4>(p, s);
  return v2_launch_t<512, 4, 2>(p, s);
}

}  // namespace cfm
