// tcgen05 GEMM core, fp16x3 scheme:  D[m, n] = sum_k A[m, k] * B[n, k]  at fp32-grade accuracy on the
// kind::f16 tensor pipe (twice the kind::tf32 rate, half the operand bytes).
//
// Every fp32 operand x is carried as two fp16 arrays,  hi = fp16(x)  and  lo = fp16((x - hi) * 2^11)
// (round to nearest: |x - hi| <= 2^-11 |x|, so the scaled residual has hi's magnitude and never goes
// subnormal before x itself is ~1e-8; x = hi + lo * 2^-11 to 2^-22 relative).  Products of fp16 numbers
// are exact in fp32, so
//     acc0 += A_hi . B_hi                     (TMEM accumulator 0)
//     acc1 += A_hi . B_lo + A_lo . B_hi       (TMEM accumulator 1, carries the common factor 2^11)
//     D     = acc0 + acc1 * 2^-11             (epilogue)
// drops only the lo.lo term (2^-22 relative): the NumPy emulation of the representation
// (tests/test_host.py::test_fp16x3_operand_split_emulation) gives 1.5e-8 of sum|a||b| at d = 784,
// 3xTF32 with truncating splits 7e-8.  Range: fp16 tops out at 65504, so producers scale rows by a power of
// two where the data is unbounded (cost matrix: per-row scales, exact) and saturate otherwise.
//
// Same persistent warp-specialised structure as gemm_tc.cuh (warp 0 TMA producer, warp 1 single-thread MMA
// issuer, warps 2-9 epilogue), with
//   * K-chunk = 64 fp16 (= the 128-byte swizzle atom), 4 k-steps of K = 16 per chunk, 3 MMAs per k-step;
//   * a stage = A_hi | A_lo (128 x 64 fp16 each) | B_hi | B_lo (TN x 64 fp16 each): 96 KB at TN = 256
//     (2 stages), 64 KB at TN = 128 (3 stages);
//   * TWO accumulators per tile: TN = 256 fills all 512 TMEM columns (single-buffered: the epilogue of a
//     tile is exposed, but the TMA ring keeps prefetching the next tile's stages meanwhile), TN = 128 is
//     double-buffered.
// Operand bytes per MMA-FLOP are half those of the 3xTF32 core, which was L2-feed-bound.
#pragma once
#include <cuda_fp16.h>

#include "gemm_tc.cuh"

namespace cfm {

constexpr int kHK = 64;                    // K-chunk in fp16 elements
constexpr int kHABytes = kTM * kHK * 2;    // 16 KB: one A (hi or lo) tile
constexpr float kH3Scale = 2048.f;         // 2^11: scale of the lo parts
constexpr float kH3InvScale = 1.f / 2048.f;

template <int TN> struct H3Cfg {
  static constexpr int kBBytes = TN * kHK * 2;
  static constexpr int kStageBytes = 2 * kHABytes + 2 * kBBytes;
  static constexpr int kStages = TN == 256 ? 2 : 3;
  static constexpr int kAccBufs = TN == 256 ? 1 : 2;      // (acc0, acc1) pairs resident in TMEM
  static constexpr uint32_t kTmemCols = 512;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 256 + 8 * 4096;
  // kind::f16: fp16 x fp16 -> fp32 accumulate, A and B K-major, M = 128, N = TN
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
};

__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// fp32 -> (hi, lo) of the fp16x3 scheme; v must already carry the producer's range scale
__device__ __forceinline__ void split_h3(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn((v - __half2float(hi)) * kH3Scale);
}
// saturating variant for unscaled producers (activations): |v| > 65504 clamps instead of becoming inf
// (cvt.rn.satfinite: one instruction per conversion)
__device__ __forceinline__ __half f2h_satfinite(float v) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
}
__device__ __forceinline__ void split_h3_sat(float v, __half& hi, __half& lo) {
  hi = f2h_satfinite(v);
  lo = f2h_satfinite((v - __half2float(hi)) * kH3Scale);
}

// two values at once: packed (hi0 | hi1 << 16) and (lo0 | lo1 << 16) via cvt.rn.satfinite.f16x2.f32
__device__ __forceinline__ void split_h3_sat_x2(float v0, float v1, uint32_t& hi2, uint32_t& lo2) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(v1), "f"(v0));  // first source -> upper half
  const __half2 h = *reinterpret_cast<const __half2*>(&hi2);
  const float r0 = (v0 - __low2float(h)) * kH3Scale, r1 = (v1 - __high2float(h)) * kH3Scale;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(r1), "f"(r0));
}

// Epilogue functor interface (called by whole warps, thread = row):
//   begin_row(row, ok)
//   store32(row0, lane, col0, const float (&acc)[32], n_rows, n_cols, tile)   acc = acc0 + acc1 * 2^-11
//   finish(lane)
template <int TN, class Epi>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_h3_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
               const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
               const TcShape p, Epi epi) {
  constexpr int kTN = TN, kStages = H3Cfg<TN>::kStages, kBBytes = H3Cfg<TN>::kBBytes;
  constexpr int kStageBytes = H3Cfg<TN>::kStageBytes, kAccBufs = H3Cfg<TN>::kAccBufs;
  constexpr uint32_t kTmemCols = H3Cfg<TN>::kTmemCols, kIdesc = H3Cfg<TN>::kIdesc;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage_base = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;   // [2] accumulator pair ready
  uint64_t* tempty = tfull + 2;        // [2] accumulator pair drained
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* epi_tiles = reinterpret_cast<float*>(smem + kStages * kStageBytes + 256);  // [kEpiWarps][32 * 32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TC_MARK(0);
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int nk = (p.d + kHK - 1) / kHK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 32 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) TC_MARK(1);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tm = t % p.tiles_m, tn = t / p.tiles_m;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sb = stage_base + stage * kStageBytes;
          mbar_expect_tx(&full[stage], kStageBytes);
          tma_load_2d(sb, &map_ah, &full[stage], kc * kHK, tm * kTM);
          tma_load_2d(sb + kHABytes, &map_al, &full[stage], kc * kHK, tm * kTM);
          tma_load_2d(sb + 2 * kHABytes, &map_bh, &full[stage], kc * kHK, tn * kTN);
          tma_load_2d(sb + 2 * kHABytes + kBBytes, &map_bl, &full[stage], kc * kHK, tn * kTN);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator pair
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(acc * 2 * kTN);  // hi.hi
      const uint32_t d1 = d0 + (uint32_t)kTN;                      // cross terms (x 2^11)
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0 && kc == 0 && t == (int)blockIdx.x) TC_MARK(2);
        if (lane == 0) {
          const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
          const uint64_t ah = umma_desc_sw128(sa), al = umma_desc_sw128(sa + kHABytes);
          const uint64_t bh = umma_desc_sw128(sa + 2 * kHABytes), bl = umma_desc_sw128(sa + 2 * kHABytes + kBBytes);
#pragma unroll
          for (int k = 0; k < kHK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 16 * 2) >> 4);  // 32 bytes per k-step, 16-byte units
            const uint32_t first = (kc | k) ? 1u : 0u;
            tc_mma_f16(d1, ah + koff, bl + koff, kIdesc, first);
            tc_mma_f16(d1, al + koff, bh + koff, kIdesc, 1u);
            tc_mma_f16(d0, ah + koff, bh + koff, kIdesc, first);
          }
          tc_commit(&empty[stage]);
          if (kc == nk - 1) tc_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == kAccBufs) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quad = warp & 3;               // TMEM lane quadrant this warp may touch
    const int half = (warp - 2) >> 2;         // which half of the tile's columns this warp drains
    constexpr int kColsPerWarp = kTN / (kEpiWarps / 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tm = t % p.tiles_m, tn = t / p.tiles_m;
      const int row0 = tm * kTM + quad * 32, row = row0 + lane;
      float* tile = epi_tiles + (warp - 2) * 1024;
      epi.begin_row(row, row < p.n0);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (warp == 2 && lane == 0 && t == (int)blockIdx.x) TC_MARK(3);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 2 * kTN + half * kColsPerWarp);
#pragma unroll 1
      for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
        uint32_t r0[32], r1[32];
        tc_ld32_nowait(taddr + c0, r0);
        tc_ld32_nowait(taddr + kTN + c0, r1);
        tc_ld_wait();
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = fmaf(__uint_as_float(r1[c]), kH3InvScale, __uint_as_float(r0[c]));
        const int col0 = tn * kTN + half * kColsPerWarp + c0;
        if (row0 < p.n0 && col0 < p.n1) epi.store32(row0, lane, col0, v, p.n0, p.n1, tile);
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (warp == 2 && lane == 0 && t == (int)blockIdx.x) TC_MARK(4);
      if (++acc == kAccBufs) { acc = 0; acc_phase ^= 1; }
    }
    epi.finish(lane);
    if (warp == 2 && lane == 0) TC_MARK(5);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------------
// (rows, d) fp16 row-major (row stride ld halves, ld % 8 == 0) -> box of (box_rows x 64 halves), 128B swizzle,
// zero fill outside [0, d) x [0, rows).  Descriptors are cached by (base, shape): see sqdist_tc.cu.
int tc_make_map_f16(CUtensorMap* m, const __half* base, int rows, int d, int64_t ld, int box_rows);

template <int TN, class Epi>
inline int launch_gemm_h3(const __half* ah, const __half* al, int n0, int64_t lda, const __half* bh,
                          const __half* bl, int n1, int64_t ldb, int d, Epi epi, cudaStream_t s) {
  CUtensorMap mah, mal, mbh, mbl;
  int rc;
  if ((rc = tc_make_map_f16(&mah, ah, n0, d, lda, kTM)) != CFM_OK) return rc;
  if ((rc = tc_make_map_f16(&mal, al, n0, d, lda, kTM)) != CFM_OK) return rc;
  if ((rc = tc_make_map_f16(&mbh, bh, n1, d, ldb, TN)) != CFM_OK) return rc;
  if ((rc = tc_make_map_f16(&mbl, bl, n1, d, ldb, TN)) != CFM_OK) return rc;
  TcShape p;
  p.n0 = n0; p.n1 = n1; p.d = d;
  p.tiles_m = (n0 + kTM - 1) / kTM;
  p.tiles_n = (n1 + TN - 1) / TN;
  p.dbg = tc_debug_buffer();
  auto kern = gemm_h3_kernel<TN, Epi>;
  constexpr size_t kSmem = H3Cfg<TN>::kSmemBytes;
  static bool attr_set = false;  // per (TN, Epi) instantiation
  if (!attr_set) {
    CFM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    attr_set = true;
  }
  int grid = p.tiles_m * p.tiles_n;
  if (grid > sm_count()) grid = sm_count();
  kern<<<grid, kTcThreads, kSmem, s>>>(mah, mal, mbh, mbl, p, epi);
  ::cfm::note_launches(1);
  CFM_CUDA_OK(cudaGetLastError());
  return CFM_OK;
}

}  // namespace cfm
