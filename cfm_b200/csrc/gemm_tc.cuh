// Reusable tcgen05 GEMM core:  D[m, n] = sum_k A[m, k] * B[n, k]  at fp32-grade accuracy via the
// 3xTF32 error-compensated split (A = A_hi + A_lo, B = B_hi + B_lo, all four K-major fp32 arrays
// whose entries are exactly TF32 numbers):  D ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  accumulated in
// ONE fp32 TMEM accumulator.  The tensor core's accumulator truncates at each of the 3*K/8
// accumulate steps, so the result is good to ~1e-6 of sum|a||b| (fp32 FMA: ~1e-7).
//
// Persistent warp-specialised kernel, one CTA per SM, 320 threads:
//   warp 0   TMA producer: 4 tiled loads per K-chunk (A_hi, A_lo: 128x32 fp32; B_hi, B_lo: 256x32
//            fp32; 128B swizzle) into a 2-stage smem ring, mbarrier complete_tx.
//   warp 1   MMA issuer (one elected lane): per K-chunk 4 k-steps x 3 tcgen05.mma (M=128, N=256,
//            K=8), accumulators in TMEM, double-buffered (2 x 256 columns = all 512) so the epilogue
//            of tile i overlaps the MMAs of tile i+1; tcgen05.commit frees smem stages / publishes
//            accumulators.
//   warps 2-9 epilogue (two per TMEM lane quadrant, half the columns each; two tcgen05.ld 32x32b.x32
//            in flight per warp; thread = row, 32 columns per load) handed to the
//            epilogue functor:  begin_row(row, ok); store32(row0, lane, col0, acc[32], n_rows, n_cols, tile)
//            -- called by the whole warp, which writes the chunk coalesced via tc_store_chunk32; finish(lane).
// Tile = 128 x 256 outputs; tiles are walked M-fastest so the B panel stays hot in L2.
// Users: sqdist_tc.cu (cost matrix, algo 2: round 1's path, kept for A/B).  The PTX wrappers below (mbarrier, TMA,
// tcgen05 fences / commit / ld, the SW128 descriptor, the coalesced chunk store) are shared with gemm_h3.cuh, the
// fp16x3 core that serves the cost matrix and the MLP by default.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cfm {

constexpr int kTM = 128, kTK = 32;  // tile rows / K-chunk (32 fp32 = 128 B swizzle atom)
constexpr int kABytes = kTM * kTK * 4;         // 16 KB
// The N-tile (TN) is a template parameter: 256 (2 smem stages of 96 KB; highest operand reuse, used by
// the cost matrix) or 128 (3 stages of 64 KB; twice the tiles, used by the MLP layers whose N is 256/784)
template <int TN> struct TcCfg {
  static constexpr int kBBytes = TN * kTK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStages = TN == 256 ? 2 : 3;
  static constexpr uint32_t kTmemCols = 2 * TN;  // two accumulator buffers (power of two: 512 / 256)
  // + one 32x32 fp32 staging tile per epilogue warp (coalesced read-out, see tc_store_chunk32)
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 256 + 8 * 4096;
  // kind::tf32, fp32 accumulate, A and B K-major, M=128, N=TN
  static constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) |
                                     ((uint32_t)(kTM >> 4) << 24);
};
constexpr int kEpiWarps = 8;                  // two per TMEM lane quadrant, each takes half the columns
constexpr int kTcThreads = 64 + 32 * kEpiWarps;  // warp 0 TMA, warp 1 MMA, warps 2.. epilogue

struct TcShape {
  int n0, n1, d;          // rows of A, rows of B (= output columns), K
  int tiles_m, tiles_n;
  unsigned long long* dbg;  // optional: 8 globaltimer checkpoints per CTA (CFM_TC_DEBUG), else null
};
__device__ __forceinline__ unsigned long long tc_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_MARK(slot) do { if (p.dbg) p.dbg[blockIdx.x * 8 + (slot)] = tc_now(); } while (0)

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);       // start address, 16-byte units
  d |= (uint64_t)0 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // layout type: SWIZZLE_128B
  return d;
}

// Coalesced write of one 32 x 32 fp32 chunk whose ROWS are spread over the lanes (TMEM read-out layout: lane =
// row, v[] = 32 consecutive columns).  Written straight from that layout a warp store instruction touches 32
// different 128-byte lines, 16 bytes each, and the LSU becomes the epilogue's bottleneck (~6 us of a 10 us tile
// epilogue).  Here the chunk is transposed through a per-warp XOR-swizzled shared-memory tile (conflict-free
// 128-bit accesses both ways), so every store instruction writes 4 full rows x 128 contiguous bytes.
// out: &D[row0][col0], ld in floats (multiple of 4, 16-byte aligned); rows >= rows_valid are not written.
__device__ __forceinline__ void tc_store_chunk32(float* tile, const float (&v)[32], float* out, int64_t ld,
                                                 int rows_valid, int lane) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    *reinterpret_cast<float4*>(tile + lane * 32 + ((g ^ (lane & 7)) << 2)) =
        make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
  __syncwarp();
  const int g = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3);
    const float4 q = *reinterpret_cast<const float4*>(tile + r * 32 + ((g ^ (r & 7)) << 2));
    if (r < rows_valid) *reinterpret_cast<float4*>(out + (int64_t)r * ld + 4 * g) = q;
  }
  __syncwarp();
}

template <int TN, class Epi>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
               const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
               const TcShape p, Epi epi) {
  constexpr int kTN = TN, kStages = TcCfg<TN>::kStages, kBBytes = TcCfg<TN>::kBBytes;
  constexpr int kStageBytes = TcCfg<TN>::kStageBytes;
  constexpr uint32_t kTmemCols = TcCfg<TN>::kTmemCols, kIdescTf32 = TcCfg<TN>::kIdesc;
  extern __shared__ __align__(1024) uint8_t smem[];
  // carve: stages | barriers | tmem ptr
  uint8_t* stage_base = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;   // [2] accumulator ready
  uint64_t* tempty = tfull + 2;        // [2] accumulator drained
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* epi_tiles = reinterpret_cast<float*>(smem + kStages * kStageBytes + 256);  // [kEpiWarps][32 * 32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TC_MARK(0);  // kernel entry
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int nk = (p.d + kTK - 1) / kTK;

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
  if (threadIdx.x == 0) TC_MARK(1);  // barriers initialised, TMEM allocated
  // Programmatic dependent launch: everything above touches no memory written by the preceding kernel, so when
  // this grid is launched with the programmatic-serialization attribute its prologue overlaps the tail of its
  // predecessor; from here on the predecessor's results are needed (operands via TMA, the device scalar t).
  // Without the attribute both instructions are no-ops.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

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
          tma_load_2d(sb, &map_ah, &full[stage], kc * kTK, tm * kTM);
          tma_load_2d(sb + kABytes, &map_al, &full[stage], kc * kTK, tm * kTM);
          tma_load_2d(sb + 2 * kABytes, &map_bh, &full[stage], kc * kTK, tn * kTN);
          tma_load_2d(sb + 2 * kABytes + kBBytes, &map_bl, &full[stage], kc * kTK, tn * kTN);
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
      mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kTN);
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0 && kc == 0 && t == (int)blockIdx.x) TC_MARK(2);  // first operand stage landed
        if (lane == 0) {
          const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
          const uint64_t ah = umma_desc_sw128(sa), al = umma_desc_sw128(sa + kABytes);
          const uint64_t bh = umma_desc_sw128(sa + 2 * kABytes), bl = umma_desc_sw128(sa + 2 * kABytes + kBBytes);
#pragma unroll
          for (int k = 0; k < kTK / 8; ++k) {
            const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);  // 32 bytes per k-step, 16-byte units
            const uint32_t first = (kc | k) ? 1u : 0u;
            tc_mma_tf32(d_tmem, al + koff, bh + koff, kIdescTf32, first);  // small terms first
            tc_mma_tf32(d_tmem, ah + koff, bl + koff, kIdescTf32, 1u);
            tc_mma_tf32(d_tmem, ah + koff, bh + koff, kIdescTf32, 1u);
          }
          tc_commit(&empty[stage]);                  // smem stage reusable once these MMAs retire
          if (kc == nk - 1) tc_commit(&tfull[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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
      if (warp == 2 && lane == 0 && t == (int)blockIdx.x) TC_MARK(3);  // first accumulator complete
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kTN + half * kColsPerWarp);
#pragma unroll 1
      for (int c0 = 0; c0 < kColsPerWarp; c0 += 64) {
        // two 32-column TMEM loads in flight, then both are consumed: twice the ILP per warp
        uint32_t r0[32], r1[32];
        tc_ld32_nowait(taddr + c0, r0);
        tc_ld32_nowait(taddr + c0 + 32, r1);
        tc_ld_wait();
        const int col0 = tn * kTN + half * kColsPerWarp + c0;
        // warp-uniform conditions: the whole warp cooperates on the store of a chunk
        if (row0 < p.n0 && col0 < p.n1) epi.store32(row0, lane, col0, r0, p.n0, p.n1, tile);
        if (row0 < p.n0 && col0 + 32 < p.n1) epi.store32(row0, lane, col0 + 32, r1, p.n0, p.n1, tile);
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (warp == 2 && lane == 0 && t == (int)blockIdx.x) TC_MARK(4);  // first tile's epilogue done
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    epi.finish(lane);
    if (warp == 2 && lane == 0) TC_MARK(5);  // all epilogues done
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------------
// (rows, d) fp32 row-major (row stride ld floats) -> box of (box_rows x 32 floats), 128B swizzle, zero OOB fill
int tc_make_map(CUtensorMap* m, const float* base, int rows, int d, int64_t ld, int box_rows);
unsigned long long* tc_debug_buffer();  // device buffer set through cfm_tc_debug_buffer(), else null
// x -> (hi, lo) TF32 split of n contiguous floats (sqdist_tc.cu)
int tc_split(const float* x, float* hi, float* lo, int64_t n, cudaStream_t s);

template <int TN, class Epi>
inline int launch_gemm_tc(const float* ah, const float* al, int n0, int64_t lda, const float* bh,
                          const float* bl, int n1, int64_t ldb, int d, Epi epi, cudaStream_t s) {
  constexpr int kTN = TN;
  CUtensorMap mah, mal, mbh, mbl;
  int rc;
  if ((rc = tc_make_map(&mah, ah, n0, d, lda, kTM)) != CFM_OK) return rc;
  if ((rc = tc_make_map(&mal, al, n0, d, lda, kTM)) != CFM_OK) return rc;
  if ((rc = tc_make_map(&mbh, bh, n1, d, ldb, kTN)) != CFM_OK) return rc;
  if ((rc = tc_make_map(&mbl, bl, n1, d, ldb, kTN)) != CFM_OK) return rc;
  TcShape p;
  p.n0 = n0; p.n1 = n1; p.d = d;
  p.tiles_m = (n0 + kTM - 1) / kTM;
  p.tiles_n = (n1 + kTN - 1) / kTN;
  p.dbg = tc_debug_buffer();
  auto kern = gemm_tc_kernel<TN, Epi>;
  constexpr size_t kSmem = TcCfg<TN>::kSmemBytes;
  CFM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
  int grid = p.tiles_m * p.tiles_n;
  if (grid > sm_count()) grid = sm_count();
  static int pdl = -1;  // CFM_TC_PDL=1: launch as a programmatic dependent of the preceding kernel in the stream
  if (pdl < 0) { const char* e = getenv("CFM_TC_PDL"); pdl = e ? atoi(e) : 0; }
  if (pdl) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = kSmem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CFM_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mah, mal, mbh, mbl, p, epi));
  } else {
    kern<<<grid, kTcThreads, kSmem, s>>>(mah, mal, mbh, mbl, p, epi);
  }
  ::cfm::note_launches(1);
  CFM_CUDA_OK(cudaGetLastError());
  return CFM_OK;
}

}  // namespace cfm
