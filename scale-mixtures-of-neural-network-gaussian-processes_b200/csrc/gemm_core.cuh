// C_tile(128x128) = sum_k A[i,k] * B[j,k]   (both operands K-contiguous, i.e. row-major A times row-major B
// transposed) - the one contraction shape on the hot path: X.X^T for the Gram, panel.panel^T for the Cholesky
// trailing update, panel.Linv^T for the TRSM.  4-stage cp.async pipeline into padded shared memory, 8 warps of
// 64x32, each k4 step = 12 LDS.64 + 32 DMMA.8x8x4 per warp; accumulators stay in registers for the epilogue.
#pragma once
#include "common.cuh"

namespace smnngp {

// Per-thread addressing of the cp.async staging, hoisted out of the k loop.  Each thread copies the same
// (row, k-chunk) slots of every slab: 16-byte chunks (aligned operands) or 8-byte chunks.
template <bool ALIGN16, int ROWS, int THREADS>
struct TileLoader {
  static constexpr int NCH = ROWS * (ALIGN16 ? 8 : 16) / THREADS;
  const double* src[NCH];   // global address of the slot at k0 = 0 (nullptr: row out of range -> always zero fill)
  int soff[NCH];            // smem offset (doubles) inside the operand tile
  int kc[NCH];              // k offset of the slot inside a slab
  __device__ __forceinline__ void init(const double* __restrict__ g, long long ld, int rows_valid, int tid) {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      int c = tid + i * THREADS;
      int r = ALIGN16 ? (c >> 3) : (c >> 4);
      kc[i] = ALIGN16 ? (c & 7) * 2 : (c & 15);
      soff[i] = r * LDK + kc[i];
      src[i] = (r < rows_valid) ? g + (long long)r * ld + kc[i] : nullptr;
    }
  }
  __device__ __forceinline__ void issue(double* s, int k0, int K, const double* dummy) const {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      int rem = K - (k0 + kc[i]);
      int bytes = 0;
      if (src[i] != nullptr && rem > 0) bytes = ALIGN16 ? (rem >= 2 ? 16 : 8) : 8;
      const double* p = bytes ? src[i] + k0 : dummy;
      if (ALIGN16) cp_async16(s + soff[i], p, bytes);
      else cp_async8(s + soff[i], p, bytes);
    }
  }
};

// Accumulator element acc[mi][ni][e] is C(row, col) of the CTA tile with
//   row = wm*64 + mi*8 + (lane>>2),  col = wn*32 + ni*8 + (lane&3)*2 + e   (wm = warp / WARPS_N, wn = warp % WARPS_N)
template <typename Cfg, bool ALIGN16>
__device__ __forceinline__ void gemm_mainloop(double (&acc)[MI][NI][2], const double* __restrict__ Ag,
                                              long long lda, int a_rows, const double* __restrict__ Bg,
                                              long long ldb, int b_rows, int K, double* smem) {
  constexpr int STAGES = Cfg::STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  const int KT = (K + BK - 1) / BK;
#pragma unroll
  for (int mi = 0; mi < MI; mi++)
#pragma unroll
    for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  TileLoader<ALIGN16, Cfg::BM, Cfg::THREADS> la;
  TileLoader<ALIGN16, Cfg::BN, Cfg::THREADS> lb;
  la.init(Ag, lda, a_rows, tid);
  lb.init(Bg, ldb, b_rows, tid);
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < KT) {
      double* st = smem + s * Cfg::STAGE_DOUBLES;
      la.issue(st, s * BK, K, Ag);
      lb.issue(st + Cfg::BM * LDK, s * BK, K, Bg);
    }
    cp_async_commit();
  }
  const int frag_off = (lane >> 2) * LDK + (lane & 3);
  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const double* As = smem + (kt % STAGES) * Cfg::STAGE_DOUBLES;
    const double* ap = As + (wm * 64) * LDK + frag_off;
    const double* bp = As + Cfg::BM * LDK + (wn * 32) * LDK + frag_off;
#pragma unroll
    for (int kk = 0; kk < BK / 4; kk++) {
      double a[MI], b[NI];
#pragma unroll
      for (int mi = 0; mi < MI; mi++) a[mi] = ap[mi * 8 * LDK + kk * 4];
#pragma unroll
      for (int ni = 0; ni < NI; ni++) b[ni] = bp[ni * 8 * LDK + kk * 4];
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) dmma8x8x4(acc[mi][ni], a[mi], b[ni]);
      if (kk == 0) {
        // refill the stage consumed in the previous iteration; issued after the first DMMA group so the
        // tensor pipe is already busy while the copy instructions go out
        int nk = kt + STAGES - 1;
        if (nk < KT) {
          double* st = smem + (nk % STAGES) * Cfg::STAGE_DOUBLES;
          la.issue(st, nk * BK, K, Ag);
          lb.issue(st + Cfg::BM * LDK, nk * BK, K, Bg);
        }
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

}  // namespace smnngp
