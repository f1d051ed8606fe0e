// C_tile(128x128) = sum_k A[i,k] * B[j,k]   (both operands K-contiguous, i.e. row-major A times row-major B
// transposed) - the one contraction shape on the hot path: X.X^T for the Gram, panel.panel^T for the Cholesky
// trailing update, panel.Linv^T for the TRSM.  4-stage cp.async pipeline into padded shared memory, 8 warps of
// 64x32, each k4 step = 12 LDS.64 + 32 DMMA.8x8x4 per warp; accumulators stay in registers for the epilogue.
#pragma once
#include "common.cuh"

namespace smnngp {

// Stage one 128 x BK operand tile.  g points at (first tile row, k = 0); rows >= rows_valid and k >= K zero-fill.
template <bool ALIGN16>
__device__ __forceinline__ void load_operand_tile(double* s, const double* __restrict__ g, long long ld,
                                                  int rows_valid, int k0, int K, int tid) {
  if (ALIGN16) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int c = tid + i * GEMM_THREADS;
      int r = c >> 3;
      int kc = (c & 7) * 2;
      int rem = K - (k0 + kc);
      int bytes = 0;
      if (r < rows_valid && rem > 0) bytes = rem >= 2 ? 16 : 8;
      const double* src = bytes ? g + (long long)r * ld + (k0 + kc) : g;
      cp_async16(s + r * LDK + kc, src, bytes);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int c = tid + i * GEMM_THREADS;
      int r = c >> 4;
      int kc = c & 15;
      int bytes = (r < rows_valid && (k0 + kc) < K) ? 8 : 0;
      const double* src = bytes ? g + (long long)r * ld + (k0 + kc) : g;
      cp_async8(s + r * LDK + kc, src, bytes);
    }
  }
}

// Accumulator element acc[mi][ni][e] is C(row, col) with
//   row = wm*64 + mi*8 + (lane>>2),  col = wn*32 + ni*8 + (lane&3)*2 + e      (wm = warp>>2, wn = warp&3)
template <bool ALIGN16>
__device__ __forceinline__ void gemm_mainloop(double (&acc)[MI][NI][2], const double* __restrict__ Ag,
                                              long long lda, int a_rows, const double* __restrict__ Bg,
                                              long long ldb, int b_rows, int K, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int KT = (K + BK - 1) / BK;
#pragma unroll
  for (int mi = 0; mi < MI; mi++)
#pragma unroll
    for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < KT) {
      double* st = smem + s * STAGE_DOUBLES;
      load_operand_tile<ALIGN16>(st, Ag, lda, a_rows, s * BK, K, tid);
      load_operand_tile<ALIGN16>(st + BM * LDK, Bg, ldb, b_rows, s * BK, K, tid);
    }
    cp_async_commit();
  }
  const int frag_off = (lane >> 2) * LDK + (lane & 3);
  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < KT) {
        double* st = smem + (nk % STAGES) * STAGE_DOUBLES;
        load_operand_tile<ALIGN16>(st, Ag, lda, a_rows, nk * BK, K, tid);
        load_operand_tile<ALIGN16>(st + BM * LDK, Bg, ldb, b_rows, nk * BK, K, tid);
      }
      cp_async_commit();
    }
    const double* As = smem + (kt % STAGES) * STAGE_DOUBLES;
    const double* ap = As + (wm * 64) * LDK + frag_off;
    const double* bp = As + BM * LDK + (wn * 32) * LDK + frag_off;
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
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

}  // namespace smnngp
