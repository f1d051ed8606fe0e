// Shared device helpers for the sm_100a FP64 path: DMMA.8x8x4 wrapper, cp.async staging, tile decode.
// FP64 has no tcgen05 kind on Blackwell; the FP64 tensor path is mma.sync.m8n8k4 (SASS DMMA.8x8x4),
// measured at 37.1 TFLOP/s register-resident on B200 (profiles/r01_fp64_peak.txt).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace smnngp {

constexpr int BM = 128;          // CTA tile rows
constexpr int BN = 128;          // CTA tile cols
constexpr int BK = 16;           // k-slab per pipeline stage
constexpr int LDK = BK + 4;      // padded smem row (doubles): 160 B rows -> conflict-free 8-row x 4-k fragment loads
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 256;  // 8 warps as 2 (M) x 4 (N); warp tile 64 x 32
constexpr int STAGE_DOUBLES = (BM + BN) * LDK;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;  // 163840
constexpr int MI = 8;            // m8 blocks per warp tile
constexpr int NI = 4;            // n8 blocks per warp tile

enum Act { ACT_RELU = 0, ACT_ERF = 1 };
enum Arch { ARCH_MLP = 0, ARCH_RESNET = 1 };
enum Kind { KIND_GAUSS = 0, KIND_STUDENT_T = 1 };
// which diagonal shift the Gram epilogue adds (values live in the device scalar block, see scal layout)
enum Shift { SHIFT_NONE = 0, SHIFT_EPS_ABS = 1, SHIFT_EPS_REL = 2, SHIFT_LIK = 3 };
enum GramOut { OUT_FULL = 0, OUT_LOWER = 1 };

// device scalar block ("scal"), doubles
constexpr int SC_TRMEAN = 0;     // tr(K)/N of the un-shifted Gram
constexpr int SC_SHIFT0 = 1;     // scal[SC_SHIFT0 + shift] = shift value: [0, eps, eps*tr/N, 1e-6*a/b]
constexpr int SC_LOGDET = 5;     // running sum of log L_ii
constexpr int SC_QUAD = 6;       // ||L^-1 y||^2
constexpr int SC_COUNT = 16;

__device__ __forceinline__ void dmma8x8x4(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src, int src_bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src, int src_bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Linear tile id -> (ti, tj).  lower != 0: only tiles with tj <= ti (region starts on the diagonal; rows may
// extend past the column count = trapezoid).  ntn = number of column tiles.
__device__ __forceinline__ void decode_tile(long long t, int ntn, int lower, int& ti, int& tj) {
  if (!lower) {
    ti = (int)(t / ntn);
    tj = (int)(t % ntn);
    return;
  }
  long long tri = (long long)ntn * (ntn + 1) / 2;
  if (t < tri) {
    long long r = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (r * (r + 1) / 2 > t) --r;
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    ti = (int)r;
    tj = (int)(t - r * (r + 1) / 2);
  } else {
    long long u = t - tri;
    ti = ntn + (int)(u / ntn);
    tj = (int)(u % ntn);
  }
}

static inline long long count_tiles(long long M, long long N, int lower) {
  long long ntm = (M + BM - 1) / BM, ntn = (N + BN - 1) / BN;
  if (!lower) return ntm * ntn;
  long long sq = ntm < ntn ? ntm : ntn;
  long long tri = sq * (sq + 1) / 2;
  // rows beyond the square part see all column tiles; (ntm < ntn cannot happen for our regions but stay safe)
  if (ntm >= ntn) return ntn * (ntn + 1) / 2 + (ntm - ntn) * ntn;
  return tri;
}

}  // namespace smnngp
