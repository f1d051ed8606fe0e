// TMA-fed, warp-specialised, persistent GEMM core for sm_100a:  C_tile(128x64) = sum_k A[i,k] * B[j,k].
//
//   CTA = 3 warpgroups (384 threads, one CTA per SM, persistent over the tile list): warpgroup 0 is the
//   producer (setmaxnreg.dec -> 40 registers), warpgroups 1 and 2 are the math groups (setmaxnreg.inc -> 232;
//   the register file is per SMSP, 16384 / (40 + 2 * 232) / 32 fits exactly one producer + two math warps).
//   The two math groups are independent (2 x 2 warps each, warp tile 64 x 32, 64 FP64 accumulators per
//   thread).  Each group owns its own 128 x 64 output tile, its own 4-stage shared-memory ring and its own
//   full / empty mbarriers, so the groups drift apart in time: while one group is in its epilogue (Gram
//   recursion or the read-modify-write of the Cholesky update) or waits for the first slab of its next tile,
//   the other group keeps the FP64 pipe busy (ping-pong).  Lane 0 of producer warp g feeds group g:
//   wait on the empty barrier, arm the full barrier with the slab's byte count, issue two cp.async.bulk.tensor
//   (UTMALDG) loads - A box 16 x 128, B box 16 x 64, 128-byte swizzle, out-of-range rows / k zero-filled by
//   the TMA unit.  The producer runs ahead across tile boundaries, so the ring is already full when a group
//   comes back from an epilogue.  No CTA-wide barrier and no address arithmetic in the math warps.
//
//   Shared-memory layout of a slab: row r of the box at byte r*128, 16-byte chunk c stored at c ^ (r & 7).
//   The k index inside an 8x8x4 DMMA is a free permutation (the same one for A and B), chosen so that the 16
//   lanes of a half-warp (4 rows x 4 k) touch 16 distinct 8-byte bank pairs:  lane j = lane & 3 reads
//   k = 8 (j & 1) + (j >> 1) + 2 kk  in k4-step kk, i.e. chunk (4 (j & 1) + kk) ^ (lane >> 2), half j >> 1.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace smnngp {

cudaError_t configure_kernel_once(const void* fn, int smem_bytes, bool carveout);   // context.cu

constexpr int TM_BM = 128, TM_BN = 64, TM_STAGES = 4;
constexpr int TM_A_BYTES = TM_BM * BK * 8;                 // 16384
constexpr int TM_B_BYTES = TM_BN * BK * 8;                 // 8192
constexpr int TM_STAGE_BYTES = TM_A_BYTES + TM_B_BYTES;    // 24576
constexpr int TM_THREADS = 384;                           // producer warpgroup + 2 math warpgroups
constexpr int TM_SMEM_BYTES = 1024 + 2 * TM_STAGES * TM_STAGE_BYTES + 256;
using TileTma = TileCfg<TM_BM, TM_BN, TM_STAGES, 1>;       // for count_tiles / decode_tile<Q = 2>

struct TmaShape {
  int M, N, K;          // output rows, cols, contraction length
  int lower;            // tile list: 1 = triangular enumeration (only tiles touching col <= row)
  long long tiles;
  // block-row-cyclic mask (multi-GPU local update, see GemmParams): rectangular enumeration, inactive tiles skipped
  int cyc_db, cyc_p, base_shift;
  // 1: both operands are rows of an UPPER-triangular matrix (zero left of the diagonal) and the tile list is the
  // lower triangle, so the contraction of tile (r0, c0) only runs over k >= r0 (A^-1 = U U^T with U = L^-T)
  int k_from_row;
  // 1: operand B is LOWER triangular (row j zero right of column j), so the contraction of the tile whose first
  // column is c0 only runs over k < c0 + 64 (panel solve with the explicit inverse: rows * inv(L)^T)
  int k_upto_col;
  // Always 0 (aggregate initialisation leaves it value-initialised).  The math warps AND it with a token derived from
  // the fragments they loaded and add the result to the address of the stage-release arrive: ptxas cannot see the
  // value of a kernel parameter, so the arrive carries a true data dependency on the loads (see tma_gemm_kernel).
  unsigned int zero;
  // > 0: L2-aware rasterisation.  Tiles are enumerated super-tile by super-tile (super_rows x 2 super_rows tiles =
  // a square of 128 super_rows elements; lower == 1: block-triangular list of super-tiles), row-major inside.  The
  // ~2 x #SM tiles in flight at any time then cover 2-3 super-tiles (a few MB of operand rows, L2 resident) instead
  // of one long strip of a tile row whose B operand streams from DRAM again for every tile row.  Tile slots outside
  // the matrix or above the diagonal are skipped (producer and math warps decode alike).
  int super_rows;
  int cyc_alt;          // block-row-cyclic mask, snake distribution: extra shift of the rows of odd local blocks
  int k_row0;           // k_from_row: global index of the output's first row (row strip of a larger matrix)
};

__host__ __device__ __forceinline__ long long tma_super_count_tiles(int M, int N, int lower, int sr) {
  const long long ntm = (M + TM_BM - 1) / TM_BM, ntn = (N + TM_BN - 1) / TM_BN;
  const long long nsr = (ntm + sr - 1) / sr, nsc = (ntn + 2 * sr - 1) / (2 * sr);
  const long long per = 2ll * sr * sr;
  if (!lower) return nsr * nsc * per;
  // block-triangular: super row I holds super columns 0 .. min(I, nsc - 1)
  long long cnt = 0;
  for (long long I = 0; I < nsr; I++) cnt += (I + 1 < nsc ? I + 1 : nsc);
  return cnt * per;
}

// Block-row-cyclic mask: number of column tiles of row tile ti that touch the active region (a prefix of the row,
// the limit grows with the row).  Only these tiles are enumerated, so the static tile -> CTA assignment stays
// balanced (a rectangular enumeration that merely skipped the inactive half left some CTAs with up to 2x the
// work of others late in the factorisation).
__host__ __device__ __forceinline__ int tma_cyc_row_tiles(const TmaShape& sh, int ti, int ntn) {
  const int rl = (ti * TM_BM + TM_BM < sh.M ? ti * TM_BM + TM_BM : sh.M) - 1;
  const int lb = rl / sh.cyc_db;
  const long long lim = (long long)rl + sh.base_shift + (long long)lb * (sh.cyc_p - 1) * sh.cyc_db +
                        ((lb & 1) ? sh.cyc_alt : 0);
  if (lim < 0) return 0;
  const long long c = lim / TM_BN + 1;
  return c < ntn ? (int)c : ntn;
}
inline long long tma_cyc_count_tiles(const TmaShape& sh) {
  const int ntm = (sh.M + TM_BM - 1) / TM_BM, ntn = (sh.N + TM_BN - 1) / TM_BN;
  long long t = 0;
  for (int ti = 0; ti < ntm; ti++) t += tma_cyc_row_tiles(sh, ti, ntn);
  return t;
}
// Tile ids handed to one warp only grow, so the row of a tile id is found by walking a cursor forward.
struct TileCursor {
  int ti = 0;
  long long base = 0;      // active tiles in rows < ti
};
// returns false for an enumeration slot that holds no tile (super-tile rasterisation only)
__device__ __forceinline__ bool tma_decode(const TmaShape& sh, long long t, int ntn, TileCursor& cur, int& ti,
                                           int& tj) {
  if (sh.super_rows > 0) {
    const int sr = sh.super_rows, sc = 2 * sr;
    const int per = sr * sc;
    const long long S = t / per;
    const int within = (int)(t - S * per);
    const int ntm = (sh.M + TM_BM - 1) / TM_BM;
    long long I, J;
    if (sh.lower) {
      const long long nsc = (ntn + sc - 1) / sc;
      const long long tri = nsc * (nsc + 1) / 2;          // super rows 0 .. nsc-1 are triangular, later ones full
      if (S < tri) {
        I = (long long)((sqrt(8.0 * (double)S + 1.0) - 1.0) * 0.5);
        while (I * (I + 1) / 2 > S) --I;
        while ((I + 1) * (I + 2) / 2 <= S) ++I;
        J = S - I * (I + 1) / 2;
      } else {
        I = nsc + (S - tri) / nsc;
        J = (S - tri) % nsc;
      }
    } else {
      const long long nsc = (ntn + sc - 1) / sc;
      I = S / nsc;
      J = S % nsc;
    }
    ti = (int)(I * sr) + within / sc;
    tj = (int)(J * sc) + within % sc;
    return ti < ntm && tj < ntn && (!sh.lower || tj <= 2 * ti + 1);
  }
  if (sh.cyc_db == 0) {
    decode_tile<2>(t, ntn, sh.lower, ti, tj);
    return true;
  }
  int cnt = tma_cyc_row_tiles(sh, cur.ti, ntn);
  while (t >= cur.base + cnt) {
    cur.base += cnt;
    cnt = tma_cyc_row_tiles(sh, ++cur.ti, ntn);
  }
  ti = cur.ti;
  tj = (int)(t - cur.base);
  return true;
}

// ---- host: tensor map for a row-major [rows, inner] FP64 operand (row pitch ld doubles) ---------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tma_encode_fn() {
  static const PFN_encodeTiled fn = []() -> PFN_encodeTiled {      // thread-safe one-time lookup (idempotent cache)
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<PFN_encodeTiled>(p);
    return nullptr;
  }();
  return fn;
}

// operands the TMA path accepts: 16-byte aligned base, even pitch
inline bool tma_operand_ok(const double* base, long long ld) {
  return tma_encode_fn() != nullptr && (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld % 2) == 0 && ld > 0;
}

inline bool make_tmap(CUtensorMap* m, const double* base, long long rows, long long inner, long long ld,
                      int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = tma_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// ---- device helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c_inner, int c_row,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c_inner), "r"(c_row), "r"(bar)
      : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t hi32(double v) { return (uint32_t)__double2hiint(v); }

// optional Epi::finish(params): run once per CTA by the 256 math threads after the last tile (e.g. cross-GPU signal)
template <class Epi, class = void>
struct epi_has_finish : std::false_type {};
template <class Epi>
struct epi_has_finish<Epi, std::void_t<decltype(&Epi::finish)>> : std::true_type {};

__device__ __forceinline__ int tma_kt_end(const TmaShape& sh, int KT, int c0) {
  if (!sh.k_upto_col) return KT;
  const int e = (c0 + TM_BN + BK - 1) / BK;
  return e < KT ? e : KT;
}

// Epi must provide:  struct Params;  static __device__ void apply(const Params&, double (&acc)[MI][NI][2],
//                    int r0, int c0, int wm, int wn, int lane);   (r0, c0 = tile origin; wm, wn in {0, 1})
template <class Epi>
__global__ void __launch_bounds__(TM_THREADS, 1)
tma_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const TmaShape sh, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;              // 1024-byte aligned (128B swizzle)
  const uint32_t bars = ring + 2 * TM_STAGES * TM_STAGE_BYTES;              // full[2][S], then empty[2][S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * TM_STAGES; i++) {
      mbar_init(bars + 8 * i, 1);                                           // full: producer arrive + tx bytes
      mbar_init(bars + 8 * (2 * TM_STAGES + i), 4);                         // empty: one arrive per math warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();

  const int KT = (sh.K + BK - 1) / BK;
  const int ntn = (sh.N + TM_BN - 1) / TM_BN;
  const long long stride = 2ll * gridDim.x;

  if (warp < 4) {
    // ===== producer warpgroup: lane 0 of warp g feeds group g =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp < 2 && lane == 0) {
      const int g = warp;
      const uint32_t gring = ring + g * TM_STAGES * TM_STAGE_BYTES;
      const uint32_t full0 = bars + 8 * (g * TM_STAGES), empty0 = bars + 8 * (2 * TM_STAGES + g * TM_STAGES);
      int s = 0;
      uint32_t ph = 0;
      TileCursor cur;
      for (long long t = 2ll * blockIdx.x + g; t < sh.tiles; t += stride) {
        int ti, tj;
        if (!tma_decode(sh, t, ntn, cur, ti, tj)) continue;
        const int r0 = ti * TM_BM, c0 = tj * TM_BN;
        const int kt_end = tma_kt_end(sh, KT, c0);
        for (int kt = sh.k_from_row ? (sh.k_row0 + r0) / BK : 0; kt < kt_end; kt++) {
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          mbar_arrive_expect_tx(full0 + 8 * s, TM_STAGE_BYTES);
          const uint32_t dst = gring + s * TM_STAGE_BYTES;
          tma_load_2d(dst, &mapA, kt * BK, r0, full0 + 8 * s);
          tma_load_2d(dst + TM_A_BYTES, &mapB, kt * BK, c0, full0 + 8 * s);
          if (++s == TM_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
    return;
  }

  // ===== math warpgroups =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int g = (warp >> 2) - 1, wg = warp & 3, wm = wg >> 1, wn = wg & 1;
  const uint32_t gring = ring + g * TM_STAGES * TM_STAGE_BYTES;
  const uint32_t full0 = bars + 8 * (g * TM_STAGES), empty0 = bars + 8 * (2 * TM_STAGES + g * TM_STAGES);
  const int g8 = lane >> 2, j = lane & 3;
  const uint32_t off0 = (uint32_t)((((4 * (j & 1)) ^ g8) << 4) | ((j >> 1) << 3));
  const uint32_t a_off = (uint32_t)((wm * 64 + g8) * 128) + off0;
  const uint32_t b_off = (uint32_t)(TM_A_BYTES + (wn * 32 + g8) * 128) + off0;
  int s = 0;
  uint32_t ph = 0;
  TileCursor cur;
  for (long long t = 2ll * blockIdx.x + g; t < sh.tiles; t += stride) {
    int ti, tj;
    if (!tma_decode(sh, t, ntn, cur, ti, tj)) continue;
    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
      for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    const int kt_end = tma_kt_end(sh, KT, tj * TM_BN);
    for (int kt = sh.k_from_row ? (sh.k_row0 + ti * TM_BM) / BK : 0; kt < kt_end; kt++) {
      mbar_wait(full0 + 8 * s, ph);
      const uint32_t st = gring + s * TM_STAGE_BYTES;
      double a[2][MI], b[2][NI];
#pragma unroll
      for (int ni = 0; ni < NI; ni++) b[0][ni] = lds_f64(st + b_off + ni * 1024);
#pragma unroll
      for (int mi = 0; mi < MI; mi++) a[0][mi] = lds_f64(st + a_off + mi * 1024);
#pragma unroll
      for (int kk = 0; kk < BK / 4; kk++) {
        const int cur = kk & 1, nxt = cur ^ 1;
        if (kk + 1 < BK / 4) {              // next k4-step's fragments are in flight while this one computes
          const uint32_t x = (uint32_t)((kk + 1) << 4);
#pragma unroll
          for (int ni = 0; ni < NI; ni++) b[nxt][ni] = lds_f64(st + (b_off ^ x) + ni * 1024);
#pragma unroll
          for (int mi = 0; mi < MI; mi++) a[nxt][mi] = lds_f64(st + (a_off ^ x) + mi * 1024);
        }
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
          for (int ni = 0; ni < NI; ni++) dmma8x8x4(acc[mi][ni], a[cur][mi], b[cur][ni]);
      }
      // Stage release carries a DATA DEPENDENCY on the slab's fragments.  An mbarrier arrive has no register
      // operand in common with the ld.shared that filled the fragments, so ptxas may issue it while those loads
      // are still in flight; the producer then re-arms the stage and the next TMA can overwrite rows a load has
      // not read yet (round 1: wrong fragments in ~0.3 % of the tiles, masked there by releasing one slab late).
      // Here the arrive's address is offset by (token & sh.zero): the token ORs the high words of every fragment
      // register written by the slab's last two k4-steps (same registers as steps 0/1: a register's writes retire
      // in order, so all of the slab's loads have landed when these are readable), is reduced over the warp
      // (every lane's loads), and sh.zero is a kernel parameter ptxas cannot fold.  The arrive can therefore not
      // issue before the slab's last ld.shared has delivered its data, independent of instruction scheduling.
      uint32_t tok = 0;
#pragma unroll
      for (int ni = 0; ni < NI; ni++) tok |= hi32(b[0][ni]) | hi32(b[1][ni]);
#pragma unroll
      for (int mi = 0; mi < MI; mi++) tok |= hi32(a[0][mi]) | hi32(a[1][mi]);
      tok = __reduce_or_sync(0xffffffffu, tok) & sh.zero;
      if (lane == 0) mbar_arrive(empty0 + 8 * s + tok);
      if (++s == TM_STAGES) { s = 0; ph ^= 1u; }
    }
    Epi::apply(ep, acc, ti * TM_BM, tj * TM_BN, wm, wn, lane);
  }
  if constexpr (epi_has_finish<Epi>::value) Epi::finish(ep);
}

template <class Epi>
cudaError_t launch_tma_gemm(cudaStream_t s, const CUtensorMap& mapA, const CUtensorMap& mapB, const TmaShape& sh,
                            const typename Epi::Params& ep, int num_sms) {
  auto kern = tma_gemm_kernel<Epi>;
  cudaError_t e = configure_kernel_once(reinterpret_cast<const void*>(kern), TM_SMEM_BYTES, false);
  if (e != cudaSuccess) return e;
  long long ctas = (sh.tiles + 1) / 2;
  if (ctas > num_sms) ctas = num_sms;
  if (ctas < 1) return cudaSuccess;
  kern<<<(unsigned)ctas, TM_THREADS, TM_SMEM_BYTES, s>>>(mapA, mapB, sh, ep);
  return cudaGetLastError();
}

int device_sm_count();      // SM count of the current device (context.cu)

}  // namespace smnngp
