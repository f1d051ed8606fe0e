// W = inv(L) for one outer panel's w x w (w <= 512) lower-triangular diagonal block, assembled from L itself and the
// inverses of its 128 x 128 diagonal blocks (both are produced by the diagonal-block factorisation, chol.cu).
//
// Why: with the full inverse the panel solve  R <- R * inv(L)^T  of ALL rows below the block is ONE launch of the TMA
// GEMM core (contraction cut at W's diagonal, TmaShape::k_upto_col) instead of a chain of 4 TRSM-as-GEMM + 3 update
// launches per 512-wide panel - on one GPU that chain sat on the main stream in front of every trailing update, on
// several GPUs it is the owner's critical path that every other rank waits for.
//
// How: block forward substitution, column strips are independent.
//   W_jj = inv(L_jj) (given),   W_ij = -inv(L_ii) * sum_{k=j}^{i-1} L_ik W_kj    (i > j, 128-blocks)
// One CTA owns an 8-column strip of block column j and walks down the block rows i = j+1 .. nb-1; the strip of W it
// has produced so far stays in shared memory (k-major, the DMMA B operand), the 128 x 128 A blocks (L_ik, then
// inv(L_ii)) stream from global memory / L2 through a 3-stage cp.async ring in 32-column slabs.  16 x 4 CTAs, no
// inter-CTA dependency, every product is an 8x8x4 DMMA.  The result is stored to every rank's W buffer (multi-GPU:
// NVLink peer stores) and, optionally, the last CTA raises the W-ready flag on every rank.
#include "context.cuh"
#include "kernels.cuh"
#include "peer_signal.cuh"

namespace smnngp {

namespace {

constexpr int AS_CS = 8;                    // columns of W per CTA
constexpr int AS_KS = 32;                   // k per staged slab
constexpr int AS_LDA = AS_KS + 4;           // 36 doubles: conflict-free 8-row x 4-k fragment loads, 16-byte rows
constexpr int AS_LDB = AS_CS + 4;           // 12 doubles: conflict-free 4-k x 8-col fragment loads
constexpr int AS_STAGES = 3;
constexpr int AS_THREADS = 256;
constexpr int AS_MAXW = LINV_BLOCKS * PB;   // 512
constexpr int AS_STAGE_DOUBLES = PB * AS_LDA;
constexpr int AS_SMEM_BYTES = (AS_MAXW * AS_LDB + PB * AS_LDB + AS_STAGES * AS_STAGE_DOUBLES) * 8;

struct AssembleParams {
  const double* L;            // w x w block, lower part valid, pitch ldl
  long long ldl;
  const double* linv;         // nb blocks of 128 x 128 (pitch 128): inv(L_ii), zero above the diagonal, identity padded
  int w;
  double* out[MAX_PEERS];     // W buffer of every rank (pitch ldw); P = 1: the local buffer only
  long long ldw;
  int P;
  PeerSignal sig;             // sig.counter == nullptr: no flag
};

// position in the CTA's sequence of A slabs: block row i, block column k (k == i: the inv(L_ii) step), slab 0..3
struct SlabCursor {
  int i, k, slab;
  __device__ __forceinline__ void advance(int j) {
    if (++slab == PB / AS_KS) {
      slab = 0;
      if (++k > i) { ++i; k = j; }
    }
  }
};

__device__ __forceinline__ void issue_slab(const AssembleParams& p, const SlabCursor& c, double* stage, int tid) {
  const double* src;
  long long ld;
  int rows_valid;
  if (c.k < c.i) {
    src = p.L + (long long)(c.i * PB) * p.ldl + c.k * PB + c.slab * AS_KS;
    ld = p.ldl;
    rows_valid = min(PB, p.w - c.i * PB);
  } else {
    src = p.linv + (long long)c.i * PB * PB + c.slab * AS_KS;
    ld = PB;
    rows_valid = PB;
  }
#pragma unroll
  for (int t = 0; t < PB * (AS_KS / 2) / AS_THREADS; t++) {
    const int ch = tid + t * AS_THREADS;
    const int row = ch >> 4, kc = (ch & 15) * 2;
    const bool ok = row < rows_valid;
    cp_async16(stage + row * AS_LDA + kc, ok ? src + (long long)row * ld + kc : p.linv, ok ? 16 : 0);
  }
}

__global__ void __launch_bounds__(AS_THREADS, 1) assemble_inverse_kernel(const AssembleParams p) {
  extern __shared__ __align__(16) double sm[];
  double* Wc = sm;                                   // [512][AS_LDB]: strip of W, row = k
  double* Tt = Wc + AS_MAXW * AS_LDB;                // [128][AS_LDB]: sum_k L_ik W_kj of the current block row
  double* ring = Tt + PB * AS_LDB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, jj = lane & 3;
  const int j = blockIdx.y;
  const int nb = (p.w + PB - 1) / PB;
  const int c0 = j * PB + blockIdx.x * AS_CS;        // first column of this CTA's strip
  const bool active = c0 < p.w;

  if (active) {
    // W_jj strip: rows of block j
    for (int e = tid; e < PB * AS_CS; e += AS_THREADS) {
      const int r = e >> 3, c = e & 7;
      Wc[(j * PB + r) * AS_LDB + c] = p.linv[(long long)j * PB * PB + r * PB + blockIdx.x * AS_CS + c];
    }
    int total = 0;                                   // slabs in the sequence
    for (int i = j + 1; i < nb; i++) total += (i - j + 1) * (PB / AS_KS);
    SlabCursor pre{j + 1, j, 0}, cur{j + 1, j, 0};
#pragma unroll
    for (int s = 0; s < AS_STAGES - 1; s++) {
      if (s < total) {
        issue_slab(p, pre, ring + s * AS_STAGE_DOUBLES, tid);
        pre.advance(j);
      }
      cp_async_commit();
    }
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const int arow = (warp * 16 + g) * AS_LDA + jj;
    for (int q = 0; q < total; q++) {
      cp_async_wait<AS_STAGES - 2>();
      __syncthreads();                               // slab q landed; Wc / Tt written in iteration q-1 are visible
      {
        const int nq = q + AS_STAGES - 1;            // refill the stage consumed in iteration q-1
        if (nq < total) {
          issue_slab(p, pre, ring + (nq % AS_STAGES) * AS_STAGE_DOUBLES, tid);
          pre.advance(j);
        }
        cp_async_commit();
      }
      const double* As = ring + (q % AS_STAGES) * AS_STAGE_DOUBLES + arow;
      const double* Bs = (cur.k < cur.i ? Wc + (cur.k * PB) * AS_LDB : Tt) + (cur.slab * AS_KS + jj) * AS_LDB + g;
#pragma unroll
      for (int kk = 0; kk < AS_KS / 4; kk++) {
        const double b = Bs[kk * 4 * AS_LDB];
        dmma8x8x4(acc[0], As[kk * 4], b);
        dmma8x8x4(acc[1], As[8 * AS_LDA + kk * 4], b);
      }
      if (cur.slab == PB / AS_KS - 1) {
        if (cur.k == cur.i - 1) {
          // block row finished accumulating: T -> shared memory, B operand of the inv(L_ii) step
#pragma unroll
          for (int mi = 0; mi < 2; mi++) {
            double* t = Tt + (warp * 16 + mi * 8 + g) * AS_LDB + 2 * jj;
            t[0] = acc[mi][0];
            t[1] = acc[mi][1];
            acc[mi][0] = acc[mi][1] = 0.0;
          }
        } else if (cur.k == cur.i) {
          // W_ij strip = -inv(L_ii) T: becomes part of the B operand of the next block rows
#pragma unroll
          for (int mi = 0; mi < 2; mi++) {
            double* t = Wc + (cur.i * PB + warp * 16 + mi * 8 + g) * AS_LDB + 2 * jj;
            t[0] = -acc[mi][0];
            t[1] = -acc[mi][1];
            acc[mi][0] = acc[mi][1] = 0.0;
          }
        }
      }
      cur.advance(j);
    }
    cp_async_wait<0>();
    __syncthreads();
    // store rows [128 j, w) of the strip (16-byte pairs) into every rank's W
    const int r_lo = j * PB, nrows = p.w - r_lo;
    for (int e = tid; e < nrows * (AS_CS / 2); e += AS_THREADS) {
      const int r = r_lo + (e >> 2), c = (e & 3) * 2;
      const double v0 = Wc[r * AS_LDB + c], v1 = Wc[r * AS_LDB + c + 1];
#pragma unroll
      for (int q = 0; q < MAX_PEERS; q++) {
        if (q >= p.P || p.out[q] == nullptr) continue;
        double* dst = p.out[q] + (long long)r * p.ldw + c0 + c;
        if (c0 + c + 1 < p.w) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        else if (c0 + c < p.w) dst[0] = v0;
      }
    }
  }
  if (p.sig.counter != nullptr) {
    __threadfence_system();
    __syncthreads();
    if (tid == 0) signal_if_last_cta(p.sig, gridDim.x * gridDim.y);
  }
}

}  // namespace

bool assemble_inverse_ok(const double* L, long long ldl, int w, long long ldw) {
  return w > 0 && w <= AS_MAXW && (reinterpret_cast<uintptr_t>(L) & 15) == 0 && (ldl % 2) == 0 && (ldw % 2) == 0;
}

cudaError_t launch_assemble_inverse(cudaStream_t s, const double* L, long long ldl, int w, const double* linv_blocks,
                                    double* const* out_ptrs, int P, long long ldw, const PeerSignal* sig) {
  if (!assemble_inverse_ok(L, ldl, w, ldw) || P < 1 || P > MAX_PEERS) return cudaErrorInvalidValue;
  AssembleParams p{};
  p.L = L; p.ldl = ldl; p.linv = linv_blocks; p.w = w; p.ldw = ldw; p.P = P;
  for (int q = 0; q < P; q++) p.out[q] = out_ptrs[q];
  if (sig != nullptr) p.sig = *sig;
  cudaError_t e = configure_kernel_once(reinterpret_cast<const void*>(assemble_inverse_kernel), AS_SMEM_BYTES, false);
  if (e != cudaSuccess) return e;
  dim3 grid(PB / AS_CS, (unsigned)((w + PB - 1) / PB));
  assemble_inverse_kernel<<<grid, AS_THREADS, AS_SMEM_BYTES, s>>>(p);
  instr().launches++;
  return cudaGetLastError();
}

}  // namespace smnngp
