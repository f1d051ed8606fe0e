// Multi-GPU panel exchange over NVLink peer memory (one process per GPU, buffers shared through CUDA IPC, peer.cu).
// Replaces the two NCCL collectives per panel of the block-row-cyclic Cholesky by stores:
//
//   owner of panel p : factors the w x w diagonal block with identity rows carried along -> L^-T, transposes it
//                      into W = inv(L_pp) (lower triangular, row-major) and stores W into EVERY rank's buffer,
//                      then raises a flag on every rank                                    (was: ncclBroadcast)
//   every rank       : ONE kernel computes its panel rows  R W^T  on the tensor pipe (TMA-fed persistent core,
//                      contraction cut at the diagonal of W) and its epilogue stores each tile into the panel
//                      buffer of every rank at the tile's GLOBAL row, plus a local-order copy that serves as the
//                      A operand of the trailing update; the last CTA raises this rank's flag on every rank
//                                                      (was: 7 launches + copy + ncclAllGather + index_select)
//   consumers        : a one-thread kernel spins on the local flags (acquire, system scope).
//
// Ordering: every math thread fences (system scope) after its last remote store, the CTA synchronises, one thread
// bumps a device counter; the CTA that sees the final count fences again and writes the flags with release
// semantics.  Flags carry a monotonically increasing sequence number, so they never need resetting.
#include "../../include/smnngp.h"

#include "gemm_core.cuh"
#include "context.cuh"
#include "kernels.cuh"
#include "peer_signal.cuh"
#include "tma_core.cuh"

namespace smnngp {

namespace {

struct ScatterParams {
  GemmParams g;                 // A = panel rows R [m, w], B = W [w, w], C = local-order copy [m, ldc]
  double* peer[MAX_PEERS];      // panel buffer of every rank: row 0 = global row c1, pitch ld_peer
  long long ld_peer;
  long long local_row0;         // local row index of R's first row (local storage order)
  long long c1, n;              // first global row held by the panel buffers; rows >= n (the y^T row) stay local
  int db, P, rank;
  long long even_off, odd_off;  // global block of local block lb = lb * P + (lb odd ? odd_off : even_off)  (cyclic: rank)
  PeerSignal sig;
};

// panel solve + scatter: epilogue of the TMA core
struct EpiScatterTma {
  using Params = ScatterParams;
  static __device__ __forceinline__ void apply(const Params& p, double (&acc)[MI][NI][2], int r0, int c0, int wm,
                                               int wn, int lane) {
    const int rbase = r0 + wm * 64 + (lane >> 2), cbase = c0 + wn * 32 + (lane & 3) * 2;
#pragma unroll
    for (int mi = 0; mi < MI; mi++) {
      const int r = rbase + mi * 8;
      if (r >= p.g.M) continue;
      const long long lr = p.local_row0 + r;
      const long long lb = lr / p.db;                                        // global row of this local row
      const long long g = (lb * p.P + ((lb & 1) ? p.odd_off : p.even_off)) * p.db + lr % p.db;
      const bool remote = g < p.n;
      const long long prow = (g - p.c1) * p.ld_peer;
#pragma unroll
      for (int ni = 0; ni < NI; ni++) {
        const int c = cbase + ni * 8;
        if (c >= p.g.N) continue;
        const double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        if (c + 1 < p.g.N) {
          *reinterpret_cast<double2*>(p.g.C + (long long)r * p.g.ldc + c) = v;
          if (remote) {
#pragma unroll
            for (int q = 0; q < MAX_PEERS; q++)
              if (q < p.P && p.peer[q] != nullptr) *reinterpret_cast<double2*>(p.peer[q] + prow + c) = v;
          }
        } else {                                                              // odd panel width: last column
          p.g.C[(long long)r * p.g.ldc + c] = v.x;
          if (remote) {
#pragma unroll
            for (int q = 0; q < MAX_PEERS; q++)
              if (q < p.P && p.peer[q] != nullptr) p.peer[q][prow + c] = v.x;
          }
        }
      }
    }
  }
  static __device__ __forceinline__ void finish(const Params& p) {
    __threadfence_system();
    asm volatile("bar.sync 1, 256;" ::: "memory");            // the 8 math warps; the producer warps have left
    if (threadIdx.x == 128) signal_if_last_cta(p.sig, gridDim.x);
  }
};

// W = (L^-T)^T for the w x w block: Ut [w, ldu] upper triangular -> every peer's W [w, ldw] (lower, zeros above)
__global__ void __launch_bounds__(256) transpose_scatter_kernel(const double* __restrict__ Ut, long long ldu, int w,
                                                                long long ldw, ScatterParams sp) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;       // output tile: rows by.., cols bx..  (W[j][k])
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int k = bx + i, j = by + tx;                        // read Ut[k][j]
    tile[i][tx] = (k < w && j < w && k <= j) ? Ut[(long long)k * ldu + j] : 0.0;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int j = by + i, k = bx + tx;
    if (j < w && k < w) {
      const double v = tile[tx][i];
#pragma unroll
      for (int q = 0; q < MAX_PEERS; q++)
        if (q < sp.P) sp.peer[q][(long long)j * ldw + k] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) signal_if_last_cta(sp.sig, gridDim.x * gridDim.y);
}

__global__ void signal_kernel(PeerSignal sg) {
  __threadfence_system();
  for (int q = 0; q < sg.P; q++)
    if (sg.flag[q] != nullptr) st_release_sys(sg.flag[q], sg.seq);
}

// one thread: wait until flags[0 .. count) >= seq.  A peer that died must not hang this GPU: after `timeout_ns`
// the wait gives up and poisons *info (INT_MAX), which turns every result of the evaluation into NaN.
__global__ void wait_flags_kernel(const unsigned long long* __restrict__ flags, int count, unsigned long long seq,
                                  unsigned long long timeout_ns, int* __restrict__ info) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i < count; i++) {
    while (ld_acquire_sys(flags + i) < seq) {
      __nanosleep(100);
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) {
        if (info != nullptr) atomicMax(info, 0x7fffffff);
        return;
      }
    }
  }
  __threadfence_system();
}

// diag block -> factorisation scratch: T [2w, ldt]: top = lower triangle of the block, bottom = identity
__global__ void stage_diag_kernel(const double* __restrict__ A, long long lda, int w, double* __restrict__ T,
                                  long long ldt) {
  const int i = blockIdx.x;                                   // 0 .. 2w-1
  for (int c = threadIdx.x; c < w; c += blockDim.x) {
    double v;
    if (i < w) v = (c <= i) ? A[(long long)i * lda + c] : 0.0;
    else v = (c == i - w) ? 1.0 : 0.0;
    T[(long long)i * ldt + c] = v;
  }
}

// cuStreamWaitValue64: the wait is executed by the stream's front-end, no SM is occupied while waiting
typedef CUresult (*PFN_streamWaitValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
PFN_streamWaitValue64 stream_wait_fn() {
  static const PFN_streamWaitValue64 fn = []() -> PFN_streamWaitValue64 {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<PFN_streamWaitValue64>(p);
    return nullptr;
  }();
  return fn;
}
// 0 = stream memory operation (default: no SM is held; the time-out is the caller's watchdog - csrc/multigpu.cu has one
// per handle, distributed.py one per driver), 1 = one-thread spin kernel with its own time-out.  Per-device knob.
int& wait_mode() { return dctx().peer_wait_mode; }

void fill_signal(PeerSignal& sg, void* const* flag_ptrs, int P, long long flag_index, unsigned long long seq,
                 unsigned int* counter) {
  sg.P = P;
  sg.seq = seq;
  sg.counter = counter;
  for (int q = 0; q < MAX_PEERS; q++)
    sg.flag[q] = (q < P && flag_ptrs[q] != nullptr) ? static_cast<unsigned long long*>(flag_ptrs[q]) + flag_index
                                                    : nullptr;
}

}  // namespace
}  // namespace smnngp

using namespace smnngp;

extern "C" {

// Factor the w x w diagonal block at A (lower part read, A itself is NOT modified) inside the scratch T [2w, w]
// with identity rows carried along: afterwards T rows 0..w-1 hold L (lower), rows w..2w-1 hold L^-T.
// logdet_dev += sum log L_ii; linv_ws: LINV_BLOCKS * 128 * 128 doubles.
int smnngp_stage_factor_diag_inv_f64(void* stream, const double* A, int64_t lda, int64_t w, double* T,
                                     double* linv_ws, double* logdet_dev, int* info_dev, int64_t gcol0) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!A || !T || !linv_ws || !logdet_dev || !info_dev || w <= 0 || w > LINV_BLOCKS * PB) return SMNNGP_EINVAL;
  (void)gcol0;
  stage_diag_kernel<<<(unsigned)(2 * w), 128, 0, s>>>(A, lda, (int)w, T, w);
  instr().launches++;
  if (cudaGetLastError() != cudaSuccess) return SMNNGP_ECUDA;
  cudaError_t e = potrf_trapezoid(s, T, w, 2 * w, w, (int)((w + PB - 1) / PB * PB), linv_ws, logdet_dev, info_dev,
                                  (long long)PB * PB, w);
  return e == cudaSuccess ? SMNNGP_OK : SMNNGP_ECUDA;
}

// W = inv(L) (row-major lower, pitch ldw) from Ut = L^-T [w, ldu] into the buffer of every rank, then flag
int smnngp_stage_scatter_inverse_f64(void* stream, const double* Ut, int64_t ldu, int64_t w, void* const* dst_ptrs,
                                     int P, int64_t ldw, void* const* flag_ptrs, int64_t flag_index, uint64_t seq,
                                     unsigned int* counter) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!Ut || !dst_ptrs || !flag_ptrs || !counter || P < 1 || P > MAX_PEERS || w <= 0) return SMNNGP_EINVAL;
  ScatterParams sp{};
  sp.P = P;
  for (int q = 0; q < P; q++) sp.peer[q] = static_cast<double*>(dst_ptrs[q]);
  fill_signal(sp.sig, flag_ptrs, P, flag_index, seq, counter);
  dim3 grid((unsigned)((w + 31) / 32), (unsigned)((w + 31) / 32));
  transpose_scatter_kernel<<<grid, 256, 0, s>>>(Ut, ldu, (int)w, ldw, sp);
  instr().launches++;
  return cudaGetLastError() == cudaSuccess ? SMNNGP_OK : SMNNGP_ECUDA;
}

// W = inv(L) (row-major lower, pitch ldw) of the factored w x w block L (pitch ldl) from L and the inverses of its
// 128 x 128 diagonal blocks, stored straight into the W buffer of every rank, then flag (trtri.cu)
int smnngp_stage_assemble_inverse_f64(void* stream, const double* L, int64_t ldl, int64_t w, const double* linv_blocks,
                                      void* const* dst_ptrs, int P, int64_t ldw, void* const* flag_ptrs,
                                      int64_t flag_index, uint64_t seq, unsigned int* counter) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!L || !linv_blocks || !dst_ptrs || P < 1 || P > MAX_PEERS || w <= 0 || !assemble_inverse_ok(L, ldl, (int)w, ldw))
    return SMNNGP_EINVAL;
  double* outs[MAX_PEERS] = {};
  for (int q = 0; q < P; q++) outs[q] = static_cast<double*>(dst_ptrs[q]);
  PeerSignal sg{};
  if (flag_ptrs != nullptr && counter != nullptr) fill_signal(sg, flag_ptrs, P, flag_index, seq, counter);
  return launch_assemble_inverse(s, L, ldl, (int)w, linv_blocks, outs, P, ldw, counter ? &sg : nullptr) == cudaSuccess
             ? SMNNGP_OK
             : SMNNGP_ECUDA;
}

int smnngp_stage_signal_f64(void* stream, void* const* flag_ptrs, int P, int64_t flag_index, uint64_t seq) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!flag_ptrs || P < 1 || P > MAX_PEERS) return SMNNGP_EINVAL;
  PeerSignal sg{};
  fill_signal(sg, flag_ptrs, P, flag_index, seq, nullptr);
  signal_kernel<<<1, 1, 0, s>>>(sg);
  instr().launches++;
  return cudaGetLastError() == cudaSuccess ? SMNNGP_OK : SMNNGP_ECUDA;
}

int smnngp_stage_wait_flags_f64(void* stream, const void* flags_local, int64_t first, int count, uint64_t seq,
                                double timeout_s, int* info_dev) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!flags_local || count < 0) return SMNNGP_EINVAL;
  if (count == 0) return SMNNGP_OK;
  if (wait_mode() == 0 && stream_wait_fn() != nullptr) {
    for (int i = 0; i < count; i++) {
      const CUdeviceptr a = reinterpret_cast<CUdeviceptr>(static_cast<const unsigned long long*>(flags_local) + first + i);
      if (stream_wait_fn()(s, a, (cuuint64_t)seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) return SMNNGP_ECUDA;
    }
    return SMNNGP_OK;
  }
  const unsigned long long tns = (unsigned long long)((timeout_s > 0.0 ? timeout_s : 10.0) * 1e9);
  wait_flags_kernel<<<1, 1, 0, s>>>(static_cast<const unsigned long long*>(flags_local) + first, count, seq, tns,
                                    info_dev);
  instr().launches++;
  return cudaGetLastError() == cudaSuccess ? SMNNGP_OK : SMNNGP_ECUDA;
}

// Push this rank's solved panel rows (local order, whole distribution blocks of db rows, pitch db = contiguous
// blocks) to their global position in every OTHER rank's panel buffer with the copy engines - no SM, full NVLink
// rate (remote 16-byte stores issued from the few SMs the look-ahead chain owns reached < 100 GB/s) - then raise
// this rank's flag on every rank.  Rows whose global index is >= n (the appended y^T row) stay local.
int smnngp_stage_push_panel_f64(void* stream, const double* Ploc, int64_t m, int64_t w, int64_t db, int P, int rank,
                                int64_t local_row0, int64_t c1, int64_t n, void* const* peer_ptrs,
                                void* const* flag_ptrs, int64_t flag_index, uint64_t seq) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!peer_ptrs || !flag_ptrs || P < 1 || P > MAX_PEERS || m < 0 || w != db || db <= 0 || rank < 0 || rank >= P ||
      (m > 0 && !Ploc) || local_row0 % db != 0)
    return SMNNGP_EINVAL;
  const size_t blk_bytes = (size_t)db * db * sizeof(double);
  // local blocks lb0, lb0+1, ... <-> global blocks (lb * P + rank); keep only rows with global index < n
  const int64_t lb0 = local_row0 / db;
  int64_t full = 0, tail_rows = 0;
  for (int64_t r = 0; r < m; r += db) {
    const int64_t g0 = ((lb0 + r / db) * P + rank) * db;
    const int64_t rows = (m - r < db) ? m - r : db;
    const int64_t ok = (g0 + rows <= n) ? rows : (n > g0 ? n - g0 : 0);
    if (ok == db) full++;
    else { tail_rows = ok; break; }
  }
  for (int q = 0; q < P; q++) {
    if (q == rank) continue;
    char* dst0 = static_cast<char*>(peer_ptrs[q]) + (size_t)((lb0 * P + rank) * db - c1) * db * sizeof(double);
    if (full > 0 &&
        cudaMemcpy2DAsync(dst0, (size_t)P * blk_bytes, Ploc, blk_bytes, blk_bytes, (size_t)full, cudaMemcpyDeviceToDevice,
                          s) != cudaSuccess)
      return SMNNGP_ECUDA;
    if (tail_rows > 0 &&
        cudaMemcpyAsync(dst0 + (size_t)full * P * blk_bytes, reinterpret_cast<const char*>(Ploc) + (size_t)full * blk_bytes,
                        (size_t)tail_rows * db * sizeof(double), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return SMNNGP_ECUDA;
  }
  return smnngp_stage_signal_f64(stream, flag_ptrs, P, flag_index, seq);
}

// 0 = cuStreamWaitValue64 (default), 1 = spin kernel with timeout
void smnngp_set_peer_wait_mode(int mode) { wait_mode() = mode ? 1 : 0; }

// Panel solve + all-gather in one kernel:  Ploc [m, ldp] = R [m, w] W^T  and the same rows stored at their global
// position (global row - c1) of every rank's panel buffer; then this rank's flag (index flag_index) is raised on
// every rank.  local_row0 = local storage index of R's first row; rows whose global index is >= n stay local.
int smnngp_stage_trsm_scatter2_f64(void* stream, const double* R, int64_t ldr, int64_t m, int64_t w, const double* W,
                                  int64_t ldw, double* Ploc, int64_t ldp, void* const* peer_ptrs, int P, int rank,
                                  int64_t db, int64_t local_row0, int64_t c1, int64_t n, int64_t ld_peer,
                                  void* const* flag_ptrs, int64_t flag_index, uint64_t seq, unsigned int* counter,
                                   int64_t even_off, int64_t odd_off) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!W || !peer_ptrs || !flag_ptrs || !counter || P < 1 || P > MAX_PEERS || m < 0 || w <= 0 ||
      db <= 0 || rank < 0 || rank >= P)
    return SMNNGP_EINVAL;
  if (m == 0) return smnngp_stage_signal_f64(stream, flag_ptrs, P, flag_index, seq);
  if (!R || !Ploc || !tma_operand_ok(R, ldr) || !tma_operand_ok(W, ldw) || (ldp & 1) || (ld_peer & 1))
    return SMNNGP_EINVAL;
  ScatterParams sp{};
  sp.g.A = R; sp.g.lda = ldr; sp.g.B = W; sp.g.ldb = ldw; sp.g.C = Ploc; sp.g.ldc = ldp;
  sp.g.M = (int)m; sp.g.N = (int)w; sp.g.K = (int)w;
  for (int q = 0; q < P; q++) sp.peer[q] = static_cast<double*>(peer_ptrs[q]);
  sp.ld_peer = ld_peer; sp.local_row0 = local_row0; sp.c1 = c1; sp.n = n;
  sp.db = (int)db; sp.P = P; sp.rank = rank; sp.even_off = even_off; sp.odd_off = odd_off;
  fill_signal(sp.sig, flag_ptrs, P, flag_index, seq, counter);
  CUtensorMap ma, mb;
  if (!make_tmap(&ma, R, m, w, ldr, TM_BM) || !make_tmap(&mb, W, w, w, ldw, TM_BN)) return SMNNGP_ECUDA;
  TmaShape sh{(int)m, (int)w, (int)w, 0, count_tiles<TileTma>(m, w, 0), 0, 1, 0, 0, 1};
  cudaError_t e = launch_tma_gemm<EpiScatterTma>(s, ma, mb, sh, sp, device_sm_count());
  instr().launches++;
  return e == cudaSuccess ? SMNNGP_OK : SMNNGP_ECUDA;
}

int smnngp_stage_trsm_scatter_f64(void* stream, const double* R, int64_t ldr, int64_t m, int64_t w, const double* W,
                                  int64_t ldw, double* Ploc, int64_t ldp, void* const* peer_ptrs, int P, int rank,
                                  int64_t db, int64_t local_row0, int64_t c1, int64_t n, int64_t ld_peer,
                                  void* const* flag_ptrs, int64_t flag_index, uint64_t seq, unsigned int* counter) {
  return smnngp_stage_trsm_scatter2_f64(stream, R, ldr, m, w, W, ldw, Ploc, ldp, peer_ptrs, P, rank, db, local_row0, c1, n,
                                        ld_peer, flag_ptrs, flag_index, seq, counter, rank, rank);
}

}  // extern "C"
