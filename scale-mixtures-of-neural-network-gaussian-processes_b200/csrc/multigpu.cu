// Multi-GPU exact-GP log marginal likelihood behind ONE C call per rank (include/smnngp.h: smnngp_mg_*,
// smnngp_lml_mg_f64): the panel loop of the block-row-cyclic Cholesky, the NVLink peer-store panel exchange and the
// final cross-rank reduction all run inside the library, enqueue-only on the caller's stream.  The caller supplies one
// handle per rank (one process per GPU - handles connected through 64-byte CUDA IPC handles the caller moves between
// the processes once - or several devices in one process, connected by pointer) and nothing else: no NCCL
// communicator, no Python.
//
// Layout (P ranks, distribution block DB = outer panel width): global block b = rows
// [b DB, (b+1) DB) of K + eps I lives on rank b mod P, full width, lower part only; global row N = y^T is carried
// through the factorisation and comes out as (L^-1 y)^T.  Every rank holds all of X and generates exactly the rows it
// owns.  Per panel p: the owner factors the DB x DB diagonal block in place and assembles its full inverse W straight
// into EVERY rank's W buffer (trtri.cu) -> every rank solves its own panel rows with one TMA GEMM whose epilogue
// stores each tile at its GLOBAL row of every rank's panel buffer (exchange.cu) -> every rank updates its trailing rows
// with one launch.  One panel of look-ahead: panel p+1 is prepared on a high-priority side stream under the bulk of
// panel p's update.  sum log L_ii, ||L^-1 y||^2 and info are exchanged the same way (every rank stores its part into
// every rank's reduce slots, the sum runs in rank order: bit-identical on every rank).
//
// A dead peer cannot hang a GPU: waits are stream memory operations (no SM is held) and a per-handle host watchdog
// thread releases them after `timeout_s` by writing the flags itself and poisoning info (results become NaN).
#include "../../include/smnngp.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "context.cuh"
#include "kernels.cuh"
#include "peer_signal.cuh"

using namespace smnngp;

namespace {

constexpr int FLAG_WORDS = 32;     // word 0: W ready; 8..15: panel ready (per source rank); 16..23: reduce slot ready
constexpr int FLAG_W = 0, FLAG_PANEL = 8, FLAG_REDUCE = 16;
constexpr int REDUCE_DOUBLES = 4;  // per source rank: sum log L_ii, ||z||^2, info (as double), pad

inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

struct TimelineMark { int panel, label; cudaEvent_t ev; };

}  // namespace

struct smnngp_mg {
  int rank = 0, P = 1, device = 0;
  long long n = 0, extra = 1, db = 512, ld = 0, mtotal = 0, nblocks = 0, mloc = 0;
  bool emulate = false, connected = false;
  std::atomic<bool> dead{false};
  // local storage
  double* a = nullptr;                    // [mloc, ld] block-row-cyclic rows
  double *tab = nullptr, *q = nullptr, *scal = nullptr, *linv = nullptr, *zvec = nullptr, *sums = nullptr;
  double* ploc[2] = {nullptr, nullptr};   // solved panel rows in local order (A operand of the update)
  unsigned int* counters = nullptr;
  size_t tab_doubles = 0;
  // peer-visible region: [W db x db | flags | 3 panel slots n x db | reduce slots P x 4]
  char* region = nullptr;
  size_t off_w = 0, off_flags = 0, off_panel = 0, slot_bytes = 0, off_reduce = 0, region_bytes = 0;
  unsigned char handle[64] = {};
  char* bases[MAX_PEERS] = {};
  bool opened[MAX_PEERS] = {};
  unsigned long long seq_base = 0;
  cudaStream_t side = nullptr, poison_stream = nullptr;
  cudaEvent_t ev_panel = nullptr, ev_a = nullptr, ev_fork = nullptr, ev_done = nullptr;
  // tuning
  int sm_reserve_override = -1;
  double timeout_s = 20.0;
  // timeline (profiling)
  bool timeline_on = false;
  std::vector<TimelineMark> marks;
  // watchdog
  std::thread dog;
  std::mutex dog_mu;
  std::condition_variable dog_cv;
  bool dog_stop = false, dog_armed = false;
  unsigned long long dog_gen = 0;         // bumped by every arm: a query that raced with a re-arm must not disarm
  std::chrono::steady_clock::time_point dog_deadline;
  int* dog_info = nullptr;

  // ---- layout (same rules as distributed.BlockRowCyclic) ----
  int owner(long long b) const { return (int)(b % P); }
  long long block_rows(long long b) const { return std::min((b + 1) * db, mtotal) - b * db; }
  long long local_rows(int r) const {
    long long s = 0;
    for (long long b = r; b < nblocks; b += P) s += block_rows(b);
    return s;
  }
  long long local_offset(long long b) const {          // local row offset of global block b on its owner
    long long s = 0;
    for (long long x = owner(b); x < b; x += P) s += block_rows(x);
    return s;
  }
  long long first_block_from(long long gb, int r) const { return gb + (((r - gb) % P) + P) % P; }
  void rows_from_block(long long gb, int r, long long& off, long long& cnt) const {
    const long long fb = first_block_from(gb, r);
    const long long all = local_rows(r);
    if (fb >= nblocks) { off = all; cnt = 0; return; }
    off = 0;
    for (long long x = r; x < fb; x += P) off += block_rows(x);
    cnt = all - off;
  }
  unsigned long long* flags_local() const { return reinterpret_cast<unsigned long long*>(region + off_flags); }
  double* w_local() const { return reinterpret_cast<double*>(region + off_w); }
  double* panel_local(int slot) const { return reinterpret_cast<double*>(region + off_panel + slot * slot_bytes); }
};

namespace {

thread_local char g_mg_err[256] = "";
int mg_fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  if (e != cudaSuccess) snprintf(g_mg_err, sizeof g_mg_err, "%s: %s", what, cudaGetErrorString(e));
  else snprintf(g_mg_err, sizeof g_mg_err, "%s", what);
  return code;
}
#define MG_CU(call)                                                        \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) return mg_fail(SMNNGP_ECUDA, #call, e__);      \
  } while (0)
#define MG_RC(call)                                                        \
  do {                                                                     \
    int rc__ = (call);                                                     \
    if (rc__ != SMNNGP_OK) return mg_fail(rc__, #call);                    \
  } while (0)

void mark(smnngp_mg* g, cudaStream_t s, int panel, int label) {
  if (!g->timeline_on) return;
  TimelineMark m{panel, label, nullptr};
  if (cudaEventCreate(&m.ev) != cudaSuccess) return;
  cudaEventRecord(m.ev, s);
  g->marks.push_back(m);
}

// host watchdog: if the armed evaluation has not finished by its deadline, release every wait of this rank by writing
// the flag words (the waits compare cyclically: a value far ahead of any sequence number satisfies them all) and
// poison info so that every result of the evaluation becomes NaN.  The handle is unusable afterwards.
void watchdog_main(smnngp_mg* g) {
  cudaSetDevice(g->device);
  std::unique_lock<std::mutex> lk(g->dog_mu);
  while (!g->dog_stop) {
    if (!g->dog_armed) {
      g->dog_cv.wait(lk);
      continue;
    }
    const unsigned long long gen = g->dog_gen;
    lk.unlock();
    const cudaError_t st = cudaEventQuery(g->ev_done);
    lk.lock();
    if (gen != g->dog_gen) continue;         // re-armed meanwhile: look again
    if (st != cudaErrorNotReady) {           // finished (or failed): nothing to release
      g->dog_armed = false;
      continue;
    }
    if (std::chrono::steady_clock::now() >= g->dog_deadline) {
      g->dead = true;
      g->dog_armed = false;
      int* info = g->dog_info;
      lk.unlock();
      static const int poison = 0x7fffffff;
      std::vector<unsigned long long> big(FLAG_WORDS, g->seq_base + (1ull << 40));
      if (info) cudaMemcpyAsync(info, &poison, sizeof(int), cudaMemcpyHostToDevice, g->poison_stream);
      cudaMemcpyAsync(g->flags_local(), big.data(), FLAG_WORDS * 8, cudaMemcpyHostToDevice, g->poison_stream);
      cudaStreamSynchronize(g->poison_stream);
      lk.lock();
      continue;
    }
    g->dog_cv.wait_for(lk, std::chrono::milliseconds(50));
  }
}

void arm_watchdog(smnngp_mg* g, cudaStream_t s, int* info_dev) {
  cudaEventRecord(g->ev_done, s);
  {
    std::lock_guard<std::mutex> lk(g->dog_mu);
    g->dog_armed = true;
    g->dog_gen++;
    g->dog_info = info_dev;
    g->dog_deadline = std::chrono::steady_clock::now() +
                      std::chrono::milliseconds((long long)(g->timeout_s * 1000.0));
  }
  g->dog_cv.notify_all();
}

void fill_ptrs(smnngp_mg* g, size_t off, void** out) {
  for (int r = 0; r < MAX_PEERS; r++) out[r] = (r < g->P && g->bases[r]) ? g->bases[r] + off : nullptr;
}

// every rank's contribution {sum log L_ii, ||z||^2, info} -> slot `rank` of every rank's reduce area, then flag
__global__ void reduce_scatter_kernel(const double* __restrict__ sums, const int* __restrict__ info, int rank,
                                      double* const p0, double* const p1, double* const p2, double* const p3,
                                      double* const p4, double* const p5, double* const p6, double* const p7, PeerSignal sg) {
  double* const dst[MAX_PEERS] = {p0, p1, p2, p3, p4, p5, p6, p7};
  const double v[REDUCE_DOUBLES] = {sums[0], sums[1], (double)info[0], 0.0};
  for (int qd = 0; qd < sg.P; qd++) {
    if (dst[qd] == nullptr) continue;
    for (int i = 0; i < REDUCE_DOUBLES; i++) dst[qd][rank * REDUCE_DOUBLES + i] = v[i];
  }
  __threadfence_system();
  for (int qd = 0; qd < sg.P; qd++)
    if (sg.flag[qd] != nullptr) st_release_sys(sg.flag[qd], sg.seq);
}

// fixed-order sum over the ranks' slots -> scal[SC_LOGDET], scal[SC_QUAD], info = max
__global__ void reduce_gather_kernel(const double* __restrict__ slots, int P, double* __restrict__ scal,
                                     int* __restrict__ info) {
  double ld = 0.0, qd = 0.0, bad = 0.0;
  for (int r = 0; r < P; r++) {
    ld += slots[r * REDUCE_DOUBLES + 0];
    qd += slots[r * REDUCE_DOUBLES + 1];
    bad = fmax(bad, slots[r * REDUCE_DOUBLES + 2]);
  }
  scal[SC_LOGDET] = ld;
  scal[SC_QUAD] = qd;
  if (bad > 0.0) atomicMax(info, (int)fmin(bad, 2147483647.0));
}

// one panel on stream s: owner factors + publishes W, everybody solves + scatters, waits for the whole panel.
// Returns this rank's rows below the diagonal block (local offset ls, count m).
int panel_step(smnngp_mg* g, cudaStream_t s, long long p, int* info_dev, long long& ls, long long& m) {
  const long long n = g->n, db = g->db;
  const int P = g->P;
  const long long c0 = p * db, c1 = std::min((p + 1) * db, n), w = c1 - c0;
  const int own = g->owner(p);
  const unsigned long long seq = g->seq_base + (unsigned long long)p + 1;
  void* w_ptrs[MAX_PEERS];
  void* flag_ptrs[MAX_PEERS];
  void* panel_ptrs[MAX_PEERS];
  const int slot = (int)(p % 3);
  fill_ptrs(g, g->off_w, w_ptrs);
  fill_ptrs(g, g->off_flags, flag_ptrs);
  fill_ptrs(g, g->off_panel + slot * g->slot_bytes, panel_ptrs);
  if (g->rank == own) {
    const long long lo = g->local_offset(p);
    double* blk = g->a + lo * g->ld + c0;
    // diagonal block in place (L_pp + inverses of its 128-blocks), then its full inverse into every rank's W
    MG_CU(potrf_trapezoid(s, blk, g->ld, w, w, (int)(cdiv(w, PB) * PB), g->linv, g->sums, info_dev,
                          (long long)PB * PB));
    PeerSignal sg{};
    sg.P = P; sg.seq = seq; sg.counter = g->counters;
    for (int r = 0; r < MAX_PEERS; r++)
      sg.flag[r] = flag_ptrs[r] ? static_cast<unsigned long long*>(flag_ptrs[r]) + FLAG_W : nullptr;
    double* outs[MAX_PEERS];
    for (int r = 0; r < MAX_PEERS; r++) outs[r] = static_cast<double*>(w_ptrs[r]);
    MG_CU(launch_assemble_inverse(s, blk, g->ld, (int)w, g->linv, outs, P, db, &sg));
  }
  mark(g, s, (int)p, 0);                                                          // diag
  if (!g->emulate || g->rank == own)
    MG_RC(smnngp_stage_wait_flags_f64(s, g->flags_local(), FLAG_W, 1, seq, g->timeout_s, info_dev));
  mark(g, s, (int)p, 1);                                                          // bcast
  if (g->rank == own) {
    ls = g->local_offset(p) + w;
    m = g->mloc - ls;
  } else {
    g->rows_from_block(p + 1, g->rank, ls, m);
  }
  double* ploc = g->ploc[p & 1];
  MG_RC(smnngp_stage_trsm_scatter_f64(s, g->a + ls * g->ld + c0, g->ld, m, w, g->w_local(), db, ploc, db, panel_ptrs, P,
                                      g->rank, db, ls, c1, n, db, flag_ptrs, FLAG_PANEL + g->rank, seq,
                                      g->counters + 4));
  mark(g, s, (int)p, 2);                                                          // trsm
  // z = L^-1 y: the carried row (global row n) sits at the end of its owner's storage
  const long long bn = n / db;
  if (m > 0 && g->rank == g->owner(bn)) {
    const long long lrow = g->local_offset(bn) + (n - bn * db);
    if (lrow >= ls)
      MG_CU(cudaMemcpyAsync(g->zvec + c0, ploc + (lrow - ls) * db, (size_t)w * 8, cudaMemcpyDeviceToDevice, s));
  }
  if (c1 >= n) return SMNNGP_OK;
  if (g->emulate) MG_RC(smnngp_stage_wait_flags_f64(s, g->flags_local(), FLAG_PANEL + g->rank, 1, seq, g->timeout_s, info_dev));
  else MG_RC(smnngp_stage_wait_flags_f64(s, g->flags_local(), FLAG_PANEL, P, seq, g->timeout_s, info_dev));
  mark(g, s, (int)p, 3);                                                          // gather
  return SMNNGP_OK;
}

int build_gram(smnngp_mg* g, cudaStream_t s, const double* X, const double* y, long long D, int nh, int act, int arch,
               const double* hp, int shift) {
  const long long n = g->n, db = g->db;
  const int n_act = std::max(n_act_applications(nh, arch), 1);
  if ((size_t)n_act * n > g->tab_doubles) {
    if (g->tab) cudaFree(g->tab);
    g->tab = nullptr;
    MG_CU(cudaMalloc(&g->tab, (size_t)n_act * n * 8));
    g->tab_doubles = (size_t)n_act * n;
  }
  MG_RC(smnngp_stage_qtable_f64(s, X, n, D, nh, act, arch, hp, g->tab, n, g->q, g->scal));
  for (long long b = g->rank; b < g->nblocks; b += g->P) {
    const long long g0 = b * db, lo = g->local_offset(b);
    const long long rows = std::min(g0 + g->block_rows(b), n) - g0;          // rows of the square part in this block
    if (rows > 0) {
      const double* xb = X + g0 * D;
      double* arow = g->a + lo * g->ld;
      if (g0 > 0)                                                          // rectangle left of the diagonal block
        MG_RC(smnngp_stage_gram_f64(s, xb, rows, X, g0, D, nh, act, arch, hp, g->tab + g0, n, g->tab, n, g->scal,
                                    SHIFT_NONE, 0, arow, g->ld));
      MG_RC(smnngp_stage_gram_f64(s, xb, rows, xb, rows, D, nh, act, arch, hp, g->tab + g0, n, g->tab + g0, n, g->scal,
                                  shift, 1, arow + g0, g->ld));
    }
    if (g0 <= n && n < g0 + g->block_rows(b))                              // the appended row y^T
      MG_CU(cudaMemcpyAsync(g->a + (lo + (n - g0)) * g->ld, y, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
  }
  return SMNNGP_OK;
}

}  // namespace

extern "C" {

const char* smnngp_mg_last_error(void) { return g_mg_err; }

int smnngp_mg_create(smnngp_mg** out, int rank, int world, int64_t n, int64_t block) {
  if (!out || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || n <= 0 || block <= 0 || block % PB != 0 ||
      block > LINV_BLOCKS * PB || n + 1 > INT32_MAX)
    return mg_fail(SMNNGP_EINVAL, "smnngp_mg_create: invalid argument");
  smnngp_mg* g = new smnngp_mg();
  g->rank = rank; g->P = world; g->n = n; g->db = block;
  cudaGetDevice(&g->device);
  g->mtotal = n + g->extra;
  g->nblocks = cdiv(g->mtotal, g->db);
  g->ld = cdiv(n, 16) * 16;
  g->mloc = g->local_rows(rank);
  g->off_w = 0;
  g->off_flags = (size_t)g->db * g->db * 8;
  g->off_panel = g->off_flags + FLAG_WORDS * 8;
  g->slot_bytes = (size_t)n * g->db * 8;
  g->off_reduce = g->off_panel + 3 * g->slot_bytes;
  g->region_bytes = g->off_reduce + (size_t)MAX_PEERS * REDUCE_DOUBLES * 8;
  void* reg = nullptr;
  if (smnngp_peer_alloc(g->region_bytes, &reg, g->handle) != SMNNGP_OK) {
    delete g;
    return mg_fail(SMNNGP_ECUDA, "smnngp_mg_create: peer region allocation / IPC export failed");
  }
  g->region = static_cast<char*>(reg);
  const long long mrows = std::max<long long>(g->mloc, 1);
  bool ok = cudaMemset(g->region + g->off_flags, 0, FLAG_WORDS * 8) == cudaSuccess &&
            cudaMemset(g->region + g->off_reduce, 0, MAX_PEERS * REDUCE_DOUBLES * 8) == cudaSuccess &&
            cudaMalloc(&g->a, (size_t)mrows * g->ld * 8) == cudaSuccess &&
            cudaMalloc(&g->q, (size_t)n * 8) == cudaSuccess && cudaMalloc(&g->scal, SC_COUNT * 8) == cudaSuccess &&
            cudaMalloc(&g->linv, (size_t)LINV_BLOCKS * PB * PB * 8) == cudaSuccess &&
            cudaMalloc(&g->zvec, (size_t)n * 8) == cudaSuccess && cudaMalloc(&g->sums, 2 * 8) == cudaSuccess &&
            cudaMalloc(&g->ploc[0], (size_t)mrows * g->db * 8) == cudaSuccess &&
            cudaMalloc(&g->ploc[1], (size_t)mrows * g->db * 8) == cudaSuccess &&
            cudaMalloc(&g->counters, 8 * sizeof(unsigned int)) == cudaSuccess &&
            cudaMemset(g->counters, 0, 8 * sizeof(unsigned int)) == cudaSuccess &&
            cudaMemset(g->zvec, 0, (size_t)n * 8) == cudaSuccess;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  ok = ok && cudaStreamCreateWithPriority(&g->side, cudaStreamNonBlocking, hi) == cudaSuccess &&
       cudaStreamCreateWithFlags(&g->poison_stream, cudaStreamNonBlocking) == cudaSuccess &&
       cudaEventCreateWithFlags(&g->ev_panel, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&g->ev_a, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&g->ev_done, cudaEventDisableTiming) == cudaSuccess;
  if (!ok || cudaDeviceSynchronize() != cudaSuccess) {
    smnngp_mg_destroy(g);
    return mg_fail(SMNNGP_ECUDA, "smnngp_mg_create: device allocation failed");
  }
  g->bases[rank] = g->region;
  if (world == 1) g->connected = true;
  g->dog = std::thread(watchdog_main, g);
  *out = g;
  return SMNNGP_OK;
}

int smnngp_mg_ipc_handle(smnngp_mg* g, unsigned char* handle_out64) {
  if (!g || !handle_out64) return mg_fail(SMNNGP_EINVAL, "smnngp_mg_ipc_handle: invalid argument");
  memcpy(handle_out64, g->handle, 64);
  return SMNNGP_OK;
}

void* smnngp_mg_region(smnngp_mg* g) { return g ? g->region : nullptr; }

// one process per GPU: handles[r * 64 .. ) = rank r's handle (the caller gathered them, e.g. with one all-gather)
int smnngp_mg_connect_ipc(smnngp_mg* g, const unsigned char* handles) {
  if (!g || !handles) return mg_fail(SMNNGP_EINVAL, "smnngp_mg_connect_ipc: invalid argument");
  for (int r = 0; r < g->P; r++) {
    if (r == g->rank) continue;
    void* p = nullptr;
    if (smnngp_peer_open(handles + 64 * r, &p) != SMNNGP_OK)
      return mg_fail(SMNNGP_ECUDA, "smnngp_mg_connect_ipc: cudaIpcOpenMemHandle failed");
    g->bases[r] = static_cast<char*>(p);
    g->opened[r] = true;
  }
  g->connected = true;
  return SMNNGP_OK;
}

// several devices in ONE process: regions[r] = smnngp_mg_region() of rank r's handle, peer_devices[r] its device
int smnngp_mg_connect_ptrs(smnngp_mg* g, void* const* regions, const int* peer_devices) {
  if (!g || !regions) return mg_fail(SMNNGP_EINVAL, "smnngp_mg_connect_ptrs: invalid argument");
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(g->device);
  for (int r = 0; r < g->P; r++) {
    if (r == g->rank) continue;
    if (peer_devices && peer_devices[r] != g->device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(peer_devices[r], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        cudaSetDevice(cur);
        return mg_fail(SMNNGP_ECUDA, "smnngp_mg_connect_ptrs: cudaDeviceEnablePeerAccess", e);
      }
      cudaGetLastError();
    }
    g->bases[r] = static_cast<char*>(regions[r]);
  }
  cudaSetDevice(cur);
  g->connected = true;
  return SMNNGP_OK;
}

// timing dry-run of ONE rank of a P-rank job on a single device: every peer aliases the local region
int smnngp_mg_connect_emulated(smnngp_mg* g) {
  if (!g) return mg_fail(SMNNGP_EINVAL, "smnngp_mg_connect_emulated: invalid argument");
  for (int r = 0; r < g->P; r++) g->bases[r] = g->region;
  g->emulate = true;
  g->connected = true;
  return SMNNGP_OK;
}

void smnngp_mg_set_timeout(smnngp_mg* g, double seconds) {
  if (g && seconds > 0.0) g->timeout_s = seconds;
}
void smnngp_mg_set_sm_reserve(smnngp_mg* g, int sms) {
  if (g) g->sm_reserve_override = sms;
}
void smnngp_mg_timeline(smnngp_mg* g, int enable) {
  if (!g) return;
  for (auto& m : g->marks) cudaEventDestroy(m.ev);
  g->marks.clear();
  g->timeline_on = enable != 0;
}
// marks recorded since smnngp_mg_timeline(g, 1): milliseconds relative to the first mark; returns the count
int smnngp_mg_timeline_read(smnngp_mg* g, int cap, int* panel_out, int* label_out, double* ms_out) {
  if (!g || g->marks.empty()) return 0;
  int k = 0;
  for (auto& m : g->marks) {
    if (k >= cap) break;
    cudaEventSynchronize(m.ev);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g->marks[0].ev, m.ev);
    panel_out[k] = m.panel;
    label_out[k] = m.label;
    ms_out[k] = ms;
    k++;
  }
  return k;
}

int smnngp_mg_destroy(smnngp_mg* g) {
  if (!g) return SMNNGP_OK;
  if (g->dog.joinable()) {
    {
      std::lock_guard<std::mutex> lk(g->dog_mu);
      g->dog_stop = true;
    }
    g->dog_cv.notify_all();
    g->dog.join();
  }
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  for (auto& m : g->marks) cudaEventDestroy(m.ev);
  for (int r = 0; r < MAX_PEERS; r++)
    if (g->opened[r]) smnngp_peer_close(g->bases[r]);
  if (g->region) smnngp_peer_free(g->region);
  for (double* p : {g->a, g->tab, g->q, g->scal, g->linv, g->zvec, g->sums, g->ploc[0], g->ploc[1]})
    if (p) cudaFree(p);
  if (g->counters) cudaFree(g->counters);
  for (cudaEvent_t e : {g->ev_panel, g->ev_a, g->ev_fork, g->ev_done})
    if (e) cudaEventDestroy(e);
  if (g->side) cudaStreamDestroy(g->side);
  if (g->poison_stream) cudaStreamDestroy(g->poison_stream);
  cudaSetDevice(cur);
  delete g;
  return SMNNGP_OK;
}

// SPR.loss (spax/models.py:93-98) on the ranks that share the handle group: out_dev[4] = {log p(y), -log p(y) / N,
// sum log L_ii, ||L^-1 y||^2}, identical on every rank.  Every rank calls with the SAME X [N, D], y [N], hp_dev [6]
// (device pointers on ITS device).  shift: SMNNGP_SHIFT_* added to the Gram diagonal (SHIFT_EPS_ABS for SPR.loss).
int smnngp_lml_mg_f64(smnngp_mg* g, void* stream, const double* X, const double* y, int64_t D, int n_hidden, int act,
                      int arch, const double* hp_dev, int kind, int shift, double* out_dev, int* info_dev) {
  if (!g || !X || !y || !hp_dev || !out_dev || !info_dev || D <= 0 || n_hidden < 0 || n_hidden > 64 ||
      (act != ACT_RELU && act != ACT_ERF) || (arch != ARCH_MLP && arch != ARCH_RESNET) ||
      (kind != KIND_GAUSS && kind != KIND_STUDENT_T) || shift < 0 || shift > 3)
    return mg_fail(SMNNGP_EINVAL, "smnngp_lml_mg_f64: invalid argument");
  if (!g->connected) return mg_fail(SMNNGP_EINVAL, "smnngp_lml_mg_f64: handle not connected to its peers");
  if (g->dead) return mg_fail(SMNNGP_ECUDA, "smnngp_lml_mg_f64: handle was poisoned by a peer time-out");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (scope.dev != g->device) return mg_fail(SMNNGP_EINVAL, "smnngp_lml_mg_f64: stream belongs to another device");
  const long long n = g->n, db = g->db;
  const int P = g->P;
  mark(g, s, -1, 7);                                                 // start (Gram stage follows)
  MG_CU(cudaMemsetAsync(info_dev, 0, sizeof(int), s));
  MG_CU(cudaMemsetAsync(g->sums, 0, 2 * sizeof(double), s));
  MG_RC(build_gram(g, s, X, y, D, n_hidden, act, arch, hp_dev, shift));
  const long long npanels = cdiv(n, db);
  // ---- right-looking factorisation with one panel of look-ahead ----
  MG_CU(cudaEventRecord(g->ev_fork, s));
  MG_CU(cudaStreamWaitEvent(g->side, g->ev_fork, 0));
  long long ls = 0, m = 0;
  MG_RC(panel_step(g, g->side, 0, info_dev, ls, m));
  MG_CU(cudaEventRecord(g->ev_panel, g->side));
  for (long long p = 0; p < npanels; p++) {
    const long long c0 = p * db, c1 = std::min((p + 1) * db, n), w = c1 - c0;
    MG_CU(cudaStreamWaitEvent(s, g->ev_panel, 0));                 // panel p is factored and gathered
    mark(g, s, (int)p, 4);                                           // main_start
    if (c1 >= n) break;
    const double* arows = g->ploc[p & 1];
    const double* pfull = g->panel_local((int)(p % 3));            // row 0 = global row c1
    const long long gb0 = g->first_block_from(p + 1, g->rank);
    const long long shiftc = gb0 * db - c1;
    const long long na = std::min(db, n - c1);                     // next panel's block column first
    if (m > 0)
      MG_RC(smnngp_stage_update_f64(s, arows, db, pfull, db, g->a + ls * g->ld + c1, g->ld, m, na, w, 1, db, P, shiftc, 0));
    mark(g, s, (int)p, 5);                                           // update_a
    MG_CU(cudaEventRecord(g->ev_a, s));
    MG_CU(cudaStreamWaitEvent(g->side, g->ev_a, 0));
    long long ls2 = 0, m2 = 0;
    MG_RC(panel_step(g, g->side, p + 1, info_dev, ls2, m2));        // look-ahead: prepare panel p + 1
    MG_CU(cudaEventRecord(g->ev_panel, g->side));
    if (m > 0 && c1 + na < n) {                                     // the rest of the trailing matrix
      // SMs the fused panel solve needs to keep pace with this update: solve flops m w^2 at ~0.2 TF/s per SM against
      // update flops 2 m ncols w at 33 TF/s -> 82.5 w / ncols, independent of m and P; x4.5 measured margin + 3
      int reserve = g->sm_reserve_override;
      if (reserve < 0) {
        reserve = P > 1 || g->emulate ? (int)(4.5 * 82.5 * (double)w / (double)(n - c1 - na)) + 3 : 0;
        if (P > 1 || g->emulate) reserve = std::min(32, std::max(2, reserve));
      }
      MG_RC(smnngp_stage_update_f64(s, arows, db, pfull + na * db, db, g->a + ls * g->ld + c1 + na, g->ld, m,
                                    n - c1 - na, w, 1, db, P, shiftc - na, reserve));
    }
    mark(g, s, (int)p, 6);                                           // update_b
    ls = ls2;
    m = m2;
  }
  // ---- ||L^-1 y||^2 on the owner of the carried row, then the cross-rank reduction through the peer slots ----
  const long long bn = n / db;
  if (g->rank == g->owner(bn)) MG_CU(launch_sumsq(s, g->zvec, n, g->sums + 1));
  {
    void* red_ptrs[MAX_PEERS];
    void* flag_ptrs[MAX_PEERS];
    fill_ptrs(g, g->off_reduce, red_ptrs);
    fill_ptrs(g, g->off_flags, flag_ptrs);
    const unsigned long long seq = g->seq_base + (unsigned long long)npanels + 1;
    PeerSignal sg{};
    sg.P = P; sg.seq = seq; sg.counter = nullptr;
    for (int r = 0; r < MAX_PEERS; r++) {
      const bool on = flag_ptrs[r] != nullptr && (!g->emulate || r == g->rank);
      sg.flag[r] = on ? static_cast<unsigned long long*>(flag_ptrs[r]) + FLAG_REDUCE + g->rank : nullptr;
      if (!on) red_ptrs[r] = nullptr;
    }
    reduce_scatter_kernel<<<1, 1, 0, s>>>(g->sums, info_dev, g->rank, (double*)red_ptrs[0], (double*)red_ptrs[1],
                                          (double*)red_ptrs[2], (double*)red_ptrs[3], (double*)red_ptrs[4],
                                          (double*)red_ptrs[5], (double*)red_ptrs[6], (double*)red_ptrs[7], sg);
    instr().launches++;
    MG_CU(cudaGetLastError());
    if (g->emulate) {
      MG_RC(smnngp_stage_wait_flags_f64(s, g->flags_local(), FLAG_REDUCE + g->rank, 1, seq, g->timeout_s, info_dev));
    } else {
      MG_RC(smnngp_stage_wait_flags_f64(s, g->flags_local(), FLAG_REDUCE, P, seq, g->timeout_s, info_dev));
    }
    const double* slots = reinterpret_cast<const double*>(g->region + g->off_reduce);
    if (g->emulate) reduce_gather_kernel<<<1, 1, 0, s>>>(slots + g->rank * REDUCE_DOUBLES, 1, g->scal, info_dev);
    else reduce_gather_kernel<<<1, 1, 0, s>>>(slots, P, g->scal, info_dev);
    instr().launches++;
    MG_CU(cudaGetLastError());
  }
  MG_CU(launch_lml_finalize(s, g->scal, hp_dev, kind, n, info_dev, out_dev));
  g->seq_base += (unsigned long long)npanels + 2;
  arm_watchdog(g, s, info_dev);
  return SMNNGP_OK;
}

}  // extern "C"
