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
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "context.cuh"
#include "kernels.cuh"
#include "peer_signal.cuh"

using namespace smnngp;

namespace {

// flag words: 0: W ready; 8..15: panel ready (per source rank); 16..23: reduce slot ready; 24..31: right-hand-side
// rows (L^-1 Y)^T ready; 32..39: predictive results ready
constexpr int FLAG_WORDS = 64;
constexpr int FLAG_W = 0, FLAG_PANEL = 8, FLAG_REDUCE = 16, FLAG_Z = 24, FLAG_RES = 32, FLAG_GRAD = 40;
constexpr int REDUCE_DOUBLES = 4;  // per source rank: sum log L_ii, ||z||^2, info (as double), pad

inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

struct TimelineMark { int panel, label; cudaEvent_t ev; };

}  // namespace

struct smnngp_mg {
  int rank = 0, P = 1, device = 0;
  long long n = 0, extra = 1, db = 512, ld = 0, mtotal = 0, nblocks = 0, mloc = 0;
  // predictive handle: rows n .. n+c-1 = Y^T (right-hand sides), rows n+c .. n+c+t-1 = test-train cross-Gram;
  // LML handle: t = 0, c = 1 (the single carried row y^T)
  long long t = 0, c = 1;
  // gradient handle (smnngp_mg_create_grad): t = n "test rows" that start as the IDENTITY and come out as U = L^-T
  bool grad = false;
  double *tab3 = nullptr, *alpha = nullptr, *strip = nullptr, *partial = nullptr, *psum = nullptr;
  size_t tab3_doubles = 0;
  long long strip_r0 = 0, strip_r1 = 0, strip_ld = 0, slots = 0;
  size_t off_u = 0, off_gslots = 0;       // peer region: full U [n, ld] (all-gathered), per-rank gradient partial sums
  long long extra_lo = 0, n_carried = 0;  // local storage keeps global order: the carried rows are its tail
  bool emulate = false, connected = false;
  std::atomic<bool> dead{false};
  // local storage
  double* a = nullptr;                    // [mloc, ld] block-row-cyclic rows
  double *tab = nullptr, *q = nullptr, *scal = nullptr, *linv = nullptr, *carried = nullptr, *sums = nullptr;
  double *tab_t = nullptr, *q_t = nullptr, *res_tmp = nullptr;      // predictive: test-point tables, local results
  size_t tab_t_doubles = 0;
  double* ploc[2] = {nullptr, nullptr};   // solved panel rows in local order (A operand of the update)
  unsigned int* counters = nullptr;
  int* info_tmp = nullptr;                // scratch status word (second factorisation of test_nll)
  size_t tab_doubles = 0;
  // peer-visible region: [W db x db | flags | 3 panel slots n x db | reduce slots P x 4 | Z c x n | results t x (c+1)]
  char* region = nullptr;
  size_t off_w = 0, off_flags = 0, off_panel = 0, slot_bytes = 0, off_reduce = 0, off_z = 0, off_res = 0,
         region_bytes = 0;
  unsigned char handle[64] = {};
  char* bases[MAX_PEERS] = {};
  bool opened[MAX_PEERS] = {};
  unsigned long long seq_base = 0;
  cudaStream_t side = nullptr, poison_stream = nullptr;
  cudaEvent_t ev_panel = nullptr, ev_a = nullptr, ev_a0 = nullptr, ev_fork = nullptr, ev_done = nullptr;
  // tuning
  int sm_reserve_override = -1;
  double reserve_margin = 4.5;            // see factor_all: SMs left to the look-ahead chain = margin * 82.5 w / ncols + 3
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

  // ---- layout: which rank owns which global block (rows [b DB, (b+1) DB)); local storage keeps a rank's blocks in
  // global order.  Three block -> rank maps, all of the form  block(LB) = LB * P + (LB odd ? odd_off : even_off)  for the
  // LB-th local block of a rank (that is the only property the kernels rely on: exchange.cu scatter epilogue, update
  // mask), tabulated once at creation:
  //   cyclic    : b mod P                                   (distributed.BlockRowCyclic, the Python drivers)
  //   snake     : odd cycles of P blocks run backwards
  //   snake_end : the same, but the cycles are aligned to the END of the matrix
  //   auto      : (default for P > 1) plain cyclic or one of the 2P phases of the snake, whichever gives the busiest rank
  //               the least modelled update work (rows x position^2 per block; the last block is usually short).
  // Why: in the lower-triangular update a block row's work grows with the square of its global index, so with the
  // plain order the rank at the end of each cycle always holds the widest block: 9 % more update flops than the mean
  // over a factorisation of 118 blocks on 8 ranks (max / mean 1.091), and every rank waits for it at every panel.
  // The snake pairs a wide block with a narrow one; aligning the cycles to the end puts the one incomplete cycle where
  // the blocks are cheap (max / mean 1.054 start-aligned, 1.0075 end-aligned at 118 blocks).
  int layout = 0;                                   // 0 cyclic, 1 snake (start-aligned), 2 snake_end, 3 auto
  int snake_shift = -1;                             // phase of the snake actually used (-1: plain cyclic)
  std::vector<int> owner_tab;                       // [nblocks]
  std::vector<long long> lb_tab;                    // [nblocks] local block index on the owner
  std::vector<std::vector<long long>> blocks;       // [P][local blocks] global block ids, ascending
  long long even_off = 0, odd_off = 0;              // this rank's block(LB) formula (see above)

  int snake_owner(long long b, int shift) const {
    const long long x = b + shift;
    const int j = (int)(x % P);
    return ((x / P) & 1) ? P - 1 - j : j;
  }
  // model of the update work of a block of rows, in the same units for both kinds of rows:
  //   square rows (and y^T / right-hand-side / test rows) at global position pos: updated by pos / db panels, each over
  //   (pos - c1) columns -> rows x pos^2 (capped at n^2 for carried rows, which span all n columns);
  //   identity rows of the gradient handle, identity index i: inactive until the panel that holds column i, then
  //   updated over the remaining (n - c1) columns of every later panel -> rows x (n - i)^2.
  // Returns the largest per-rank load of an assignment.
  double block_weight(long long b) const {
    const double pos = ((double)b + 0.5) * (double)db;
    double wgt;
    if (grad && pos > (double)(n + c)) {
      const double rest = std::max(0.0, (double)n - (pos - (double)(n + c)));
      wgt = rest * rest;
    } else {
      wgt = std::min(pos, (double)n) * std::min(pos, (double)n);
    }
    return (double)block_rows(b) * wgt;
  }
  double max_load(int shift) const {
    std::vector<double> w((size_t)P, 0.0);
    for (long long b = 0; b < nblocks; b++) w[shift < 0 ? (int)(b % P) : snake_owner(b, shift)] += block_weight(b);
    return *std::max_element(w.begin(), w.end());
  }
  bool build_layout() {
    // candidates: plain cyclic (shift -1) and the 2P phases of the snake; "auto" keeps the one whose busiest rank has
    // the least modelled work (the same deterministic choice on every rank)
    snake_shift = -1;
    if (layout == 1) snake_shift = 0;
    if (layout == 2) snake_shift = (int)((2 * P - nblocks % (2 * P)) % (2 * P));
    // two ranks: plain cyclic measured faster than any snake (1174 vs 1189 ms at C3 on one box, 1185 ms for the phase
    // the model prefers) - the model's spread is < 1 % there and the snake's two-in-a-row ownership costs more
    if (layout == 3 && P > 2) {
      double best = max_load(-1);
      for (int sft = 0; sft < 2 * P; sft++) {
        const double m = max_load(sft);
        if (m < best * (1.0 - 1e-12)) { best = m; snake_shift = sft; }
      }
    }
    owner_tab.assign((size_t)nblocks, 0);
    lb_tab.assign((size_t)nblocks, 0);
    blocks.assign((size_t)P, {});
    for (long long b = 0; b < nblocks; b++) owner_tab[b] = snake_shift < 0 ? (int)(b % P) : snake_owner(b, snake_shift);
    for (long long b = 0; b < nblocks; b++) {
      lb_tab[b] = (long long)blocks[owner_tab[b]].size();
      blocks[owner_tab[b]].push_back(b);
    }
    const auto& mine = blocks[rank];
    even_off = mine.empty() ? rank : mine[0];
    odd_off = mine.size() > 1 ? mine[1] - P : even_off;
    for (size_t LB = 0; LB < mine.size(); LB++)
      if (mine[LB] != (long long)LB * P + ((LB & 1) ? odd_off : even_off)) return false;
    return true;
  }
  long long blk(long long LB, int r) const {        // LB-th local block of rank r; past the end: a block id >= nblocks
    return LB < (long long)blocks[r].size() ? blocks[r][LB] : nblocks + LB;
  }
  int owner(long long b) const { return owner_tab[b]; }
  long long block_rows(long long b) const { return std::min((b + 1) * db, mtotal) - b * db; }
  long long local_rows(int r) const {
    long long s = 0;
    for (long long b : blocks[r]) s += block_rows(b);
    return s;
  }
  long long local_offset(long long b) const {          // local row offset of global block b on its owner
    const auto& v = blocks[owner_tab[b]];
    long long s = 0;
    for (long long LB = 0; LB < lb_tab[b]; LB++) s += block_rows(v[LB]);
    return s;
  }
  long long first_block_from(long long gb, int r) const {   // smallest block >= gb owned by r (>= nblocks: none)
    for (long long b : blocks[r])
      if (b >= gb) return b;
    return nblocks;
  }
  void rows_from_block(long long gb, int r, long long& off, long long& cnt) const {
    off = 0;
    for (long long b : blocks[r]) {
      if (b >= gb) break;
      off += block_rows(b);
    }
    cnt = local_rows(r) - off;
  }
  // extra column shift of the rows of ODD local blocks (counted from the update's first row, which belongs to block
  // gb0) in the update mask: consecutive local blocks are alternately closer / further apart than P (GemmParams)
  long long cyc_alt(long long gb0) const {
    if (gb0 >= nblocks) return 0;
    const auto& v = blocks[owner_tab[gb0]];
    const long long LB0 = lb_tab[gb0];
    return LB0 + 1 < (long long)v.size() ? (v[LB0 + 1] - gb0 - P) * db : 0;
  }
  // gradient handle: identity row i is still e_i (zero in every column < i) until the panel that contains column i,
  // so while panel [c0, c1) is processed only local rows with global index < n + c + c1 take part
  long long active_prefix(long long c1) const {
    const long long limit = grad ? n + c + c1 : mtotal;
    long long cnt = 0;
    for (long long b : blocks[rank]) {
      const long long g0 = b * db, g1 = g0 + block_rows(b);
      if (g0 >= limit) break;
      cnt += std::min(g1, limit) - g0;
    }
    return cnt;
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
// rows_ready (optional): event after which ALL of this rank's rows of block column p are up to date; the owner's
// diagonal block only needs its own rows, which the caller updates first (the stream already waited for them).
int panel_step(smnngp_mg* g, cudaStream_t s, long long p, int* info_dev, long long& ls, long long& m,
               cudaEvent_t rows_ready = nullptr) {
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
  if (rows_ready != nullptr) MG_CU(cudaStreamWaitEvent(s, rows_ready, 0));
  if (!g->emulate || g->rank == own)
    MG_RC(smnngp_stage_wait_flags_f64(s, g->flags_local(), FLAG_W, 1, seq, g->timeout_s, info_dev));
  mark(g, s, (int)p, 1);                                                          // bcast
  if (g->rank == own) {
    ls = g->local_offset(p) + w;
    m = g->mloc - ls;
  } else {
    g->rows_from_block(p + 1, g->rank, ls, m);
  }
  if (g->grad) m = std::max<long long>(0, std::min(m, g->active_prefix(c1) - ls));
  double* ploc = g->ploc[p & 1];
  MG_RC(smnngp_stage_trsm_scatter2_f64(s, g->a + ls * g->ld + c0, g->ld, m, w, g->w_local(), db, ploc, db, panel_ptrs, P,
                                       g->rank, db, ls, c1, n, db, flag_ptrs, FLAG_PANEL + g->rank, seq,
                                       g->counters + 4, g->even_off, g->odd_off));
  mark(g, s, (int)p, 2);                                                          // trsm
  // carried rows (global index >= n: y^T / Y^T rows, test-train cross-Gram rows) are the tail of the local storage:
  // their solved panel columns only exist in ploc - keep them
  {
    const long long k = g->extra_lo - ls;                      // first carried row inside this rank's panel rows
    if (m > k && k >= 0)
      MG_CU(cudaMemcpy2DAsync(g->carried + c0, (size_t)g->ld * 8, ploc + k * db, (size_t)db * 8, (size_t)w * 8,
                              (size_t)(m - k), cudaMemcpyDeviceToDevice, s));
  }
  if (c1 >= n) return SMNNGP_OK;
  if (g->emulate) MG_RC(smnngp_stage_wait_flags_f64(s, g->flags_local(), FLAG_PANEL + g->rank, 1, seq, g->timeout_s, info_dev));
  else MG_RC(smnngp_stage_wait_flags_f64(s, g->flags_local(), FLAG_PANEL, P, seq, g->timeout_s, info_dev));
  mark(g, s, (int)p, 3);                                                          // gather
  return SMNNGP_OK;
}

int build_gram(smnngp_mg* g, cudaStream_t s, const double* X, const double* y, const double* xt, long long D, int nh,
               int act, int arch, const double* hp, int shift) {
  const long long n = g->n, db = g->db;
  const int n_act = std::max(n_act_applications(nh, arch), 1);
  if ((size_t)n_act * n > g->tab_doubles) {
    if (g->tab) cudaFree(g->tab);
    g->tab = nullptr;
    MG_CU(cudaMalloc(&g->tab, (size_t)n_act * n * 8));
    g->tab_doubles = (size_t)n_act * n;
  }
  MG_RC(smnngp_stage_qtable_f64(s, X, n, D, nh, act, arch, hp, g->tab, n, g->q, g->scal));
  if (xt != nullptr && g->t > 0) {
    if ((size_t)n_act * g->t > g->tab_t_doubles) {
      if (g->tab_t) cudaFree(g->tab_t);
      g->tab_t = nullptr;
      MG_CU(cudaMalloc(&g->tab_t, (size_t)n_act * g->t * 8));
      g->tab_t_doubles = (size_t)n_act * g->t;
    }
    MG_RC(smnngp_stage_qtable_f64(s, xt, g->t, D, nh, act, arch, hp, g->tab_t, g->t, g->q_t, nullptr));
  }
  for (long long LB = 0; g->blk(LB, g->rank) < g->nblocks; LB++) {
    const long long b = g->blk(LB, g->rank);
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
    const long long g1 = g0 + g->block_rows(b);
    if (g1 > n) {                                                          // this block holds carried rows
      // right-hand sides: global rows n .. n+c-1 = the columns of Y [N, C] (LML: the single row y^T)
      for (long long j = std::max(g0, n); j < std::min(g1, n + g->c); j++) {
        double* dst = g->a + (lo + (j - g0)) * g->ld;
        if (g->c == 1) MG_CU(cudaMemcpyAsync(dst, y, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
        else MG_CU(cudaMemcpy2DAsync(dst, 8, y + (j - n), (size_t)g->c * 8, 8, (size_t)n, cudaMemcpyDeviceToDevice, s));
      }
      // test points held by this block: rows of the test-train cross-Gram
      const long long t0 = std::max(g0, n + g->c) - (n + g->c), t1 = g1 - (n + g->c);
      if (xt != nullptr && t1 > t0)
        MG_RC(smnngp_stage_gram_f64(s, xt + t0 * D, t1 - t0, X, n, D, nh, act, arch, hp, g->tab_t + t0, g->t, g->tab, n,
                                    g->scal, SHIFT_NONE, 0, g->a + (lo + (n + g->c + t0 - g0)) * g->ld, g->ld));
      if (g->grad && t1 > t0) {                                            // identity rows t0 .. t1-1
        double* rows = g->a + (lo + (n + g->c + t0 - g0)) * g->ld;
        MG_CU(cudaMemsetAsync(rows, 0, (size_t)(t1 - t0) * g->ld * 8, s));
        MG_CU(launch_set_ones_diag(s, rows + t0, g->ld, t1 - t0));
      }
    }
  }
  return SMNNGP_OK;
}

}  // namespace

extern "C" {

const char* smnngp_mg_last_error(void) { return g_mg_err; }

static int mg_create_impl(smnngp_mg** out, int rank, int world, int64_t n, int64_t block, int64_t t, int64_t c,
                          bool grad = false) {
  if (!out || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || n <= 0 || block <= 0 || block % PB != 0 ||
      block > LINV_BLOCKS * PB || t < 0 || c < 1 || n + t + c > INT32_MAX)
    return mg_fail(SMNNGP_EINVAL, "smnngp_mg_create: invalid argument");
  smnngp_mg* g = new smnngp_mg();
  g->rank = rank; g->P = world; g->n = n; g->db = block; g->t = t; g->c = c; g->grad = grad;
  g->extra = t + c;
  cudaGetDevice(&g->device);
  g->mtotal = n + g->extra;
  g->nblocks = cdiv(g->mtotal, g->db);
  {
    const char* lay = getenv("SMNNGP_MG_LAYOUT");     // cyclic | snake | snake_end | auto (default)
    g->layout = world == 1 ? 0 : 3;
    if (lay != nullptr && strcmp(lay, "cyclic") == 0) g->layout = 0;
    if (lay != nullptr && strcmp(lay, "snake") == 0 && world > 1) g->layout = 1;
    if (lay != nullptr && strcmp(lay, "snake_end") == 0 && world > 1) g->layout = 2;
    if (!g->build_layout()) {                         // (cannot happen for the three maps above)
      g->layout = 0;
      g->build_layout();
    }
  }
  g->ld = cdiv(n, 16) * 16;
  g->mloc = g->local_rows(rank);
  for (long long LB = 0; g->blk(LB, rank) < g->nblocks; LB++) {      // local rows with global index < n
    const long long b = g->blk(LB, rank);
    const long long g0 = b * g->db, g1 = g0 + g->block_rows(b);
    g->extra_lo += std::max<long long>(0, std::min<long long>(g1, n) - g0);
  }
  g->n_carried = g->mloc - g->extra_lo;
  g->off_w = 0;
  g->off_flags = (size_t)g->db * g->db * 8;
  g->off_panel = g->off_flags + FLAG_WORDS * 8;
  g->slot_bytes = (size_t)n * g->db * 8;
  g->off_reduce = g->off_panel + 3 * g->slot_bytes;
  g->off_z = g->off_reduce + (size_t)MAX_PEERS * REDUCE_DOUBLES * 8;
  g->off_res = g->off_z + (t > 0 ? (size_t)c * n * 8 : 0);
  g->off_gslots = g->off_res + ((t > 0 && !grad) ? (size_t)t * (c + 1) * 8 : 0);
  g->off_u = (g->off_gslots + (size_t)MAX_PEERS * 4 * 8 + 255) & ~(size_t)255;
  g->region_bytes = g->off_u + (grad ? (size_t)n * g->ld * 8 : 0) + 256;
  void* reg = nullptr;
  if (smnngp_peer_alloc(g->region_bytes, &reg, g->handle) != SMNNGP_OK) {
    delete g;
    return mg_fail(SMNNGP_ECUDA, "smnngp_mg_create: peer region allocation / IPC export failed");
  }
  g->region = static_cast<char*>(reg);
  const long long mrows = std::max<long long>(g->mloc, 1), crow = std::max<long long>(g->n_carried, 1);
  bool ok = cudaMemset(g->region + g->off_flags, 0, FLAG_WORDS * 8) == cudaSuccess &&
            cudaMemset(g->region + g->off_reduce, 0, MAX_PEERS * REDUCE_DOUBLES * 8) == cudaSuccess &&
            cudaMalloc(&g->a, (size_t)mrows * g->ld * 8) == cudaSuccess &&
            cudaMalloc(&g->q, (size_t)n * 8) == cudaSuccess && cudaMalloc(&g->scal, SC_COUNT * 8) == cudaSuccess &&
            cudaMalloc(&g->linv, (size_t)LINV_BLOCKS * PB * PB * 8) == cudaSuccess &&
            cudaMalloc(&g->carried, (size_t)crow * g->ld * 8) == cudaSuccess &&
            cudaMalloc(&g->sums, 2 * 8) == cudaSuccess &&
            cudaMalloc(&g->ploc[0], (size_t)mrows * g->db * 8) == cudaSuccess &&
            cudaMalloc(&g->ploc[1], (size_t)mrows * g->db * 8) == cudaSuccess &&
            cudaMalloc(&g->counters, 8 * sizeof(unsigned int)) == cudaSuccess &&
            cudaMalloc(&g->info_tmp, sizeof(int)) == cudaSuccess &&
            cudaMemset(g->counters, 0, 8 * sizeof(unsigned int)) == cudaSuccess &&
            cudaMemset(g->carried, 0, (size_t)crow * g->ld * 8) == cudaSuccess;
  if (ok && t > 0 && !grad)
    ok = cudaMalloc(&g->q_t, (size_t)t * 8) == cudaSuccess &&
         cudaMalloc(&g->res_tmp, (size_t)t * (c + 1) * 8) == cudaSuccess;
  if (ok && grad) {
    // every rank contracts one row strip [r0, r1) x [0, r1) of the lower triangle of A^-1 with dK/dtheta; strips of
    // equal area: r ~ N sqrt(k / P), aligned to the 128-row tile
    auto cut = [&](int k) {
      long long r = (long long)(std::sqrt((double)k / (double)world) * (double)n / 128.0 + 0.5) * 128;
      return k >= world ? (long long)n : std::min<long long>(r, n);
    };
    g->strip_r0 = cut(rank);
    g->strip_r1 = cut(rank + 1);
    g->strip_ld = cdiv(std::max<long long>(g->strip_r1, 1), 16) * 16;
    g->slots = grad_partial_slots(n);
    const long long srows = std::max<long long>(g->strip_r1 - g->strip_r0, 1);
    ok = cudaMalloc(&g->alpha, (size_t)n * 8) == cudaSuccess &&
         cudaMalloc(&g->strip, (size_t)srows * g->strip_ld * 8) == cudaSuccess &&
         cudaMalloc(&g->partial, (size_t)g->slots * 4 * 8) == cudaSuccess &&
         cudaMalloc(&g->psum, 4 * 8) == cudaSuccess &&
         cudaMemset(g->region + g->off_u, 0, (size_t)n * g->ld * 8) == cudaSuccess;
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  ok = ok && cudaStreamCreateWithPriority(&g->side, cudaStreamNonBlocking, hi) == cudaSuccess &&
       cudaStreamCreateWithFlags(&g->poison_stream, cudaStreamNonBlocking) == cudaSuccess &&
       cudaEventCreateWithFlags(&g->ev_panel, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&g->ev_a, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&g->ev_a0, cudaEventDisableTiming) == cudaSuccess &&
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

int smnngp_mg_create(smnngp_mg** out, int rank, int world, int64_t n, int64_t block) {
  return mg_create_impl(out, rank, world, n, block, 0, 1);
}
// handle for smnngp_predict_mg_f64 / smnngp_test_nll_mg_f64: T test points and C right-hand sides ride along as extra
// global rows of the same layout
// handle for smnngp_lml_grad_mg_f64: the N rows of the identity ride through the factorisation (they come out as
// U = L^-T, all-gathered afterwards: every rank needs 8 N^2 bytes for it)
int smnngp_mg_create_grad(smnngp_mg** out, int rank, int world, int64_t n, int64_t block) {
  return mg_create_impl(out, rank, world, n, block, n, 1, true);
}
int smnngp_mg_create_predict(smnngp_mg** out, int rank, int world, int64_t n, int64_t t, int64_t c, int64_t block) {
  if (t <= 0) return mg_fail(SMNNGP_EINVAL, "smnngp_mg_create_predict: t must be positive");
  return mg_create_impl(out, rank, world, n, block, t, c);
}

// Pure host function (no CUDA call): the block -> rank map a handle with these parameters would use.
// layout: 0 cyclic, 1 snake, 2 snake_end, 3 auto.  owner_out [nblocks]; returns nblocks (< 0: invalid argument).
// load_out [world] (optional): modelled update work per rank (the model the "auto" map minimises), for tests / diagnostics.
int64_t smnngp_mg_layout(int world, int64_t n, int64_t extra_rows, int64_t block, int layout, int* owner_out,
                         double* load_out) {
  if (world < 1 || world > MAX_PEERS || n <= 0 || extra_rows < 0 || block <= 0 || layout < 0 || layout > 3) return -1;
  smnngp_mg g;
  g.rank = 0; g.P = world; g.n = n; g.db = block; g.extra = extra_rows;
  g.mtotal = n + extra_rows;
  g.nblocks = cdiv(g.mtotal, g.db);
  g.layout = world == 1 ? 0 : layout;
  for (int r = 0; r < world; r++) {          // the formula check is per rank: run it for every rank
    g.rank = r;
    if (!g.build_layout()) return -2;
  }
  if (owner_out)
    for (long long b = 0; b < g.nblocks; b++) owner_out[b] = g.owner_tab[b];
  if (load_out) {
    for (int r = 0; r < world; r++) load_out[r] = 0.0;
    for (long long b = 0; b < g.nblocks; b++) load_out[g.owner_tab[b]] += g.block_weight(b);
  }
  return g.nblocks;
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
void smnngp_mg_set_reserve_margin(smnngp_mg* g, double margin) {
  if (g && margin > 0.0) g->reserve_margin = margin;
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
  for (double* p : {g->a, g->tab, g->q, g->scal, g->linv, g->carried, g->sums, g->ploc[0], g->ploc[1], g->tab_t, g->q_t,
                    g->res_tmp, g->tab3, g->alpha, g->strip, g->partial, g->psum})
    if (p) cudaFree(p);
  if (g->counters) cudaFree(g->counters);
  if (g->info_tmp) cudaFree(g->info_tmp);
  for (cudaEvent_t e : {g->ev_panel, g->ev_a, g->ev_a0, g->ev_fork, g->ev_done})
    if (e) cudaEventDestroy(e);
  if (g->side) cudaStreamDestroy(g->side);
  if (g->poison_stream) cudaStreamDestroy(g->poison_stream);
  cudaSetDevice(cur);
  delete g;
  return SMNNGP_OK;
}

}  // extern "C" (helpers below are internal)

namespace {

// Gram rows of this rank + right-looking factorisation with one panel of look-ahead (see file comment)
int factor_all(smnngp_mg* g, cudaStream_t s, const double* X, const double* y, const double* xt, long long D, int nh,
               int act, int arch, const double* hp_dev, int shift, int* info_dev) {
  const long long n = g->n, db = g->db;
  const int P = g->P;
  mark(g, s, -1, 7);                                                 // start (Gram stage follows)
  MG_CU(cudaMemsetAsync(info_dev, 0, sizeof(int), s));
  MG_CU(cudaMemsetAsync(g->sums, 0, 2 * sizeof(double), s));
  MG_RC(build_gram(g, s, X, y, xt, D, nh, act, arch, hp_dev, shift));
  if (g->grad)     // rows that never become active keep their zeros (U is upper triangular)
    MG_CU(cudaMemsetAsync(g->carried, 0, (size_t)std::max<long long>(g->n_carried, 1) * g->ld * 8, s));
  const long long npanels = cdiv(n, db);
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
    const long long alt = g->cyc_alt(gb0);
    // The rank that factors the NEXT diagonal block updates that block's rows first and lets its chain start on them
    // (ev_a0); the rest of its rows of the block column follow (ev_a), and its panel solve waits for those.
    const bool own_next = g->owner(p + 1) == g->rank && gb0 == p + 1 && m > na;
    cudaEvent_t rows_ready = nullptr;
    if (own_next) {
      MG_RC(smnngp_stage_update2_f64(s, arows, db, pfull, db, g->a + ls * g->ld + c1, g->ld, na, na, w, 1, db, P, shiftc,
                                     alt, 0));
      MG_CU(cudaEventRecord(g->ev_a0, s));
      const long long gb1 = g->first_block_from(p + 2, g->rank);
      MG_RC(smnngp_stage_update2_f64(s, arows + na * db, db, pfull, db, g->a + (ls + na) * g->ld + c1, g->ld, m - na, na,
                                     w, 1, db, P, gb1 * db - c1, g->cyc_alt(gb1), 0));
      rows_ready = g->ev_a;
    } else if (m > 0) {
      MG_RC(smnngp_stage_update2_f64(s, arows, db, pfull, db, g->a + ls * g->ld + c1, g->ld, m, na, w, 1, db, P, shiftc,
                                     alt, 0));
    }
    mark(g, s, (int)p, 5);                                           // update_a
    MG_CU(cudaEventRecord(g->ev_a, s));
    MG_CU(cudaStreamWaitEvent(g->side, own_next ? g->ev_a0 : g->ev_a, 0));
    long long ls2 = 0, m2 = 0;
    MG_RC(panel_step(g, g->side, p + 1, info_dev, ls2, m2, rows_ready));   // look-ahead: prepare panel p + 1
    MG_CU(cudaEventRecord(g->ev_panel, g->side));
    if (m > 0 && c1 + na < n) {                                     // the rest of the trailing matrix
      // SMs the fused panel solve needs to keep pace with this update: solve flops m w^2 at ~0.2 TF/s per SM against
      // update flops 2 m ncols w at 33 TF/s -> 82.5 w / ncols, independent of m and P; x4.5 measured margin + 3
      int reserve = g->sm_reserve_override;
      if (reserve < 0) {
        reserve = P > 1 || g->emulate ? (int)(g->reserve_margin * 82.5 * (double)w / (double)(n - c1 - na)) + 3 : 0;
        if (P > 1 || g->emulate) reserve = std::min(32, std::max(2, reserve));
      }
      MG_RC(smnngp_stage_update2_f64(s, arows, db, pfull + na * db, db, g->a + ls * g->ld + c1 + na, g->ld, m,
                                     n - c1 - na, w, 1, db, P, shiftc - na, alt, reserve));
    }
    mark(g, s, (int)p, 6);                                           // update_b
    ls = ls2;
    m = m2;
  }
  return SMNNGP_OK;
}

void signal_for(smnngp_mg* g, int word, unsigned long long seq, PeerSignal& sg, void** data_ptrs, size_t data_off) {
  void* flag_ptrs[MAX_PEERS];
  fill_ptrs(g, g->off_flags, flag_ptrs);
  if (data_ptrs) fill_ptrs(g, data_off, data_ptrs);
  sg = PeerSignal{};
  sg.P = g->P; sg.seq = seq; sg.counter = nullptr;
  for (int r = 0; r < MAX_PEERS; r++) {
    const bool on = flag_ptrs[r] != nullptr && (!g->emulate || r == g->rank);
    sg.flag[r] = on ? static_cast<unsigned long long*>(flag_ptrs[r]) + word + g->rank : nullptr;
    if (!on && data_ptrs) data_ptrs[r] = nullptr;
  }
}

int wait_all(smnngp_mg* g, cudaStream_t s, int word, unsigned long long seq, int* info_dev) {
  if (g->emulate) return smnngp_stage_wait_flags_f64(s, g->flags_local(), word + g->rank, 1, seq, g->timeout_s, info_dev);
  return smnngp_stage_wait_flags_f64(s, g->flags_local(), word, g->P, seq, g->timeout_s, info_dev);
}

// {sum log L_ii, ||z||^2, info} of every rank -> scal[SC_LOGDET], scal[SC_QUAD], info (max), same on every rank
int reduce_all(smnngp_mg* g, cudaStream_t s, unsigned long long seq, int* info_dev) {
  void* red_ptrs[MAX_PEERS];
  PeerSignal sg;
  signal_for(g, FLAG_REDUCE, seq, sg, red_ptrs, g->off_reduce);
  reduce_scatter_kernel<<<1, 1, 0, s>>>(g->sums, info_dev, g->rank, (double*)red_ptrs[0], (double*)red_ptrs[1],
                                        (double*)red_ptrs[2], (double*)red_ptrs[3], (double*)red_ptrs[4],
                                        (double*)red_ptrs[5], (double*)red_ptrs[6], (double*)red_ptrs[7], sg);
  instr().launches++;
  MG_CU(cudaGetLastError());
  MG_RC(wait_all(g, s, FLAG_REDUCE, seq, info_dev));
  const double* slots = reinterpret_cast<const double*>(g->region + g->off_reduce);
  if (g->emulate) reduce_gather_kernel<<<1, 1, 0, s>>>(slots + g->rank * REDUCE_DOUBLES, 1, g->scal, info_dev);
  else reduce_gather_kernel<<<1, 1, 0, s>>>(slots, g->P, g->scal, info_dev);
  instr().launches++;
  MG_CU(cudaGetLastError());
  return SMNNGP_OK;
}

// rows [rows x width] (pitch lds) -> the same offset of every rank's buffer (pitch ldd); optional flag afterwards
struct PeerCopy {
  const double* src;
  long long lds, ldd, rows, width;
  double* dst[MAX_PEERS];
  int P;
};
__global__ void peer_copy_kernel(const PeerCopy pc) {
  const long long total = pc.rows * pc.width;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / pc.width, c = e - r * pc.width;
    const double v = pc.src[r * pc.lds + c];
#pragma unroll
    for (int q = 0; q < MAX_PEERS; q++)
      if (q < pc.P && pc.dst[q] != nullptr) pc.dst[q][r * pc.ldd + c] = v;
  }
}
__global__ void fold_info_kernel(const int* __restrict__ src, int* __restrict__ dst) {
  if (*src != 0) atomicMax(dst, *src);
}
__global__ void flag_kernel(PeerSignal sg) {
  __threadfence_system();
  for (int q = 0; q < sg.P; q++)
    if (sg.flag[q] != nullptr) st_release_sys(sg.flag[q], sg.seq);
}

int check_call(smnngp_mg* g, cudaStream_t s, const Enter& scope, const char* who) {
  if (!g->connected) return mg_fail(SMNNGP_EINVAL, "handle not connected to its peers");
  if (g->dead) return mg_fail(SMNNGP_ECUDA, "handle was poisoned by a peer time-out");
  if (scope.dev != g->device) return mg_fail(SMNNGP_EINVAL, "stream belongs to another device than the handle");
  (void)s; (void)who;
  return SMNNGP_OK;
}

bool valid_stack_mg(int n_hidden, int act, int arch) {
  return n_hidden >= 0 && n_hidden <= 64 && (act == ACT_RELU || act == ACT_ERF) && (arch == ARCH_MLP || arch == ARCH_RESNET);
}

}  // namespace

extern "C" {

// SPR.loss (spax/models.py:93-98) on the ranks that share the handle group: out_dev[4] = {log p(y), -log p(y) / N,
// sum log L_ii, ||L^-1 y||^2}, identical on every rank.  Every rank calls with the SAME X [N, D], y [N], hp_dev [6]
// (device pointers on ITS device).  shift: SMNNGP_SHIFT_* added to the Gram diagonal (SHIFT_EPS_ABS for SPR.loss).
int smnngp_lml_mg_f64(smnngp_mg* g, void* stream, const double* X, const double* y, int64_t D, int n_hidden, int act,
                      int arch, const double* hp_dev, int kind, int shift, double* out_dev, int* info_dev) {
  if (!g || !X || !y || !hp_dev || !out_dev || !info_dev || D <= 0 || !valid_stack_mg(n_hidden, act, arch) ||
      (kind != KIND_GAUSS && kind != KIND_STUDENT_T) || shift < 0 || shift > 3 || g->t != 0 || g->c != 1)
    return mg_fail(SMNNGP_EINVAL, "smnngp_lml_mg_f64: invalid argument (needs a handle from smnngp_mg_create)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  MG_RC(check_call(g, s, scope, "smnngp_lml_mg_f64"));
  const long long n = g->n, npanels = cdiv(n, g->db);
  MG_RC(factor_all(g, s, X, y, nullptr, D, n_hidden, act, arch, hp_dev, shift, info_dev));
  // ||L^-1 y||^2 on the owner of the carried row, then the cross-rank reduction through the peer slots
  if (g->n_carried > 0) MG_CU(launch_sumsq(s, g->carried, n, g->sums + 1));
  MG_RC(reduce_all(g, s, g->seq_base + (unsigned long long)npanels + 1, info_dev));
  MG_CU(launch_lml_finalize(s, g->scal, hp_dev, kind, n, info_dev, out_dev));
  g->seq_base += (unsigned long long)npanels + 4;
  arm_watchdog(g, s, info_dev);
  return SMNNGP_OK;
}

// SPR.loss AND its gradient w.r.t. {w_std, b_std, last_w_std, eps, alpha, beta} - what objax.GradValues(model.loss, vars)
// computes in the reference's training step (experiments/regression/train.py:62-66) - on the ranks of the handle group.
// out_dev[4] as smnngp_lml_mg_f64, grad_dev[6] = d loss / d hp, identical on every rank.
//   1. factorisation with the N identity rows carried along (rows of U = L^-T, block-cyclic like everything else; a
//      row only takes part once its column has been reached: N^3/3 extra flop, not N^3);
//   2. z = L^-1 y is stored to every rank; every rank's U rows are copied to every rank (copy engines over NVLink,
//      right of the diagonal only); a = A^-1 y = U z on every rank;
//   3. every rank forms ONE row strip of A^-1 = U U^T (equal-area strips of the lower triangle, contraction from the
//      diagonal on) and contracts it with dK/dtheta in the dual Gram pass of grad.cu restricted to that strip;
//   4. the four partial sums per rank go through peer slots, are added in rank order and finalised (closed forms for
//      the inverse-gamma parameters).
int smnngp_lml_grad_mg_f64(smnngp_mg* g, void* stream, const double* X, const double* y, int64_t D, int n_hidden,
                           int act, int arch, const double* hp_dev, int kind, double* out_dev, double* grad_dev,
                           int* info_dev) {
  if (!g || !X || !y || !hp_dev || !out_dev || !grad_dev || !info_dev || D <= 0 || !valid_stack_mg(n_hidden, act, arch) ||
      (kind != KIND_GAUSS && kind != KIND_STUDENT_T) || !g->grad)
    return mg_fail(SMNNGP_EINVAL, "smnngp_lml_grad_mg_f64: invalid argument (needs a handle from smnngp_mg_create_grad)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  MG_RC(check_call(g, s, scope, "smnngp_lml_grad_mg_f64"));
  const long long n = g->n, db = g->db, ld = g->ld, npanels = cdiv(n, db);
  const int P = g->P;
  const unsigned long long seq0 = g->seq_base + (unsigned long long)npanels;
  const int n_act = std::max(n_act_applications(n_hidden, arch), 1);
  if ((size_t)3 * n_act * n > g->tab3_doubles) {
    if (g->tab3) cudaFree(g->tab3);
    g->tab3 = nullptr;
    MG_CU(cudaMalloc(&g->tab3, (size_t)3 * n_act * n * 8));
    g->tab3_doubles = (size_t)3 * n_act * n;
  }
  // ---- 1. value: factorisation with y^T and the identity rows carried ----
  MG_RC(factor_all(g, s, X, y, nullptr, D, n_hidden, act, arch, hp_dev, SMNNGP_SHIFT_EPS_ABS, info_dev));
  if (g->owner(n / db) == g->rank) MG_CU(launch_sumsq(s, g->carried, n, g->sums + 1));      // carried row 0 = z^T
  MG_RC(reduce_all(g, s, seq0 + 1, info_dev));
  MG_CU(launch_lml_finalize(s, g->scal, hp_dev, kind, n, info_dev, out_dev));
  MG_CU(launch_qtable_dual(s, X, D, (int)n, (int)D, n_hidden, act, arch, hp_dev, g->tab3, n));
  double* zloc = reinterpret_cast<double*>(g->region + g->off_z);
  double* Uloc = reinterpret_cast<double*>(g->region + g->off_u);
  // ---- 2. z and U to every rank ----
  {
    void* zp[MAX_PEERS];
    void* up[MAX_PEERS];
    PeerSignal sg;
    signal_for(g, FLAG_Z, seq0 + 2, sg, zp, g->off_z);
    fill_ptrs(g, g->off_u, up);
    long long row = 0;                                               // index into the carried rows (global order)
    for (long long LB = 0; g->blk(LB, g->rank) < g->nblocks; LB++) {
      const long long b = g->blk(LB, g->rank);
      const long long g0 = b * db, g1 = g0 + g->block_rows(b);
      if (g1 <= n) continue;
      const long long first = std::max(g0, n);
      if (g0 <= n && n < g1) {                                       // this block holds z^T
        PeerCopy pc{};
        pc.src = g->carried + row * ld; pc.lds = ld; pc.ldd = n; pc.rows = 1; pc.width = n; pc.P = P;
        for (int r = 0; r < MAX_PEERS; r++) pc.dst[r] = zp[r] ? static_cast<double*>(zp[r]) : nullptr;
        peer_copy_kernel<<<64, 256, 0, s>>>(pc);
        instr().launches++;
        MG_CU(cudaGetLastError());
      }
      const long long t0 = std::max(g0, n + 1) - (n + 1), t1 = g1 - (n + 1);   // identity rows of this block
      if (t1 > t0) {
        const double* src = g->carried + (row + (n + 1 + t0 - first)) * ld;
        const long long cfirst = t0 & ~1ll;                          // U is upper triangular: columns >= t0 only
        for (int r = 0; r < P; r++) {
          if (up[r] == nullptr || (g->emulate && r != g->rank)) continue;
          MG_CU(cudaMemcpy2DAsync(static_cast<double*>(up[r]) + t0 * ld + cfirst, (size_t)ld * 8, src + cfirst,
                                  (size_t)ld * 8, (size_t)(n - cfirst) * 8, (size_t)(t1 - t0), cudaMemcpyDeviceToDevice, s));
        }
      }
      row += g1 - first;
    }
    flag_kernel<<<1, 1, 0, s>>>(sg);
    instr().launches++;
    MG_CU(cudaGetLastError());
    MG_RC(wait_all(g, s, FLAG_Z, seq0 + 2, info_dev));
  }
  MG_CU(launch_upper_gemv(s, Uloc, ld, zloc, n, g->alpha));          // a = U z, on every rank
  // ---- 3. this rank's strip of A^-1 = U U^T and its contraction with dK/dtheta ----
  const long long r0 = g->strip_r0, r1 = g->strip_r1;
  if (r1 > r0) {
    GemmParams u{};
    u.A = Uloc + r0 * ld; u.lda = ld;
    u.B = Uloc;           u.ldb = ld;
    u.C = g->strip;       u.ldc = g->strip_ld;
    u.M = (int)(r1 - r0); u.N = (int)r1; u.K = (int)n; u.lower = 1;
    u.cyc_db = 1 << 30; u.cyc_p = 1; u.base_shift = (int)r0;       // local row r may touch columns <= r0 + r
    u.k_from_row = 1; u.k_row0 = (int)r0;
    MG_CU(launch_gemm_store_lower(s, u));
  }
  MG_CU(launch_grad_gram_strip(s, X, n, D, r0, r1 - r0, n_hidden, act, arch, hp_dev, g->tab3, n, g->strip, g->strip_ld,
                               g->alpha, g->scal + SC_QUAD, kind, g->partial, g->slots));
  MG_CU(launch_sum_slots(s, g->partial, g->slots, g->psum));
  // ---- 4. partial sums of every rank -> every rank, fixed-order total, closed forms ----
  {
    void* gp[MAX_PEERS];
    PeerSignal sg;
    signal_for(g, FLAG_GRAD, seq0 + 3, sg, gp, g->off_gslots);
    PeerCopy pc{};
    pc.src = g->psum; pc.lds = 4; pc.ldd = 4; pc.rows = 1; pc.width = 4; pc.P = P;
    for (int r = 0; r < MAX_PEERS; r++) pc.dst[r] = gp[r] ? static_cast<double*>(gp[r]) + 4 * g->rank : nullptr;
    peer_copy_kernel<<<1, 32, 0, s>>>(pc);
    flag_kernel<<<1, 1, 0, s>>>(sg);
    instr().launches += 2;
    MG_CU(cudaGetLastError());
    MG_RC(wait_all(g, s, FLAG_GRAD, seq0 + 3, info_dev));
    const double* slots = reinterpret_cast<const double*>(g->region + g->off_gslots);
    if (g->emulate) MG_CU(launch_grad_finalize(s, slots + 4 * g->rank, 1, hp_dev, g->scal + SC_QUAD, kind, n, info_dev, grad_dev));
    else MG_CU(launch_grad_finalize(s, slots, P, hp_dev, g->scal + SC_QUAD, kind, n, info_dev, grad_dev));
  }
  g->seq_base += (unsigned long long)npanels + 6;
  arm_watchdog(g, s, info_dev);
  return SMNNGP_OK;
}

// NNGPKernel.predict (spax/kernels.py:29-32; neural_tangents gradient_descent_mse_ensemble) on the ranks of the handle
// group: mean_out [T, C], var_out [T] (= diag of the posterior covariance), identical on every rank.  Y [N, C]
// row-major, Xt [T, D].  The C right-hand sides and the T test-train cross-Gram rows ride through the distributed
// factorisation as extra global rows (never exchanged); (L^-1 Y)^T is then stored to every rank, every rank finishes
// ITS test points and stores the results to every rank.  shift = SMNNGP_SHIFT_EPS_REL for the reference semantics.
int smnngp_predict_mg_f64(smnngp_mg* g, void* stream, const double* X, const double* Y, const double* Xt, int64_t D,
                          int n_hidden, int act, int arch, const double* hp_dev, int shift, double* mean_out,
                          double* var_out, int* info_dev) {
  if (!g || !X || !Y || !Xt || !hp_dev || !mean_out || !var_out || !info_dev || D <= 0 ||
      !valid_stack_mg(n_hidden, act, arch) || shift < 0 || shift > 3 || g->t <= 0)
    return mg_fail(SMNNGP_EINVAL, "smnngp_predict_mg_f64: invalid argument (needs a handle from smnngp_mg_create_predict)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  MG_RC(check_call(g, s, scope, "smnngp_predict_mg_f64"));
  const long long n = g->n, db = g->db, T = g->t, C = g->c, npanels = cdiv(n, db);
  const unsigned long long seq0 = g->seq_base + (unsigned long long)npanels;
  MG_RC(factor_all(g, s, X, Y, Xt, D, n_hidden, act, arch, hp_dev, shift, info_dev));
  MG_RC(reduce_all(g, s, seq0 + 1, info_dev));                      // info of every rank (a non-PD block anywhere)
  double* zloc = reinterpret_cast<double*>(g->region + g->off_z);
  double* rloc = reinterpret_cast<double*>(g->region + g->off_res);  // [T, C] means then [T] variances
  // (1) right-hand-side rows (L^-1 Y)^T this rank carried -> every rank's Z [C, n]
  {
    void* zp[MAX_PEERS];
    PeerSignal sg;
    signal_for(g, FLAG_Z, seq0 + 2, sg, zp, g->off_z);
    long long row = 0;                                               // index into the carried rows (global order)
    for (long long LB = 0; g->blk(LB, g->rank) < g->nblocks; LB++) {
      const long long b = g->blk(LB, g->rank);
      const long long g0 = b * db, g1 = g0 + g->block_rows(b);
      if (g1 <= n) continue;
      const long long j0 = std::max(g0, n), j1 = std::min(g1, n + C);   // right-hand sides in this block
      if (j1 > j0) {
        PeerCopy pc{};
        pc.src = g->carried + row * g->ld; pc.lds = g->ld; pc.ldd = n; pc.rows = j1 - j0; pc.width = n; pc.P = g->P;
        for (int r = 0; r < MAX_PEERS; r++) pc.dst[r] = zp[r] ? static_cast<double*>(zp[r]) + (j0 - n) * n : nullptr;
        peer_copy_kernel<<<148, 256, 0, s>>>(pc);
        instr().launches++;
        MG_CU(cudaGetLastError());
      }
      row += g1 - std::max(g0, n);
    }
    flag_kernel<<<1, 1, 0, s>>>(sg);
    instr().launches++;
    MG_CU(cudaGetLastError());
    MG_RC(wait_all(g, s, FLAG_Z, seq0 + 2, info_dev));
  }
  // (2) predictive moments of this rank's test points -> every rank's result buffer
  {
    void* rp[MAX_PEERS];
    PeerSignal sg;
    signal_for(g, FLAG_RES, seq0 + 3, sg, rp, g->off_res);
    long long row = 0;
    for (long long LB = 0; g->blk(LB, g->rank) < g->nblocks; LB++) {
      const long long b = g->blk(LB, g->rank);
      const long long g0 = b * db, g1 = g0 + g->block_rows(b);
      if (g1 <= n) continue;
      const long long first = std::max(g0, n);
      const long long t0 = std::max(g0, n + C) - (n + C), t1 = g1 - (n + C);
      if (t1 > t0) {
        const double* V = g->carried + (row + (n + C + t0 - first)) * g->ld;
        double* mean_l = g->res_tmp + t0 * C;
        double* var_l = g->res_tmp + T * C + t0;
        MG_CU(launch_predict_finalize(s, V, g->ld, zloc, n, g->q_t + t0, (int)(t1 - t0), (int)C, n, info_dev, mean_l,
                                      var_l));
        PeerCopy pm{};
        pm.src = mean_l; pm.lds = C; pm.ldd = C; pm.rows = t1 - t0; pm.width = C; pm.P = g->P;
        PeerCopy pv{};
        pv.src = var_l; pv.lds = t1 - t0; pv.ldd = t1 - t0; pv.rows = 1; pv.width = t1 - t0; pv.P = g->P;
        for (int r = 0; r < MAX_PEERS; r++) {
          pm.dst[r] = rp[r] ? static_cast<double*>(rp[r]) + t0 * C : nullptr;
          pv.dst[r] = rp[r] ? static_cast<double*>(rp[r]) + T * C + t0 : nullptr;
        }
        peer_copy_kernel<<<64, 256, 0, s>>>(pm);
        peer_copy_kernel<<<16, 256, 0, s>>>(pv);
        instr().launches += 2;
        MG_CU(cudaGetLastError());
      }
      row += g1 - first;
    }
    flag_kernel<<<1, 1, 0, s>>>(sg);
    instr().launches++;
    MG_CU(cudaGetLastError());
    MG_RC(wait_all(g, s, FLAG_RES, seq0 + 3, info_dev));
  }
  MG_CU(cudaMemcpyAsync(mean_out, rloc, (size_t)T * C * 8, cudaMemcpyDeviceToDevice, s));
  MG_CU(cudaMemcpyAsync(var_out, rloc + T * C, (size_t)T * 8, cudaMemcpyDeviceToDevice, s));
  g->seq_base += (unsigned long long)npanels + 4;
  arm_watchdog(g, s, info_dev);
  return SMNNGP_OK;
}

// SPR.test_nll (spax/models.py:100-120) on the handle group: g_pred from smnngp_mg_create_predict(..., t, c = 1, ...),
// g_lik from smnngp_mg_create (same n / block / group).  (1) predictive with the relative regulariser, (2) for the
// Student-t likelihood the scale d = 2a + y^T ((b/a) K + 1e-6 I)^-1 y (spax/likelihoods.py:60-61) from a second
// distributed factorisation of K + 1e-6 (a/b) I, (3) the closed-form tail.  nll_out_dev [1], mean_out [T], var_out [T].
int smnngp_test_nll_mg_f64(smnngp_mg* g_pred, smnngp_mg* g_lik, void* stream, const double* X, const double* y,
                           const double* Xt, const double* yt, int64_t D, int n_hidden, int act, int arch,
                           const double* hp_dev, int kind, double y_mean, double y_std, double* nll_out_dev,
                           double* mean_out, double* var_out, int* info_dev) {
  if (!g_pred || !X || !y || !Xt || !yt || !hp_dev || !nll_out_dev || !mean_out || !var_out || !info_dev || g_pred->c != 1 || (kind != KIND_GAUSS && kind != KIND_STUDENT_T) ||
      (kind == KIND_STUDENT_T && (!g_lik || g_lik->n != g_pred->n)))
    return mg_fail(SMNNGP_EINVAL, "smnngp_test_nll_mg_f64: invalid argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MG_RC(smnngp_predict_mg_f64(g_pred, stream, X, y, Xt, D, n_hidden, act, arch, hp_dev, SMNNGP_SHIFT_EPS_REL, mean_out,
                              var_out, info_dev));
  const double* quad2 = mean_out;                                   // unused for the Gaussian likelihood
  if (kind == KIND_STUDENT_T) {
    // second factorisation: K + 1e-6 (a/b) I, ||L2^-1 y||^2 ends up in g_lik's scalar block.  It gets its own status
    // word (factor_all clears the word it is given), folded into info_dev afterwards.
    Enter scope(s);
    MG_RC(check_call(g_lik, s, scope, "smnngp_test_nll_mg_f64"));
    const long long n = g_lik->n, npanels = cdiv(n, g_lik->db);
    MG_RC(factor_all(g_lik, s, X, y, nullptr, D, n_hidden, act, arch, hp_dev, SMNNGP_SHIFT_LIK, g_lik->info_tmp));
    if (g_lik->n_carried > 0) MG_CU(launch_sumsq(s, g_lik->carried, n, g_lik->sums + 1));
    MG_RC(reduce_all(g_lik, s, g_lik->seq_base + (unsigned long long)npanels + 1, g_lik->info_tmp));
    fold_info_kernel<<<1, 1, 0, s>>>(g_lik->info_tmp, info_dev);
    instr().launches++;
    MG_CU(cudaGetLastError());
    g_lik->seq_base += (unsigned long long)npanels + 4;
    arm_watchdog(g_lik, s, info_dev);
    quad2 = g_lik->scal + SC_QUAD;
  }
  MG_CU(launch_test_nll_finalize(s, mean_out, var_out, yt, (int)g_pred->t, g_pred->n, y_mean, y_std, hp_dev, kind, quad2,
                                 info_dev, nullptr, nll_out_dev));
  return SMNNGP_OK;
}

}  // extern "C"
