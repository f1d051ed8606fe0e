// Per-device library context.  Everything that used to be process-global mutable state (look-ahead side stream and
// events, host arena of the *_host_f64 entry points, instrumentation, tuning knobs, the "this kernel's dynamic
// shared-memory limit has been raised" flags, the SM count) lives here, one instance per device ordinal, created
// lazily.  Entry points open an `Enter` scope on the caller's stream: it resolves the stream's device (so a caller
// whose current device differs from the stream's - single-process multi-device frameworks such as JAX - still
// launches on the right GPU), makes it current for the duration of the call, and serialises host-side enqueueing on
// that device's context (CUDA event record / wait pairs on the shared side stream are ordered by enqueue time, so
// serialised enqueues from several host threads / streams stay correct).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <unordered_set>
#include <vector>

#include "kernels.cuh"

namespace smnngp {

struct EvPair { cudaEvent_t a, b; };

struct DeviceCtx {
  int device = -1;
  int sms = 148;
  std::recursive_mutex mu;
  // kernels whose MaxDynamicSharedMemorySize / carve-out has been configured on THIS device
  std::unordered_set<const void*> configured;
  // instrumentation (bench.py)
  Instrumentation instr;
  std::vector<EvPair> ev;
  int ev_used = 0;
  // tuning knobs (smnngp_set_*): apply to the device that is current when they are set
  int tile_variant = 0, lookahead = 1, la_reserve[2] = {8, 0}, panel_width = 0, fused_panel = 1, gram_super = 8;
  long long tail_cols = 2048;   // look-ahead factorisation (512-wide panels) hands the last <= tail_cols columns to the serial path
  long long gram_super_min_bytes = 40000000;   // B operand (M x D doubles) below this: plain row-major walk
  int peer_wait_mode = 0;       // smnngp_set_peer_wait_mode (exchange.cu)
  long long* potf2_clk = nullptr;
  // look-ahead side stream of the fused factorisation
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_diag = nullptr, ev_a = nullptr;
  bool la_ok = false;
  // grow-only device arena + stream of the *_host_f64 entry points
  void* arena = nullptr;
  size_t arena_bytes = 0;
  cudaStream_t arena_stream = nullptr;
};

// context of the device that is current on the calling thread
DeviceCtx& dctx();
DeviceCtx& dctx_of(int device);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize [, carve-out 100]) once per (kernel, device)
cudaError_t configure_kernel_once(const void* fn, int smem_bytes, bool carveout);

// RAII scope of one C-ABI call (see file comment)
struct Enter {
  int prev = -1, dev = -1;
  DeviceCtx* ctx = nullptr;
  explicit Enter(cudaStream_t s);
  ~Enter();
  Enter(const Enter&) = delete;
  Enter& operator=(const Enter&) = delete;
};

}  // namespace smnngp
