// Per-device context registry (context.cuh) + the accessors the launch code uses.
#include "context.cuh"

namespace smnngp {

namespace {
constexpr int MAX_DEVICES = 64;
std::mutex g_registry_mu;
DeviceCtx* g_ctx[MAX_DEVICES] = {};
}  // namespace

DeviceCtx& dctx_of(int device) {
  if (device < 0 || device >= MAX_DEVICES) device = 0;
  std::lock_guard<std::mutex> lk(g_registry_mu);
  DeviceCtx*& c = g_ctx[device];
  if (c == nullptr) {
    c = new DeviceCtx();
    c->device = device;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) c->sms = sms;
  }
  return *c;
}

DeviceCtx& dctx() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return dctx_of(dev);
}

cudaError_t configure_kernel_once(const void* fn, int smem_bytes, bool carveout) {
  DeviceCtx& c = dctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if (c.configured.count(fn)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  if (carveout) {
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
  }
  c.configured.insert(fn);
  return cudaSuccess;
}

Enter::Enter(cudaStream_t s) {
  cudaGetDevice(&prev);
  dev = prev;
  // the legacy / per-thread default streams belong to whatever device is current
  // (not while the stream is being captured into a graph: cudaStreamGetDevice is not a capturable call and would
  // invalidate the capture; a capture was begun on the current device anyway)
  if (s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    int sd = -1;
    if (cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone &&
        cudaStreamGetDevice(s, &sd) == cudaSuccess && sd >= 0)
      dev = sd;
  }
  if (dev != prev) cudaSetDevice(dev);
  ctx = &dctx_of(dev);
  ctx->mu.lock();
}

Enter::~Enter() {
  ctx->mu.unlock();
  if (dev != prev && prev >= 0) cudaSetDevice(prev);
}

// ---- accessors (names kept from the process-global version) --------------------------------------------------
Instrumentation& instr() { return dctx().instr; }
int& tile_variant() { return dctx().tile_variant; }
int& lookahead_mode() { return dctx().lookahead; }
int& gram_super_rows() { return dctx().gram_super; }
int* lookahead_reserve() { return dctx().la_reserve; }
long long*& potf2_clock_buffer() { return dctx().potf2_clk; }
int device_sm_count() { return dctx().sms; }

void instr_reset() {
  DeviceCtx& c = dctx();
  c.instr.launches = 0;
  c.instr.update_flops = 0.0;
  c.instr.update_alg_flops = 0.0;
  c.ev_used = 0;
}

void instr_begin_update(cudaStream_t s, double alg_flops) {
  DeviceCtx& c = dctx();
  if (c.ev_used >= 4096) return;
  if (c.ev_used >= (int)c.ev.size()) {
    EvPair p{};
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
    c.ev.push_back(p);
  }
  c.instr.update_alg_flops += alg_flops;
  cudaEventRecord(c.ev[c.ev_used].a, s);
}

void instr_end_update(cudaStream_t s) {
  DeviceCtx& c = dctx();
  if (c.ev_used >= (int)c.ev.size()) return;
  cudaEventRecord(c.ev[c.ev_used].b, s);
  c.ev_used++;
}

double instr_collect_update_ms(int* n_out) {
  DeviceCtx& c = dctx();
  double total = 0.0;
  for (int i = 0; i < c.ev_used; i++) {
    cudaEventSynchronize(c.ev[i].b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c.ev[i].a, c.ev[i].b) == cudaSuccess) total += ms;
  }
  if (n_out) *n_out = c.ev_used;
  return total;
}

}  // namespace smnngp
