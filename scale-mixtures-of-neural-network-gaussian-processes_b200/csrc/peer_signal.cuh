// Cross-GPU completion signal shared by the kernels that store into peer buffers (exchange.cu, trtri.cu).
#pragma once
#include "common.cuh"

namespace smnngp {

constexpr int MAX_PEERS = 8;

struct PeerSignal {
  unsigned long long* flag[MAX_PEERS];   // flag word of THIS source on every destination rank (nullptr: skip)
  unsigned long long seq;
  unsigned int* counter;                 // device-local CTA counter (zero between launches); nullptr: no signal
  int P;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// called by ONE thread per CTA after the CTA's threads have fenced (system scope) and synchronised;
// total_ctas = number of CTAs of the launch that make this call
__device__ __forceinline__ void signal_if_last_cta(const PeerSignal& sg, unsigned int total_ctas) {
  if (sg.counter == nullptr) return;
  __threadfence_system();
  const unsigned int prev = atomicAdd(sg.counter, 1u);
  if (prev == total_ctas - 1) {
    *sg.counter = 0u;                    // next launch on this stream starts from zero
    __threadfence_system();
    for (int q = 0; q < sg.P; q++)
      if (sg.flag[q] != nullptr) st_release_sys(sg.flag[q], sg.seq);
  }
}

}  // namespace smnngp
