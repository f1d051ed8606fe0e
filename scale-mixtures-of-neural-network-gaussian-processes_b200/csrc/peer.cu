// Peer-memory plumbing for the multi-GPU path (one process per GPU): device buffers that every rank of the node
// can address directly over NVLink / NVSwitch.  Allocation + CUDA IPC handle export / import only; the kernels
// that store into peer buffers live in stages.cu (fused panel solve + all-gather).
#include "../../include/smnngp.h"

#include <cuda_runtime.h>

extern "C" {

// cudaMalloc'ed (IPC-exportable, never from a pool) buffer + its 64-byte IPC handle
int smnngp_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out64) {
  if (!ptr_out || !handle_out64 || bytes == 0) return SMNNGP_EINVAL;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return SMNNGP_ECUDA;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
    cudaFree(p);
    return SMNNGP_ECUDA;
  }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  for (int i = 0; i < 64; i++) handle_out64[i] = reinterpret_cast<unsigned char*>(&h)[i];
  *ptr_out = p;
  return SMNNGP_OK;
}

int smnngp_peer_open(const unsigned char* handle64, void** ptr_out) {
  if (!handle64 || !ptr_out) return SMNNGP_EINVAL;
  cudaIpcMemHandle_t h;
  for (int i = 0; i < 64; i++) reinterpret_cast<unsigned char*>(&h)[i] = handle64[i];
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return SMNNGP_ECUDA;
  *ptr_out = p;
  return SMNNGP_OK;
}

int smnngp_peer_close(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? SMNNGP_OK : SMNNGP_ECUDA; }
int smnngp_peer_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? SMNNGP_OK : SMNNGP_ECUDA; }

}  // extern "C"
