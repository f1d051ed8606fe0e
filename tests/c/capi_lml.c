/* Pure-C caller of libsmnngp.so (no Python, no torch): what a C / cgo / JNI / XLA-FFI host would do.
 *
 *   capi_lml <data.bin> N D [world]
 *
 * data.bin = X [N, D] doubles followed by y [N] doubles.  Prints
 *   loss_host  <v>            SPR.loss through the host-buffer entry point (smnngp_lml_host_f64)
 *   loss_dev   <v>            the same through the device entry point on an explicit stream (smnngp_lml_f64)
 *   loss_mg <world> <v>       (world > 1, needs that many GPUs) smnngp_lml_mg_f64: one handle and one host thread per
 *                             device in THIS process, peers connected by pointer - exercises the per-device contexts
 *                             and the re-entrancy of the library from several threads at once
 * Build:  gcc -O2 -std=c11 capi_lml.c -I<repo>/include -I/usr/local/cuda/include -L<libdir> -lsmnngp \
 *             -L/usr/local/cuda/lib64 -lcudart -lpthread -o capi_lml
 */
#define _POSIX_C_SOURCE 200809L
#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smnngp.h"

#define CK(call)                                                                  \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e__));         \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

static const double HP[6] = {1.0, 1e-8, 1.0, 1e-6, 2.0, 2.0}; /* reference CLI defaults, regression/train.py:37-45 */
enum { N_HIDDEN = 3, ACT_RELU = 0, ARCH_MLP = 0, KIND_STUDENT_T = 1 };

typedef struct {
  int rank, world;
  int64_t n, d, block;
  const double *x, *y;
  smnngp_mg* g;
  pthread_barrier_t* bar;
  void** regions;
  int* devices;
  double loss;
  int info, rc;
} rank_arg;

static void* rank_main(void* p) {
  rank_arg* a = (rank_arg*)p;
  a->rc = 1;
  CK(cudaSetDevice(a->rank));
  if (smnngp_mg_create(&a->g, a->rank, a->world, a->n, a->block) != SMNNGP_OK) {
    fprintf(stderr, "rank %d: smnngp_mg_create: %s\n", a->rank, smnngp_mg_last_error());
    exit(3);
  }
  a->regions[a->rank] = smnngp_mg_region(a->g);
  a->devices[a->rank] = a->rank;
  pthread_barrier_wait(a->bar);
  if (smnngp_mg_connect_ptrs(a->g, a->regions, a->devices) != SMNNGP_OK) {
    fprintf(stderr, "rank %d: smnngp_mg_connect_ptrs: %s\n", a->rank, smnngp_mg_last_error());
    exit(3);
  }
  double *dx, *dy, *dhp, *dout;
  int* dinfo;
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CK(cudaMalloc((void**)&dx, (size_t)a->n * a->d * 8));
  CK(cudaMalloc((void**)&dy, (size_t)a->n * 8));
  CK(cudaMalloc((void**)&dhp, 6 * 8));
  CK(cudaMalloc((void**)&dout, 4 * 8));
  CK(cudaMalloc((void**)&dinfo, sizeof(int)));
  CK(cudaMemcpyAsync(dx, a->x, (size_t)a->n * a->d * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dy, a->y, (size_t)a->n * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dhp, HP, 6 * 8, cudaMemcpyHostToDevice, s));
  pthread_barrier_wait(a->bar);
  double out[4];
  for (int rep = 0; rep < 2; rep++) { /* second evaluation: sequence numbers / buffers are reused */
    if (smnngp_lml_mg_f64(a->g, s, dx, dy, a->d, N_HIDDEN, ACT_RELU, ARCH_MLP, dhp, KIND_STUDENT_T,
                          SMNNGP_SHIFT_EPS_ABS, dout, dinfo) != SMNNGP_OK) {
      fprintf(stderr, "rank %d: smnngp_lml_mg_f64: %s\n", a->rank, smnngp_mg_last_error());
      exit(3);
    }
    CK(cudaMemcpyAsync(out, dout, 4 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&a->info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  a->loss = out[1];
  pthread_barrier_wait(a->bar);
  smnngp_mg_destroy(a->g);
  cudaFree(dx); cudaFree(dy); cudaFree(dhp); cudaFree(dout); cudaFree(dinfo);
  cudaStreamDestroy(s);
  a->rc = 0;
  return NULL;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s data.bin N D [world]\n", argv[0]);
    return 1;
  }
  const int64_t n = atoll(argv[2]), d = atoll(argv[3]);
  const int world = argc > 4 ? atoi(argv[4]) : 1;
  double* x = (double*)malloc((size_t)n * d * 8);
  double* y = (double*)malloc((size_t)n * 8);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(x, 8, (size_t)n * d, f) != (size_t)n * d || fread(y, 8, (size_t)n, f) != (size_t)n) {
    fprintf(stderr, "cannot read %s\n", argv[1]);
    return 1;
  }
  fclose(f);
  if (smnngp_abi_version() != SMNNGP_ABI_VERSION) return 4;

  /* (1) host-buffer entry point */
  double out[4];
  int info = -1;
  CK(cudaSetDevice(0));
  if (smnngp_lml_host_f64(x, y, n, d, N_HIDDEN, ACT_RELU, ARCH_MLP, HP, KIND_STUDENT_T, out, &info) != SMNNGP_OK) {
    fprintf(stderr, "smnngp_lml_host_f64: %s\n", smnngp_last_error());
    return 3;
  }
  printf("loss_host %.17g info %d\n", out[1], info);

  /* (2) device entry point on the caller's stream with a caller-owned workspace */
  {
    double *dx, *dy, *dhp, *dout;
    int* dinfo;
    void* ws;
    cudaStream_t s;
    const size_t wsb = smnngp_lml_workspace_bytes(n, d, N_HIDDEN, ARCH_MLP);
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaMalloc((void**)&dx, (size_t)n * d * 8));
    CK(cudaMalloc((void**)&dy, (size_t)n * 8));
    CK(cudaMalloc((void**)&dhp, 6 * 8));
    CK(cudaMalloc((void**)&dout, 4 * 8));
    CK(cudaMalloc((void**)&dinfo, sizeof(int)));
    CK(cudaMalloc(&ws, wsb));
    CK(cudaMemcpyAsync(dx, x, (size_t)n * d * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dy, y, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dhp, HP, 6 * 8, cudaMemcpyHostToDevice, s));
    if (smnngp_lml_f64(s, dx, dy, n, d, N_HIDDEN, ACT_RELU, ARCH_MLP, dhp, KIND_STUDENT_T, ws, wsb, dout, dinfo) !=
        SMNNGP_OK) {
      fprintf(stderr, "smnngp_lml_f64: %s\n", smnngp_last_error());
      return 3;
    }
    CK(cudaMemcpyAsync(out, dout, 4 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    printf("loss_dev %.17g info %d\n", out[1], info);
    cudaFree(dx); cudaFree(dy); cudaFree(dhp); cudaFree(dout); cudaFree(dinfo); cudaFree(ws);
    cudaStreamDestroy(s);
  }
  smnngp_host_release();

  /* (3) multi-GPU: one thread and one handle per device of this process */
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (world > 1 && ndev >= world) {
    pthread_t th[8];
    rank_arg args[8];
    void* regions[8] = {0};
    int devices[8] = {0};
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)world);
    const int64_t block = n / world >= 4096 ? 512 : (n / world >= 1024 ? 256 : 128);
    for (int r = 0; r < world; r++) {
      rank_arg a = {r, world, n, d, block, x, y, NULL, &bar, regions, devices, 0.0, -1, 1};
      args[r] = a;
      pthread_create(&th[r], NULL, rank_main, &args[r]);
    }
    for (int r = 0; r < world; r++) pthread_join(th[r], NULL);
    for (int r = 0; r < world; r++) {
      if (args[r].rc != 0) return 5;
      printf("loss_mg %d rank %d %.17g info %d\n", world, r, args[r].loss, args[r].info);
    }
  } else if (world > 1) {
    printf("loss_mg skipped: %d devices\n", ndev);
  }
  free(x);
  free(y);
  return 0;
}
