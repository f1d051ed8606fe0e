// Library bar for the multi-GPU factorisation: cusolverMgPotrf (FP64, lower, 1-D column block-cyclic over the
// visible GPUs of one process) on an SPD matrix of order N.  Prints the time and N^3/3 TFLOP/s.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a cusolvermg_potrf.cu -lcusolverMg -lcusolver -lcublas -o cusolvermg_potrf
//   ./cusolvermg_potrf [N] [block] [reps]
// Development / measurement tool only (profiles/): never linked into the product library.
#include <cuda_runtime.h>
#include <cusolverMg.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 2; } \
  } while (0)
#define SK(x)                                                                \
  do {                                                                       \
    cusolverStatus_t s_ = (x);                                               \
    if (s_ != CUSOLVER_STATUS_SUCCESS) { printf("%s: status %d\n", #x, (int)s_); return 3; } \
  } while (0)

// local column lc of device dev holds global column ((lc / T) * P + dev) * T + lc % T:  A = (N - 1) I + 1 1^T
__global__ void fill_kernel(double* a, long long n, long long lcols, int T, int P, int dev) {
  const long long lc = blockIdx.y;
  const long long gc = ((lc / T) * P + dev) * T + lc % T;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x)
    a[lc * n + r] = (r == gc) ? (double)n : 1.0;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 60000;
  const int T = argc > 2 ? atoi(argv[2]) : 256;
  const int reps = argc > 3 ? atoi(argv[3]) : 2;
  int P = 0;
  CK(cudaGetDeviceCount(&P));
  std::vector<int> devs(P);
  for (int i = 0; i < P; i++) devs[i] = i;
  for (int i = 0; i < P; i++) {
    CK(cudaSetDevice(i));
    for (int j = 0; j < P; j++)
      if (i != j) cudaDeviceEnablePeerAccess(j, 0);
    cudaGetLastError();
  }
  cusolverMgHandle_t h;
  SK(cusolverMgCreate(&h));
  SK(cusolverMgDeviceSelect(h, P, devs.data()));
  cudaLibMgGrid_t grid;
  SK(cusolverMgCreateDeviceGrid(&grid, 1, P, devs.data(), CUDALIBMG_GRID_MAPPING_COL_MAJOR));
  cudaLibMgMatrixDesc_t desc;
  SK(cusolverMgCreateMatrixDesc(&desc, N, N, N, T, CUDA_R_64F, grid));
  const long long nblocks = (N + T - 1) / T;
  std::vector<void*> dA(P), dW(P);
  std::vector<long long> lcols(P);
  for (int d = 0; d < P; d++) {
    long long nb_d = nblocks / P + (d < nblocks % P ? 1 : 0);
    lcols[d] = nb_d * T;                         // whole blocks (the last global block may be partial: padded)
    CK(cudaSetDevice(d));
    CK(cudaMalloc(&dA[d], (size_t)lcols[d] * N * sizeof(double)));
  }
  int64_t lwork = 0;
  SK(cusolverMgPotrf_bufferSize(h, CUBLAS_FILL_MODE_LOWER, N, dA.data(), 1, 1, desc, CUDA_R_64F, &lwork));
  for (int d = 0; d < P; d++) {
    CK(cudaSetDevice(d));
    CK(cudaMalloc(&dW[d], (size_t)lwork * sizeof(double)));
  }
  double best = 1e30;
  int info = -1;
  for (int r = 0; r < reps + 1; r++) {
    for (int d = 0; d < P; d++) {
      CK(cudaSetDevice(d));
      dim3 g(64, (unsigned)lcols[d]);
      fill_kernel<<<g, 256>>>((double*)dA[d], N, lcols[d], T, P, d);
    }
    for (int d = 0; d < P; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
    CK(cudaSetDevice(0));
    auto t0 = std::chrono::steady_clock::now();
    SK(cusolverMgPotrf(h, CUBLAS_FILL_MODE_LOWER, N, dA.data(), 1, 1, desc, CUDA_R_64F, dW.data(), lwork, &info));
    for (int d = 0; d < P; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (r > 0 && ms < best) best = ms;
    printf("cusolverMgPotrf N=%d P=%d T=%d run %d: %.2f ms  info %d\n", N, P, T, r, ms, info);
    fflush(stdout);
  }
  printf("{\"library\": \"cusolverMgPotrf\", \"N\": %d, \"gpus\": %d, \"block\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"info\": %d}\n",
         N, P, T, best, (double)N * N * N / 3.0 / (best * 1e-3) * 1e-12, info);
  return 0;
}
