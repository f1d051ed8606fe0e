// FP64 denominators for the roofline on B200 (sm_100a): register-resident DMMA.8x8x4 issue rate,
// DFMA issue rate, both together (do they share a pipe?), FP64 transcendental cost, and the library
// bars (cuBLAS Dgemm / Dsyrk, cuSOLVER Dpotrf).  Measurement tool only - not part of the product path.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo fp64_peak.cu -o fp64_peak -lcublas -lcusolver
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double seed) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = 0; c1[i] = 0; }
  double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double seed) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = seed + threadIdx.x * 1e-9, b = seed * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  if (s == 123.456) out[0] = s;
}

// even warps DMMA, odd warps DFMA: if total time ~ max(t_dmma, t_dfma) the pipes are separate
template <int NACC>
__global__ void k_mixed(double* out, int iters_mma, int iters_fma, double seed) {
  int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = seed + threadIdx.x * 1e-9, b = seed * 1e-3;
    for (int it = 0; it < iters_fma; it++) {
#pragma unroll
      for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
  } else {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = 0; c1[i] = 0; }
    double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
    for (int it = 0; it < iters_mma; it++) {
#pragma unroll
      for (int i = 0; i < NACC; i++) dmma(c0[i], c1[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  }
  if (s == 123.456) out[0] = s;
}

// transcendental cost per element: op 0 = relu arc-cosine step (sqrt+atan2), 1 = erf step (asin + rsqrt-ish), 2 = sqrt only, 3 = div
template <int OP>
__global__ void k_trans(double* out, int iters, double seed) {
  double x = seed * (1.0 + 1e-6 * threadIdx.x), q1 = 1.0 + 1e-7 * blockIdx.x, q2 = 1.1, acc = 0;
  for (int it = 0; it < iters; it++) {
    double k = x;
    if (OP == 0) {
      double s = sqrt(fmax(q1 * q2 - k * k, 0.0));
      double th = atan2(s, k);
      k = s * 0.15915494309189535 + (0.5 - th * 0.15915494309189535) * k;
    } else if (OP == 1) {
      k = 0.6366197723675814 * asin(2.0 * k / sqrt((1 + 2 * q1) * (1 + 2 * q2)));
    } else if (OP == 2) {
      k = sqrt(fmax(q1 * q2 - k * k, 0.0));
    } else if (OP == 3) {
      k = q1 / (k + 2.0);
    } else if (OP == 4) {
      k = acos(k * rsqrt(q1 * q2));
    }
    acc += k;
    x = x * 0.999 + 1e-4;
  }
  if (acc == 123.456) out[0] = acc;
}

static float time_it(void (*launch)(void*), void* ctx, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(ctx); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(ctx); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

struct Cfg { int blocks, threads, iters; double* out; };

#define RUN_DMMA(NACC) { Cfg c{blocks, threads, iters, dout}; \
  float ms = time_it([](void* p){ Cfg* c=(Cfg*)p; k_dmma<NACC><<<c->blocks,c->threads>>>(c->out,c->iters,1.0); }, &c, 3); \
  double flop = 2.0*8*8*4*(double)NACC*iters*(threads/32)*blocks; \
  printf("DMMA  blocks/SM=%d warps/blk=%2d nacc=%2d : %8.3f ms  %7.2f TFLOP/s\n", blocks/nsm, threads/32, NACC, ms, flop/ms*1e-9); }
#define RUN_DFMA(NACC) { Cfg c{blocks, threads, iters, dout}; \
  float ms = time_it([](void* p){ Cfg* c=(Cfg*)p; k_dfma<NACC><<<c->blocks,c->threads>>>(c->out,c->iters,1.0); }, &c, 3); \
  double flop = 2.0*32*(double)NACC*iters*(threads/32)*blocks; \
  printf("DFMA  blocks/SM=%d warps/blk=%2d nacc=%2d : %8.3f ms  %7.2f TFLOP/s\n", blocks/nsm, threads/32, NACC, ms, flop/ms*1e-9); }

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  int nsm = prop.multiProcessorCount;
  printf("device %s SMs=%d clock=%d kHz smem/SM=%zu\n", prop.name, nsm, prop.clockRate, prop.sharedMemPerMultiprocessor);
  double* dout; CK(cudaMalloc(&dout, 1024));
  int iters = 20000;
  for (int wpb : {4, 8, 16, 32}) {
    int threads = wpb * 32; int blocks = nsm;
    RUN_DMMA(1) RUN_DMMA(4) RUN_DMMA(16) RUN_DMMA(32)
  }
  { int threads = 256, blocks = 2 * nsm; RUN_DMMA(16) }
  for (int wpb : {4, 8, 16, 32}) {
    int threads = wpb * 32; int blocks = nsm;
    RUN_DFMA(1) RUN_DFMA(4) RUN_DFMA(16)
  }
  // mixed
  for (int ratio : {0, 1, 2, 4}) {
    struct M { int blocks, im, ifm; double* out; } m{nsm, iters, iters * ratio * 4, dout};
    // per iteration: dmma warp does 16 DMMA (16*512 flop), dfma warp does 16 DFMA*32 lanes*2 = 1024 flop
    float ms = time_it([](void* p){ M* m=(M*)p; k_mixed<16><<<m->blocks,512>>>(m->out,m->im,m->ifm,1.0); }, &m, 3);
    double f_mma = 2.0*256*16*(double)m.im*8*nsm, f_fma = 2.0*32*16*(double)m.ifm*8*nsm;
    printf("MIXED 8 dmma warps + 8 dfma warps (fma iters x%d): %8.3f ms  dmma %7.2f TF/s + dfma %7.2f TF/s\n", ratio*4, ms, f_mma/ms*1e-9, f_fma/ms*1e-9);
  }
  // transcendental
  {
    int blocks = nsm * 4, threads = 512, it2 = 2000;
    const char* names[] = {"relu-step(sqrt+atan2)", "erf-step(asin+sqrt+div)", "sqrt", "div", "acos+rsqrt"};
    float ms[5];
    struct T { int b, t, i; double* o; } t{blocks, threads, it2, dout};
    ms[0] = time_it([](void* p){ T* t=(T*)p; k_trans<0><<<t->b,t->t>>>(t->o,t->i,0.3); }, &t, 3);
    ms[1] = time_it([](void* p){ T* t=(T*)p; k_trans<1><<<t->b,t->t>>>(t->o,t->i,0.3); }, &t, 3);
    ms[2] = time_it([](void* p){ T* t=(T*)p; k_trans<2><<<t->b,t->t>>>(t->o,t->i,0.3); }, &t, 3);
    ms[3] = time_it([](void* p){ T* t=(T*)p; k_trans<3><<<t->b,t->t>>>(t->o,t->i,0.3); }, &t, 3);
    ms[4] = time_it([](void* p){ T* t=(T*)p; k_trans<4><<<t->b,t->t>>>(t->o,t->i,0.3); }, &t, 3);
    for (int i = 0; i < 5; i++) {
      double ev = (double)blocks * threads * it2;
      printf("TRANS %-26s: %8.3f ms  %8.2f Geval/s  (%.1f ns/eval/SM-thread-slot)\n", names[i], ms[i], ev / ms[i] * 1e-6, ms[i]*1e6/it2);
    }
  }
  // cuBLAS
  cublasHandle_t h; cublasCreate(&h);
  auto gemm = [&](int M, int N, int K) {
    double *A, *B, *C; CK(cudaMalloc(&A, (size_t)M*K*8)); CK(cudaMalloc(&B, (size_t)N*K*8)); CK(cudaMalloc(&C, (size_t)M*N*8));
    CK(cudaMemset(A, 0, (size_t)M*K*8)); CK(cudaMemset(B, 0, (size_t)N*K*8)); CK(cudaMemset(C, 0, (size_t)M*N*8));
    double al = -1.0, be = 1.0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // column-major C(MxN) = A^T(MxK) * B(KxN) with A stored KxM, i.e. both operands K-contiguous (our row-major A*B^T)
    for (int w = 0; w < 2; w++) cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, M, N, K, &al, A, K, B, K, &be, C, M);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, M, N, K, &al, A, K, B, K, &be, C, M); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf("cuBLAS Dgemm TN M=%d N=%d K=%d : %8.3f ms %7.2f TFLOP/s\n", M, N, K, best, 2.0*M*N*K/best*1e-9);
    // sustained: 3 s back-to-back
    if (K >= 4096) {
      int n = (int)(3000.0 / best) + 1; cudaEventRecord(e0);
      for (int r = 0; r < n; r++) cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, M, N, K, &al, A, K, B, K, &be, C, M);
      cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("cuBLAS Dgemm TN sustained %d calls: %7.2f TFLOP/s\n", n, 2.0*M*N*K*n/ms*1e-9);
    }
    cudaFree(A); cudaFree(B); cudaFree(C);
  };
  gemm(8192, 8192, 8192);
  gemm(16384, 16384, 256);
  gemm(16384, 16384, 512);
  gemm(16384, 16384, 784);
  {
    int N = 16384;
    for (int K : {256, 512}) {
      double *A, *C; CK(cudaMalloc(&A, (size_t)N*K*8)); CK(cudaMalloc(&C, (size_t)N*N*8));
      CK(cudaMemset(A, 0, (size_t)N*K*8)); CK(cudaMemset(C, 0, (size_t)N*N*8));
      double al = -1.0, be = 1.0; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int w = 0; w < 2; w++) cublasDsyrk(h, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, N, K, &al, A, K, &be, C, N);
      CK(cudaDeviceSynchronize()); float best = 1e30f;
      for (int r = 0; r < 5; r++) { cudaEventRecord(e0); cublasDsyrk(h, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, N, K, &al, A, K, &be, C, N); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
      printf("cuBLAS Dsyrk N=%d K=%d : %8.3f ms %7.2f TFLOP/s (N*(N+1)*K flop)\n", N, K, best, (double)N*(N+1)*K/best*1e-9);
      cudaFree(A); cudaFree(C);
    }
  }
  // cuSOLVER potrf
  {
    cusolverDnHandle_t sh; cusolverDnCreate(&sh);
    for (int N : {8192, 16384, 32768}) {
      double* A; CK(cudaMalloc(&A, (size_t)N*N*8));
      std::vector<double> hdiag(N, (double)N);
      int lwork; cusolverDnDpotrf_bufferSize(sh, CUBLAS_FILL_MODE_LOWER, N, A, N, &lwork);
      double* work; CK(cudaMalloc(&work, (size_t)lwork*8)); int* info; CK(cudaMalloc(&info, 4));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30f;
      for (int r = 0; r < 3; r++) {
        CK(cudaMemset(A, 0, (size_t)N*N*8));
        CK(cudaMemcpy2D(A, (size_t)(N+1)*8, hdiag.data(), 8, 8, N, cudaMemcpyHostToDevice));
        cudaEventRecord(e0); cusolverDnDpotrf(sh, CUBLAS_FILL_MODE_LOWER, N, A, N, work, lwork, info); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      int hinfo; cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
      printf("cuSOLVER Dpotrf N=%d : %8.3f ms %7.2f TFLOP/s (N^3/3) info=%d\n", N, best, (double)N*N*N/3/best*1e-9, hinfo);
      cudaFree(A); cudaFree(work); cudaFree(info);
    }
  }
  return 0;
}
