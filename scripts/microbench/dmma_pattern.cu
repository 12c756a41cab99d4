// Microbenchmark 2: the DMMA issue pattern of process_batch (distinct A/B/C registers, B fragments
// from shared memory), vs warps per SM.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHI, int NBAT, bool SMEMB>
__global__ void k(double* out, int iters) {
  constexpr int NB = CHI / 8, KB = CHI / 4;
  __shared__ double sB[KB * NB * 32];
  for (int i = threadIdx.x; i < KB * NB * 32; i += blockDim.x) sB[i] = 1.0 / (1 + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double a[NBAT][KB];
#pragma unroll
  for (int b = 0; b < NBAT; ++b)
#pragma unroll
    for (int j = 0; j < KB; ++j) a[b][j] = 1.0 + 1e-3 * (threadIdx.x + b + j);
  for (int it = 0; it < iters; ++it) {
    double bf[KB * NB];
#pragma unroll
    for (int j = 0; j < KB * NB; ++j) bf[j] = SMEMB ? sB[j * 32 + lane] : 1.0 / (1 + j + it);
    double acc[NBAT][KB];
#pragma unroll
    for (int b = 0; b < NBAT; ++b)
#pragma unroll
      for (int j = 0; j < KB; ++j) acc[b][j] = 0.0;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
      for (int nbp = 0; nbp < NB; ++nbp)
#pragma unroll
        for (int b = 0; b < NBAT; ++b) dmma(acc[b][2 * nbp], acc[b][2 * nbp + 1], a[b][kb], bf[kb * NB + nbp]);
#pragma unroll
    for (int b = 0; b < NBAT; ++b)
#pragma unroll
      for (int j = 0; j < KB; ++j) a[b][j] = acc[b][j] * 1e-3;
  }
  double s = 0;
#pragma unroll
  for (int b = 0; b < NBAT; ++b)
#pragma unroll
    for (int j = 0; j < KB; ++j) s += a[b][j];
  if (s == 12345.678) out[0] = s;
}

template <int CHI, int NBAT, bool SMEMB>
void run(int warps_per_sm, int sms, double* d) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  k<CHI, NBAT, SMEMB><<<sms, warps_per_sm * 32>>>(d, iters);
  cudaEventRecord(e0);
  k<CHI, NBAT, SMEMB><<<sms, warps_per_sm * 32>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 512.0 * (CHI / 4) * (CHI / 8) * NBAT * iters * (double)sms * warps_per_sm;
  printf("CHI %2d NBAT %d smemB %d warps/SM %2d : %7.2f TFLOP/s\n", CHI, NBAT, (int)SMEMB, warps_per_sm, fl / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* d; cudaMalloc(&d, 64);
  for (int w : {4, 8, 12, 16}) {
    run<16, 4, false>(w, p.multiProcessorCount, d);
    run<16, 4, true>(w, p.multiProcessorCount, d);
    run<16, 2, true>(w, p.multiProcessorCount, d);
    run<16, 1, true>(w, p.multiProcessorCount, d);
    run<32, 4, true>(w, p.multiProcessorCount, d);
  }
  return 0;
}
