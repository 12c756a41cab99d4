// Microbenchmark: DMMA.8x8x4 throughput vs warps per SM and independent accumulator chains per warp.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_issue dmma_issue.cu && ./dmma_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, int iters) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 / (1 + (threadIdx.x & 7));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <int ILP>
void run(int warps_per_sm, int sms, double* d) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 14;
  k<ILP><<<sms, warps_per_sm * 32>>>(d, iters);
  cudaEventRecord(e0);
  k<ILP><<<sms, warps_per_sm * 32>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 512.0 * ILP * iters * (double)sms * warps_per_sm;
  printf("warps/SM %2d  ILP %d : %7.2f TFLOP/s\n", warps_per_sm, ILP, fl / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* d; cudaMalloc(&d, 64);
  for (int w : {4, 8, 12, 16, 32}) {
    run<1>(w, p.multiProcessorCount, d);
    run<2>(w, p.multiProcessorCount, d);
    run<4>(w, p.multiProcessorCount, d);
    run<8>(w, p.multiProcessorCount, d);
  }
  return 0;
}
