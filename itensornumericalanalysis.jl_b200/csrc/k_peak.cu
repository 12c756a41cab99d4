// FP64 roofline denominators measured on the device (MEASURED_PEAKS.json carries
// only HBM and bf16): a DFMA issue-rate loop for the vector pipe and an mma.sync m8n8k4 f64 loop
// (SASS DMMA.8x8x4) for the tensor pipe.  Both are register-only, so they are upper bounds no
// real kernel with operand traffic can exceed.
#include "ttn_internal.h"

namespace ttn {

__global__ void __launch_bounds__(512) dfma_peak_kernel(double* out, int iters, double seed) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.6789) out[0] = s; // never true; keeps the loop alive
}

__global__ void __launch_bounds__(512) dmma_peak_kernel(double* out, int iters, double seed) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = seed + (threadIdx.x & 31) * 1e-3, b = 1.0 / (1 + (threadIdx.x & 7));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.6789) out[0] = s;
}

int measure_fp64_peak(int device, double* dfma, double* dmma) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_error("no such CUDA device");
    return TTN_ERR_CUDA;
  }
  TTN_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TTN_CUDA(cudaGetDeviceProperties(&prop, device));
  double* d_out;
  TTN_CUDA(cudaMalloc(&d_out, 64));
  cudaEvent_t e0, e1;
  TTN_CUDA(cudaEventCreate(&e0));
  TTN_CUDA(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 2, threads = 512;
  const int iters = 1 << 15;
  double best_f = 0, best_m = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(d_out, iters, 1.0);
    cudaEventRecord(e1);
    TTN_CUDA(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * (double)iters * blocks * threads;
    if (rep) best_f = std::max(best_f, fl / (ms * 1e-3) / 1e12);
    cudaEventRecord(e0);
    dmma_peak_kernel<<<blocks, threads>>>(d_out, iters / 4, 1.0);
    cudaEventRecord(e1);
    TTN_CUDA(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 512.0 * 8 * (double)(iters / 4) * blocks * (threads / 32);
    if (rep) best_m = std::max(best_m, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  TTN_CUDA(cudaGetLastError());
  *dfma = best_f;
  *dmma = best_m;
  return TTN_OK;
}

} // namespace ttn
