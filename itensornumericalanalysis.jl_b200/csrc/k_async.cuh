// PTX wrappers: mbarrier, TMA bulk copy (cp.async.bulk -> SASS UBLKCP), shared-memory vector access,
// FP64 tensor-core MMA (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; tcgen05 has no FP64 kind).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ttn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA bulk copy global -> shared, completion reported to an mbarrier as transaction bytes.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
  return r;
}

__device__ __forceinline__ double lds64(uint32_t addr) {
  double r;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
// 256-bit read-only global load (sm_100: LDG.E.256): a lane that walks its own row issues half as many
// loads — and L1 wavefronts — as with 128-bit ones.  p must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// 128-bit read-only global load
__device__ __forceinline__ double2 ldg128(const double* p) {
  double2 r;
  asm("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
// 16-byte asynchronous copy global -> shared past L1 (SASS LDGSTS.E.BYPASS.128)
__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// D(8x8) += A(8x4) * B(4x8).  Fragments (PTX ISA, m8n8k4 .f64): lane = 4*g + t;
//   A[row g][col t], B[row t][col g], C/D[row g][cols 2t, 2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
#ifdef TTN_NO_DMMA   // experiment: measure everything but the tensor work
  c0 = a; c1 = b; return;
#endif
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

} // namespace ttn
