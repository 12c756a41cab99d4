// Shared by the DMMA chain kernels (k_chain_ring.cu, k_chain_team.cu) and their host side
// (k_chain_mma.cu): state-row swizzle, packed-stream constants, digit-table limits.
#pragma once
#include "k_async.cuh"
#include "k_digits.cuh"

namespace ttn {

constexpr int kMmaMaxStages = 16;
constexpr int kMaxClasses = 16;

// stream bits per chain position holding `nsl` slices (<= 32)
__host__ __device__ constexpr int slice_bits(int nsl) {
  return nsl <= 1 ? 0 : (nsl <= 2 ? 1 : (nsl <= 4 ? 2 : (nsl <= 8 ? 3 : (nsl <= 16 ? 4 : 5))));
}

template <int CHI>
__device__ __forceinline__ uint32_t row_chunk(uint32_t state_base, int row, int chunk) {
  constexpr int CPR = CHI / 2;                      // 16-byte chunks per row
  constexpr int RP = (CPR >= 8) ? 1 : 8 / CPR;      // rows per 128 bytes
  constexpr int SW = (CPR >= 8) ? 7 : CPR - 1;
  return state_base + (uint32_t)row * (CHI * 8) + (uint32_t)((chunk ^ ((row / RP) & SW)) << 4);
}

// compact base-2 digit entry for the branch-free fast path (16 bytes -> one LDS.128)
struct __align__(16) Digit2 {
  double thr1;   // |index_value_to_scalar(ind, 1)|
  uint32_t sh;   // shift inside the word
  uint32_t wv;   // (site << 16) | (word << 8) | stride
};

constexpr int kFeMaxSites = 160; // static shared-memory copies of the digit tables (within the 12 KB the
constexpr int kFeMaxThr = 640;   // launchers reserve); larger networks take the chain / generic kernels

// K1 of one site index for NP points of a thread when the network has non-binary site indices (base 3 / 4,
// test/test_realitensorfunction.jl:89-104): the tabulated greedy loop of k_digits.cuh on the global-memory digit
// table (warp-uniform addresses: one broadcast load per entry), digit value v placed as v * stride at the site's
// stream position.  step = distance between two points of the thread.
template <int NP>
__device__ __forceinline__ void k1_generic_site(const DigitTable& dg, const CoordSource& src, int e_i, int64_t p0, int64_t step,
                                                double (&x)[NP], uint64_t (&w0)[NP], uint64_t (&w1)[NP], int* err) {
  const DigitEntry e = dg.entries[e_i];
  const double* thr = dg.thr + e.thr_off;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int v = src.digits ? given_digit(src, p0 + k * step, dg.n_sites, e.site, e.base, err) : greedy_digit(x[k], thr, e.base);
    const uint64_t bb = (uint64_t)(uint32_t)(v * e.stride) << e.shift;
    if (e.word) w1[k] += bb;
    else w0[k] += bb;
  }
}

__device__ __forceinline__ int greedy_digit_smem(double& x, const double* thr, int base) {
  int v = base - 1;
  double t = thr[v];
  while (v > 0 && !(x >= t)) {
    --v;
    t = thr[v];
  }
  x = __dsub_rn(x, t);
  return v;
}

} // namespace ttn
