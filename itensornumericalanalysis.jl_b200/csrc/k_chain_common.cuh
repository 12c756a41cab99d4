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

// (w1:w0) += val << (64 * hi + sh) as ONE 128-bit number: a radix-3 field (several digits, each adding v * 3^i) may
// straddle the two stream words, and the partial sums must carry across the boundary.
__device__ __forceinline__ void stream_add128(uint64_t& w0, uint64_t& w1, uint64_t val, uint32_t sh, bool hi) {
  if (hi) {
    w1 += val << sh;
  } else {
    const uint64_t lo = val << sh;
    w0 += lo;
    w1 += (sh ? (val >> (64 - sh)) : 0ull) + (w0 < lo ? 1ull : 0ull);
  }
}

// base 3 / 4 digit entry for the branch-free form of the greedy loop (32 bytes).  With strictly increasing
// thresholds thr[1] < thr[2] < thr[3] the loop "largest v with x >= thr[v]" (abstractindexmap.jl:121-138) picks
// v = (x >= thr[1]) + (x >= thr[2]) + (x >= thr[3]); thresholds a base does not have are +inf.
struct __align__(16) Digit4 {
  double t1, t2, t3;
  uint32_t sh;   // (stride << 8) | shift inside the word   (radix-3 fields: stride up to 3^12)
  uint32_t wv;   // (base << 24) | (site << 16) | (word << 8)
};
__device__ __forceinline__ Digit4 make_digit4(const DigitTable& dg, int i) {
  const DigitEntry e = dg.entries[i];
  Digit4 d4;
  d4.t1 = e.base > 1 ? dg.thr[e.thr_off + 1] : __longlong_as_double(0x7ff0000000000000ll);
  d4.t2 = e.base > 2 ? dg.thr[e.thr_off + 2] : __longlong_as_double(0x7ff0000000000000ll);
  d4.t3 = e.base > 3 ? dg.thr[e.thr_off + 3] : __longlong_as_double(0x7ff0000000000000ll);
  d4.sh = ((uint32_t)e.stride << 8) | (uint32_t)e.shift;
  d4.wv = ((uint32_t)e.base << 24) | ((uint32_t)e.site << 16) | ((uint32_t)e.word << 8);
  return d4;
}
// One step of the loop for one point in the team-sorted kernel (few warps per SM: K1 is a LATENCY chain there): the chosen
// value v, x updated.  Base 2 / 3: both candidate residuals are formed beside the compares (x - t_v is only USED for the
// chosen v, so it is the loop's own __dsub_rn) — one FP64 operation and two selects per digit instead of compare -> select
// -> subtract: 40-site base-3 chain 4.1 -> 5.0 G points/s.  Base 4 keeps select-then-subtract (a third speculative
// subtraction costs the FP64 pipe more than the shorter chain gains), and so does the table kernel (k1_digit4 below), whose
// many resident warps make K1 a THROUGHPUT matter: there the speculative form measured 19-30 % slower.
__device__ __forceinline__ uint32_t digit4_step(double& x, const Digit4& e, bool base4) {
  const bool g1 = x >= e.t1, g2 = x >= e.t2;
  if (base4) { // warp-uniform
    const bool g3 = x >= e.t3;
    x = __dsub_rn(x, g3 ? e.t3 : (g2 ? e.t2 : (g1 ? e.t1 : 0.0)));
    return (uint32_t)g1 + (uint32_t)g2 + (uint32_t)g3;
  }
  const double x1 = __dsub_rn(x, e.t1), x2 = __dsub_rn(x, e.t2);
  x = g2 ? x2 : (g1 ? x1 : x); // no threshold reached: x - thr[0] = x - 0 = x
  return (uint32_t)g1 + (uint32_t)g2;
}
template <int NP>
__device__ __forceinline__ void k1_digit4(const Digit4 e, const DigitTable& dg, const CoordSource& src, int64_t p0, int64_t step,
                                          double (&x)[NP], uint64_t (&w0)[NP], uint64_t (&w1)[NP], int* err) {
  const uint32_t stride = e.sh >> 8, sh = e.sh & 0xffu;
  const bool hi = ((e.wv >> 8) & 0xffu) != 0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    uint32_t v;
    if (src.digits) {
      v = (uint32_t)given_digit(src, p0 + k * step, dg.n_sites, (int)((e.wv >> 16) & 0xffu), (int)(e.wv >> 24), err);
    } else {
      const bool g1 = x[k] >= e.t1, g2 = x[k] >= e.t2, g3 = x[k] >= e.t3;
      v = (uint32_t)g1 + (uint32_t)g2 + (uint32_t)g3;
      x[k] = __dsub_rn(x[k], g3 ? e.t3 : (g2 ? e.t2 : (g1 ? e.t1 : 0.0)));
    }
    stream_add128(w0[k], w1[k], (uint64_t)(v * stride), sh, hi);
  }
}
// The entries [e0, e1) of one coordinate for NP points: digits that land in the same stream field (a radix-3 group field
// holds up to 12 of them, each adding v * 3^i) are summed in a 32-bit register and inserted into the 128-bit stream once
// per field instead of once per digit.  (Partial sums of a field whose digits belong to several coordinates are simply
// added one after the other.)
template <int NP>
__device__ __forceinline__ void k1_digit4_run(const Digit4* s_d4, int e0, int e1, const DigitTable& dg, const CoordSource& src,
                                              int64_t p0, int64_t step, double (&x)[NP], uint64_t (&w0)[NP], uint64_t (&w1)[NP],
                                              int* err) {
  if (e0 >= e1) return;
  uint32_t acc[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) acc[k] = 0u;
  uint32_t cur = (s_d4[e0].sh & 0xffu) | (s_d4[e0].wv & 0xff00u); // (shift, word) of the open field
  for (int e_i = e0; e_i < e1; ++e_i) {
    const Digit4 e = s_d4[e_i];
    const uint32_t pos = (e.sh & 0xffu) | (e.wv & 0xff00u);
    if (pos != cur) {
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        stream_add128(w0[k], w1[k], (uint64_t)acc[k], cur & 0xffu, (cur >> 8) != 0u);
        acc[k] = 0u;
      }
      cur = pos;
    }
    const uint32_t stride = e.sh >> 8;
    const bool base4 = (e.wv >> 24) > 3u;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const uint32_t v = src.digits ? (uint32_t)given_digit(src, p0 + k * step, dg.n_sites, (int)((e.wv >> 16) & 0xffu), (int)(e.wv >> 24), err)
                                    : digit4_step(x[k], e, base4);
      acc[k] += v * stride;
    }
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) stream_add128(w0[k], w1[k], (uint64_t)acc[k], cur & 0xffu, (cur >> 8) != 0u);
}

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
    stream_add128(w0[k], w1[k], (uint64_t)(uint32_t)(v * e.stride), (uint32_t)e.shift, e.word != 0);
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

// K1 mode of the kernels with a fused K1 for a description with non-binary site indices: 1 = branch-free Digit4
// form (every site index has dimension <= 4 and strictly increasing thresholds), 2 = tabulated greedy loop on
// the global-memory digit table (any base), 0 = all binary.
inline int k1_generic_mode(const ttn_desc* d) {
  bool binary = true, fast = true;
  for (int s = 0; s < d->n_sites; ++s) {
    binary = binary && d->site_dim[s] == 2;
    fast = fast && d->site_dim[s] <= 4;
    for (int v = 1; v < d->site_dim[s]; ++v) fast = fast && d->thr[d->thr_ptr[s] + v] > d->thr[d->thr_ptr[s] + v - 1];
  }
  return binary ? 0 : (fast ? 1 : 2);
}

} // namespace ttn
