// K1 — digit extraction, shared by every kernel (one implementation => one bit-exactness proof).
//
// Restates the greedy loop set_ind_values! (src/IndexMaps/abstractindexmap.jl:121-138):
//   for ind in sorted_inds: ind_val = dim(ind)-1; while !(x_rn >= |place(ind, ind_val)|) ind_val--;
//                           x_rn -= |place(ind, ind_val)|
// with the place values |index_value_to_scalar| tabulated by the caller (thr[]).  Only FP64
// compares and one FP64 subtract (__dsub_rn: never contracted into an FMA) per digit, so the
// result is bit-identical to the reference for any base and any digit numbering.
#pragma once
#include "ttn_internal.h"

namespace ttn {

__device__ __forceinline__ double load_coord(const CoordSource& src, int64_t p, int c) {
  if (src.digits) return 0.0; // index-setting mode: no coordinates
  if (src.grid) {
    // grid_points (src/IndexMaps/realindexmap.jl:78-86): x = i * (a / b^L); Cartesian product over
    // the coordinate slots, slot 0 slowest.
    int64_t idx = src.first + p;
    int64_t stride = 1;
    for (int cc = src.n_coords - 1; cc > c; --cc) stride *= src.count[cc];
    int64_t i = (idx / stride) % src.count[c];
    return __dmul_rn((double)i, src.step[c]);
  }
  return src.layout == TTN_LAYOUT_AOS ? __ldg(src.coords + p * src.n_coords + c)
                                      : __ldg(src.coords + (int64_t)c * src.npts + p);
}

// Returns the chosen value and updates x (x_rn).  thr[0] == 0, so the loop ends at v == 0 for
// every x >= 0.
__device__ __forceinline__ int greedy_digit(double& x, const double* __restrict__ thr, int base) {
  int v = base - 1;
  double t = __ldg(thr + v);
  while (v > 0 && !(x >= t)) {
    --v;
    t = __ldg(thr + v);
  }
  x = __dsub_rn(x, t);
  return v;
}

// Index-setting mode (the inner loop of TCI): the caller supplies the value of every
// site index; out-of-range values are flagged (bit 1 of *err) instead of indexing past a tensor.
__device__ __forceinline__ int given_digit(const CoordSource& src, int64_t p, int n_sites, int site, int base, int* err) {
  if (p >= src.npts) return 0; // padding points of the last tile
  int v = src.digits[p * n_sites + site];
  if (v >= base) {
    atomicOr(err, 2);
    v = 0;
  }
  return v;
}

// Fused quadrature functional of one evaluated point.
__device__ __forceinline__ void accumulate_point(const CoordSource& src, int64_t p, double vr, double vi, double& sr,
                                                 double& si) {
  if (src.reduce_mode == TTN_REDUCE_ABS2) {
    sr = fma(vr, vr, sr);
    sr = fma(vi, vi, sr);
  } else if (src.reduce_mode == TTN_REDUCE_WEIGHTED) {
    const double w = __ldg(src.weights + p);
    sr = fma(w, vr, sr);
    si = fma(w, vi, si);
  } else {
    sr += vr;
    si += vi;
  }
}

// Domain check: the reference never terminates for x < 0 or NaN; we flag it.
__device__ __forceinline__ bool coord_in_domain(double x) { return x >= 0.0; }

} // namespace ttn
