// TTN_ACCURACY_REFINED — the opt-in pass that makes the 1e-12 bar hold at the MAXIMUM, not only at p99.9.
//
// Why a second pass and not a more careful first one: an FP64 leaf-to-root contraction carries an absolute
// error of a few ulp of the intermediate norms (~1e-15 * rms(f)) whatever the evaluation order — the CPU
// restatement of the reference's own belief-propagation arithmetic has the same tail (Accuracy).  In
// the floored relative metric of SURVEY 8(d), |v - ref| / max(|ref|, 1e-3 rms), that only exceeds 1e-12 where
// |f| << rms(f), i.e. where the contraction cancels.  So after the FP64 kernels have written the values of a
// chunk:
//   1. refine_stats_kernel    sum |f|^2 over the chunk (deterministic per-CTA partials)
//   2. refine_select_kernel   compacts the indices of the points with |f| < tau * rms (tau = 0.02: ~1.6 % of
//                             the points of a random network) — no host round trip, the count stays on the GPU
//   3. refine_dd_kernel       re-evaluates exactly those points in double-double arithmetic (106-bit
//                             significand: error-free two_prod via FMA + two_sum), one CTA per point, threads
//                             across the output elements of each pairwise contraction, and overwrites them
//   4. refine_reduce_kernel   the fused functionals (TTN_REDUCE_*) are then accumulated from the final values
// The digits of a re-evaluated point come from the same greedy_digit as everywhere else, the contraction order
// is the flop rule's (children in ascending id), for any tree / base / degree, real or complex.
// Reference path this restates: src/itensornetworkfunction.jl:84-106 (project + scalar).
#include <algorithm>

#include "k_digits.cuh"

namespace ttn {

// ---- double-double helpers (explicit rounding intrinsics: nothing here may be contracted or reassociated)
struct dd {
  double hi, lo;
};
__device__ __forceinline__ void dd_mac_d(dd& acc, const dd m, const double s) {  // acc += m * s  (s exact double)
  const double p = __dmul_rn(m.hi, s);
  double e = __fma_rn(m.hi, s, -p);
  e = __fma_rn(m.lo, s, e);
  const double t = __dadd_rn(acc.hi, p);
  const double bb = __dsub_rn(t, acc.hi);
  const double err = __dadd_rn(__dsub_rn(acc.hi, __dsub_rn(t, bb)), __dsub_rn(p, bb));
  acc.lo = __dadd_rn(acc.lo, __dadd_rn(err, e));
  acc.hi = t;
}
__device__ __forceinline__ void dd_mac(dd& acc, const dd m, const dd s) {  // acc += m * s
  const double p = __dmul_rn(m.hi, s.hi);
  double e = __fma_rn(m.hi, s.hi, -p);
  e = __fma_rn(m.hi, s.lo, e);
  e = __fma_rn(m.lo, s.hi, e);
  const double t = __dadd_rn(acc.hi, p);
  const double bb = __dsub_rn(t, acc.hi);
  const double err = __dadd_rn(__dsub_rn(acc.hi, __dsub_rn(t, bb)), __dsub_rn(p, bb));
  acc.lo = __dadd_rn(acc.lo, __dadd_rn(err, e));
  acc.hi = t;
}
__device__ __forceinline__ dd dd_norm(const dd a) {  // fast two-sum renormalisation
  const double s = __dadd_rn(a.hi, a.lo);
  return dd{s, __dsub_rn(a.lo, __dsub_rn(s, a.hi))};
}
__device__ __forceinline__ dd dd_neg(const dd a) { return dd{-a.hi, -a.lo}; }

// ---- 1. sum |f|^2 of the chunk
__global__ void __launch_bounds__(256) refine_stats_kernel(const double* __restrict__ out, int64_t n_doubles,
                                                           double* __restrict__ partial) {
  double a = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_doubles; i += (int64_t)gridDim.x * 256) {
    const double v = out[i];
    a = fma(v, v, a);
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// ---- 2. compaction of the points to re-evaluate.  d_sel[0] = count (zeroed by the launcher), d_sel[1 + i] = point
__global__ void __launch_bounds__(256) refine_select_kernel(const double* __restrict__ out, int64_t npts, int nc,
                                                            const double* __restrict__ partial, int n_partial,
                                                            double tau, int32_t* __restrict__ d_sel) {
  __shared__ double s_thr2;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_partial; ++i) s += partial[i];  // fixed order: every CTA gets the same threshold
    s_thr2 = tau * tau * (s / (double)npts);
  }
  __syncthreads();
  const double thr2 = s_thr2;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < npts; p += (int64_t)gridDim.x * 256) {
    double m2 = out[p * nc] * out[p * nc];
    if (nc == 2) m2 = fma(out[2 * p + 1], out[2 * p + 1], m2);
    if (!(m2 >= thr2)) {  // also catches NaN
      const int32_t k = atomicAdd(d_sel, 1);
      d_sel[1 + k] = (int32_t)p;
    }
  }
}

// ---- 3. double-double re-evaluation, one CTA per selected point
// Workspace per CTA (global, L1/L2 resident): messages [msg_total] and two intermediates [max_inter], each
// element a dd (complex: two dd), plus the slice index of every vertex.
template <bool CPLX>
__global__ void __launch_bounds__(128)
    refine_dd_kernel(TreeDev t, DigitTable dg, CoordSource src, const int32_t* __restrict__ d_sel,
                     double* __restrict__ out, dd* __restrict__ work, int64_t work_per_cta, int32_t* __restrict__ wsl,
                     int* err) {
  constexpr int NC = CPLX ? 2 : 1;
  const int G = blockDim.x, tid = threadIdx.x;
  dd* msgs = work + (int64_t)blockIdx.x * work_per_cta;
  dd* bufA = msgs + (int64_t)NC * t.msg_total;
  dd* bufB = bufA + (int64_t)NC * t.max_inter;
  int32_t* sl = wsl + (int64_t)blockIdx.x * t.n_vertices;
  const int n_sel = d_sel[0];
  for (int k = blockIdx.x; k < n_sel; k += gridDim.x) {
    const int64_t p = d_sel[1 + k];
    // K1 (same greedy loop, same tables): one thread per coordinate slot; slots touch disjoint digits but may
    // share a vertex (Real + Imag index on one vertex), hence the shared-memory-free atomicAdd on sl[]
    for (int v = tid; v < t.n_vertices; v += G) sl[v] = 0;
    __syncthreads();
    for (int c = tid; c < dg.n_coords; c += G) {
      double x = load_coord(src, p, c);
      if (!coord_in_domain(x)) x = 0.0;  // already flagged by the FP64 pass
      for (int e_i = dg.coord_ptr[c]; e_i < dg.coord_ptr[c + 1]; ++e_i) {
        const DigitEntry e = dg.entries[e_i];
        const int v = src.digits ? given_digit(src, p, dg.n_sites, e.site, e.base, err)
                                 : greedy_digit(x, dg.thr + e.thr_off, e.base);
        atomicAdd(sl + e.vertex, v * e.stride);
      }
    }
    __syncthreads();
    for (int oi = 0; oi < t.n_vertices; ++oi) {
      const int v = t.post[oi];
      const int64_t ssize = t.slice_size[v];
      const double* S = t.tensors + (t.tensor_off[v] + (int64_t)sl[v] * ssize) * NC;
      dd* mv = msgs + (int64_t)NC * t.msg_off[v];
      const int c0 = t.child_ptr[v], c1 = t.child_ptr[v + 1];
      if (c0 == c1) {
        for (int64_t i = tid; i < ssize * NC; i += G) mv[i] = dd{__ldg(S + i), 0.0};
      } else {
        const dd* cur = nullptr;  // null: the (exact, double) tensor slice
        int64_t rest = ssize;
        dd* nxt = bufA;
        for (int ci = c0; ci < c1; ++ci) {
          const int c = t.child[ci];
          const int na = t.link_dim[c];
          rest /= na;
          const dd* m = msgs + (int64_t)NC * t.msg_off[c];
          dd* dst = (ci == c1 - 1) ? mv : nxt;
          for (int64_t i = tid; i < rest; i += G) {
            dd ar{0.0, 0.0}, ai{0.0, 0.0};
            for (int a = 0; a < na; ++a) {
              const int64_t e = (int64_t)a * rest + i;
              if (CPLX) {
                const dd mr = m[2 * a], mi = m[2 * a + 1];
                if (cur) {
                  const dd sr = cur[2 * e], si = cur[2 * e + 1];
                  dd_mac(ar, mr, sr);
                  dd_mac(ar, dd_neg(mi), si);
                  dd_mac(ai, mr, si);
                  dd_mac(ai, mi, sr);
                } else {
                  const double sr = __ldg(S + 2 * e), si = __ldg(S + 2 * e + 1);
                  dd_mac_d(ar, mr, sr);
                  dd_mac_d(ar, dd_neg(mi), si);
                  dd_mac_d(ai, mr, si);
                  dd_mac_d(ai, mi, sr);
                }
              } else {
                if (cur) dd_mac(ar, m[a], cur[e]);
                else dd_mac_d(ar, m[a], __ldg(S + e));
              }
            }
            if (CPLX) {
              dst[2 * i] = dd_norm(ar);
              dst[2 * i + 1] = dd_norm(ai);
            } else {
              dst[i] = dd_norm(ar);
            }
          }
          __syncthreads();
          cur = dst;
          nxt = (dst == bufA) ? bufB : bufA;
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      const dd* r = msgs + (int64_t)NC * t.msg_off[t.root];
      if (CPLX) {
        out[2 * p] = r[0].hi;
        out[2 * p + 1] = r[1].hi;
      } else {
        out[p] = r[0].hi;
      }
    }
    __syncthreads();
  }
}

// ---- 4. fused functionals from the final values (deterministic per-CTA partials, as in the FP64 kernels)
__global__ void __launch_bounds__(256) refine_reduce_kernel(CoordSource src, const double* __restrict__ out, int nc,
                                                            double* __restrict__ partial) {
  double sr = 0.0, si = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < src.npts; p += (int64_t)gridDim.x * 256)
    accumulate_point(src, p, out[p * nc], nc == 2 ? out[2 * p + 1] : 0.0, sr, si);
  __shared__ double sh[2][256];
  sh[0][threadIdx.x] = sr;
  sh[1][threadIdx.x] = si;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = sh[0][0];
    partial[2 * blockIdx.x + 1] = sh[1][0];
  }
}

// Runs steps 1-3 (and 4 when d_partial != nullptr) on stream s for the chunk described by src, whose FP64 values
// are in d_out.  d_sel must hold src.npts + 1 int32.  *n_launches is incremented.
int launch_refine(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double tau, int32_t* d_sel,
                  double* d_partial, int* n_partial, cudaStream_t s, int* n_launches) {
  const int NC = p->info.is_complex ? 2 : 1;
  const int64_t npts = src.npts;
  *n_partial = 0;
  if (npts == 0) return TTN_OK;
  const int nb = (int)std::min<int64_t>((npts + 255) / 256, (int64_t)p->sm_count * 8);
  // the partial-sum buffer holds 3 * (8 * SMs + 8) doubles: the stats take the last third
  double* stat_partial = st.d_partial + (size_t)2 * (p->sm_count * 8 + 8);
  TTN_CUDA(cudaMemsetAsync(d_sel, 0, sizeof(int32_t), s));
  refine_stats_kernel<<<nb, 256, 0, s>>>(d_out, npts * NC, stat_partial);
  refine_select_kernel<<<nb, 256, 0, s>>>(d_out, npts, NC, stat_partial, nb, tau, d_sel);
  // double-double pass: persistent CTAs, workspace per CTA
  const TreeDev& t = p->tree;
  const int64_t work_per_cta = (int64_t)NC * (t.msg_total + 2 * t.max_inter);
  int grid = p->sm_count * 8;
  const size_t budget = (size_t)1 << 30;
  const size_t per_cta = sizeof(dd) * (size_t)work_per_cta + sizeof(int32_t) * (size_t)t.n_vertices;
  grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)grid, budget / per_cta));
  const size_t dd_bytes = (sizeof(dd) * (size_t)work_per_cta * grid + 255) / 256 * 256;
  const size_t need = dd_bytes + sizeof(int32_t) * (size_t)t.n_vertices * grid;
  if (st.refine_bytes < need) {
    if (st.d_refine) cudaFree(st.d_refine);
    st.d_refine = nullptr;
    st.refine_bytes = 0;
    TTN_CUDA(cudaMalloc(&st.d_refine, need));
    st.refine_bytes = need;
  }
  dd* work = reinterpret_cast<dd*>(st.d_refine);
  int32_t* wsl = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(st.d_refine) + dd_bytes);
  // threads per point: as many as the widest pairwise contraction has output elements (32..128)
  int64_t widest = 1;
  for (int v = 0; v < t.n_vertices; ++v) {
    const int c0 = p->child_ptr[v], c1 = p->child_ptr[v + 1];
    widest = std::max<int64_t>(widest, c0 == c1 ? p->slice_size[v] : p->slice_size[v] / p->link_dim[p->child[c0]]);
  }
  const int G = widest >= 128 ? 128 : (widest >= 64 ? 64 : 32);
  if (p->info.is_complex)
    refine_dd_kernel<true><<<grid, G, 0, s>>>(p->tree, p->digits, src, d_sel, d_out, work, work_per_cta, wsl, p->d_err);
  else
    refine_dd_kernel<false><<<grid, G, 0, s>>>(p->tree, p->digits, src, d_sel, d_out, work, work_per_cta, wsl, p->d_err);
  *n_launches += 3;
  if (d_partial) {
    refine_reduce_kernel<<<nb, 256, 0, s>>>(src, d_out, NC, d_partial);
    *n_partial = nb;
    *n_launches += 1;
  }
  TTN_CUDA(cudaGetLastError());
  return TTN_OK;
}

} // namespace ttn
