// Generic tree kernel + standalone digit kernel + partial-sum reduction.
//
// generic_kernel: any tree, any vertex degree, any site/link dimensions, real or complex.
// One thread per point; the per-point messages and intermediates live in an HBM workspace laid
// out element-major ([element][thread]) so that every access is coalesced across the warp.  It is
// the completeness path (arbitrary `uniform_tree`s, several site indices per vertex, base 3, ...);
// the chain and DMMA kernels are the fast paths for the shapes BASELINE.json names.
//
// Per vertex v (children c_1..c_k in ascending id, parent link p) it performs exactly the
// pairwise contractions of the flop rule in SURVEY §8(d):
//   R_1[c_2..c_k,p] = sum_{c_1} m_{c_1}[c_1] * S[c_1, c_2..c_k, p]   (S = digit-selected slice,
//   R_j            = sum_{c_j} m_{c_j}[c_j] * R_{j-1}[c_j, ...]        project(), itensornetworkfunction.jl:84-94)
// leaves copy their slice.  The root's message (length 1) is the value
// (scalar(tn), src/itensornetworkfunction.jl:105).
#include "k_digits.cuh"

namespace ttn {

__global__ void digits_kernel(DigitTable dg, CoordSource src, uint8_t* __restrict__ out, int* err) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= src.npts) return;
  for (int c = 0; c < dg.n_coords; ++c) {
    double x = load_coord(src, p, c);
    if (!coord_in_domain(x)) {
      atomicOr(err, 1);
      x = 0.0;
    }
    for (int k = dg.coord_ptr[c]; k < dg.coord_ptr[c + 1]; ++k) {
      DigitEntry e = dg.entries[k];
      int v = src.digits ? given_digit(src, p, dg.n_sites, e.site, e.base, err) : greedy_digit(x, dg.thr + e.thr_off, e.base);
      out[p * dg.n_sites + e.site] = (uint8_t)v;
    }
  }
}

int launch_digits(ttn_plan* p, const CoordSource& src, uint8_t* d_digits, cudaStream_t s) {
  if (src.npts == 0) return TTN_OK;
  int nt = 256;
  int64_t nb = (src.npts + nt - 1) / nt;
  digits_kernel<<<(unsigned)nb, nt, 0, s>>>(p->digits, src, d_digits, p->d_err);
  TTN_CUDA(cudaGetLastError());
  return TTN_OK;
}

template <bool CPLX>
__global__ void __launch_bounds__(256)
    generic_kernel(TreeDev t, DigitTable dg, CoordSource src, double* __restrict__ out,
                   int32_t* __restrict__ wsl, double* __restrict__ wd, int64_t T, int* err,
                   double* __restrict__ partial, int do_sum) {
  constexpr int NC = CPLX ? 2 : 1;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double sum_re = 0.0, sum_im = 0.0;
  double* msgs = wd + tid;                       // element e, component k at msgs[(NC*e+k)*T]
  double* bufA = wd + (int64_t)NC * t.msg_total * T + tid;
  double* bufB = bufA + (int64_t)NC * t.max_inter * T;
  int32_t* sl = wsl + tid;                       // slice of vertex v at sl[v*T]

  for (int64_t p = tid; p < src.npts; p += T) {
    // ---- K1: digits -> per-vertex slice index
    for (int v = 0; v < t.n_vertices; ++v) sl[(int64_t)v * T] = 0;
    for (int c = 0; c < dg.n_coords; ++c) {
      double x = load_coord(src, p, c);
      if (!coord_in_domain(x)) {
        atomicOr(err, 1);
        x = 0.0;
      }
      for (int k = dg.coord_ptr[c]; k < dg.coord_ptr[c + 1]; ++k) {
        DigitEntry e = dg.entries[k];
        int v = src.digits ? given_digit(src, p, dg.n_sites, e.site, e.base, err) : greedy_digit(x, dg.thr + e.thr_off, e.base);
        sl[(int64_t)e.vertex * T] += v * e.stride;
      }
    }
    // ---- leaf-to-root contraction
    for (int oi = 0; oi < t.n_vertices; ++oi) {
      const int v = t.post[oi];
      const int64_t ssize = t.slice_size[v];
      const double* S = t.tensors + (t.tensor_off[v] + (int64_t)sl[(int64_t)v * T] * ssize) * NC;
      double* mv = msgs + (int64_t)NC * t.msg_off[v] * T;
      const int c0 = t.child_ptr[v], c1 = t.child_ptr[v + 1];
      if (c0 == c1) {
        for (int64_t i = 0; i < ssize * NC; ++i) mv[i * T] = __ldg(S + i);
        continue;
      }
      const double* cur = S;
      int64_t cstride = 1;
      int64_t rest = ssize;
      double* nxt = bufA;
      for (int ci = c0; ci < c1; ++ci) {
        const int c = t.child[ci];
        const int na = t.link_dim[c];
        rest /= na;
        const double* m = msgs + (int64_t)NC * t.msg_off[c] * T;
        double* dst = (ci == c1 - 1) ? mv : nxt;
        for (int64_t i = 0; i < rest; ++i) {
          double ar = 0.0, ai = 0.0;
          for (int a = 0; a < na; ++a) {
            const int64_t e = (int64_t)a * rest + i;
            if (CPLX) {
              const double mr = m[(2 * a) * T], mi = m[(2 * a + 1) * T];
              const double sr = cur[(2 * e) * cstride], si = cur[(2 * e + 1) * cstride];
              ar = fma(mr, sr, ar);
              ar = fma(-mi, si, ar);
              ai = fma(mr, si, ai);
              ai = fma(mi, sr, ai);
            } else {
              ar = fma(m[a * T], cur[e * cstride], ar);
            }
          }
          if (CPLX) {
            dst[(2 * i) * T] = ar;
            dst[(2 * i + 1) * T] = ai;
          } else {
            dst[i * T] = ar;
          }
        }
        cur = dst;
        cstride = T;
        nxt = (dst == bufA) ? bufB : bufA;
      }
    }
    const double* r = msgs + (int64_t)NC * t.msg_off[t.root] * T;
    const double vr = r[0], vi = CPLX ? r[T] : 0.0;
    if (out) {
      if (CPLX) {
        out[2 * p] = vr;
        out[2 * p + 1] = vi;
      } else {
        out[p] = vr;
      }
    }
    accumulate_point(src, p, vr, vi, sum_re, sum_im);
  }
  if (do_sum) {
    __shared__ double sh[2][256];
    sh[0][threadIdx.x] = sum_re;
    sh[1][threadIdx.x] = sum_im;
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
}

int launch_generic(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out,
                   double* d_partial, int* n_partial, cudaStream_t s) {
  *n_partial = 0;
  if (src.npts == 0) return TTN_OK;
  const int NC = p->info.is_complex ? 2 : 1;
  const int nt = 256;
  const size_t per_thread = sizeof(int32_t) * (size_t)p->tree.n_vertices +
                            sizeof(double) * NC * (size_t)(p->tree.msg_total + 2 * p->tree.max_inter);
  const size_t budget = (size_t)2 << 30;
  int64_t max_threads = (int64_t)(budget / per_thread) / nt * nt;
  if (max_threads < nt) max_threads = nt;
  int64_t T = (src.npts + nt - 1) / nt * nt;
  const int64_t full = (int64_t)p->sm_count * 8 * nt;
  if (T > full) T = full;
  if (T > max_threads) T = max_threads;
  const size_t int_bytes = ((sizeof(int32_t) * (size_t)p->tree.n_vertices * T) + 255) / 256 * 256;
  const size_t need = int_bytes + sizeof(double) * NC * (size_t)(p->tree.msg_total + 2 * p->tree.max_inter) * T;
  if (st.work_bytes < need) {
    if (st.d_work) cudaFree(st.d_work);
    st.d_work = nullptr;
    st.work_bytes = 0;
    TTN_CUDA(cudaMalloc(&st.d_work, need));
    st.work_bytes = need;
  }
  int32_t* wsl = reinterpret_cast<int32_t*>(st.d_work);
  double* wd = reinterpret_cast<double*>(reinterpret_cast<char*>(st.d_work) + int_bytes);
  const int nb = (int)(T / nt);
  const int do_sum = d_partial != nullptr;
  if (p->info.is_complex)
    generic_kernel<true><<<nb, nt, 0, s>>>(p->tree, p->digits, src, d_out, wsl, wd, T, p->d_err, d_partial, do_sum);
  else
    generic_kernel<false><<<nb, nt, 0, s>>>(p->tree, p->digits, src, d_out, wsl, wd, T, p->d_err, d_partial, do_sum);
  TTN_CUDA(cudaGetLastError());
  *n_partial = do_sum ? nb : 0;
  return TTN_OK;
}

// Deterministic final reduction of per-CTA partial sums: one block, fixed-order tree.
__global__ void sum_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ sum) {
  __shared__ double sh[2][256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    a += partial[2 * i];
    b += partial[2 * i + 1];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sum[0] = sh[0][0];
    sum[1] = sh[1][0];
  }
}

int launch_sum_partials(ttn_plan* p, const double* d_partial, int n_partial, int nc, double* d_sum,
                        cudaStream_t s) {
  (void)p;
  (void)nc;
  if (n_partial <= 0) return TTN_OK;
  sum_partials_kernel<<<1, 256, 0, s>>>(d_partial, n_partial, d_sum);
  TTN_CUDA(cudaGetLastError());
  return TTN_OK;
}

} // namespace ttn
