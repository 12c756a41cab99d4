// K8 — prefix-shared evaluation of a chain on a FULL dyadic grid (BASELINE config 4:
// "evaluated on a 16384^2 grid, with a summed-grid quadrature reduction").
//
// On the full grid every digit string occurs exactly once, so instead of contracting the chain
// once per point (L*chi^2 MACs per point) the chain is EXPANDED level by level: the state after
// the first k sites is shared by all 2^(L-k) points with that digit prefix,
//     level_k[2 i + d] = level_{k-1}[i] * M_k[d],           sum_k 2^k chi^2  ~  2 * 2^L * chi^2  MACs
// i.e. ~L/2 times fewer flops (26x for config 4), no sorting, no gather/scatter, and every 8-row
// DMMA group of a level uses the same B matrix.
//
// One kernel, `grid_expand_kernel<CHI, D, NBAT, FINAL>`, expands D levels depth-first IN REGISTERS:
// a warp loads an 8*NBAT-row group of level a from HBM, walks the 2^D descendants with a stack of
// D+1 register tiles (the D fragment of a site is the A fragment of the next, B rows pre-permuted,
// as in k_chain_mma.cu) and either stores the level-(a+D) rows (intermediate passes) or applies the
// root vector and writes / accumulates the 2^(D+1) grid values per row (final pass).  The B
// fragments of the pass's D sites stay resident in shared memory.  The grid's linear point index
// is assembled from per-position weights, so any digit arrangement (interleaved, per-tooth, ...)
// and any number of coordinates works.  The roofline of this kernel is reported on the flops it
// EXECUTES (ttn_opts.flops_executed), not on the per-point flop rule.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "k_async.cuh"
#include "k_digits.cuh"

namespace ttn {

struct GridShareArgs {
  const double* in;        // level-a rows [n_in][CHI]
  double* out_rows;        // level-(a+D) rows (intermediate passes)
  double* out_vals;        // grid values (final pass), may be null
  const double* frags;     // B fragments of the pass's first site: [D][2][CHI*CHI]
  const double* root;      // [nout][2][CHI] (final pass)
  int64_t n_in;            // rows of level a
  int32_t a_bits;          // number of digit bits encoded in a level-a row index
  int32_t nout;
  int64_t w_prefix[64];    // grid-index weight of the j-th most significant bit of a level-a row index
  int64_t w_level[8];      // weights of the D sites of this pass
  int64_t w_root;          // weight of the root digit
  CoordSource src;         // reduce mode / weights
};

template <int CHI, int NBAT>
struct Tile {
  double v[NBAT][CHI / 4];
};

template <int CHI, int NBAT>
__device__ __forceinline__ void tile_mma(Tile<CHI, NBAT>& dst, const Tile<CHI, NBAT>& a, uint32_t bb) {
  constexpr int NB = CHI / 8, KB = CHI / 4;
#pragma unroll
  for (int b = 0; b < NBAT; ++b)
#pragma unroll
    for (int j = 0; j < KB; ++j) dst.v[b][j] = 0.0;
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
    double bf[NB];
#pragma unroll
    for (int nbp = 0; nbp < NB; ++nbp) bf[nbp] = lds64(bb + (uint32_t)((kb * NB + nbp) * 32) * 8u);
#pragma unroll
    for (int nbp = 0; nbp < NB; ++nbp)
#pragma unroll
      for (int b = 0; b < NBAT; ++b) dmma884(dst.v[b][2 * nbp], dst.v[b][2 * nbp + 1], a.v[b][kb], bf[nbp]);
  }
}

template <int CHI, int D, int NBAT, bool FINAL, int J>
struct Expand {
  __device__ __forceinline__ static void run(const GridShareArgs& A, const Tile<CHI, NBAT>& tin, uint32_t bsm,
                                             const int64_t (&row)[NBAT], const int64_t (&gidx)[NBAT], int g, int tq,
                                             double& sr, double& si, const double* sroot) {
#pragma unroll 1
    for (int d = 0; d < 2; ++d) {
      Tile<CHI, NBAT> tout;
      tile_mma<CHI, NBAT>(tout, tin, bsm + (uint32_t)((J * 2 + d) * CHI * CHI) * 8u);
      int64_t row2[NBAT], gidx2[NBAT];
#pragma unroll
      for (int b = 0; b < NBAT; ++b) {
        row2[b] = row[b] * 2 + d;
        gidx2[b] = gidx[b] + (d ? A.w_level[J] : 0);
      }
      Expand<CHI, D, NBAT, FINAL, J + 1>::run(A, tout, bsm, row2, gidx2, g, tq, sr, si, sroot);
    }
  }
};

template <int CHI, int D, int NBAT, bool FINAL>
struct Expand<CHI, D, NBAT, FINAL, D> {
  __device__ __forceinline__ static void run(const GridShareArgs& A, const Tile<CHI, NBAT>& t, uint32_t,
                                             const int64_t (&row)[NBAT], const int64_t (&gidx)[NBAT], int g, int tq,
                                             double& sr, double& si, const double* sroot) {
    constexpr int NB = CHI / 8;
    if (!FINAL) {
#pragma unroll
      for (int b = 0; b < NBAT; ++b) {
        if (row[b] >= 0) {
          double* dst = A.out_rows + (size_t)row[b] * CHI + 2 * tq;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb)
            *reinterpret_cast<double2*>(dst + 8 * nb) = make_double2(t.v[b][2 * nb], t.v[b][2 * nb + 1]);
        }
      }
    } else {
      // root: value(d_r) = row . R[d_r]; the lane holds columns 8nb+2tq, +1 of row g
#pragma unroll
      for (int b = 0; b < NBAT; ++b) {
#pragma unroll 1
        for (int dr = 0; dr < 2; ++dr) {
          double o0 = 0.0, o1 = 0.0;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int col = 8 * nb + 2 * tq;
            o0 = fma(t.v[b][2 * nb], sroot[dr * CHI + col], o0);
            o0 = fma(t.v[b][2 * nb + 1], sroot[dr * CHI + col + 1], o0);
            if (A.nout == 2) {
              o1 = fma(t.v[b][2 * nb], sroot[(2 + dr) * CHI + col], o1);
              o1 = fma(t.v[b][2 * nb + 1], sroot[(2 + dr) * CHI + col + 1], o1);
            }
          }
          o0 += __shfl_xor_sync(0xffffffffu, o0, 1);
          o0 += __shfl_xor_sync(0xffffffffu, o0, 2);
          if (A.nout == 2) {
            o1 += __shfl_xor_sync(0xffffffffu, o1, 1);
            o1 += __shfl_xor_sync(0xffffffffu, o1, 2);
          }
          if (tq == 0 && row[b] >= 0) {
            const int64_t p = gidx[b] + (dr ? A.w_root : 0);
            if (A.out_vals) {
              if (A.nout == 2) reinterpret_cast<double2*>(A.out_vals)[p] = make_double2(o0, o1);
              else A.out_vals[p] = o0;
            }
            accumulate_point(A.src, p, o0, o1, sr, si);
          }
        }
      }
    }
  }
};

template <int CHI, int D, int NBAT, bool FINAL>
__global__ void __launch_bounds__(256, 1)
    grid_expand_kernel(GridShareArgs A, double* __restrict__ partial, int do_sum) {
  constexpr int NB = CHI / 8;
  extern __shared__ __align__(128) unsigned char smem[];
  double* sB = reinterpret_cast<double*>(smem);                 // [D][2][CHI*CHI]
  double* sroot = sB + (size_t)(D > 0 ? D : 1) * 2 * CHI * CHI; // [nout][2][CHI]
  __shared__ double red[2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < D * 2 * CHI * CHI; i += 256) sB[i] = A.frags[i];
  if (FINAL)
    for (int i = tid; i < A.nout * 2 * CHI; i += 256) sroot[i] = A.root[i];
  __syncthreads();
  const uint32_t bsm = smem_u32(sB) + (uint32_t)lane * 8u;
  const int g = lane >> 2, tq = lane & 3;
  double sr = 0.0, si = 0.0;
  const int64_t n_groups = (A.n_in + 8 * NBAT - 1) / (8 * NBAT);
  for (int64_t grp = (int64_t)blockIdx.x * 8 + warp; grp < n_groups; grp += (int64_t)gridDim.x * 8) {
    Tile<CHI, NBAT> t0;
    int64_t row[NBAT], gidx[NBAT];
#pragma unroll
    for (int b = 0; b < NBAT; ++b) {
      const int64_t r = (grp * NBAT + b) * 8 + g;
      const bool valid = r < A.n_in;
      row[b] = valid ? r : -((int64_t)1 << 40); // stays negative through the D doublings: never stored
      gidx[b] = 0;
      const double* srcp = A.in + (size_t)(valid ? r : 0) * CHI + 2 * tq;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const double2 v = valid ? *reinterpret_cast<const double2*>(srcp + 8 * nb) : make_double2(0.0, 0.0);
        t0.v[b][2 * nb] = v.x;
        t0.v[b][2 * nb + 1] = v.y;
      }
      if (FINAL && valid) {
        int64_t acc = 0;
        for (int j = 0; j < A.a_bits; ++j)
          if ((r >> (A.a_bits - 1 - j)) & 1) acc += A.w_prefix[j];
        gidx[b] = acc;
      }
    }
    Expand<CHI, D, NBAT, FINAL, 0>::run(A, t0, bsm, row, gidx, g, tq, sr, si, sroot);
  }
  if (FINAL && do_sum) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_down_sync(0xffffffffu, sr, o);
      si += __shfl_down_sync(0xffffffffu, si, o);
    }
    if (lane == 0) {
      red[0][warp] = sr;
      red[1][warp] = si;
    }
    __syncthreads();
    if (tid == 0) {
      double x = 0.0, y = 0.0;
      for (int w = 0; w < 8; ++w) {
        x += red[0][w];
        y += red[1][w];
      }
      partial[2 * blockIdx.x] = x;
      partial[2 * blockIdx.x + 1] = y;
    }
  }
}

// ------------------------------------------------------------------------------ host side

// Eligibility + per-position weights.  Requires the DMMA chain images (build_chain_mma) and: every
// vertex carries exactly one binary site index whose thresholds are exactly 2^-k with digit numbers
// 1..L_c per coordinate slot.
int build_grid_share(ttn_plan* p, const ttn_desc* d) {
  p->gshare_ok = false;
  if (!p->cmma_ok || !p->cmma_plain_ok || p->cmma_plain.nsl != 2 || !p->all_base2 || d->n_vertices < 2) return TTN_OK;
  const int n = d->n_vertices;
  for (int v = 0; v < n; ++v)
    if (p->nslices[v] != 2 || d->site_ptr[v + 1] - d->site_ptr[v] != 1) return TTN_OK;
  if (n > 62) return TTN_OK;
  std::vector<int> Lc(d->n_coords, 0);
  for (int s = 0; s < d->n_sites; ++s) Lc[d->site_coord[s]] = std::max(Lc[d->site_coord[s]], d->site_digit[s]);
  {
    std::vector<int> cnt(d->n_coords, 0);
    for (int s = 0; s < d->n_sites; ++s) {
      const int c = d->site_coord[s], k = d->site_digit[s];
      if (k < 1 || d->thr[d->thr_ptr[s] + 1] != std::ldexp(1.0, -k)) return TTN_OK;
      cnt[c]++;
    }
    for (int c = 0; c < d->n_coords; ++c)
      if (cnt[c] != Lc[c] || Lc[c] < 1 || Lc[c] > 40) return TTN_OK;
  }
  // chain order (position 0 = leaf ... n-1 = root) and per-position (coordinate, digit)
  std::vector<int> order(n);
  {
    int v = d->root;
    for (int pos = n - 1; pos >= 0; --pos) {
      order[pos] = v;
      if (pos > 0) v = p->child[p->child_ptr[v]];
    }
  }
  p->gs_coord.assign(n, 0);
  p->gs_digit.assign(n, 0);
  for (int pos = 0; pos < n; ++pos) {
    const int s = d->site_ptr[order[pos]];
    p->gs_coord[pos] = d->site_coord[s];
    p->gs_digit[pos] = d->site_digit[s];
  }
  p->gs_L = Lc;
  p->gshare_ok = true;
  return TTN_OK;
}

bool grid_share_applicable(const ttn_plan* p, const CoordSource& src) {
  if (!p->gshare_ok || !src.grid || src.first != 0) return false;
  int64_t total = 1;
  for (int c = 0; c < src.n_coords; ++c) {
    if (src.count[c] != ((int64_t)1 << p->gs_L[c]) || src.step[c] != std::ldexp(1.0, -p->gs_L[c])) return false;
    total *= src.count[c];
  }
  if (src.npts != total) return false;
  const int n = p->cmma_plain.n_vertices;
  if (n - 1 > 40) return false; // 2^40 points and more: not a realistic dense grid
  return true;
}

template <int CHI, int D, int NBAT, bool FINAL>
static int launch_expand(const GridShareArgs& A, int sm_count, double* d_partial, int do_sum, int* grid_out, cudaStream_t s) {
  const size_t smem = ((size_t)std::max(D, 1) * 2 * CHI * CHI + 4 * CHI) * 8;
  auto kern = grid_expand_kernel<CHI, D, NBAT, FINAL>;
  TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_groups = (A.n_in + 8 * NBAT - 1) / (8 * NBAT);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_groups + 7) / 8, (int64_t)sm_count));
  kern<<<grid, 256, smem, s>>>(A, d_partial, do_sum);
  TTN_CUDA(cudaGetLastError());
  *grid_out = grid;
  return TTN_OK;
}

template <int CHI, int NBAT>
static int launch_expand_d(int D, bool final_, const GridShareArgs& A, int sm_count, double* d_partial, int do_sum,
                           int* grid_out, cudaStream_t s) {
#define TTN_GS_CASE(DD)                                                                                         \
  case DD:                                                                                                      \
    return final_ ? launch_expand<CHI, DD, NBAT, true>(A, sm_count, d_partial, do_sum, grid_out, s)             \
                  : launch_expand<CHI, DD, NBAT, false>(A, sm_count, d_partial, do_sum, grid_out, s);
  switch (D) {
    TTN_GS_CASE(0)
    TTN_GS_CASE(1)
    TTN_GS_CASE(2)
    TTN_GS_CASE(3)
    TTN_GS_CASE(4)
    TTN_GS_CASE(5)
  }
#undef TTN_GS_CASE
  set_error("grid-share kernel: bad pass depth");
  return TTN_ERR_UNSUPPORTED;
}

int launch_grid_share(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double* d_partial,
                      int* n_partial, cudaStream_t s, int* n_launches, double* flops_executed) {
  *n_partial = 0;
  const ChainMmaDev& c = p->cmma_plain;
  const int n = c.n_vertices, CHI = c.chi;
  const int n_mid = n - 2;
  constexpr int DMAX = 5;
  // passes over the middle sites: a short first pass, then full passes of DMAX; the last one is fused
  // with the root
  std::vector<int> passes;
  {
    int rem = n_mid;
    const int first = rem % DMAX;
    if (first > 0 || rem == 0) passes.push_back(first);
    rem -= first;
    while (rem > 0) {
      passes.push_back(DMAX);
      rem -= DMAX;
    }
  }
  // weights of the chain positions in the grid's linear point index (slot 0 slowest)
  std::vector<int64_t> w(n);
  {
    std::vector<int64_t> stride(src.n_coords, 1);
    for (int cc = src.n_coords - 2; cc >= 0; --cc) stride[cc] = stride[cc + 1] * src.count[cc + 1];
    for (int pos = 0; pos < n; ++pos)
      w[pos] = stride[p->gs_coord[pos]] << (p->gs_L[p->gs_coord[pos]] - p->gs_digit[pos]);
  }
  // ping-pong level buffers: the largest stored level is the input of the last pass
  const int last_D = passes.back();
  const int64_t max_rows = (int64_t)1 << (1 + n_mid - last_D);
  const size_t buf_b = (size_t)std::max<int64_t>(max_rows, 16) * CHI * 8;
  if (2 * buf_b > ((size_t)24 << 30)) {
    set_error("grid-share kernel: level buffers would exceed 24 GB");
    return TTN_ERR_UNSUPPORTED;
  }
  if (st.gemm_bytes < 2 * buf_b) {
    if (st.d_gemm) cudaFree(st.d_gemm);
    st.d_gemm = nullptr;
    st.gemm_bytes = 0;
    TTN_CUDA(cudaMalloc(&st.d_gemm, 2 * buf_b));
    st.gemm_bytes = 2 * buf_b;
  }
  double* bufA = reinterpret_cast<double*>(st.d_gemm);
  double* bufB = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(st.d_gemm) + buf_b);
  // level 0: the two leaf slices
  TTN_CUDA(cudaMemcpyAsync(bufA, c.leaf, sizeof(double) * 2 * CHI, cudaMemcpyDeviceToDevice, s));
  int64_t rows = 2;
  int bits = 1, site = 0;
  double flops = 0.0;
  const int do_sum = d_partial != nullptr;
  double *in = bufA, *outb = bufB;
  for (size_t pi = 0; pi < passes.size(); ++pi) {
    const int D = passes[pi];
    const bool final_ = pi + 1 == passes.size();
    GridShareArgs A{};
    A.in = in;
    A.out_rows = outb;
    A.out_vals = final_ ? d_out : nullptr;
    A.frags = c.frags + (size_t)site * 2 * CHI * CHI;
    A.root = c.root;
    A.n_in = rows;
    A.a_bits = bits;
    A.nout = c.nout;
    for (int j = 0; j < bits; ++j) A.w_prefix[j] = w[j];
    for (int j = 0; j < D; ++j) A.w_level[j] = w[bits + j];
    A.w_root = w[n - 1];
    A.src = src;
    int grid = 0, rc;
    if (CHI == 32) rc = launch_expand_d<32, 1>(D, final_, A, p->sm_count, d_partial, do_sum, &grid, s);
    else if (CHI == 16) rc = launch_expand_d<16, 2>(D, final_, A, p->sm_count, d_partial, do_sum, &grid, s);
    else rc = launch_expand_d<8, 2>(D, final_, A, p->sm_count, d_partial, do_sum, &grid, s);
    if (rc) return rc;
    *n_launches += 1;
    for (int j = 1; j <= D; ++j) flops += 2.0 * CHI * CHI * (double)(rows << j);
    if (final_) {
      flops += 2.0 * CHI * c.nout * (double)(rows << (D + 1));
      *n_partial = do_sum ? grid : 0;
    }
    rows <<= D;
    bits += D;
    site += D;
    std::swap(in, outb);
  }
  *flops_executed = flops;
  return TTN_OK;
}

} // namespace ttn
