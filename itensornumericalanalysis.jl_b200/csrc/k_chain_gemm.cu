// K6 — wide chains (real-embedded row width 64 / 128 / 256, e.g. BASELINE config 5: complex chi = 128):
// the per-site, per-slice product over digit-grouped points is a real dense GEMM.
//
// A chunk of PC points is processed site by site with the state S[PC][W] (doubles) in HBM/L2
// (ping-pong buffers).  Per chunk:
//   gemm_digits_kernel   K1 for every point -> slice index per chain position, slices[pos][p]
//   gemm_classify_kernel per site: counting sort of the chunk's points by selected slice ->
//                        row lists + class offsets + tile offsets (device side, no host sync)
//   gemm_leaf_kernel     S[p] = L[d_0(p)]
//   gemm_site_kernel     per site: for every class d, S_out[rows_d] = S_in[rows_d] * E_d  (W x W),
//                        a gathered/scattered FP64 GEMM on DMMA.8x8x4: CTA tile 128 x 128, BK = 16,
//                        3-stage cp.async pipeline, warp tile 32 x 64 (32 accumulator fragments),
//                        E_d pre-permuted on the host into B-fragment order (conflict-free LDS.64)
//   gemm_root_kernel     out[p] = S[p] . R[d_root(p)]
// Complex networks are embedded as real ones of twice the width, M -> [[Re, Im], [-Im, Re]].
// Arithmetic intensity per site: 2W flop per 16 B of state traffic (W = 256: 32 flop/B), far above
// the ridge, so the state living in HBM does not bound it; the FP64 tensor pipe does.
#include <algorithm>
#include <cstring>

#include "k_async.cuh"
#include "k_digits.cuh"

namespace ttn {

constexpr int GBM = 128, GBN = 128, GBK = 16, GSTAGES = 3;
constexpr int GA_STRIDE = GBK + 4; // doubles per A row in shared memory (160 B: conflict-free fragment loads)
constexpr int kGemmMaxCls = 16;     // slices per stream position (pair-merged vertices: 4 x 4)
constexpr int kGemmOff = kGemmMaxCls + 1; // per-site stride of the class / tile offset arrays

__global__ void gemm_digits_kernel(DigitTable dg, CoordSource src, int64_t p0, int pc, const int32_t* __restrict__ pos_of_vertex,
                                   int n_pos, uint8_t* __restrict__ slices, int* err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pc) return;
  const int64_t p = p0 + i;
  // slices[pos][i]; positions without a site index stay 0
  for (int pos = 0; pos < n_pos; ++pos) slices[(size_t)pos * pc + i] = 0;
  if (p >= src.npts) return;
  for (int c = 0; c < dg.n_coords; ++c) {
    double x = load_coord(src, p, c);
    if (!coord_in_domain(x)) {
      atomicOr(err, 1);
      x = 0.0;
    }
    for (int k = dg.coord_ptr[c]; k < dg.coord_ptr[c + 1]; ++k) {
      const DigitEntry e = dg.entries[k];
      const int v = src.digits ? given_digit(src, p, dg.n_sites, e.site, e.base, err) : greedy_digit(x, dg.thr + e.thr_off, e.base);
      const int pv = pos_of_vertex[e.vertex]; // stream position | (bit shift inside the position) << 16
      slices[(size_t)(pv & 0xffff) * pc + i] += (uint8_t)((v * e.stride) << (pv >> 16));
    }
  }
}

// one block per middle site: counting sort of the chunk by slice
__global__ void __launch_bounds__(1024)
    gemm_classify_kernel(const uint8_t* __restrict__ slices, int pc, int nsl, uint32_t* __restrict__ lists,
                         int* __restrict__ cls_off, int* __restrict__ tile_off, int n_nblk) {
  const int site = blockIdx.x; // middle site index t (position t + 1)
  const uint8_t* sl = slices + (size_t)(site + 1) * pc;
  uint32_t* list = lists + (size_t)site * pc;
  __shared__ int cnt[kGemmMaxCls], cursor[kGemmMaxCls];
  if (threadIdx.x < kGemmMaxCls) cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < pc; i += blockDim.x) atomicAdd(&cnt[sl[i]], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0, trun = 0;
    for (int c = 0; c < nsl; ++c) {
      cursor[c] = run;
      cls_off[site * kGemmOff + c] = run;
      tile_off[site * kGemmOff + c] = trun;
      run += cnt[c];
      trun += (cnt[c] + GBM - 1) / GBM * n_nblk;
    }
    cls_off[site * kGemmOff + nsl] = run;
    tile_off[site * kGemmOff + nsl] = trun;
  }
  __syncthreads();
  // order inside a class is irrelevant for the result of a row (each row is an independent product)
  for (int i = threadIdx.x; i < pc; i += blockDim.x) {
    const int pos = atomicAdd(&cursor[sl[i]], 1);
    list[pos] = (uint32_t)i;
  }
}

// S[p] = L[idx(p)], idx = sum_q slices[q][p] * nsl^q over the first `npos` stream positions (npos = 1:
// the plain leaf vectors; npos > 1: the leaf table of build_gemm_tables)
__global__ void gemm_leaf_kernel(const uint8_t* __restrict__ slices, int pc, const double* __restrict__ leaf, int W,
                                 double* __restrict__ S, int npos, int nsl) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // one double2 per thread
  const int per_row = W / 2;
  if (idx >= (int64_t)pc * per_row) return;
  const int i = (int)(idx / per_row), j = (int)(idx % per_row);
  uint32_t t = 0;
  for (int q = npos - 1; q >= 0; --q) t = t * (uint32_t)nsl + slices[(size_t)q * pc + i];
  const double2 v = *reinterpret_cast<const double2*>(leaf + (size_t)t * W + 2 * j);
  reinterpret_cast<double2*>(S)[idx] = v;
}

// enumerated digit settings for the table builds: slices[q][i] = (i / nsl^q) % nsl
__global__ void gemm_enum_kernel(uint8_t* __restrict__ slices, int pc, int npos, int nsl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pc) return;
  uint32_t r = (uint32_t)i;
  for (int q = 0; q < npos; ++q) {
    slices[(size_t)q * pc + i] = (uint8_t)(r % (uint32_t)nsl);
    r /= (uint32_t)nsl;
  }
}

// out[p] = S[p] . R[idx(p)], idx = sum_q slices[last - q][p] * nsl^q over the last `npos` stream positions
// (the root position is the lowest digit); `rows` = number of rows of one output's table
__global__ void gemm_root_kernel(const uint8_t* __restrict__ slices_last, int pc, int64_t p0, int64_t npts,
                                 const double* __restrict__ root, int W, int nsl, int npos, int64_t rows, int nout,
                                 int n_vertices,
                                 const double* __restrict__ S, double* __restrict__ out, double* __restrict__ partial,
                                 int do_sum, CoordSource src) {
  // one warp per point
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  double o0 = 0.0, o1 = 0.0;
  bool live = false;
  if (warp < pc && p0 + warp < npts) {
    live = true;
    const double* row = S + (size_t)warp * W;
    if (n_vertices > 1) {
      uint32_t t = 0;
      for (int q = npos - 1; q >= 0; --q) t = t * (uint32_t)nsl + slices_last[warp - (int64_t)q * pc];
      const double* R0 = root + (size_t)t * W;
      const double* R1 = R0 + (size_t)rows * W;
      for (int j = lane; j < W; j += 32) {
        o0 = fma(row[j], __ldg(R0 + j), o0);
        if (nout == 2) o1 = fma(row[j], __ldg(R1 + j), o1);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        o0 += __shfl_down_sync(0xffffffffu, o0, o);
        o1 += __shfl_down_sync(0xffffffffu, o1, o);
      }
    } else {
      o0 = row[0];
      o1 = nout == 2 ? row[W / 2] : 0.0;
    }
    if (lane == 0 && out) {
      if (nout == 2) reinterpret_cast<double2*>(out)[p0 + warp] = make_double2(o0, o1);
      else out[p0 + warp] = o0;
    }
  }
  if (do_sum) {
    __shared__ double sh[2][8];
    const int w = threadIdx.x >> 5;
    if (lane == 0) {
      double a0 = 0.0, a1 = 0.0;
      if (live) accumulate_point(src, p0 + warp, o0, o1, a0, a1);
      sh[0][w] = a0;
      sh[1][w] = a1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, b = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
        a += sh[0][k];
        b += sh[1][k];
      }
      partial[2 * blockIdx.x] = a;
      partial[2 * blockIdx.x + 1] = b;
    }
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// S_out[rows of class d] = S_in[rows of class d] * E_d for one site; E in B-fragment order
// frags[d][kb][nb][lane] = E_d[4 kb + (lane & 3)][8 nb + (lane >> 2)].
template <int W>
__global__ void __launch_bounds__(256, 1)
    gemm_site_kernel(const double* __restrict__ Sin, double* __restrict__ Sout, const uint32_t* __restrict__ list,
                     const int* __restrict__ cls_off, const int* __restrict__ tile_off, const double* __restrict__ frags,
                     int nsl) {
  constexpr int NNB = W / GBN;  // n-blocks per row block
  constexpr int NBW = W / 8;    // 8-column blocks per full row
  extern __shared__ __align__(128) unsigned char smem[];
  double* As = reinterpret_cast<double*>(smem);                                 // [GSTAGES][GBM][GA_STRIDE]
  double* Bs = As + (size_t)GSTAGES * GBM * GA_STRIDE;                          // [GSTAGES][GBK/4][GBN/8][32]
  __shared__ uint32_t rowid[GBM];

  const int tile = blockIdx.x;
  int c = -1;
  for (int k = 0; k < nsl; ++k)
    if (tile >= tile_off[k] && tile < tile_off[k + 1]) c = k;
  if (c < 0) return; // surplus CTA (the grid is an upper bound)
  const int lt = tile - tile_off[c];
  const int mblk = lt / NNB, nblk = lt % NNB;
  const int row0 = cls_off[c] + mblk * GBM, row_end = cls_off[c + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < GBM) rowid[tid] = list[min(row0 + tid, row_end - 1)];
  __syncthreads();

  const double* E = frags + (size_t)c * W * W;
  const uint32_t as_base = smem_u32(As), bs_base = smem_u32(Bs);
  constexpr int A_STAGE_B = GBM * GA_STRIDE * 8, B_STAGE_B = (GBK / 4) * (GBN / 8) * 32 * 8;

  auto load_stage = [&](int stage, int kc) {
    // A: GBM rows x GBK doubles (128 B per row = 8 chunks of 16 B): 1024 chunks, 4 per thread
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ch = tid + q * 256;
      const int r = ch >> 3, cc = ch & 7;
      cp_async16(as_base + stage * A_STAGE_B + (uint32_t)(r * GA_STRIDE + cc * 2) * 8u,
                 Sin + (size_t)rowid[r] * W + kc * GBK + cc * 2);
    }
    // B: for each of the GBK/4 k-blocks a contiguous run of GBN/8 fragments (4 KB)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kb = q; // GBK / 4 == 4
      const double* srcp = E + ((size_t)(kc * (GBK / 4) + kb) * NBW + nblk * (GBN / 8)) * 32 + tid * 2;
      cp_async16(bs_base + stage * B_STAGE_B + (uint32_t)((kb * (GBN / 8)) * 32 + tid * 2) * 8u, srcp);
    }
    cp_async_commit();
  };

  constexpr int NKC = W / GBK;
  const int wm = warp & 3, wn = warp >> 2; // 4 x 2 warps, warp tile 32 x 64
  const int g = lane >> 2, t = lane & 3;
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  load_stage(0, 0);
  if (NKC > 1) load_stage(1, 1);
  for (int kc = 0; kc < NKC; ++kc) {
    if (kc + 2 < NKC) {
      load_stage((kc + 2) % GSTAGES, kc + 2);
      cp_async_wait<2>();
    } else if (kc + 1 < NKC) {
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int st = kc % GSTAGES;
    const uint32_t a_st = as_base + st * A_STAGE_B, b_st = bs_base + st * B_STAGE_B;
#pragma unroll
    for (int k4 = 0; k4 < GBK / 4; ++k4) {
      double af[4], bf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        af[i] = lds64(a_st + (uint32_t)((wm * 32 + i * 8 + g) * GA_STRIDE + k4 * 4 + t) * 8u);
#pragma unroll
      for (int j = 0; j < 8; ++j) bf[j] = lds64(b_st + (uint32_t)((k4 * (GBN / 8) + wn * 8 + j) * 32 + lane) * 8u);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads(); // the stage is overwritten two iterations later
  }
  // epilogue: scatter the rows (quads write 64 contiguous bytes)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = wm * 32 + i * 8 + g;
    if (row0 + r < row_end) {
      double* dst = Sout + (size_t)rowid[r] * W + nblk * GBN + wn * 64 + 2 * t;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<double2*>(dst + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

// ------------------------------------------------------------------------------ host side

static int gemm_width(int w) {
  for (int o : {128, 256}) // the CTA tile is 128 columns wide
    if (w <= o) return o;
  return 0;
}

static int build_gemm_tables(ttn_plan* p, int tabL, int tabR, const double* d_frags_rev);

int build_chain_gemm(ttn_plan* p, const ttn_desc* d) {
  p->cgemm_ok = false;
  if (!p->is_chain) return TTN_OK;
  const int n = d->n_vertices;
  const bool cplx = d->is_complex != 0;
  const int NC = cplx ? 2 : 1;
  int maxchi = 1, maxsl = 1;
  for (int v = 0; v < n; ++v) {
    maxchi = std::max(maxchi, d->link_dim[v]);
    maxsl = std::max(maxsl, p->nslices[v]);
  }
  const int width = (cplx ? 2 : 1) * maxchi;
  if (width <= 32 || n < 3) return TTN_OK; // narrow chains belong to the shared-memory DMMA kernels
  const int W = gemm_width(width);
  if (W == 0 || maxsl > 8) return TTN_OK;
  const int H = W / 2, NSL = maxsl;

  std::vector<int> order(n), pos_of(n);
  {
    int v = d->root;
    for (int pos = n - 1; pos >= 0; --pos) {
      order[pos] = v;
      if (pos > 0) v = p->child[p->child_ptr[v]];
    }
    for (int pos = 0; pos < n; ++pos) pos_of[order[pos]] = pos;
  }
  const double* T = reinterpret_cast<const double*>(d->tensors);
  auto elem = [&](int v, int64_t idx, double* re, double* im) {
    *re = T[(d->tensor_ptr[v] + idx) * NC];
    *im = cplx ? T[(d->tensor_ptr[v] + idx) * NC + 1] : 0.0;
  };
  const int nout = cplx ? 2 : 1;
  const size_t M = (size_t)W * W;
  // Pair merging (same idea as build_chain_mma's group merging): when the slice indices are bit
  // fields and two vertices together have <= 16 slices, chain positions (2m, 2m+1) are contracted
  // into one stream position at plan time, M[s_0 + NSL0 s_1] = E_0[s_0] E_1[s_1]: half as many
  // GEMMs per point.  An odd chain gets an identity vertex in front of the root.
  const int bits0 = NSL == 2 ? 1 : (NSL == 4 ? 2 : -1);
  bool merge = bits0 > 0 && n >= 4;
  if (const char* e = getenv("TTN_MMA_MERGE")) merge = merge && atoi(e) >= 2;
  const int n_pad = merge ? (n + 1) / 2 * 2 : n;             // chain positions incl. identity padding
  const int n_steps0 = n_pad - 2;                            // unmerged middle positions
  if (n >= 2) pos_of[order[n - 1]] = n_pad - 1;
  std::vector<double> leaf((size_t)NSL * W, 0.0), root((size_t)nout * NSL * W, 0.0);
  {
    const int v = order[0], b = d->link_dim[v];
    for (int s = 0; s < p->nslices[v]; ++s)
      for (int j = 0; j < b; ++j) {
        double re, im;
        elem(v, (int64_t)s * b + j, &re, &im);
        leaf[(size_t)s * W + j] = re;
        if (cplx) leaf[(size_t)s * W + H + j] = im;
      }
  }
  {
    const int v = order[n - 1], a = d->link_dim[order[n - 2]];
    for (int s = 0; s < p->nslices[v]; ++s)
      for (int i = 0; i < a; ++i) {
        double re, im;
        elem(v, (int64_t)s * a + i, &re, &im);
        root[(size_t)s * W + i] = re;
        if (cplx) {
          root[(size_t)s * W + H + i] = -im;
          root[((size_t)NSL + s) * W + i] = im;
          root[((size_t)NSL + s) * W + H + i] = re;
        }
      }
  }
  // row-major site matrices of middle position t (slice s): W x W, zero padded
  auto site_matrix = [&](int tI, int s, double* E) {
    std::fill(E, E + M, 0.0);
    if (tI >= n - 2) { // identity padding
      if (s == 0)
        for (int i = 0; i < W; ++i) E[(size_t)i * W + i] = 1.0;
      return;
    }
    const int v = order[tI + 1], a = d->link_dim[order[tI]], b = d->link_dim[v];
    if (s >= p->nslices[v]) return;
    for (int i = 0; i < a; ++i)
      for (int j = 0; j < b; ++j) {
        double re, im;
        elem(v, ((int64_t)s * a + i) * b + j, &re, &im);
        E[(size_t)i * W + j] = re;
        if (cplx) {
          E[(size_t)i * W + H + j] = im;
          E[(size_t)(H + i) * W + j] = -im;
          E[(size_t)(H + i) * W + H + j] = re;
        }
      }
  };
  auto to_frags = [&](const double* E, double* Fs) {
    const int NBW = W / 8;
    for (int kb = 0; kb < W / 4; ++kb)
      for (int nb = 0; nb < NBW; ++nb)
        for (int ln = 0; ln < 32; ++ln)
          Fs[((size_t)kb * NBW + nb) * 32 + ln] = E[(size_t)(4 * kb + (ln & 3)) * W + 8 * nb + (ln >> 2)];
  };
  auto matmul = [&](const double* A, const double* B, double* C) { // C = A B, rows of A that are zero are skipped
    std::fill(C, C + M, 0.0);
    for (int i = 0; i < W; ++i)
      for (int k = 0; k < W; ++k) {
        const double a = A[(size_t)i * W + k];
        if (a == 0.0) continue;
        const double* Bk = B + (size_t)k * W;
        double* Ci = C + (size_t)i * W;
        for (int j = 0; j < W; ++j) Ci[j] += a * Bk[j];
      }
  };
  const int NSLm = merge ? NSL * NSL : NSL;
  const int n_steps = merge ? (n_pad - 4) / 2 : n_steps0;
  const size_t per_site = (size_t)NSLm * M;
  // leaf / root tables (build_gemm_tables below): how many middle positions each side absorbs.  A table has
  // NSLm^(1 + tab) rows of W doubles; budget 2^TTN_GEMM_TABLE_BITS rows (default 16, 0 = no tables).
  int tabL = 0, tabR = 0;
  {
    int tb = 16;
    if (const char* e = getenv("TTN_GEMM_TABLE_BITS")) tb = std::min(atoi(e), 20);
    int mmax = 0;
    for (double rows = NSLm; tb > 0 && rows * NSLm <= std::ldexp(1.0, tb) && rows * NSLm * W * 8 <= 160e6; rows *= NSLm) ++mmax;
    tabL = std::min(mmax, n_steps / 2);
    tabR = std::min(mmax, n_steps - tabL);
    if (NSLm < 2) tabL = tabR = 0;
  }
  auto transpose = [&](const double* A, double* At) {
    for (int i = 0; i < W; ++i)
      for (int j = 0; j < W; ++j) At[(size_t)j * W + i] = A[(size_t)i * W + j];
  };
  double* d_frags_rev = nullptr;
  if (tabR > 0) {
    TTN_CUDA(cudaMalloc(&d_frags_rev, per_site * tabR * 8));
    p->allocs.push_back(d_frags_rev);
  }
  double* d_frags;
  TTN_CUDA(cudaMalloc(&d_frags, std::max<size_t>(per_site * n_steps, 1) * 8));
  p->allocs.push_back(d_frags);
  double flops_exec = 0.0;
  std::vector<double> step_flops; // per (merged) middle position
  const double fmul = cplx ? 8.0 : 2.0;
  auto dim_in = [&](int tI) { return tI < n - 2 ? d->link_dim[order[tI]] : d->link_dim[order[n - 2]]; };
  auto dim_out = [&](int tI) { return tI < n - 2 ? d->link_dim[order[tI + 1]] : d->link_dim[order[n - 2]]; };
  {
    std::vector<double> F(per_site), Frev(tabR > 0 ? per_site : 1), Ea(M), Eb(M), Ec(M);
    if (!merge) {
      for (int tI = 0; tI < n_steps; ++tI) {
        std::fill(F.begin(), F.end(), 0.0);
        for (int s = 0; s < NSL; ++s) {
          site_matrix(tI, s, Ea.data());
          to_frags(Ea.data(), F.data() + (size_t)s * M);
        }
        flops_exec += fmul * dim_in(tI) * dim_out(tI);
        step_flops.push_back(fmul * dim_in(tI) * dim_out(tI));
        TTN_CUDA(cudaMemcpy(d_frags + per_site * tI, F.data(), per_site * 8, cudaMemcpyHostToDevice));
        if (tI >= n_steps - tabR) { // transposed image for the right-to-left table build
          for (int sl = 0; sl < NSL; ++sl) {
            site_matrix(tI, sl, Ea.data());
            transpose(Ea.data(), Eb.data());
            to_frags(Eb.data(), F.data() + (size_t)sl * M);
          }
          TTN_CUDA(cudaMemcpy(d_frags_rev + per_site * (n_steps - 1 - tI), F.data(), per_site * 8, cudaMemcpyHostToDevice));
        }
      }
    } else {
      // leaf' [s0 + NSL s1] = leaf[s0] E_0[s1];  root' [s0 + NSL s1] = E_last[s0] root[s1]
      std::vector<double> leaf2((size_t)NSLm * W, 0.0), root2((size_t)nout * NSLm * W, 0.0);
      for (int s1 = 0; s1 < NSL; ++s1) {
        site_matrix(0, s1, Ea.data());
        for (int s0 = 0; s0 < NSL; ++s0)
          for (int i = 0; i < W; ++i) {
            const double a = leaf[(size_t)s0 * W + i];
            if (a == 0.0) continue;
            for (int j = 0; j < W; ++j) leaf2[(size_t)(s0 + NSL * s1) * W + j] += a * Ea[(size_t)i * W + j];
          }
      }
      for (int s0 = 0; s0 < NSL; ++s0) {
        site_matrix(n_steps0 - 1, s0, Ea.data());
        for (int o = 0; o < nout; ++o)
          for (int s1 = 0; s1 < NSL; ++s1)
            for (int i = 0; i < W; ++i) {
              double acc = 0.0;
              for (int j = 0; j < W; ++j) acc += Ea[(size_t)i * W + j] * root[((size_t)o * NSL + s1) * W + j];
              root2[((size_t)o * NSLm + (s0 + NSL * s1)) * W + i] = acc;
            }
      }
      leaf.swap(leaf2);
      root.swap(root2);
      for (int m = 0; m < n_steps; ++m) { // middle positions (2m+1, 2m+2) of the unmerged chain
        const int ta = 2 * m + 1, tb = 2 * m + 2;
        for (int s1 = 0; s1 < NSL; ++s1) {
          site_matrix(tb, s1, Eb.data());
          for (int s0 = 0; s0 < NSL; ++s0) {
            site_matrix(ta, s0, Ea.data());
            matmul(Ea.data(), Eb.data(), Ec.data());
            to_frags(Ec.data(), F.data() + (size_t)(s0 + NSL * s1) * M);
            if (m >= n_steps - tabR) {
              transpose(Ec.data(), Ea.data());
              to_frags(Ea.data(), Frev.data() + (size_t)(s0 + NSL * s1) * M);
            }
          }
        }
        flops_exec += fmul * dim_in(ta) * dim_out(tb);
        step_flops.push_back(fmul * dim_in(ta) * dim_out(tb));
        TTN_CUDA(cudaMemcpy(d_frags + per_site * m, F.data(), per_site * 8, cudaMemcpyHostToDevice));
        if (m >= n_steps - tabR)
          TTN_CUDA(cudaMemcpy(d_frags_rev + per_site * (n_steps - 1 - m), Frev.data(), per_site * 8, cudaMemcpyHostToDevice));
      }
    }
  }
  p->cgemm_flops_exec = flops_exec + fmul * (merge ? dim_in(n_steps0 - 1) : d->link_dim[order[n - 2]]);
  // stream position and bit shift of every vertex
  for (int v = 0; v < n; ++v) {
    const int pos = pos_of[v];
    pos_of[v] = merge ? ((pos / 2) | ((bits0 * (pos & 1)) << 16)) : pos;
  }
  double *d_leaf, *d_root;
  int32_t* d_pos;
  TTN_CUDA(cudaMalloc(&d_leaf, leaf.size() * 8));
  p->allocs.push_back(d_leaf);
  TTN_CUDA(cudaMalloc(&d_root, root.size() * 8));
  p->allocs.push_back(d_root);
  TTN_CUDA(cudaMalloc(&d_pos, sizeof(int32_t) * n));
  p->allocs.push_back(d_pos);
  TTN_CUDA(cudaMemcpy(d_leaf, leaf.data(), leaf.size() * 8, cudaMemcpyHostToDevice));
  TTN_CUDA(cudaMemcpy(d_root, root.data(), root.size() * 8, cudaMemcpyHostToDevice));
  TTN_CUDA(cudaMemcpy(d_pos, pos_of.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
  ChainGemmDev& c = p->cgemm;
  c.n_vertices = n;
  c.n_steps = n_steps;
  c.n_pos = merge ? n_pad / 2 : n;
  c.merged = merge ? 2 : 0;
  c.nsl = NSLm;
  c.W = W;
  c.nout = nout;
  c.leaf = d_leaf;
  c.root = d_root;
  c.frags = d_frags;
  c.pos_of_vertex = d_pos;
  c.tab_L = c.tab_R = 0;
  c.leaf_tab = c.root_tab = nullptr;
  p->cgemm_ok = true;
  if (tabL + tabR > 0) {
    int rc = build_gemm_tables(p, tabL, tabR, d_frags_rev);
    if (rc) return rc;
    double fe = fmul * (merge ? dim_in(n_steps0 - 1) : d->link_dim[order[n - 2]]);
    for (int tI = tabL; tI < n_steps - tabR; ++tI) fe += step_flops[tI];
    p->cgemm_flops_exec = fe;
  }
  return TTN_OK;
}

template <int W>
static int launch_site(const double* Sin, double* Sout, const uint32_t* list, const int* cls_off, const int* tile_off,
                       const double* frags, int nsl, int pc, cudaStream_t s) {
  constexpr size_t smem = (size_t)GSTAGES * (GBM * GA_STRIDE + (GBK / 4) * (GBN / 8) * 32) * 8;
  auto kern = gemm_site_kernel<W>;
  TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (pc / GBM + nsl) * (W / GBN); // upper bound on the number of tiles
  kern<<<grid, 256, smem, s>>>(Sin, Sout, list, cls_off, tile_off, frags, nsl);
  TTN_CUDA(cudaGetLastError());
  return TTN_OK;
}

// Leaf / root tables of the GEMM-regime chain kernel.  The state after the first 1 + tabL positions depends
// only on their slices: all nsl^(1+tabL) states are computed here, once, by running the ordinary per-site
// kernels over the enumerated slice settings; likewise the co-vectors of the last 1 + tabR positions, by
// running the transposed sites right to left from the root vectors.  An evaluation then gathers one row of
// each table per point and runs only the middle sites as GEMMs.
static int build_gemm_tables(ttn_plan* p, int tabL, int tabR, const double* d_frags_rev) {
  ChainGemmDev& c = p->cgemm;
  const int W = c.W, NS = c.nsl;
  auto ipow = [&](int e) {
    int64_t r = 1;
    for (int q = 0; q < e; ++q) r *= NS;
    return r;
  };
  const int64_t rowsL = ipow(1 + tabL), rowsR = ipow(1 + tabR);
  const int PC = (int)((std::max(rowsL, rowsR) + GBM - 1) / GBM * GBM);
  const int mmax = std::max(tabL, tabR);
  double *S0 = nullptr, *S1 = nullptr, *d_ltab = nullptr, *d_rtab = nullptr;
  uint8_t* slices = nullptr;
  uint32_t* lists = nullptr;
  int* offs = nullptr;
  TTN_CUDA(cudaMalloc(&S0, (size_t)PC * W * 8));
  TTN_CUDA(cudaMalloc(&S1, (size_t)PC * W * 8));
  TTN_CUDA(cudaMalloc(&slices, (size_t)(mmax + 1) * PC));
  TTN_CUDA(cudaMalloc(&lists, (size_t)std::max(mmax, 1) * PC * 4));
  TTN_CUDA(cudaMalloc(&offs, ((size_t)std::max(mmax, 1) * kGemmOff * 2 + 16) * 4));
  int* cls_off = offs;
  int* tile_off = offs + (size_t)std::max(mmax, 1) * kGemmOff + 8;
  TTN_CUDA(cudaMalloc(&d_ltab, (size_t)rowsL * W * 8));
  p->allocs.push_back(d_ltab);
  TTN_CUDA(cudaMalloc(&d_rtab, (size_t)c.nout * rowsR * W * 8));
  p->allocs.push_back(d_rtab);
  int rc = TTN_OK;
  auto run = [&](const double* leaf, const double* frags, int m, double* dst, int64_t rows) {
    // positions 0..m of the enumerated settings: leaf vectors, then m sites
    gemm_enum_kernel<<<(PC + 255) / 256, 256>>>(slices, PC, m + 1, NS);
    if (m > 0) gemm_classify_kernel<<<m, 1024>>>(slices, PC, NS, lists, cls_off, tile_off, W / GBN);
    gemm_leaf_kernel<<<(unsigned)(((int64_t)PC * (W / 2) + 255) / 256), 256>>>(slices, PC, leaf, W, S0, 1, NS);
    double *Sin = S0, *Sout = S1;
    for (int t = 0; t < m && rc == TTN_OK; ++t) {
      const double* fr = frags + (size_t)t * NS * W * W;
      if (W == 128) rc = launch_site<128>(Sin, Sout, lists + (size_t)t * PC, cls_off + t * kGemmOff, tile_off + t * kGemmOff, fr, NS, PC, 0);
      else rc = launch_site<256>(Sin, Sout, lists + (size_t)t * PC, cls_off + t * kGemmOff, tile_off + t * kGemmOff, fr, NS, PC, 0);
      std::swap(Sin, Sout);
    }
    if (rc == TTN_OK && cudaMemcpy(dst, Sin, (size_t)rows * W * 8, cudaMemcpyDeviceToDevice) != cudaSuccess) rc = TTN_ERR_CUDA;
  };
  run(c.leaf, c.frags, tabL, d_ltab, rowsL);
  for (int o = 0; o < c.nout && rc == TTN_OK; ++o)
    run(c.root + (size_t)o * NS * W, d_frags_rev, tabR, d_rtab + (size_t)o * rowsR * W, rowsR);
  if (rc == TTN_OK && cudaDeviceSynchronize() != cudaSuccess) rc = TTN_ERR_CUDA;
  cudaFree(S0);
  cudaFree(S1);
  cudaFree(slices);
  cudaFree(lists);
  cudaFree(offs);
  if (rc == TTN_ERR_CUDA) {
    set_error(std::string("GEMM chain tables: ") + cudaGetErrorString(cudaGetLastError()));
    return rc;
  }
  if (rc) return rc;
  c.tab_L = tabL;
  c.tab_R = tabR;
  c.leaf_tab = d_ltab;
  c.root_tab = d_rtab;
  return TTN_OK;
}

int launch_chain_gemm(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double* d_partial,
                      int* n_partial, cudaStream_t s, int* n_launches) {
  *n_partial = 0;
  if (src.npts == 0) return TTN_OK;
  if (!p->cgemm_ok) {
    set_error("GEMM chain kernel requested but the network is not a wide chain (32 < width <= 256)");
    return TTN_ERR_UNSUPPORTED;
  }
  const ChainGemmDev& c = p->cgemm;
  const int W = c.W, n_pos = c.n_pos;
  // points per chunk: (PC/128 + nsl) * (W/128) tiles per site should fill a whole number of waves
  const int nnb = W / GBN;
  const int waves = nnb == 2 ? 7 : 4;
  const int PC = (waves * p->sm_count / nnb - std::max(c.nsl, 8)) * GBM; // every class may end in a partial tile
  // workspace: two state buffers, slices, lists, offsets
  const size_t state_b = (size_t)PC * W * 8;
  const size_t slices_b = ((size_t)n_pos * PC + 255) / 256 * 256;
  const size_t lists_b = (size_t)std::max(c.n_steps, 1) * PC * 4;
  const size_t offs_b = (size_t)std::max(c.n_steps, 1) * kGemmOff * 4 * 2 + 256;
  const size_t need = 2 * state_b + slices_b + lists_b + offs_b;
  if (st.gemm_bytes < need) {
    if (st.d_gemm) cudaFree(st.d_gemm);
    st.d_gemm = nullptr;
    st.gemm_bytes = 0;
    TTN_CUDA(cudaMalloc(&st.d_gemm, need));
    st.gemm_bytes = need;
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(st.d_gemm);
  double* S0 = reinterpret_cast<double*>(base);
  double* S1 = reinterpret_cast<double*>(base + state_b);
  uint8_t* slices = base + 2 * state_b;
  uint32_t* lists = reinterpret_cast<uint32_t*>(base + 2 * state_b + slices_b);
  int* cls_off = reinterpret_cast<int*>(base + 2 * state_b + slices_b + lists_b);
  int* tile_off = cls_off + (size_t)std::max(c.n_steps, 1) * kGemmOff + 8;
  const int do_sum = d_partial != nullptr;
  const int64_t n_chunks = (src.npts + PC - 1) / PC;
  const int root_blocks = PC * 32 / 256;
  double* big_partial = nullptr;
  if (do_sum) { // per-block partials of every chunk, reduced to ONE entry of the caller's array at the end
    const size_t needp = sizeof(double) * 2 * (size_t)n_chunks * root_blocks;
    if (st.partial2_bytes < needp) {
      if (st.d_partial2) cudaFree(st.d_partial2);
      st.d_partial2 = nullptr;
      st.partial2_bytes = 0;
      TTN_CUDA(cudaMalloc(&st.d_partial2, needp));
      st.partial2_bytes = needp;
    }
    big_partial = st.d_partial2;
  }
  for (int64_t ck = 0; ck < n_chunks; ++ck) {
    const int64_t p0 = ck * PC;
    const int pc = PC; // dead points of the last chunk are evaluated at x = 0 and never stored
    gemm_digits_kernel<<<(pc + 255) / 256, 256, 0, s>>>(p->digits, src, p0, pc, c.pos_of_vertex, n_pos, slices, p->d_err);
    if (c.n_steps > 0)
      gemm_classify_kernel<<<c.n_steps, 1024, 0, s>>>(slices, pc, c.nsl, lists, cls_off, tile_off, W / GBN);
    gemm_leaf_kernel<<<(unsigned)(((int64_t)pc * (W / 2) + 255) / 256), 256, 0, s>>>(slices, pc, c.leaf_tab ? c.leaf_tab : c.leaf, W, S0,
                                                                                     1 + c.tab_L, c.nsl);
    *n_launches += 3;
    double *Sin = S0, *Sout = S1;
    for (int t = c.tab_L; t < c.n_steps - c.tab_R; ++t) {
      const double* fr = c.frags + (size_t)t * c.nsl * W * W;
      int rc;
      if (W == 128) rc = launch_site<128>(Sin, Sout, lists + (size_t)t * pc, cls_off + t * kGemmOff, tile_off + t * kGemmOff, fr, c.nsl, pc, s);
      else rc = launch_site<256>(Sin, Sout, lists + (size_t)t * pc, cls_off + t * kGemmOff, tile_off + t * kGemmOff, fr, c.nsl, pc, s);
      if (rc) return rc;
      std::swap(Sin, Sout);
      *n_launches += 1;
    }
    int64_t root_rows = c.nsl;
    for (int q = 0; q < c.tab_R; ++q) root_rows *= c.nsl;
    gemm_root_kernel<<<root_blocks, 256, 0, s>>>(slices + (size_t)(n_pos - 1) * pc, pc, p0, src.npts,
                                                 c.root_tab ? c.root_tab : c.root, W, c.nsl, 1 + c.tab_R, root_rows, c.nout,
                                                 c.n_vertices, Sin, d_out, do_sum ? big_partial + 2 * ck * root_blocks : nullptr,
                                                 do_sum, src);
    *n_launches += 1;
    TTN_CUDA(cudaGetLastError());
  }
  if (do_sum) {
    int rc = launch_sum_partials(p, big_partial, (int)(n_chunks * root_blocks), c.nout, d_partial, s);
    if (rc) return rc;
    *n_launches += 1;
  }
  *n_partial = do_sum ? 1 : 0;
  return TTN_OK;
}

} // namespace ttn
