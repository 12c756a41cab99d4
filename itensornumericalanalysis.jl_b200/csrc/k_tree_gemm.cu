// K7 — trees with at most two children per vertex and bond dimension up to 64 (BASELINE config 3:
// binary tree, chi = 64): the network is contracted VERTEX BY VERTEX over a chunk of PC points, the
// messages M_v[PC][W] living in HBM/L2 (one vertex's slices stay L2-resident while the
// whole chunk streams through it).  Per vertex, for the points that selected slice d:
//   leaf            M_v[p]    = T_v[d]                                   (row copy)
//   one child  c    M_v[p,:]  = M_c[p,:] * T_v[d]            (W x W)      grouped GEMM, K = W
//   two children    M_v[p,n]  = sum_{a,b} M_a[p,a] M_b[p,b] T_v[d][a,b,n]  Khatri-Rao GEMM, K = W^2:
//                   the A operand M_a[p,a]*M_b[p,b] is never materialised: for every `a` the b-sum
//                   M_b[p,:] T_v[d][a,:,n] runs on DMMA into a second accumulator tile, which is then
//                   folded in with M_a[p,a] (one DFMA per 8 DMMAs): b first, a second, like a nested
//                   contraction (2W additions deep instead of W^2)
//   root            value = full contraction with a parent dimension of 1  (small dot kernels)
// CTA tile: 128 rows x W columns; both child row blocks stay in shared memory for the whole K loop,
// only the tensor T_v[d] (B operand, fragment order) streams through a 3-stage cp.async ring.
// This is exactly the flop rule of SURVEY §8(d): 2*W^3 per degree-3 vertex (+2*W^2 absorbed).
#include <algorithm>
#include <cstring>

#include "k_async.cuh"
#include "k_digits.cuh"

namespace ttn {

constexpr int TBM = 128, TBK = 16, TSTAGES = 3;

__device__ __forceinline__ void t_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void t_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void t_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// slices[v][i] = slice index of vertex v at point i of the chunk.  A vertex with exactly one site index gets one
// direct store; only the vertices listed in zero_v (no site index, or several) are zero-filled and accumulated.
__global__ void tree_digits_kernel(DigitTable dg, CoordSource src, int64_t p0, int pc, int n_vertices,
                                   uint8_t* __restrict__ slices, int* err, const int32_t* __restrict__ zero_v, int n_zero) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pc) return;
  const int64_t p = p0 + i;
  if (p >= src.npts) { // padding points of the last chunk: slice 0 everywhere (table rows / classes stay in range)
    for (int v = 0; v < n_vertices; ++v) slices[(size_t)v * pc + i] = 0;
    return;
  }
  for (int k = 0; k < n_zero; ++k) slices[(size_t)zero_v[k] * pc + i] = 0;
  for (int c = 0; c < dg.n_coords; ++c) {
    double x = load_coord(src, p, c);
    if (!coord_in_domain(x)) {
      atomicOr(err, 1);
      x = 0.0;
    }
    for (int k = dg.coord_ptr[c]; k < dg.coord_ptr[c + 1]; ++k) {
      const DigitEntry e = dg.entries[k];
      const int v = src.digits ? given_digit(src, p, dg.n_sites, e.site, e.base, err) : greedy_digit(x, dg.thr + e.thr_off, e.base);
      if (e.pad_) slices[(size_t)e.vertex * pc + i] = (uint8_t)(v * e.stride);
      else slices[(size_t)e.vertex * pc + i] += (uint8_t)(v * e.stride);
    }
  }
}

// one block per vertex: counting sort of the chunk by the slice selected at that vertex.  Counts and list
// positions are warp-aggregated (one ballot per class, one shared-memory atomic per warp and class): with one
// atomic per POINT on 2..8 counters the 1024 threads of the block serialised, and this kernel — not the GEMMs —
// bounded narrow trees (60-vertex comb, W = 16: 1 ms of the 1.3 ms a 3e5-point chunk took).
__global__ void __launch_bounds__(1024)
    tree_classify_kernel(const uint8_t* __restrict__ slices, int pc, const int32_t* __restrict__ nslices,
                         uint32_t* __restrict__ lists, int* __restrict__ cls_off, int* __restrict__ tile_off) {
  const int v = blockIdx.x;
  const int nsl = nslices[v];
  const uint8_t* sl = slices + (size_t)v * pc;
  uint32_t* list = lists + (size_t)v * pc;
  __shared__ int cnt[8], cursor[8];
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  if (threadIdx.x < 8) cnt[threadIdx.x] = 0;
  __syncthreads();
  if (nsl > 1) {
    // 4 independent loads per thread and trip: the block walks the chunk alone, so the global-load latency of
    // every trip is exposed
    int local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i0 = threadIdx.x - lane; i0 < pc; i0 += 4 * blockDim.x) { // warp-uniform bounds
      int sv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x + lane;
        sv[u] = i < pc ? (int)sl[i] : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < nsl) local[c] += __popc(__ballot_sync(0xffffffffu, sv[u] == c));
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < nsl && local[c]) atomicAdd(&cnt[c], local[c]);
    }
  } else if (threadIdx.x == 0) {
    cnt[0] = pc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0, trun = 0;
    for (int c = 0; c < nsl; ++c) {
      cursor[c] = run;
      cls_off[v * 9 + c] = run;
      tile_off[v * 9 + c] = trun;
      run += cnt[c];
      trun += (cnt[c] + TBM - 1) / TBM;
    }
    for (int c = nsl; c <= 8; ++c) {
      cls_off[v * 9 + c] = run;
      tile_off[v * 9 + c] = trun;
    }
  }
  __syncthreads();
  if (nsl > 1) {
    for (int i0 = threadIdx.x - lane; i0 < pc; i0 += 4 * blockDim.x) {
      int sv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x + lane;
        sv[u] = i < pc ? (int)sl[i] : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x + lane;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < nsl) {
            const uint32_t m = __ballot_sync(0xffffffffu, sv[u] == c);
            if (m) { // warp-uniform
              int base = 0;
              if (lane == 0) base = atomicAdd(&cursor[c], __popc(m));
              base = __shfl_sync(0xffffffffu, base, 0);
              if (sv[u] == c) list[base + __popc(m & lt)] = (uint32_t)i;
            }
          }
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < pc; i += blockDim.x) list[i] = (uint32_t)i;
  }
}

// ---- classification of the evaluation path: only the vertices that run as GEMMs need row lists, a vertex may
// be a MERGED run of up to 4 single-child vertices (class = their slice bits side by side, <= 16 classes), and
// the chunk is split over many blocks (count -> offsets -> scatter); one block per vertex walking the chunk alone
// took a third of a narrow tree's time (profiles/r01_launches_tree_comb_summary.txt).
constexpr int TCLS = 16, TOFF = TCLS + 1, TSEG = 4096;
struct TgClass {
  int32_t v;       // list / offset slot (the top vertex of a merged run)
  int32_t nsl;     // classes
  int32_t n_mem;   // member vertices
  int32_t mem_v[4], mem_shift[4];
};

__device__ __forceinline__ int tg_class_of(const uint8_t* __restrict__ slices, int pc, const TgClass& g, int i) {
  int c = 0;
  for (int m = 0; m < g.n_mem; ++m) c |= (int)slices[(size_t)g.mem_v[m] * pc + i] << g.mem_shift[m];
  return c;
}

__global__ void __launch_bounds__(256) tree_count_kernel(const uint8_t* __restrict__ slices, int pc, const TgClass* __restrict__ gv,
                                                         int* __restrict__ counts) {
  const TgClass g = gv[blockIdx.y];
  __shared__ int cnt[TCLS];
  if (threadIdx.x < TCLS) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int i_end = min(pc, (int)(blockIdx.x + 1) * TSEG);
  int local = 0; // lane c keeps the warp's count of class c
  for (int i0 = blockIdx.x * TSEG + (threadIdx.x - lane); i0 < i_end; i0 += 256) {
    const int i = i0 + lane;
    const int c = i < i_end ? tg_class_of(slices, pc, g, i) : -1;
    for (int k = 0; k < g.nsl; ++k) {
      const int n = __popc(__ballot_sync(0xffffffffu, c == k));
      if (lane == k) local += n;
    }
  }
  if (lane < g.nsl && local) atomicAdd(&cnt[lane], local);
  __syncthreads();
  if (threadIdx.x < g.nsl && cnt[threadIdx.x]) atomicAdd(&counts[blockIdx.y * TOFF + threadIdx.x], cnt[threadIdx.x]);
}

__global__ void tree_offsets_kernel(const TgClass* __restrict__ gv, int nv, const int* __restrict__ counts,
                                    int* __restrict__ cls_off, int* __restrict__ tile_off, int* __restrict__ cursor) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= nv) return;
  const TgClass g = gv[y];
  int run = 0, trun = 0;
  for (int c = 0; c <= TCLS; ++c) {
    cls_off[g.v * TOFF + c] = run;
    tile_off[g.v * TOFF + c] = trun;
    if (c < TCLS) cursor[y * TOFF + c] = run;
    const int n = c < g.nsl ? counts[y * TOFF + c] : 0;
    run += n;
    trun += (n + TBM - 1) / TBM;
  }
}

__global__ void __launch_bounds__(256) tree_scatter_kernel(const uint8_t* __restrict__ slices, int pc, const TgClass* __restrict__ gv,
                                                           int* __restrict__ cursor, uint32_t* __restrict__ lists) {
  const TgClass g = gv[blockIdx.y];
  uint32_t* list = lists + (size_t)g.v * pc;
  __shared__ int cnt[TCLS], base[TCLS];
  if (threadIdx.x < TCLS) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const int i_beg = blockIdx.x * TSEG, i_end = min(pc, i_beg + TSEG);
  if (g.nsl == 1) {
    for (int i = i_beg + threadIdx.x; i < i_end; i += 256) list[i] = (uint32_t)i;
    return;
  }
  int local = 0;
  for (int i0 = i_beg + (threadIdx.x - lane); i0 < i_end; i0 += 256) {
    const int i = i0 + lane;
    const int c = i < i_end ? tg_class_of(slices, pc, g, i) : -1;
    for (int k = 0; k < g.nsl; ++k) {
      const int n = __popc(__ballot_sync(0xffffffffu, c == k));
      if (lane == k) local += n;
    }
  }
  if (lane < g.nsl && local) atomicAdd(&cnt[lane], local);
  __syncthreads();
  if (threadIdx.x < g.nsl) {
    base[threadIdx.x] = cnt[threadIdx.x] ? atomicAdd(&cursor[blockIdx.y * TOFF + threadIdx.x], cnt[threadIdx.x]) : 0;
    cnt[threadIdx.x] = 0;
  }
  __syncthreads();
  for (int i0 = i_beg + (threadIdx.x - lane); i0 < i_end; i0 += 256) {
    const int i = i0 + lane;
    const int c = i < i_end ? tg_class_of(slices, pc, g, i) : -1;
    for (int k = 0; k < g.nsl; ++k) {
      const uint32_t m = __ballot_sync(0xffffffffu, c == k);
      if (m) { // warp-uniform
        int b = 0;
        if (lane == 0) b = atomicAdd(&cnt[k], __popc(m));
        b = __shfl_sync(0xffffffffu, b, 0);
        if (c == k) list[base[k] + b + __popc(m & lt)] = (uint32_t)i;
      }
    }
  }
}

// root with two children and W <= 32: one THREAD per point, the root tensor of every slice in shared memory
// (value = sum_a M_a[a] sum_b T_d[a][b] M_b[b]); the warp-per-point kernel below spent a quarter of a narrow
// tree's time here.  (A DMMA variant — S = M_b x T_d^T over tiles of 128 consecutive points with a dot-product epilogue —
// was measured slower: 78 vs 63 us per 3e5-point chunk; both are bound by the latency of two row gathers per point, and
// this kernel keeps more blocks resident.)
template <int W>
__global__ void __launch_bounds__(128) tree_root2_kernel(const double* __restrict__ Ma, const double* __restrict__ Mb,
                                                         const uint8_t* __restrict__ sl, int pc, int64_t p0, int64_t npts,
                                                         const double* __restrict__ T, int nsl, double* __restrict__ out,
                                                         double* __restrict__ partial, int do_sum, CoordSource src,
                                                         const uint32_t* __restrict__ ta, const uint32_t* __restrict__ tb) {
  extern __shared__ __align__(16) double sT[]; // [nsl][W][W]
  for (int i = threadIdx.x; i < nsl * W * W; i += 128) sT[i] = T[i];
  __syncthreads();
  const int i = blockIdx.x * 128 + threadIdx.x;
  const bool live = i < pc && p0 + i < npts;
  double v = 0.0;
  if (live) {
    double mb[W];
    const double2* rb = reinterpret_cast<const double2*>(Mb + (size_t)(tb ? tb[i] : (uint32_t)i) * W);
#pragma unroll
    for (int j = 0; j < W / 2; ++j) {
      const double2 t = rb[j];
      mb[2 * j] = t.x;
      mb[2 * j + 1] = t.y;
    }
    const double* Td = sT + (size_t)sl[i] * W * W;
    const double* ra = Ma + (size_t)(ta ? ta[i] : (uint32_t)i) * W;
#pragma unroll 4
    for (int a = 0; a < W; ++a) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < W; ++b) s = fma(Td[a * W + b], mb[b], s);
      v = fma(ra[a], s, v);
    }
    if (out) out[p0 + i] = v;
  }
  if (do_sum) {
    __shared__ double sh[128];
    double a0 = 0.0, a1 = 0.0;
    if (live) accumulate_point(src, p0 + i, v, 0.0, a0, a1);
    sh[threadIdx.x] = a0;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0;
      for (int k = 0; k < 128; ++k) a += sh[k]; // fixed order: deterministic
      partial[2 * blockIdx.x] = a;
      partial[2 * blockIdx.x + 1] = 0.0;
    }
  }
}

__global__ void tree_leaf_kernel(const uint8_t* __restrict__ sl, int pc, const double* __restrict__ T, int W,
                                 double* __restrict__ M) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // one double2 per thread
  const int per_row = W / 2;
  if (idx >= (int64_t)pc * per_row) return;
  const int i = (int)(idx / per_row), j = (int)(idx % per_row);
  reinterpret_cast<double2*>(M)[idx] = *reinterpret_cast<const double2*>(T + (size_t)sl[i] * W + 2 * j);
}

// M_v[rows of class d] = KhatriRao(M_a, M_b)[rows] * T_d   (NCH == 2)   or   M_c[rows] * T_d (NCH == 1)
// frags[d][kb][nb][lane] = T_d[k = 4 kb + (lane & 3)][n = 8 nb + (lane >> 2)],  k = a * W + b.
// CPLX (two children of a COMPLEX network): rows are interleaved (re, im) pairs of W / 2 complex entries; the
// embedded tensor of one complex `a` is a W x W real matrix (k = a W + b', b' = 2 b + re/im), so the b-sum is the same
// DMMA loop and a fragment's column pair (2t, 2t + 1) is one complex number: the fold with M_a[p, a] is a complex
// multiply-add inside the thread.  Single-child vertices of complex networks are plain real GEMMs on the embedded
// matrix [[re, im], [-im, re]] and need no kernel support.
template <int W, int NCH, bool CPLX = false>
__global__ void __launch_bounds__(256, 1)
    tree_vertex_kernel(const double* __restrict__ Ma, const double* __restrict__ Mb, double* __restrict__ Mout,
                       const uint32_t* __restrict__ list, const int* __restrict__ cls_off,
                       const int* __restrict__ tile_off, const double* __restrict__ frags, int nsl,
                       const uint32_t* __restrict__ ta, const uint32_t* __restrict__ tb) {
  // ta / tb != nullptr: child a / b is a TABULATED subtree — Ma / Mb is its message table and ta / tb holds the
  // table row of every point of the chunk (tree_tabidx_kernel): the rows are gathered straight from the table
  // instead of being copied into a message buffer first
  constexpr int RS = W + 4;            // row stride (doubles) of the child blocks in shared memory
  constexpr int NT = W / 16;           // 8-column tiles per warp (warp tile 32 x W/2)
  constexpr int KTOT = (NCH == 2) ? (CPLX ? W * W / 2 : W * W) : W;
  constexpr int NKC = KTOT / TBK;
  constexpr int BCHUNKS = W / TBK;     // K chunks per value of `a`
  constexpr int B_STAGE_D = TBK * W;   // doubles per B stage (TBK/4 k-blocks x W/8 n-blocks x 32 lanes)
  extern __shared__ __align__(128) unsigned char smem[];
  double* Sa = reinterpret_cast<double*>(smem);          // [TBM][RS]  (NCH == 2 only)
  double* Sb = Sa + (NCH == 2 ? TBM * RS : 0);           // [TBM][RS]  second child, or the only child
  double* Bs = Sb + TBM * RS;                            // [TSTAGES][B_STAGE_D]
  __shared__ uint32_t rowid[TBM], rowa[TBM], rowb[TBM];

  const int tile = blockIdx.x;
  int c = -1;
  for (int k = 0; k < nsl; ++k)
    if (tile >= tile_off[k] && tile < tile_off[k + 1]) c = k;
  if (c < 0) return;
  const int mblk = tile - tile_off[c];
  const int row0 = cls_off[c] + mblk * TBM, row_end = cls_off[c + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < TBM) {
    const uint32_t r = list[min(row0 + tid, row_end - 1)];
    rowid[tid] = r;
    rowb[tid] = tb ? tb[r] : r;
    if (NCH == 2) rowa[tid] = ta ? ta[r] : r;
  }
  __syncthreads();

  const double* E = frags + (size_t)c * KTOT * W;
  const uint32_t bs_base = smem_u32(Bs);
  auto load_B = [&](int stage, int kc) {
    const double* srcp = E + (size_t)kc * B_STAGE_D;
#pragma unroll
    for (int q = 0; q < B_STAGE_D / 2 / 256; ++q)
      t_cp_async16(bs_base + (uint32_t)(stage * B_STAGE_D + (tid + q * 256) * 2) * 8u, srcp + (tid + q * 256) * 2);
    if (B_STAGE_D / 2 < 256 && tid < B_STAGE_D / 2)
      t_cp_async16(bs_base + (uint32_t)(stage * B_STAGE_D + tid * 2) * 8u, srcp + tid * 2);
  };
  // group 0: the child row blocks (gathered) + the first B stage
  {
    constexpr int CPR = W / 2; // 16-byte chunks per row
    for (int ch = tid; ch < TBM * CPR; ch += 256) {
      const int r = ch / CPR, cc = ch % CPR;
      t_cp_async16(smem_u32(Sb) + (uint32_t)(r * RS + cc * 2) * 8u, Mb + (size_t)rowb[r] * W + cc * 2);
      if (NCH == 2) t_cp_async16(smem_u32(Sa) + (uint32_t)(r * RS + cc * 2) * 8u, Ma + (size_t)rowa[r] * W + cc * 2);
    }
    load_B(0, 0);
    t_cp_async_commit();
    if (NKC > 1) {
      load_B(1, 1);
      t_cp_async_commit();
    }
  }

  const int wm = warp & 3, wn = warp >> 2; // 4 x 2 warps, warp tile 32 x (W / 2)
  const int g = lane >> 2, t = lane & 3;
  double acc[4][NT][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const uint32_t sa_base = smem_u32(Sa), sb_base = smem_u32(Sb);
  double ma[4] = {1.0, 1.0, 1.0, 1.0}, mi[4] = {0.0, 0.0, 0.0, 0.0};
  double inner[NCH == 2 ? 4 : 1][NT][2]; // partial sums over b for the current a (two children)

  for (int kc = 0; kc < NKC; ++kc) {
    if (kc + 2 < NKC) {
      load_B((kc + 2) % TSTAGES, kc + 2);
      t_cp_async_commit();
      t_cp_async_wait<2>();
    } else if (kc + 1 < NKC) {
      t_cp_async_wait<1>();
    } else {
      t_cp_async_wait<0>();
    }
    __syncthreads();
    const int a = kc / BCHUNKS, b0 = (kc % BCHUNKS) * TBK;
    if (NCH == 2 && (kc % BCHUNKS) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (CPLX) {
          const double2 z = lds128(sa_base + (uint32_t)((wm * 32 + i * 8 + g) * RS + 2 * a) * 8u);
          ma[i] = z.x;
          mi[i] = z.y;
        } else {
          ma[i] = lds64(sa_base + (uint32_t)((wm * 32 + i * 8 + g) * RS + a) * 8u);
        }
      }
    }
    const uint32_t b_st = bs_base + (uint32_t)((kc % TSTAGES) * B_STAGE_D) * 8u;
    if (NCH == 2 && (kc % BCHUNKS) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) inner[i][j][0] = inner[i][j][1] = 0.0;
    }
#pragma unroll
    for (int k4 = 0; k4 < TBK / 4; ++k4) {
      double af[4], bf[NT];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = lds64(sb_base + (uint32_t)((wm * 32 + i * 8 + g) * RS + b0 + k4 * 4 + t) * 8u);
#pragma unroll
      for (int j = 0; j < NT; ++j) bf[j] = lds64(b_st + (uint32_t)((k4 * (W / 8) + wn * NT + j) * 32 + lane) * 8u);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if (NCH == 2) dmma884(inner[i][j][0], inner[i][j][1], af[i], bf[j]);
          else dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    // two children: the b-sum of one `a` is complete -> acc += M_a[p, a] * inner.  Summing b first and a
    // second is the order a nested contraction uses: 2 W additions deep instead of W^2
    if (NCH == 2 && (kc % BCHUNKS) == BCHUNKS - 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          acc[i][j][0] = fma(ma[i], inner[i][j][0], acc[i][j][0]);
          acc[i][j][1] = fma(ma[i], inner[i][j][1], acc[i][j][1]);
          if (CPLX) {
            acc[i][j][0] = fma(-mi[i], inner[i][j][1], acc[i][j][0]);
            acc[i][j][1] = fma(mi[i], inner[i][j][0], acc[i][j][1]);
          }
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = wm * 32 + i * 8 + g;
    if (row0 + r < row_end) {
      double* dst = Mout + (size_t)rowid[r] * W + wn * (W / 2) + 2 * t;
#pragma unroll
      for (int j = 0; j < NT; ++j) *reinterpret_cast<double2*>(dst + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

// root: parent dimension 1.  One warp per point.
//   NCH == 2: value = sum_a M_a[a] * (sum_b T_d[a][b] M_b[b]);  NCH == 1: value = sum_a M_c[a] T_d[a];  NCH == 0: T_d.
__global__ void tree_root_kernel(int nch, const double* __restrict__ Ma, const double* __restrict__ Mb,
                                 const uint8_t* __restrict__ sl, int pc, int64_t p0, int64_t npts,
                                 const double* __restrict__ T, int W, double* __restrict__ out,
                                 double* __restrict__ partial, int do_sum, CoordSource src) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  double v = 0.0;
  bool live = false;
  if (warp < pc && p0 + warp < npts) {
    live = true;
    const int d = sl[warp];
    if (nch == 0) {
      v = T[d];
    } else if (nch == 1) {
      const double* Td = T + (size_t)d * W;
      for (int a = lane; a < W; a += 32) v = fma(Mb[(size_t)warp * W + a], __ldg(Td + a), v);
    } else {
      const double* Td = T + (size_t)d * W * W;
      const double* mb = Mb + (size_t)warp * W;
      for (int a = lane; a < W; a += 32) {
        double s = 0.0;
        for (int b = 0; b < W; ++b) s = fma(__ldg(Td + (size_t)a * W + b), mb[b], s);
        v = fma(Ma[(size_t)warp * W + a], s, v);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0 && out) out[p0 + warp] = v;
  }
  if (do_sum) {
    __shared__ double sh[8];
    if (lane == 0) {
      double a0 = 0.0, a1 = 0.0;
      if (live) accumulate_point(src, p0 + warp, v, 0.0, a0, a1);
      sh[threadIdx.x >> 5] = a0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) a += sh[k];
      partial[2 * blockIdx.x] = a;
      partial[2 * blockIdx.x + 1] = 0.0;
    }
  }
}

// root of a COMPLEX network: one thread per point, root tensor [slice][a][b] as (re, im) pairs with row strides
// H = W / 2 complex entries, read through the read-only cache (threads of a warp share the few slices).
__global__ void __launch_bounds__(128) tree_root_c_kernel(int nch, const double* __restrict__ Ma, const double* __restrict__ Mb,
                                                          const uint8_t* __restrict__ sl, int pc, int64_t p0, int64_t npts,
                                                          const double* __restrict__ T, int W, double* __restrict__ out,
                                                          double* __restrict__ partial, int do_sum, CoordSource src,
                                                          const uint32_t* __restrict__ ta, const uint32_t* __restrict__ tb) {
  const int H = W / 2;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const bool live = i < pc && p0 + i < npts;
  double vr = 0.0, vi = 0.0;
  if (live) {
    const int d = sl[i];
    if (nch == 0) {
      vr = __ldg(T + 2 * d);
      vi = __ldg(T + 2 * d + 1);
    } else if (nch == 1) {
      const double2* Td = reinterpret_cast<const double2*>(T) + (size_t)d * H;
      const double2* m = reinterpret_cast<const double2*>(Mb + (size_t)(tb ? tb[i] : (uint32_t)i) * W);
      for (int a = 0; a < H; ++a) {
        const double2 x = m[a], t = __ldg(Td + a);
        vr = fma(x.x, t.x, vr);
        vr = fma(-x.y, t.y, vr);
        vi = fma(x.x, t.y, vi);
        vi = fma(x.y, t.x, vi);
      }
    } else {
      const double2* Td = reinterpret_cast<const double2*>(T) + (size_t)d * H * H;
      const double2* ma_ = reinterpret_cast<const double2*>(Ma + (size_t)(ta ? ta[i] : (uint32_t)i) * W);
      const double2* mb_ = reinterpret_cast<const double2*>(Mb + (size_t)(tb ? tb[i] : (uint32_t)i) * W);
      for (int a = 0; a < H; ++a) {
        double sr = 0.0, si = 0.0;
        for (int b = 0; b < H; ++b) {
          const double2 x = mb_[b], t = __ldg(Td + (size_t)a * H + b);
          sr = fma(x.x, t.x, sr);
          sr = fma(-x.y, t.y, sr);
          si = fma(x.x, t.y, si);
          si = fma(x.y, t.x, si);
        }
        const double2 z = ma_[a];
        vr = fma(z.x, sr, vr);
        vr = fma(-z.y, si, vr);
        vi = fma(z.x, si, vi);
        vi = fma(z.y, sr, vi);
      }
    }
    if (out) reinterpret_cast<double2*>(out)[p0 + i] = make_double2(vr, vi);
  }
  if (do_sum) {
    __shared__ double sh[2][128];
    double a0 = 0.0, a1 = 0.0;
    if (live) accumulate_point(src, p0 + i, vr, vi, a0, a1);
    sh[0][threadIdx.x] = a0;
    sh[1][threadIdx.x] = a1;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
      for (int k = 0; k < 128; ++k) { // fixed order: deterministic
        a += sh[0][k];
        b += sh[1][k];
      }
      partial[2 * blockIdx.x] = a;
      partial[2 * blockIdx.x + 1] = b;
    }
  }
}

// ---- subtree message tables (build_tree_tables): the message a subtree sends upwards depends only
// on the digits inside the subtree; for subtrees with few bits every message is tabulated at plan time.
// vs[3 j .. 3 j + 2] = (vertex, bit offset in the table index, slice mask) of the j-th subtree vertex.
__global__ void tree_enum_kernel(uint8_t* __restrict__ slices, int pc, const int32_t* __restrict__ vs, int ns) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pc) return;
  for (int j = 0; j < ns; ++j) slices[(size_t)vs[3 * j] * pc + i] = (uint8_t)((i >> vs[3 * j + 1]) & vs[3 * j + 2]);
}

__global__ void tree_table_kernel(const uint8_t* __restrict__ slices, int pc, const int32_t* __restrict__ vs, int ns,
                                  const double* __restrict__ table, int W, double* __restrict__ M) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // one double2 per thread
  const int per_row = W / 2;
  if (idx >= (int64_t)pc * per_row) return;
  const int i = (int)(idx / per_row), j = (int)(idx % per_row);
  uint32_t t = 0;
  for (int q = 0; q < ns; ++q) t |= (uint32_t)slices[(size_t)vs[3 * q] * pc + i] << vs[3 * q + 1];
  reinterpret_cast<double2*>(M)[idx] = __ldg(reinterpret_cast<const double2*>(table + (size_t)t * W) + j);
}

// table row of every point of the chunk, for all tables of the plan at once: grid (chunk / 256, tables)
struct TgTabRef {
  const int32_t* vs;
  int32_t ns;
};
__global__ void tree_tabidx_kernel(const uint8_t* __restrict__ slices, int pc, const TgTabRef* __restrict__ tabs,
                                   uint32_t* __restrict__ tabidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pc) return;
  const TgTabRef tr = tabs[blockIdx.y];
  uint32_t t = 0;
  for (int q = 0; q < tr.ns; ++q) t |= (uint32_t)slices[(size_t)tr.vs[3 * q] * pc + i] << tr.vs[3 * q + 1];
  tabidx[(size_t)blockIdx.y * pc + i] = t;
}

// ------------------------------------------------------------------------------ host side

static int build_tree_tables(ttn_plan* p, const ttn_desc* d);
static int build_tree_merge(ttn_plan* p, const ttn_desc* d);

int build_tree_gemm(ttn_plan* p, const ttn_desc* d) {
  p->tgemm_ok = false;
  p->tg_tab_of.clear();
  const int n = d->n_vertices;
  if (n < 2) return TTN_OK;
  const bool cplx = d->is_complex != 0;
  const int NC = cplx ? 2 : 1;
  int maxchi = 1;
  for (int v = 0; v < n; ++v) {
    maxchi = std::max(maxchi, d->link_dim[v]);
    if (p->child_ptr[v + 1] - p->child_ptr[v] > 2 || p->nslices[v] > 8) return TTN_OK;
  }
  // row width in doubles: chi (real) or 2 chi (complex: interleaved (re, im) pairs)
  const int wreal = NC * maxchi;
  if (wreal > 64) return TTN_OK;
  const int W = wreal <= 16 ? 16 : (wreal <= 32 ? 32 : 64);
  const int H = W / NC; // entries per row
  TreeGemmDev& g = p->tgemm;
  g.n_vertices = n;
  g.W = W;
  g.root = d->root;
  const double* T = reinterpret_cast<const double*>(d->tensors);
  p->tg_frag_off.assign(n, 0);
  std::vector<double> blob;
  // element (re, im) of vertex v at flat index idx
  auto el = [&](int v, size_t idx, double* re, double* im) {
    *re = T[(d->tensor_ptr[v] + idx) * NC];
    *im = cplx ? T[(d->tensor_ptr[v] + idx) * NC + 1] : 0.0;
  };
  for (int v = 0; v < n; ++v) {
    const int nch = p->child_ptr[v + 1] - p->child_ptr[v];
    const int ns = p->nslices[v];
    const bool is_root = v == d->root;
    const int pdim = d->link_dim[v];
    const int ca = nch >= 1 ? d->link_dim[p->child[p->child_ptr[v]]] : 1;
    const int cb = nch == 2 ? d->link_dim[p->child[p->child_ptr[v] + 1]] : 1;
    p->tg_frag_off[v] = (int64_t)blob.size();
    double re, im;
    if (is_root) {
      // plain padded layout [slice][a][b] (b only for two children), parent dim 1; complex: (re, im) pairs, strides H
      const size_t per = (nch == 2 ? (size_t)H * H : (nch == 1 ? (size_t)H : 1)) * NC;
      std::vector<double> R(per * ns, 0.0);
      for (int s = 0; s < ns; ++s) {
        double* Rs = R.data() + (size_t)s * per;
        if (nch == 0) {
          el(v, s, &re, &im);
          Rs[0] = re;
          if (cplx) Rs[1] = im;
        } else if (nch == 1) {
          for (int a = 0; a < ca; ++a) {
            el(v, (size_t)s * ca + a, &re, &im);
            Rs[(size_t)a * NC] = re;
            if (cplx) Rs[(size_t)a * NC + 1] = im;
          }
        } else {
          for (int a = 0; a < ca; ++a)
            for (int b = 0; b < cb; ++b) {
              el(v, ((size_t)s * ca + a) * cb + b, &re, &im);
              Rs[((size_t)a * H + b) * NC] = re;
              if (cplx) Rs[((size_t)a * H + b) * NC + 1] = im;
            }
        }
      }
      blob.insert(blob.end(), R.begin(), R.end());
    } else if (nch == 0) {
      std::vector<double> L((size_t)ns * W, 0.0);
      for (int s = 0; s < ns; ++s)
        for (int j = 0; j < pdim; ++j) {
          el(v, (size_t)s * pdim + j, &re, &im);
          L[(size_t)s * W + (size_t)j * NC] = re;
          if (cplx) L[(size_t)s * W + (size_t)j * NC + 1] = im;
        }
      blob.insert(blob.end(), L.begin(), L.end());
    } else {
      // fragment order over K = (a, b') (or a') and N = parent index n'; complex entries are embedded as
      // [[re, im], [-im, re]] on (b', n') (two children) / (a', n') (one child), rows and columns interleaved
      const int K = nch == 2 ? H * W : W;
      std::vector<double> E((size_t)K * W), F((size_t)ns * K * W, 0.0);
      auto put = [&](size_t krow, int q, double r_, double i_) {
        // krow: row of the (re part); real networks: one row, one column
        if (!cplx) {
          E[krow * W + q] = r_;
        } else {
          E[krow * W + 2 * q] = r_;
          E[krow * W + 2 * q + 1] = i_;
          E[(krow + 1) * W + 2 * q] = -i_;
          E[(krow + 1) * W + 2 * q + 1] = r_;
        }
      };
      for (int s = 0; s < ns; ++s) {
        std::fill(E.begin(), E.end(), 0.0);
        if (nch == 2) {
          for (int a = 0; a < ca; ++a)
            for (int b = 0; b < cb; ++b)
              for (int q = 0; q < pdim; ++q) {
                el(v, (((size_t)s * ca + a) * cb + b) * pdim + q, &re, &im);
                put((size_t)a * W + (size_t)b * NC, q, re, im);
              }
        } else {
          for (int a = 0; a < ca; ++a)
            for (int q = 0; q < pdim; ++q) {
              el(v, ((size_t)s * ca + a) * pdim + q, &re, &im);
              put((size_t)a * NC, q, re, im);
            }
        }
        double* Fs = F.data() + (size_t)s * K * W;
        for (int kb = 0; kb < K / 4; ++kb)
          for (int nb = 0; nb < W / 8; ++nb)
            for (int ln = 0; ln < 32; ++ln)
              Fs[((size_t)kb * (W / 8) + nb) * 32 + ln] = E[(size_t)(4 * kb + (ln & 3)) * W + 8 * nb + (ln >> 2)];
      }
      blob.insert(blob.end(), F.begin(), F.end());
    }
  }
  double* d_blob;
  TTN_CUDA(cudaMalloc(&d_blob, std::max<size_t>(blob.size(), 2) * 8));
  p->allocs.push_back(d_blob);
  TTN_CUDA(cudaMemcpy(d_blob, blob.data(), blob.size() * 8, cudaMemcpyHostToDevice));
  g.blob = d_blob;
  int32_t* d_ns;
  TTN_CUDA(cudaMalloc(&d_ns, sizeof(int32_t) * n));
  p->allocs.push_back(d_ns);
  TTN_CUDA(cudaMemcpy(d_ns, p->nslices.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
  g.nslices = d_ns;
  p->tgemm_ok = true;
  if (int rc = build_tree_tables(p, d)) return rc;
  return build_tree_merge(p, d);
}

template <int W, int NCH, bool CPLX = false>
static int launch_vertex(const double* Ma, const double* Mb, double* Mout, const uint32_t* list, const int* cls_off,
                         const int* tile_off, const double* frags, int nsl, int pc, cudaStream_t s,
                         const uint32_t* ta = nullptr, const uint32_t* tb = nullptr) {
  constexpr size_t smem = ((size_t)(NCH == 2 ? 2 : 1) * TBM * (W + 4) + (size_t)TSTAGES * TBK * W) * 8;
  auto kern = tree_vertex_kernel<W, NCH, CPLX>;
  TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = pc / TBM + nsl;
  kern<<<grid, 256, smem, s>>>(Ma, Mb, Mout, list, cls_off, tile_off, frags, nsl, ta, tb);
  TTN_CUDA(cudaGetLastError());
  return TTN_OK;
}

template <int W>
static int launch_vertex_w(int nch, const double* Ma, const double* Mb, double* Mout, const uint32_t* list,
                           const int* cls_off, const int* tile_off, const double* frags, int nsl, int pc, cudaStream_t s,
                           const uint32_t* ta = nullptr, const uint32_t* tb = nullptr, bool cplx = false) {
  if (nch == 2 && cplx) return launch_vertex<W, 2, true>(Ma, Mb, Mout, list, cls_off, tile_off, frags, nsl, pc, s, ta, tb);
  return nch == 2 ? launch_vertex<W, 2>(Ma, Mb, Mout, list, cls_off, tile_off, frags, nsl, pc, s, ta, tb)
                  : launch_vertex<W, 1>(Ma, Mb, Mout, list, cls_off, tile_off, frags, nsl, pc, s, ta, tb);
}

int launch_tree_gemm(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double* d_partial,
                     int* n_partial, cudaStream_t s, int* n_launches) {
  *n_partial = 0;
  if (src.npts == 0) return TTN_OK;
  if (!p->tgemm_ok) {
    set_error("tree GEMM kernel requested but the network is not a real tree with <= 2 children per vertex and chi <= 64");
    return TTN_ERR_UNSUPPORTED;
  }
  const TreeGemmDev& g = p->tgemm;
  const int n = g.n_vertices, W = g.W;
  // chunk size: the per-vertex tile count (<= PC/128 + 8 classes) fills a whole number of waves.  Narrow rows
  // (W = 16, 32) take 4x / 2x the points per chunk: their per-vertex launches are otherwise too small to
  // cover the launch latency (measured: a 60-vertex comb tree at chi = 8 and chi = 16 ran at the same
  // 258 M points/s); the message workspace per chunk stays what a W = 64 tree of the same size takes.
  // The message workspace is n x PC x W doubles: keep it under 8 GB for trees of many vertices.
  const int64_t pc_mem = std::max<int64_t>((int64_t)p->sm_count * TBM, (int64_t)(8e9 / ((double)n * W * 8)) / TBM * TBM);
  const int PC = (int)std::min<int64_t>(std::min<int64_t>((int64_t)(4 * (64 / W) * p->sm_count - 8) * TBM, pc_mem),
                                        (src.npts + TBM - 1) / TBM * TBM);
  const size_t msg_b = (size_t)PC * W * 8;
  const size_t slices_b = ((size_t)n * PC + 255) / 256 * 256;
  const size_t lists_b = (size_t)n * PC * 4;
  const size_t offs_b = (size_t)n * TOFF * 4 * 4 + 256; // cls_off, tile_off, counts, cursor
  const int n_tab = (int)p->tg_tab.size();
  const size_t tabidx_b = (size_t)n_tab * PC * 4;
  const size_t need = (size_t)n * msg_b + slices_b + lists_b + offs_b + tabidx_b;
  if (st.gemm_bytes < need) {
    if (st.d_gemm) cudaFree(st.d_gemm);
    st.d_gemm = nullptr;
    st.gemm_bytes = 0;
    TTN_CUDA(cudaMalloc(&st.d_gemm, need));
    st.gemm_bytes = need;
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(st.d_gemm);
  double* msgs = reinterpret_cast<double*>(base);
  uint8_t* slices = base + (size_t)n * msg_b;
  uint32_t* lists = reinterpret_cast<uint32_t*>(slices + slices_b);
  int* cls_off = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(lists) + lists_b);
  int* tile_off = cls_off + (size_t)n * TOFF;
  int* counts = tile_off + (size_t)n * TOFF;
  int* cursor = counts + (size_t)n * TOFF;
  uint32_t* tabidx = reinterpret_cast<uint32_t*>(base + (size_t)n * msg_b + slices_b + lists_b + offs_b);
  const int do_sum = d_partial != nullptr;
  const int64_t n_chunks = (src.npts + PC - 1) / PC;
  const int root_nch = p->child_ptr[g.root + 1] - p->child_ptr[g.root];
  const bool cplx = p->info.is_complex != 0;
  const bool fast_root = cplx || (root_nch == 2 && W <= 32 && (size_t)p->nslices[g.root] * W * W * 8 <= 96 * 1024);
  const int root_blocks = fast_root ? (PC + 127) / 128 : (PC * 32 + 255) / 256;
  double* big_partial = nullptr;
  if (do_sum) {
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
  auto M = [&](int v) { return msgs + (size_t)v * PC * W; };
  // a tabulated child is read straight from its table through the per-point row index (no message copy), unless
  // its consumer is the warp-per-point root kernel
  auto tab_direct = [&](int c) {
    return c >= 0 && !p->tg_tab_of.empty() && p->tg_tab_of[c] >= 0 && (p->parent[c] != g.root || fast_root);
  };
  auto msg_of = [&](int c) -> const double* { return tab_direct(c) ? p->tg_tab[p->tg_tab_of[c]] : M(c); };
  auto idx_of = [&](int c) -> const uint32_t* { return tab_direct(c) ? tabidx + (size_t)p->tg_tab_of[c] * PC : nullptr; };
  const TgClass* gv = reinterpret_cast<const TgClass*>(p->tg_gv);
  const int ngv = p->tg_ngv;
  const int nseg = (PC + TSEG - 1) / TSEG;
  for (int64_t ck = 0; ck < n_chunks; ++ck) {
    const int64_t p0 = ck * PC;
    tree_digits_kernel<<<(PC + 255) / 256, 256, 0, s>>>(p->digits, src, p0, PC, n, slices, p->d_err, p->tg_zero_v, p->tg_n_zero);
    *n_launches += 1;
    if (n_tab > 0) {
      tree_tabidx_kernel<<<dim3((PC + 255) / 256, n_tab), 256, 0, s>>>(slices, PC, reinterpret_cast<const TgTabRef*>(p->tg_tabrefs), tabidx);
      *n_launches += 1;
    }
    if (ngv > 0) {
      TTN_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)ngv * TOFF, s));
      tree_count_kernel<<<dim3(nseg, ngv), 256, 0, s>>>(slices, PC, gv, counts);
      tree_offsets_kernel<<<(ngv + 63) / 64, 64, 0, s>>>(gv, ngv, counts, cls_off, tile_off, cursor);
      tree_scatter_kernel<<<dim3(nseg, ngv), 256, 0, s>>>(slices, PC, gv, cursor, lists);
      *n_launches += 3;
    }
    for (int oi = 0; oi < n; ++oi) {
      const int v = p->post[oi];
      const int nch = p->child_ptr[v + 1] - p->child_ptr[v];
      const double* blob = g.blob + p->tg_frag_off[v];
      const int ca = nch >= 1 ? p->child[p->child_ptr[v]] : -1;
      const int cb = nch == 2 ? p->child[p->child_ptr[v] + 1] : -1;
      const int tab = p->tg_tab_of.empty() ? -1 : p->tg_tab_of[v];
      if (tab == -2) continue; // inside a tabulated subtree
      if (tab >= 0) {
        if (!tab_direct(v)) {
          tree_table_kernel<<<(unsigned)(((int64_t)PC * (W / 2) + 255) / 256), 256, 0, s>>>(slices, PC, p->tg_tab_vs[tab], p->tg_tab_ns[tab],
                                                                                            p->tg_tab[tab], W, M(v));
          *n_launches += 1;
        }
        continue;
      }
      const int role = p->tg_role[v];
      if (role == 1) continue; // absorbed into the merged run that ends higher up
      if (v == g.root) {
        double* part = do_sum ? big_partial + 2 * ck * root_blocks : nullptr;
        if (cplx) {
          tree_root_c_kernel<<<root_blocks, 128, 0, s>>>(nch, nch == 2 ? msg_of(ca) : nullptr, nch >= 1 ? msg_of(nch == 2 ? cb : ca) : nullptr,
                                                         slices + (size_t)v * PC, PC, p0, src.npts, blob, W, d_out, part, do_sum, src,
                                                         nch == 2 ? idx_of(ca) : nullptr, nch >= 1 ? idx_of(nch == 2 ? cb : ca) : nullptr);
        } else if (fast_root) {
          const size_t sm = (size_t)p->nslices[v] * W * W * 8;
          if (W == 16) {
            TTN_CUDA(cudaFuncSetAttribute(tree_root2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            tree_root2_kernel<16><<<root_blocks, 128, sm, s>>>(msg_of(ca), msg_of(cb), slices + (size_t)v * PC, PC, p0, src.npts, blob,
                                                               p->nslices[v], d_out, part, do_sum, src, idx_of(ca), idx_of(cb));
          } else {
            TTN_CUDA(cudaFuncSetAttribute(tree_root2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            tree_root2_kernel<32><<<root_blocks, 128, sm, s>>>(msg_of(ca), msg_of(cb), slices + (size_t)v * PC, PC, p0, src.npts, blob,
                                                               p->nslices[v], d_out, part, do_sum, src, idx_of(ca), idx_of(cb));
          }
        } else {
          tree_root_kernel<<<root_blocks, 256, 0, s>>>(nch, nch == 2 ? M(ca) : nullptr, nch == 2 ? M(cb) : (nch == 1 ? M(ca) : nullptr),
                                                       slices + (size_t)v * PC, PC, p0, src.npts, blob, W, d_out, part, do_sum, src);
        }
      } else if (nch == 0) {
        tree_leaf_kernel<<<(unsigned)(((int64_t)PC * (W / 2) + 255) / 256), 256, 0, s>>>(slices + (size_t)v * PC, PC, blob, W, M(v));
      } else {
        // role 2: a merged run of single-child vertices — one GEMM with the run's class matrices, fed by the message below it
        const int vb = role == 2 ? p->tg_in[v] : (nch == 2 ? cb : ca);
        const double* Ma = nch == 2 ? msg_of(ca) : nullptr;
        const uint32_t* ia = nch == 2 ? idx_of(ca) : nullptr;
        const double* Mb = msg_of(vb);
        const uint32_t* ib = idx_of(vb);
        const double* fr = role == 2 ? p->tg_mblob + p->tg_mfrag_off[v] : blob;
        const int nsl = role == 2 ? p->tg_mnsl[v] : p->nslices[v];
        int rc;
        if (W == 16) rc = launch_vertex_w<16>(nch, Ma, Mb, M(v), lists + (size_t)v * PC, cls_off + v * TOFF, tile_off + v * TOFF, fr, nsl, PC, s, ia, ib, cplx);
        else if (W == 32) rc = launch_vertex_w<32>(nch, Ma, Mb, M(v), lists + (size_t)v * PC, cls_off + v * TOFF, tile_off + v * TOFF, fr, nsl, PC, s, ia, ib, cplx);
        else rc = launch_vertex_w<64>(nch, Ma, Mb, M(v), lists + (size_t)v * PC, cls_off + v * TOFF, tile_off + v * TOFF, fr, nsl, PC, s, ia, ib, cplx);
        if (rc) return rc;
      }
      *n_launches += 1;
    }
    TTN_CUDA(cudaGetLastError());
  }
  if (do_sum) {
    int rc = launch_sum_partials(p, big_partial, (int)(n_chunks * root_blocks), 1, d_partial, s);
    if (rc) return rc;
    *n_launches += 1;
  }
  *n_partial = do_sum ? 1 : 0;
  return TTN_OK;
}

// Evaluation plan: which vertices run as GEMMs, and which runs of single-child vertices are merged.  A run of k
// consecutive single-child vertices (a tooth of a comb above its subtree table, the backbone of a chain-like tree)
// acts on the message below it as ONE of 2^(bits) W x W matrices, bits = the slice bits of its members: the products
// are formed here once (long double, rounded once) and the run costs one GEMM launch and 2 W^2 flops per point
// instead of k of each — the tree kernel's version of the chain kernels' group merging (k_chain_mma.cu).
static int build_tree_merge(ttn_plan* p, const ttn_desc* d) {
  const int n = d->n_vertices;
  const TreeGemmDev& g = p->tgemm;
  const int W = g.W;
  p->tg_role.assign(n, 0);
  p->tg_in.assign(n, -1);
  p->tg_mnsl.assign(n, 0);
  p->tg_mfrag_off.assign(n, 0);
  auto nchild = [&](int v) { return p->child_ptr[v + 1] - p->child_ptr[v]; };
  auto tab_of = [&](int v) { return p->tg_tab_of.empty() ? -1 : p->tg_tab_of[v]; };
  auto bits_of = [&](int v) {
    int b = 0;
    while ((1 << b) < p->nslices[v]) ++b;
    return b;
  };
  auto mergeable = [&](int v) {
    const int ns = p->nslices[v];
    return v != d->root && nchild(v) == 1 && tab_of(v) == -1 && (ns & (ns - 1)) == 0;
  };
  const bool merge_on = !(getenv("TTN_TREE_MERGE") && atoi(getenv("TTN_TREE_MERGE")) == 0);
  const double* T = reinterpret_cast<const double*>(d->tensors);
  std::vector<double> mblob;
  typedef long double ld;
  for (int oi = 0; oi < n && merge_on; ++oi) { // post order: a run is met at its bottom member first
    const int b0 = p->post[oi];
    if (!mergeable(b0) || p->tg_role[b0] != 0) continue;
    std::vector<int> mem{b0};
    int bits = bits_of(b0);
    for (int u = d->parent[b0]; u >= 0 && mergeable(u) && (int)mem.size() < 4 && bits + bits_of(u) <= 4; u = d->parent[u]) {
      mem.push_back(u);
      bits += bits_of(u);
    }
    if (mem.size() < 2) continue;
    const int top = mem.back(), ncls = 1 << bits;
    const int in_v = p->child[p->child_ptr[b0]];
    // class matrices: row vector (child dim of the bottom member) x E_m0[s0] x E_m1[s1] ... (parent dim of the top);
    // complex entries embedded as [[re, im], [-im, re]] with interleaved rows / columns (the embedding of a product
    // is the product of the embeddings)
    const int NCm = d->is_complex ? 2 : 1;
    std::vector<double> F((size_t)ncls * W * W, 0.0);
    for (int c = 0; c < ncls; ++c) {
      std::vector<ld> cur((size_t)W * W, 0.0L);
      int rows = NCm * d->link_dim[in_v], cols = rows;
      for (int i = 0; i < rows; ++i) cur[(size_t)i * W + i] = 1.0L;
      int shift = 0;
      for (int m : mem) {
        const int bm = bits_of(m), sm = (c >> shift) & ((1 << bm) - 1);
        shift += bm;
        const int ca = cols / NCm, pd = d->link_dim[m];
        const double* Tm = T + (d->tensor_ptr[m] + (size_t)sm * ca * pd) * NCm;
        std::vector<ld> Em((size_t)W * W, 0.0L); // embedded member matrix
        for (int k = 0; k < ca; ++k)
          for (int q = 0; q < pd; ++q) {
            const ld re = Tm[((size_t)k * pd + q) * NCm], im = NCm == 2 ? Tm[((size_t)k * pd + q) * 2 + 1] : 0.0L;
            if (NCm == 1) {
              Em[(size_t)k * W + q] = re;
            } else {
              Em[(size_t)(2 * k) * W + 2 * q] = re;
              Em[(size_t)(2 * k) * W + 2 * q + 1] = im;
              Em[(size_t)(2 * k + 1) * W + 2 * q] = -im;
              Em[(size_t)(2 * k + 1) * W + 2 * q + 1] = re;
            }
          }
        std::vector<ld> nxt((size_t)W * W, 0.0L);
        for (int i = 0; i < rows; ++i)
          for (int k = 0; k < NCm * ca; ++k) {
            const ld a = cur[(size_t)i * W + k];
            if (a == 0.0L) continue;
            for (int q = 0; q < NCm * pd; ++q) nxt[(size_t)i * W + q] += a * Em[(size_t)k * W + q];
          }
        cur.swap(nxt);
        cols = NCm * pd;
      }
      double* Fs = F.data() + (size_t)c * W * W;
      for (int kb = 0; kb < W / 4; ++kb)
        for (int nb = 0; nb < W / 8; ++nb)
          for (int ln = 0; ln < 32; ++ln)
            Fs[((size_t)kb * (W / 8) + nb) * 32 + ln] = (double)cur[(size_t)(4 * kb + (ln & 3)) * W + 8 * nb + (ln >> 2)];
    }
    for (size_t i = 0; i + 1 < mem.size(); ++i) p->tg_role[mem[i]] = 1;
    p->tg_role[top] = 2;
    p->tg_in[top] = in_v;
    p->tg_mnsl[top] = ncls;
    p->tg_mfrag_off[top] = (int64_t)mblob.size();
    mblob.insert(mblob.end(), F.begin(), F.end());
    const double fmac = d->is_complex ? 8.0 : 2.0;
    for (int m : mem) p->tgemm_flops_exec -= fmac * d->link_dim[p->child[p->child_ptr[m]]] * d->link_dim[m];
    p->tgemm_flops_exec += fmac * d->link_dim[in_v] * d->link_dim[top];
  }
  if (!mblob.empty()) {
    TTN_CUDA(cudaMalloc(&p->tg_mblob, mblob.size() * 8));
    p->allocs.push_back(p->tg_mblob);
    TTN_CUDA(cudaMemcpy(p->tg_mblob, mblob.data(), mblob.size() * 8, cudaMemcpyHostToDevice));
  }
  // the vertices that run as GEMMs (row lists needed): merged tops with their members, other inner vertices as is
  std::vector<TgClass> gv;
  for (int v = 0; v < n; ++v) {
    if (v == d->root || nchild(v) == 0 || tab_of(v) != -1 || p->tg_role[v] == 1) continue;
    TgClass c{};
    c.v = v;
    if (p->tg_role[v] == 2) {
      std::vector<int> mem; // walk down from the top to the bottom member
      for (int u = v; u != p->tg_in[v]; u = p->child[p->child_ptr[u]]) mem.push_back(u);
      std::reverse(mem.begin(), mem.end());
      int shift = 0;
      for (int m : mem) {
        c.mem_v[c.n_mem] = m;
        c.mem_shift[c.n_mem] = shift;
        c.n_mem++;
        shift += bits_of(m);
      }
      c.nsl = p->tg_mnsl[v];
    } else {
      c.nsl = p->nslices[v];
      c.n_mem = 1;
      c.mem_v[0] = v;
      c.mem_shift[0] = 0;
    }
    gv.push_back(c);
  }
  {
    std::vector<int32_t> zv;
    for (int v = 0; v < n; ++v)
      if (d->site_ptr[v + 1] - d->site_ptr[v] != 1) zv.push_back(v);
    p->tg_n_zero = (int)zv.size();
    if (!zv.empty()) {
      TTN_CUDA(cudaMalloc(&p->tg_zero_v, zv.size() * sizeof(int32_t)));
      p->allocs.push_back(p->tg_zero_v);
      TTN_CUDA(cudaMemcpy(p->tg_zero_v, zv.data(), zv.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
  }
  if (!p->tg_tab.empty()) {
    std::vector<TgTabRef> refs(p->tg_tab.size());
    for (size_t i = 0; i < refs.size(); ++i) refs[i] = TgTabRef{p->tg_tab_vs[i], p->tg_tab_ns[i]};
    TTN_CUDA(cudaMalloc(&p->tg_tabrefs, refs.size() * sizeof(TgTabRef)));
    p->allocs.push_back(p->tg_tabrefs);
    TTN_CUDA(cudaMemcpy(p->tg_tabrefs, refs.data(), refs.size() * sizeof(TgTabRef), cudaMemcpyHostToDevice));
  }
  p->tg_ngv = (int)gv.size();
  if (!gv.empty()) {
    TTN_CUDA(cudaMalloc(&p->tg_gv, gv.size() * sizeof(TgClass)));
    p->allocs.push_back(p->tg_gv);
    TTN_CUDA(cudaMemcpy(p->tg_gv, gv.data(), gv.size() * sizeof(TgClass), cudaMemcpyHostToDevice));
  }
  return TTN_OK;
}

// Subtree message tables.  sub_bits(v) = sum of log2(nslices) over v's subtree; every non-root vertex
// whose subtree has >= 2 vertices, only power-of-two slice counts and sub_bits <= the budget (2^bits
// rows of W doubles <= 32 MB, bits <= 16; TTN_TREE_TABLE_BITS overrides, 0 disables) and whose parent
// does not qualify becomes a TABLE vertex: its 2^bits possible messages are computed here, once, by
// running the ordinary vertex-by-vertex kernels over the enumerated digit settings, and an evaluation
// replaces the whole subtree by one row gather per point (tree_table_kernel).
static int build_tree_tables(ttn_plan* p, const ttn_desc* d) {
  const int n = d->n_vertices;
  const TreeGemmDev& g = p->tgemm;
  const int W = g.W;
  // rows per table: W = 16: 2^20, W = 32: 2^19 (128 MB: a whole 20-vertex tooth of a comb is ONE row gather), W = 64: 2^16
  // (32 MB; config 3's subtrees jump from 15 to 31 bits); all tables of a plan together <= 1 GB, and the build workspace
  // (every message of the largest subtree over all its settings) must fit half of the free device memory
  int budget = 20;
  while (budget > 0 && ((size_t)1 << budget) * W * 8 > ((size_t)(W <= 32 ? 128 : 32) << 20)) --budget;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
    cudaGetLastError();
    free_b = (size_t)8 << 30;
  }
  if (const char* e = getenv("TTN_TREE_TABLE_BITS")) budget = std::min(atoi(e), 20);
  p->tg_tab_of.assign(n, -1);
  p->tg_tab.clear();
  p->tg_tab_vs.clear();
  p->tg_tab_ns.clear();
  p->tg_tab_bits.clear();
  // executed flops per point (the flop rule over the vertices that still run)
  auto vertex_flops = [&](int v) {
    const int nch = p->child_ptr[v + 1] - p->child_ptr[v];
    const double pd = v == d->root ? 1.0 : d->link_dim[v];
    if (nch == 0) return 0.0;
    const double fm = d->is_complex ? 8.0 : 2.0; // flops per (complex) multiply-add
    const double ca = d->link_dim[p->child[p->child_ptr[v]]];
    if (nch == 1) return fm * ca * pd;
    const double cb = d->link_dim[p->child[p->child_ptr[v] + 1]];
    return fm * (ca * cb * pd + cb * pd);
  };
  p->tgemm_flops_exec = 0.0;
  for (int v = 0; v < n; ++v) p->tgemm_flops_exec += vertex_flops(v);
  if (budget <= 0) return TTN_OK;
  std::vector<int> bits(n, 0), size(n, 1), frontier;
  std::vector<char> ok(n, 1);
  for (;; --budget) { // lower the per-table budget until all tables together stay under 1 GB
    if (budget <= 0) return TTN_OK;
    std::fill(ok.begin(), ok.end(), 1);
    std::fill(size.begin(), size.end(), 1);
    for (int oi = 0; oi < n; ++oi) { // post order: children first
      const int v = p->post[oi];
      const int ns = p->nslices[v];
      if (ns & (ns - 1)) ok[v] = 0;
      int b = 0;
      while ((1 << b) < ns) ++b;
      bits[v] = b;
      for (int c = p->child_ptr[v]; c < p->child_ptr[v + 1]; ++c) {
        const int u = p->child[c];
        bits[v] += bits[u];
        size[v] += size[u];
        ok[v] = ok[v] && ok[u];
      }
      if (bits[v] > budget) ok[v] = 0;
    }
    frontier.clear();
    size_t total = 0;
    for (int v = 0; v < n; ++v) {
      if (v == d->root || !ok[v] || size[v] < 2) continue;
      const int par = d->parent[v];
      if (par >= 0 && par != d->root && ok[par]) continue; // the parent's table covers it
      frontier.push_back(v);
      total += ((size_t)1 << bits[v]) * W * 8;
    }
    size_t ws = 0; // build workspace of the largest table at this budget
    for (int v : frontier) ws = std::max(ws, (size_t)std::max(TBM, 1 << bits[v]) * ((size_t)size[v] * W * 8 + (size_t)n * 5));
    if (total <= ((size_t)1 << 30) && ws <= free_b / 2) break;
  }
  if (frontier.empty()) return TTN_OK;
  // workspace for the largest table
  int maxb = 0;
  for (int v : frontier) maxb = std::max(maxb, bits[v]);
  const int PCmax = std::max(TBM, 1 << maxb);
  int maxsz = 0;
  for (int v : frontier) maxsz = std::max(maxsz, size[v]);
  double* d_msgs = nullptr;
  uint8_t* d_slices = nullptr;
  uint32_t* d_lists = nullptr;
  int* d_offs = nullptr;
  TTN_CUDA(cudaMalloc(&d_msgs, (size_t)maxsz * PCmax * W * 8));
  TTN_CUDA(cudaMalloc(&d_slices, (size_t)n * PCmax));
  TTN_CUDA(cudaMalloc(&d_lists, (size_t)n * PCmax * 4));
  TTN_CUDA(cudaMalloc(&d_offs, ((size_t)n * 9 * 2 + 32) * 4));
  int rc = TTN_OK;
  for (size_t fi = 0; fi < frontier.size() && rc == TTN_OK; ++fi) {
    const int v = frontier[fi];
    const int PC = std::max(TBM, 1 << bits[v]);
    // subtree vertices in post order, their message slots and bit offsets
    std::vector<int> sub, slot(n, -1);
    std::vector<int32_t> vs;
    {
      std::vector<char> in(n, 0);
      in[v] = 1;
      // a vertex is in the subtree iff its parent is (walk the post order backwards: parents first)
      for (int oi = n - 1; oi >= 0; --oi) {
        const int u = p->post[oi];
        if (u != v && d->parent[u] >= 0 && in[d->parent[u]]) in[u] = 1;
      }
      int off = 0;
      for (int oi = 0; oi < n; ++oi) {
        const int u = p->post[oi];
        if (!in[u]) continue;
        slot[u] = (int)sub.size();
        sub.push_back(u);
        const int ns = p->nslices[u];
        int b = 0;
        while ((1 << b) < ns) ++b;
        vs.push_back(u);
        vs.push_back(off);
        vs.push_back(ns - 1);
        off += b;
      }
    }
    int32_t* d_vs;
    if (cudaMalloc(&d_vs, vs.size() * 4) != cudaSuccess) {
      rc = TTN_ERR_CUDA;
      break;
    }
    p->allocs.push_back(d_vs);
    cudaMemcpy(d_vs, vs.data(), vs.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(d_slices, 0, (size_t)n * PC);
    int* cls_off = d_offs;
    int* tile_off = d_offs + (size_t)n * 9 + 16;
    tree_enum_kernel<<<(PC + 255) / 256, 256>>>(d_slices, PC, d_vs, (int)sub.size());
    tree_classify_kernel<<<n, 1024>>>(d_slices, PC, g.nslices, d_lists, cls_off, tile_off);
    auto M = [&](int u) { return d_msgs + (size_t)slot[u] * PC * W; };
    for (int u : sub) {
      const int nch = p->child_ptr[u + 1] - p->child_ptr[u];
      const double* blob = g.blob + p->tg_frag_off[u];
      if (nch == 0) {
        tree_leaf_kernel<<<(unsigned)(((int64_t)PC * (W / 2) + 255) / 256), 256>>>(d_slices + (size_t)u * PC, PC, blob, W, M(u));
      } else {
        const int ca = p->child[p->child_ptr[u]];
        const int cb = nch == 2 ? p->child[p->child_ptr[u] + 1] : -1;
        const double* Ma = nch == 2 ? M(ca) : nullptr;
        const double* Mb = nch == 2 ? M(cb) : M(ca);
        const bool cx = d->is_complex != 0;
        if (W == 16) rc = launch_vertex_w<16>(nch, Ma, Mb, M(u), d_lists + (size_t)u * PC, cls_off + u * 9, tile_off + u * 9, blob, p->nslices[u], PC, 0, nullptr, nullptr, cx);
        else if (W == 32) rc = launch_vertex_w<32>(nch, Ma, Mb, M(u), d_lists + (size_t)u * PC, cls_off + u * 9, tile_off + u * 9, blob, p->nslices[u], PC, 0, nullptr, nullptr, cx);
        else rc = launch_vertex_w<64>(nch, Ma, Mb, M(u), d_lists + (size_t)u * PC, cls_off + u * 9, tile_off + u * 9, blob, p->nslices[u], PC, 0, nullptr, nullptr, cx);
        if (rc) break;
      }
    }
    if (rc) break;
    double* d_tab;
    const size_t rows = (size_t)1 << bits[v];
    if (cudaMalloc(&d_tab, rows * W * 8) != cudaSuccess) {
      rc = TTN_ERR_CUDA;
      break;
    }
    p->allocs.push_back(d_tab);
    cudaMemcpy(d_tab, M(v), rows * W * 8, cudaMemcpyDeviceToDevice);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      rc = TTN_ERR_CUDA;
      break;
    }
    const int ti = (int)p->tg_tab.size();
    p->tg_tab.push_back(d_tab);
    p->tg_tab_vs.push_back(d_vs);
    p->tg_tab_ns.push_back((int)sub.size());
    p->tg_tab_bits.push_back(bits[v]);
    for (int u : sub) {
      p->tg_tab_of[u] = u == v ? ti : -2;
      p->tgemm_flops_exec -= vertex_flops(u);
    }
  }
  cudaFree(d_msgs);
  cudaFree(d_slices);
  cudaFree(d_lists);
  cudaFree(d_offs);
  if (rc == TTN_ERR_CUDA) set_error(std::string("tree tables: ") + cudaGetErrorString(cudaGetLastError()));
  return rc;
}

} // namespace ttn
