// K9 — narrow chains (real chi <= 4, complex chi <= 2): the HBM-bound end of the path.
//
// At small bond dimension the contraction of one point (project + scalar,
// src/itensornetworkfunction.jl:84-106) is a handful of flops against 8 bytes per coordinate read
// and 8 / 16 bytes written, so the kernel has to stream points at HBM speed.  A point picks one
// slice per vertex, hence a GROUP of k consecutive chain vertices acts on the state as one of
// 2^(k b) tiny matrices (b = slice bits per vertex).  The plan tabulates those products (long
// double, rounded once) for groups of up to 16 stream bits:
//     leaf group  -> 2^gL row vectors      v   = Lt[s_0]
//     middle groups -> 2^g  H x H matrices  v  <- v * Mt_g[s_g]
//     root group  -> 2^gR column vectors   out = v . Rt[s_last]
// sized so that ALL tables of the chain sit in the shared memory of one CTA (<= 200 KB).  The
// kernel is one persistent CTA per SM: tables are loaded once, then every thread streams several
// points per tile — coordinates with coalesced vector loads issued one tile ahead; K1 (k_digits.cuh
// semantics) as the bits of floor(x 2^L), shifted into place when a coordinate's digits sit on
// consecutive stream bits and deposited by a 6-step shift/select network when they are scattered
// (interleaved dimensions, Real + Imag index on one vertex), into a 128-bit packed slice stream;
// one table lookup per group with the state in registers; one coalesced store.
//
// Table layouts (make_table_image picks per plan):
//   replicated — every 16-byte chunk of an entry once per lane of a quarter-warp in one 128-byte line
//                (8-byte entries: per lane of a half-warp): no lookup has a bank conflict.  A 60-bit
//                chi = 1 chain: 8 lookups of 7-8 bits, 5.7-6.0 TB/s of the 24 algorithmic B/point;
//                chi = 2: 10 lookups of 6 bits, shared-memory pipe at its conflict-free rate;
//   plain      — entries back to back, their 16-byte chunks XOR-swizzled by the entry index (when the
//                replicated tables do not fit: chi = 4 on long chains).
// Measurements, ncu summaries and the history of the kernel: DESIGN.md, profiles/r01_table_*.txt.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "k_chain_common.cuh"

namespace ttn {

// ------------------------------------------------------------------------------ device side

// Table entries of C = 2, 4 or 8 16-byte chunks (H x H matrices, wide vectors): entry s starts at bank group
// (s C) mod 8, so chunk j of EVERY entry would sit in the same few banks and the lanes of a quarter-warp —
// which read chunk j of different entries at once — would serialise (measured: 128-byte entries ran 3x
// slower than the conflict model predicts).  Chunk j is therefore stored at position j ^ x(s), with x taken
// from the entry-index bits just above those that pick the start group: a random s then spreads every
// chunk over all 8 bank groups.  Same function on the host (image builder) and the device.
template <int N>
__host__ __device__ __forceinline__ uint32_t entry_swizzle(uint32_t s) {
  constexpr int C = N / 2;
  if constexpr (C <= 1) return 0u;
  else if constexpr (C == 2) return (s >> 2) & 1u;
  else if constexpr (C == 4) return (s >> 1) & 3u;
  else return s & 7u;
}

// dst = entry s (N doubles) of the table at shared address `base` (128-byte aligned)
template <int N>
__device__ __forceinline__ void lds_entry(uint32_t base, uint32_t s, double (&dst)[N]) {
  if constexpr (N == 1) {
    dst[0] = lds64(base + s * 8u);
  } else {
    static_assert(N % 2 == 0 && N <= 16, "entry length");
    const uint32_t a = (base + s * (uint32_t)(N * 8)) ^ (entry_swizzle<N>(s) << 4); // entry-aligned: + == ^
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const double2 t = lds128(a ^ (16u * i));
      dst[2 * i] = t.x;
      dst[2 * i + 1] = t.y;
    }
  }
}

// Parallel bit deposit (the "expand" network of Hacker's Delight 7-5, 64-bit): bit j of x moves to the j-th
// lowest set bit of m0.  The six step masks mv depend only on m0 and come from the host (expand_masks).
__device__ __forceinline__ uint64_t expand64(uint64_t x, const uint64_t* mv, uint64_t m0) {
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    const uint64_t m = mv[i];
    const uint64_t t = x << (1 << i);
    x = (x & ~m) | (t & m);
  }
  return x & m0;
}

__device__ __forceinline__ double2 ldg_nc_f64x2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// W2: the packed slice stream needs its second 64-bit word (more than 64 slice bits); chains of up to 64
// bits (config 2's layout: 2 x 30) keep the stream in one register pair and shift half as much.
// NCV: coordinates fetched with vector loads one tile ahead — 2: 2-D AoS input (one 128-bit load per point),
// 1: 1-D input (one 64-bit load per point), 0: any layout / grid generator / index-setting mode (load_coord).
// REP: tables in the replicated, conflict-free layout (make_table_image).
template <int H, bool CPLX, int NT, int MINB, int PPT, int NCV, bool W2, bool REP>
__global__ void __launch_bounds__(NT, MINB)
    chain_table_kernel(ChainTabDev ct, DigitTable dg, CoordSource src, double* __restrict__ out, int* err,
                       double* __restrict__ partial, int do_sum) {
  constexpr int E = CPLX ? 2 : 1, HE = H * E, MM = H * H * E;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ double red[2][NT / 32];
  __shared__ Digit2 s_d2[kFeMaxSites];
  // The base-3 / base-4 digit table lives in DYNAMIC shared memory behind the group tables, only for networks that need
  // it.  As a static array (5 KB) it pushed the chi = 1 bench instance — 192 KB of tables + 2.9 KB static + 1 KB reserved —
  // past the 196 KB shared-memory carveout into the 228 KB one, i.e. from 60 KB of L1 to 28 KB, and the HBM-bound
  // kernel ran 17 % slower (0.470 vs 0.400 ms per 1e8 points; scripts/microbench/ab_table.sh against the round-start build).
  Digit4* s_d4 = reinterpret_cast<Digit4*>(smem + (((size_t)ct.total_doubles * 8 + 15) & ~(size_t)15));
  __shared__ int s_cptr[TTN_MAX_COORDS + 1];

  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i <= dg.n_coords; i += NT) s_cptr[i] = dg.coord_ptr[i];
  for (int i = tid; i < (ct.k1_generic == 1 ? dg.n_sites : 0); i += NT) s_d4[i] = make_digit4(dg, i);
  for (int i = tid; i < (ct.k1_generic == 0 ? dg.n_sites : 0); i += NT) {
    const DigitEntry e = dg.entries[i];
    Digit2 d2;
    d2.thr1 = dg.thr[e.thr_off + 1];
    d2.sh = (uint32_t)e.shift;
    d2.wv = ((uint32_t)e.site << 16) | ((uint32_t)e.word << 8) | (uint32_t)e.stride;
    s_d2[i] = d2;
  }
  {
    const double2* g = reinterpret_cast<const double2*>(ct.image);
    double2* s = reinterpret_cast<double2*>(smem);
    for (int i = tid; i < ct.total_doubles / 2; i += NT) s[i] = __ldg(g + i);
  }
  __syncthreads();

  // REP: this lane's copy inside every 128-byte line
  const uint32_t sbase = smem_u32(smem) + (REP ? (HE > 1 ? (uint32_t)(lane & 7) * 16u : (uint32_t)(lane & 15) * 8u) : 0u);
  const int G = ct.n_groups;
  auto entry = [&](uint32_t base, uint32_t s, auto& dst) {
    constexpr int N = sizeof(dst) / sizeof(double);
    if constexpr (REP) {
      if constexpr (N == 1) dst[0] = lds64(base + (s << 7));
      else {
        const uint32_t a = base + s * (uint32_t)(N / 2 * 128);
#pragma unroll
        for (int j = 0; j < N / 2; ++j) {
          const double2 t = lds128(a + 128u * j);
          dst[2 * j] = t.x;
          dst[2 * j + 1] = t.y;
        }
      }
    } else {
      lds_entry<N>(base, s, dst);
    }
  };
  constexpr int64_t TILE = (int64_t)NT * PPT;
  const int64_t n_tiles = (src.npts + TILE - 1) / TILE;
  double sum_re = 0.0, sum_im = 0.0;

  // K1 for coordinate slot c of the PPT points of this thread (same arithmetic as compute_words of
  // the team-sorted kernel): run fast path or the tabulated greedy loop, bits OR-ed into the stream
  uint64_t w0[PPT], w1[PPT];
  // Domain violations (x < 0, NaN: the reference's loop never terminates) are only FLAGGED here — one atomic
  // per thread at the end of the kernel; the call then fails with TTN_ERR_DOMAIN.  Such an x yields digit 0
  // everywhere (the compares are false, the conversion saturates at 0), so no table index goes out of range.
  bool bad = false;
  auto add_coord = [&](int c, double (&x)[PPT], int64_t p0) {
#pragma unroll
    for (int k = 0; k < PPT; ++k) bad = bad || !coord_in_domain(x[k]);
    if (ct.run_kind[c] > 0 && !src.digits) {
      // digits = bits of floor(x 2^L), x >= 1 saturating to all ones (see build_chain_mma).  Every branch below
      // is uniform and sits OUTSIDE the per-point loops.
      const int L = ct.run_L[c], plow = ct.run_plow[c];
      const double scale = ct.run_scale[c];
      const bool rev = ct.run_rev[c] != 0;
      uint64_t q[PPT];
      if (L <= 31) {
        // 32-bit conversion: cvt.rzi.u32.f64 saturates (x 2^L >= 2^32 -> 0xffffffff, negative / NaN -> 0), so one
        // unsigned min with 2^L - 1 gives the all-ones pattern of x >= 1; one 32-bit BREV where the run is reversed
        const uint32_t qmax = (1u << L) - 1u;
        const int rs = 32 - L;
        if (rev) {
#pragma unroll
          for (int k = 0; k < PPT; ++k) q[k] = __brev(min(__double2uint_rz(x[k] * scale), qmax)) >> rs;
        } else {
#pragma unroll
          for (int k = 0; k < PPT; ++k) q[k] = min(__double2uint_rz(x[k] * scale), qmax);
        }
      } else {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          unsigned long long t = (unsigned long long)(x[k] * scale);
          t = x[k] >= 1.0 ? ((1ull << L) - 1ull) : t;
          q[k] = rev ? (__brevll(t) >> (64 - L)) : t;
        }
      }
      if (ct.run_kind[c] == 2) {
        // digits scattered over the stream (interleaved dimensions, Real + Imag index on one vertex)
        const int cm = c & (kTabMaskCoords - 1);
#pragma unroll
        for (int k = 0; k < PPT; ++k) w0[k] += expand64(q[k], ct.exp_mv[cm][0], ct.exp_m[cm][0]);
        if (W2 && ct.exp_m[cm][1] != 0) {
          const int nlo = ct.exp_nlo[cm];
#pragma unroll
          for (int k = 0; k < PPT; ++k) w1[k] += expand64(q[k] >> nlo, ct.exp_mv[cm][1], ct.exp_m[cm][1]);
        }
      } else if (plow < 64) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) w0[k] += q[k] << plow;
        if (W2 && plow + L > 64) {
#pragma unroll
          for (int k = 0; k < PPT; ++k) w1[k] += q[k] >> (64 - plow);
        }
      } else if (W2) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) w1[k] += q[k] << (plow - 64);
      }
    } else if (ct.k1_generic == 1) {
      for (int e_i = s_cptr[c]; e_i < s_cptr[c + 1]; ++e_i) k1_digit4<PPT>(s_d4[e_i], dg, src, p0, NT, x, w0, w1, err);
    } else if (ct.k1_generic) {
      for (int e_i = s_cptr[c]; e_i < s_cptr[c + 1]; ++e_i) k1_generic_site<PPT>(dg, src, e_i, p0, NT, x, w0, w1, err);
    } else {
      for (int e_i = s_cptr[c]; e_i < s_cptr[c + 1]; ++e_i) {
        const Digit2 e = s_d2[e_i];
        const uint32_t stride = e.wv & 0xffu;
        const bool hi = ((e.wv >> 8) & 0xffu) != 0;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          const bool ge = src.digits ? (given_digit(src, p0 + (int64_t)k * NT, dg.n_sites, (int)(e.wv >> 16), 2, err) != 0)
                                     : (x[k] >= e.thr1);
          x[k] = __dsub_rn(x[k], ge ? e.thr1 : 0.0);
          const uint64_t bb = (uint64_t)(ge ? stride : 0u) << e.sh;
          if (W2 && hi) w1[k] += bb;
          else w0[k] += bb;
        }
      }
    }
  };
  // vector coordinate loads, issued one tile ahead
  double2 xnext[NCV ? PPT : 1];
  auto fetch_next = [&](int64_t tile_) {
    const double* cp = src.coords + (NCV == 2 ? 2 : 1) * (tile_ * TILE + tid);
    if ((tile_ + 1) * TILE <= src.npts) { // whole tile inside the batch: no per-point predicates
#pragma unroll
      for (int k = 0; k < (NCV ? PPT : 1); ++k) {
        if constexpr (NCV == 2) xnext[k] = ldg_nc_f64x2(cp + 2 * k * NT);
        else xnext[k] = make_double2(__ldg(cp + k * NT), 0.0);
      }
    } else {
#pragma unroll
      for (int k = 0; k < (NCV ? PPT : 1); ++k) {
        const int64_t p = tile_ * TILE + (int64_t)k * NT + tid;
        xnext[k] = make_double2(0.0, 0.0);
        if (tile_ < n_tiles && p < src.npts) {
          if constexpr (NCV == 2) xnext[k] = ldg_nc_f64x2(cp + 2 * k * NT);
          else xnext[k].x = __ldg(cp + k * NT);
        }
      }
    }
  };
  if (NCV) fetch_next(blockIdx.x);

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p0 = tile * TILE + tid; // point k of this thread: p0 + k * NT
#pragma unroll
    for (int k = 0; k < PPT; ++k) w0[k] = w1[k] = 0;
    if (NCV) {
      double xa[PPT], xb[PPT];
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        xa[k] = xnext[k].x;
        xb[k] = xnext[k].y;
      }
      fetch_next(tile + gridDim.x);
      add_coord(0, xa, p0);
      if (NCV == 2) add_coord(1, xb, p0);
    } else {
      for (int c = 0; c < dg.n_coords; ++c) {
        double x[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          const int64_t p = p0 + (int64_t)k * NT;
          x[k] = p < src.npts ? load_coord(src, p, c) : 0.0;
        }
        add_coord(c, x, p0);
      }
    }

    if (src.dbg_stream) { // test hook: the fused K1's digits, compared as integers by the tests
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int64_t p = p0 + (int64_t)k * NT;
        if (p < src.npts) {
          src.dbg_stream[2 * p] = w0[k];
          src.dbg_stream[2 * p + 1] = W2 ? w1[k] : 0ull;
        }
      }
    }

    // slice index of the group at stream offset `off` (lb bits): one funnel shift + mask on a one-word
    // stream; the two-word stream is consumed by shifting (groups are visited in stream order)
    auto take = [&](int k, int off, int lb) -> uint32_t {
      if constexpr (W2) {
        const uint32_t s = (uint32_t)w0[k] & ((1u << lb) - 1u);
        w0[k] = (w0[k] >> lb) | (w1[k] << (64 - lb));
        w1[k] >>= lb;
        return s;
      } else {
        return (uint32_t)(w0[k] >> off) & ((1u << lb) - 1u);
      }
    };
    int off = 0;
    // ---- leaf group
    double v[PPT][HE];
    {
      const int lb = ct.gbits[0];
      const uint32_t base = sbase + 8u * (uint32_t)ct.goff[0];
#pragma unroll
      for (int k = 0; k < PPT; ++k) entry(base, take(k, off, lb), v[k]);
      off += lb;
    }
    // ---- middle groups: v <- v * M[s]
    for (int g = 1; g < G - 1; ++g) {
      const int lb = ct.gbits[g];
      const uint32_t base = sbase + 8u * (uint32_t)ct.goff[g];
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        double m[MM], a[HE];
        entry(base, take(k, off, lb), m);
#pragma unroll
        for (int j = 0; j < HE; ++j) a[j] = 0.0;
        if constexpr (!CPLX) {
#pragma unroll
          for (int i = 0; i < H; ++i)
#pragma unroll
            for (int j = 0; j < H; ++j) a[j] = fma(v[k][i], m[i * H + j], a[j]);
        } else {
#pragma unroll
          for (int i = 0; i < H; ++i)
#pragma unroll
            for (int j = 0; j < H; ++j) {
              const double mr = m[(i * H + j) * 2], mi = m[(i * H + j) * 2 + 1];
              a[2 * j] = fma(v[k][2 * i], mr, a[2 * j]);
              a[2 * j] = fma(-v[k][2 * i + 1], mi, a[2 * j]);
              a[2 * j + 1] = fma(v[k][2 * i], mi, a[2 * j + 1]);
              a[2 * j + 1] = fma(v[k][2 * i + 1], mr, a[2 * j + 1]);
            }
        }
#pragma unroll
        for (int j = 0; j < HE; ++j) v[k][j] = a[j];
      }
      off += lb;
    }
    // ---- root group: out = v . R[s]
    double res[PPT][2];
    {
      const int lb = ct.gbits[G - 1];
      const uint32_t base = sbase + 8u * (uint32_t)ct.goff[G - 1];
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        double r[HE];
        entry(base, take(k, off, lb), r);
        double o0 = 0.0, o1 = 0.0;
        if constexpr (!CPLX) {
#pragma unroll
          for (int i = 0; i < H; ++i) o0 = fma(v[k][i], r[i], o0);
        } else {
#pragma unroll
          for (int i = 0; i < H; ++i) {
            o0 = fma(v[k][2 * i], r[2 * i], o0);
            o0 = fma(-v[k][2 * i + 1], r[2 * i + 1], o0);
            o1 = fma(v[k][2 * i], r[2 * i + 1], o1);
            o1 = fma(v[k][2 * i + 1], r[2 * i], o1);
          }
        }
        res[k][0] = o0;
        res[k][1] = o1;
      }
    }
    if (out) {
      const bool full = (tile + 1) * TILE <= src.npts;
      double* op = out + (CPLX ? 2 : 1) * p0;
      if (full) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          if (CPLX) *reinterpret_cast<double2*>(op + 2 * k * NT) = make_double2(res[k][0], res[k][1]);
          else op[k * NT] = res[k][0];
        }
      } else {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          if (p0 + (int64_t)k * NT < src.npts) {
            if (CPLX) *reinterpret_cast<double2*>(op + 2 * k * NT) = make_double2(res[k][0], res[k][1]);
            else op[k * NT] = res[k][0];
          }
        }
      }
    }
    if (do_sum) {
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int64_t p = p0 + (int64_t)k * NT;
        if (p < src.npts) accumulate_point(src, p, res[k][0], res[k][1], sum_re, sum_im);
      }
    }
  }
  if (bad) atomicOr(err, 1);

  if (do_sum) {
    // deterministic: fixed-order shuffle tree per warp, then thread 0 adds the warp partials in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum_re += __shfl_down_sync(0xffffffffu, sum_re, o);
      sum_im += __shfl_down_sync(0xffffffffu, sum_im, o);
    }
    if (lane == 0) {
      red[0][tid >> 5] = sum_re;
      red[1][tid >> 5] = sum_im;
    }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < NT / 32; ++w) {
        a += red[0][w];
        b += red[1][w];
      }
      partial[2 * blockIdx.x] = a;
      partial[2 * blockIdx.x + 1] = b;
    }
  }
}

// ------------------------------------------------------------------------------ host side

namespace {
struct cld {
  long double re = 0.0L, im = 0.0L;
};
inline cld cmul(const cld& a, const cld& b) { return cld{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

// one chain position: slice-selected a x b matrix (leaf: a = 1, root: b = 1), slices padded to 2^bits0
struct PosMat {
  int a = 1, b = 1;
  std::vector<cld> e; // [(s * a + i) * b + j]
};


} // namespace

// Host image of the group tables (pure function of the description: also behind the debug symbol
// ttn_debug_table_image, which the CPU tests walk with numpy against an independent contraction).
struct TableImage {
  int H = 0, cplx = 0, bits0 = 0, n = 0;
  int rep = 0;                     // 1: chi = 1 tables replicated per lane (conflict-free layout)
  std::vector<int> gbits, goff;    // stream bits / offset (doubles) per group; group 0 = leaf, last = root
  std::vector<double> image;
  std::vector<int> pos_of;         // chain position of every vertex (0 = leaf)
  double flops_exec = 0.0;
};

static bool make_table_image(const ttn_desc* d, size_t budget_bytes, bool allow_rep, TableImage* out) {
  const int n = d->n_vertices;
  const bool cplx = d->is_complex != 0;
  const int NC = cplx ? 2 : 1;
  if (n < 2 || d->n_sites < 1 || d->n_sites > kFeMaxSites) return false;
  // chain order from the parent array: every vertex has at most one child
  std::vector<int> child(n, -1), nsl(n, 1);
  for (int v = 0; v < n; ++v) {
    const int q = d->parent[v];
    if (q < 0) continue;
    if (child[q] >= 0) return false;
    child[q] = v;
  }
  std::vector<int> order(n), pos_of(n);
  {
    int v = d->root;
    for (int pos = n - 1; pos >= 0; --pos) {
      if (v < 0) return false;
      order[pos] = v;
      v = child[v];
    }
    for (int pos = 0; pos < n; ++pos) pos_of[order[pos]] = pos;
  }
  int maxchi = 1, maxsl = 1;
  for (int v = 0; v < n; ++v) {
    // slice indices are bit fields of the stream: binary site indices (1 or 2 per vertex), or ONE site index of
    // dimension 3 or 4 (base 3 / 4 digits) in a 2-bit field (slice 3 of a base-3 vertex is an all-zero matrix that
    // no point selects)
    for (int si = d->site_ptr[v]; si < d->site_ptr[v + 1]; ++si) {
      if (d->site_dim[si] < 2 || d->site_dim[si] > 4) return false;
      if (d->site_dim[si] > 2 && d->site_ptr[v + 1] - d->site_ptr[v] != 1) return false;
      nsl[v] *= d->site_dim[si];
    }
    maxchi = std::max(maxchi, d->link_dim[v]);
    maxsl = std::max(maxsl, nsl[v]);
  }
  if (maxsl > 4 || maxsl < 2) return false;
  if (maxchi > (cplx ? 2 : 4)) return false;
  const int H = maxchi <= 1 ? 1 : (maxchi <= 2 ? 2 : 4);
  const int bits0 = maxsl == 2 ? 1 : 2, S0 = 1 << bits0;
  const int B = n * bits0;
  if (B > 128) return false;
  const int E = cplx ? 2 : 1;
  const size_t ev = (size_t)H * E, em = (size_t)H * H * E;

  // ---- group sizes: fewest groups whose tables fit the budget; among those the smallest image
  std::vector<int> gb;
  {
    const size_t budget = budget_bytes / 8;
    for (int G = 2; G <= kTabMaxGroups && gb.empty(); ++G) {
      size_t best = 0;
      for (int tm = bits0; tm <= 16; tm += bits0) {
        const int rem = B - (G - 2) * tm;
        if (rem < 2 * bits0) break;
        const int gL = ((rem / bits0 + 1) / 2) * bits0, gR = rem - gL;
        if (gL > 16 || gR < bits0) continue;
        const size_t size = (((size_t)1 << gL) + ((size_t)1 << gR)) * ev + (size_t)(G - 2) * ((size_t)1 << tm) * em +
                            (size_t)G * 15; // + alignment padding
        if (size <= budget && (gb.empty() || size < best)) {
          best = size;
          gb.assign(G, tm);
          gb[0] = gL;
          gb[G - 1] = gR;
        }
        if (G == 2) break;
      }
    }
    if (gb.empty()) return false;
  }
  // Replicated layout.  The lookups of a warp hit random banks: measured 6.35 wavefronts per LDS.64 (conflict-free: 2)
  // and 10.4 per LDS.128 (conflict-free: 4), which makes the kernel LSU-bound.  Here every 16-byte chunk of an entry is
  // stored once per lane of a quarter-warp in one 128-byte line (8-byte entries: once per lane of a half-warp), chunk j
  // of entry s of copy c at byte 128 (s C + j) + 16 c, so lane l always reads bank group l mod 8 and NO lookup
  // conflicts.  An entry then costs C x 128 bytes: groups shrink and a point takes more — but 2.6-3x cheaper —
  // lookups.  Chosen when the estimated LSU cost (wavefronts + 3 per lookup for its instructions) drops by > 10 %.
  bool rep = false;
  if (allow_rep) {
    const size_t units = budget_bytes / 128;
    const size_t lv = std::max<size_t>(1, ev / 2), lm = std::max<size_t>(1, em / 2); // 128-byte lines per entry
    const double wv_plain = ev == 1 ? 6.35 : 10.4 * (double)(ev / 2), wm_plain = em == 1 ? 6.35 : 10.4 * (double)(em / 2);
    const double wv_rep = ev == 1 ? 2.0 : 4.0 * (double)(ev / 2), wm_rep = em == 1 ? 2.0 : 4.0 * (double)(em / 2);
    const double plain_cost = 2.0 * (wv_plain + 3.0) + (double)(gb.size() - 2) * (wm_plain + 3.0);
    const int npos = B / bits0;
    std::vector<int> best;
    for (int Gr = 2; Gr <= kTabMaxGroups && best.empty(); ++Gr) {
      const int nm = Gr - 2;
      size_t best_size = 0;
      for (int pL = 1; pL * bits0 <= 16; ++pL)
        for (int pR = std::max(1, pL - 1); pR <= pL; ++pR) {
          const int rem = npos - pL - pR;
          if (rem < nm || (nm == 0 && rem != 0)) continue;
          std::vector<int> cand(Gr);
          cand[0] = pL * bits0;
          cand[Gr - 1] = pR * bits0;
          size_t size = (((size_t)1 << cand[0]) + ((size_t)1 << cand[Gr - 1])) * lv;
          bool ok = true;
          for (int i = 0; i < nm; ++i) { // middle bits as even as possible
            cand[1 + i] = (rem / nm + (i < rem % nm ? 1 : 0)) * bits0;
            ok = ok && cand[1 + i] <= 16;
            size += ((size_t)1 << cand[1 + i]) * lm;
          }
          if (!ok || size > units) continue;
          if (best.empty() || size < best_size) {
            best = cand;
            best_size = size;
          }
        }
    }
    if (!best.empty()) {
      const double rep_cost = 2.0 * (wv_rep + 3.0) + (double)(best.size() - 2) * (wm_rep + 3.0);
      if (rep_cost < 0.9 * plain_cost) {
        gb = best;
        rep = true;
      }
    }
  }
  const int G = (int)gb.size();

  // ---- per-position matrices
  const double* T = reinterpret_cast<const double*>(d->tensors);
  std::vector<PosMat> pm(n);
  for (int pos = 0; pos < n; ++pos) {
    const int v = order[pos];
    PosMat& m = pm[pos];
    m.a = pos == 0 ? 1 : d->link_dim[order[pos - 1]];
    m.b = d->link_dim[v]; // 1 at the root
    m.e.assign((size_t)S0 * m.a * m.b, cld{});
    if (d->tensor_ptr[v + 1] - d->tensor_ptr[v] != (int64_t)nsl[v] * m.a * m.b) return false;
    for (int s = 0; s < nsl[v]; ++s)
      for (int i = 0; i < m.a * m.b; ++i) {
        const int64_t idx = (d->tensor_ptr[v] + (int64_t)s * m.a * m.b + i) * NC;
        m.e[(size_t)s * m.a * m.b + i] = cld{(long double)T[idx], cplx ? (long double)T[idx + 1] : 0.0L};
      }
  }

  // ---- tables
  out->gbits = gb;
  out->goff.assign(G, 0);
  size_t total = 0;
  for (int g = 0; g < G; ++g) {
    out->goff[g] = (int)total;
    {
      const size_t e_g = (g == 0 || g == G - 1) ? ev : em; // doubles per entry; replicated: 16 doubles per line
      total += ((size_t)1 << gb[g]) * (rep ? (size_t)16 * std::max<size_t>(1, e_g / 2) : e_g);
    }
    total = (total + 15) & ~(size_t)15; // 128-byte alignment of every table (entry_swizzle, bank groups)
  }
  out->image.assign(total, 0.0);
  double flops = 0.0;
  int c0 = 0;
  for (int g = 0; g < G; ++g) {
    const int k = gb[g] / bits0;
    const int rows = pm[c0].a;
    std::vector<cld> cur = pm[c0].e;
    size_t n_cur = S0;
    int cols = pm[c0].b;
    for (int i = 1; i < k; ++i) {
      const PosMat& A = pm[c0 + i];
      if (A.a != cols) return false;
      std::vector<cld> nxt(n_cur * S0 * rows * A.b);
      for (int bsl = 0; bsl < S0; ++bsl)
        for (size_t s = 0; s < n_cur; ++s) {
          const size_t idx = s + ((size_t)bsl << (bits0 * i));
          for (int r = 0; r < rows; ++r)
            for (int j = 0; j < A.b; ++j) {
              cld acc;
              for (int mm = 0; mm < cols; ++mm) {
                const cld t = cmul(cur[(s * rows + r) * cols + mm], A.e[((size_t)bsl * A.a + mm) * A.b + j]);
                acc.re += t.re;
                acc.im += t.im;
              }
              nxt[(idx * rows + r) * A.b + j] = acc;
            }
        }
      cur.swap(nxt);
      n_cur *= S0;
      cols = A.b;
    }
    // cur: [2^gbits][rows][cols]; leaf group rows = 1, root group cols = 1
    const bool is_leaf = g == 0, is_root = g == G - 1;
    if ((is_leaf && rows != 1) || (is_root && cols != 1)) return false;
    const size_t esz = (is_leaf || is_root) ? ev : em;
    double* dst = out->image.data() + out->goff[g];
    for (size_t s = 0; s < n_cur; ++s)
      for (int r = 0; r < rows; ++r)
        for (int j = 0; j < cols; ++j) {
          const cld x = cur[(s * rows + r) * cols + j];
          const size_t at = is_leaf ? (size_t)j : (is_root ? (size_t)r : (size_t)r * H + j);
          // element e of the entry lives in 16-byte chunk e / 2, stored at chunk position (e / 2) ^ swizzle(s)
          const uint32_t sw = esz == 16 ? entry_swizzle<16>((uint32_t)s) : esz == 8 ? entry_swizzle<8>((uint32_t)s)
                              : esz == 4 ? entry_swizzle<4>((uint32_t)s) : 0u;
          auto put = [&](size_t e, double val) { dst[s * esz + (((e >> 1) ^ sw) << 1) + (e & 1)] = val; };
          if (rep) {
            // element e of the entry: chunk e / 2 -> line s C + e / 2, 8 copies 16 bytes apart (a single-double
            // entry: 16 copies 8 bytes apart)
            const size_t C = std::max<size_t>(1, esz / 2);
            auto put_rep = [&](size_t e, double val) {
              double* line = dst + (s * C + e / 2) * 16;
              if (esz == 1) for (int c = 0; c < 16; ++c) line[c] = val;
              else for (int c = 0; c < 8; ++c) line[2 * c + (e & 1)] = val;
            };
            put_rep(at * E, (double)x.re);
            if (cplx) put_rep(at * E + 1, (double)x.im);
            continue;
          }
          put(at * E, (double)x.re);
          if (cplx) put(at * E + 1, (double)x.im);
        }
    if (!is_leaf) flops += (cplx ? 8.0 : 2.0) * rows * cols;
    c0 += k;
  }
  out->H = H;
  out->rep = rep ? 1 : 0;
  out->cplx = cplx ? 1 : 0;
  out->bits0 = bits0;
  out->n = n;
  out->pos_of = pos_of;
  out->flops_exec = flops;
  return true;
}

// Step masks of the 64-bit expand network (Hacker's Delight 7-5) for target mask m.
static void expand_masks(uint64_t m, uint64_t (&mv_out)[6]) {
  uint64_t mk = ~m << 1;
  for (int i = 0; i < 6; ++i) {
    uint64_t mp = mk ^ (mk << 1);
    mp ^= mp << 2;
    mp ^= mp << 4;
    mp ^= mp << 8;
    mp ^= mp << 16;
    mp ^= mp << 32;
    const uint64_t mv = mp & m;
    mv_out[i] = mv;
    m = (m ^ mv) | (mv >> (1 << i));
    mk &= ~mp;
  }
}
static uint64_t expand_host(uint64_t x, const uint64_t (&mv)[6], uint64_t m0) {
  for (int i = 5; i >= 0; --i) {
    const uint64_t t = x << (1 << i);
    x = (x & ~mv[i]) | (t & mv[i]);
  }
  return x & m0;
}
// the network against a bit-by-bit deposit: all ones, every single bit, a few patterns
static bool expand_selfcheck(uint64_t m0, const uint64_t (&mv)[6]) {
  auto naive = [&](uint64_t x) {
    uint64_t r = 0;
    int j = 0;
    for (int b = 0; b < 64; ++b)
      if (m0 >> b & 1) {
        if (x >> j & 1) r |= 1ull << b;
        ++j;
      }
    return r;
  };
  const int n = __builtin_popcountll(m0);
  std::vector<uint64_t> xs = {0ull, ~0ull, 0x5555555555555555ull, 0x123456789abcdef1ull, 0xfedcba9876543210ull};
  for (int j = 0; j < n; ++j) xs.push_back(1ull << j);
  for (uint64_t x : xs) {
    const uint64_t xm = n >= 64 ? x : (x & ((1ull << n) - 1ull));
    if (expand_host(xm, mv, m0) != naive(xm)) return false;
  }
  return true;
}

static size_t table_budget_bytes() {
  if (const char* e = getenv("TTN_TABLE_KB")) return (size_t)std::max(8, std::min(atoi(e), 200)) * 1024;
  return (size_t)200 * 1024;
}

int build_chain_table(ttn_plan* p, const ttn_desc* d) {
  p->ctab_ok = false;
  if (!p->is_chain) return TTN_OK;
  TableImage im;
  const bool allow_rep = !(getenv("TTN_TABLE_REP") && atoi(getenv("TTN_TABLE_REP")) == 0);
  if (!make_table_image(d, table_budget_bytes(), allow_rep, &im)) return TTN_OK;
  ChainTabDev& c = p->ctab;
  c = ChainTabDev{};
  c.n_groups = (int)im.gbits.size();
  c.H = im.H;
  c.cplx = im.cplx;
  c.rep = im.rep;
  c.k1_generic = k1_generic_mode(d);
  c.total_doubles = (int)im.image.size();
  for (int g = 0; g < c.n_groups; ++g) {
    c.gbits[g] = im.gbits[g];
    c.goff[g] = im.goff[g];
  }
  double* d_img;
  TTN_CUDA(cudaMalloc(&d_img, std::max<size_t>(im.image.size() * 8, 16)));
  p->allocs.push_back(d_img);
  TTN_CUDA(cudaMemcpy(d_img, im.image.data(), im.image.size() * 8, cudaMemcpyHostToDevice));
  c.image = d_img;
  p->ctab_flops_exec = im.flops_exec;
  p->ctab_bits0 = im.bits0;

  // own copy of the digit table: (word, shift) = bit position of the vertex in THIS kernel's stream
  const int bits0 = im.bits0;
  p->digits_tab = p->digits;
  std::vector<DigitEntry> ent(d->n_sites);
  TTN_CUDA(cudaMemcpy(ent.data(), p->digits.entries, sizeof(DigitEntry) * d->n_sites, cudaMemcpyDeviceToHost));
  for (auto& e : ent) {
    const int bitpos = im.pos_of[e.vertex] * bits0;
    e.word = bitpos / 64;
    e.shift = bitpos % 64;
  }
  DigitEntry* d_ent;
  TTN_CUDA(cudaMalloc(&d_ent, sizeof(DigitEntry) * d->n_sites));
  p->allocs.push_back(d_ent);
  TTN_CUDA(cudaMemcpy(d_ent, ent.data(), sizeof(DigitEntry) * d->n_sites, cudaMemcpyHostToDevice));
  p->digits_tab.entries = d_ent;

  // K1 run fast path per coordinate slot.  Conditions (checked bitwise): every site index of the slot is binary,
  // the digit numbers are exactly 1..L with thresholds exactly 2^-k, L <= 53, and the stream bits of the digits
  // are strictly increasing or strictly decreasing in the digit number.  Then the greedy loop
  // (abstractindexmap.jl:121-138) yields digit k = bit (L-k) of floor(x 2^L), exactly (build_chain_mma has the
  // argument).  Consecutive bits -> kind 1 (one shift); anything else monotone -> kind 2 (bit deposit).
  {
    std::vector<int32_t> cptr(d->n_coords + 1);
    TTN_CUDA(cudaMemcpy(cptr.data(), p->digits.coord_ptr, sizeof(int32_t) * (d->n_coords + 1), cudaMemcpyDeviceToHost));
    for (int cidx = 0; cidx < d->n_coords; ++cidx) {
      const int L = cptr[cidx + 1] - cptr[cidx];
      if (L < 1 || L > 53) continue;
      bool ok = true, inc = true, dec = true, contiguous = true;
      std::vector<int> bp(L);
      for (int k = 0; k < L && ok; ++k) {
        const DigitEntry& e = ent[cptr[cidx] + k];
        ok = ok && e.base == 2 && (e.stride == 1 || e.stride == 2);
        ok = ok && d->site_digit[e.site] == k + 1 && d->thr[e.thr_off + 1] == std::ldexp(1.0, -(k + 1));
        bp[k] = e.word * 64 + e.shift + (e.stride == 2 ? 1 : 0);
        if (k >= 1) {
          inc = inc && bp[k] > bp[k - 1];
          dec = dec && bp[k] < bp[k - 1];
          contiguous = contiguous && std::abs(bp[k] - bp[k - 1]) == 1;
        }
      }
      if (!ok || !(inc || dec)) continue;
      const bool rev = inc; // digit 1 on the LOWEST stream bit: reverse the bits of floor(x 2^L)
      if (!contiguous) {
        if (cidx >= kTabMaskCoords) continue;
        uint64_t m[2] = {0, 0};
        for (int k = 0; k < L; ++k) m[bp[k] / 64] |= 1ull << (bp[k] % 64);
        bool good = true;
        for (int w = 0; w < 2; ++w) {
          expand_masks(m[w], c.exp_mv[cidx][w]);
          c.exp_m[cidx][w] = m[w];
          good = good && expand_selfcheck(m[w], c.exp_mv[cidx][w]);
        }
        c.exp_nlo[cidx] = __builtin_popcountll(m[0]);
        if (!good) continue; // never observed; the tabulated loop is always correct
      }
      c.run_kind[cidx] = contiguous ? 1 : 2;
      c.run_L[cidx] = L;
      c.run_rev[cidx] = rev ? 1 : 0;
      c.run_plow[cidx] = rev ? bp[0] : bp[L - 1];
      c.run_scale[cidx] = std::ldexp(1.0, L);
    }
  }
  p->ctab_ok = true;
  return TTN_OK;
}

template <int H, bool CPLX, int NT, int MINB, int PPT, int NCV, bool W2, bool REP>
static int launch_tab_inst(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                           cudaStream_t s) {
  const ChainTabDev& c = p->ctab;
  const size_t smem = (((size_t)c.total_doubles * 8 + 15) & ~(size_t)15) +
                      (c.k1_generic == 1 ? (size_t)p->digits_tab.n_sites * sizeof(Digit4) : 0); // tables [+ Digit4 entries]
  auto kern = chain_table_kernel<H, CPLX, NT, MINB, PPT, NCV, W2, REP>;
  TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  TTN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
  per_sm = std::max(1, std::min(per_sm, MINB));
  const int64_t n_tiles = (src.npts + (int64_t)NT * PPT - 1) / ((int64_t)NT * PPT);
  const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)p->sm_count * per_sm);
  const int do_sum = d_partial != nullptr;
  kern<<<grid, NT, smem, s>>>(c, p->digits_tab, src, d_out, p->d_err, d_partial, do_sum);
  TTN_CUDA(cudaGetLastError());
  *n_partial = do_sum ? grid : 0;
  return TTN_OK;
}

template <int H, bool CPLX, int PPT, int NCV, bool W2>
static int launch_tab_variant(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                              cudaStream_t s) {
  // One persistent CTA per SM: 512 threads x 2 PPT points per thread and tile (up to 128 registers: 8 points
  // in flight per thread for chi = 1).  Measured 5-10 % faster than 1024 threads x PPT points and the same as
  // 768 x 1.5 PPT (scripts/table_sweep.py history in DESIGN.md); only this shape is instantiated.
  if (p->ctab.rep) return launch_tab_inst<H, CPLX, 512, 1, 2 * PPT, NCV, W2, true>(p, src, d_out, d_partial, n_partial, s);
  return launch_tab_inst<H, CPLX, 512, 1, 2 * PPT, NCV, W2, false>(p, src, d_out, d_partial, n_partial, s);
}

template <int H, bool CPLX, int PPT>
static int launch_tab_shape(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                            cudaStream_t s) {
  const bool vec = !src.digits && !src.grid && src.coords;
  const bool aos2 = vec && src.layout == TTN_LAYOUT_AOS && src.n_coords == 2 && (reinterpret_cast<uintptr_t>(src.coords) & 15u) == 0;
  const bool one = vec && src.n_coords == 1;
  int bits = 0;
  for (int g = 0; g < p->ctab.n_groups; ++g) bits += p->ctab.gbits[g];
  const bool w2 = bits > 64;
#define TTN_TAB_GO(NCV_)                                                                                    \
  return w2 ? launch_tab_variant<H, CPLX, PPT, NCV_, true>(p, src, d_out, d_partial, n_partial, s)         \
            : launch_tab_variant<H, CPLX, PPT, NCV_, false>(p, src, d_out, d_partial, n_partial, s)
  if (aos2) TTN_TAB_GO(2);
  if (one) TTN_TAB_GO(1);
  TTN_TAB_GO(0);
#undef TTN_TAB_GO
}

int launch_chain_table(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double* d_partial,
                       int* n_partial, cudaStream_t s) {
  (void)st;
  *n_partial = 0;
  if (src.npts == 0) return TTN_OK;
  if (!p->ctab_ok) {
    set_error("table kernel requested but the network is not a narrow binary chain");
    return TTN_ERR_UNSUPPORTED;
  }
  const ChainTabDev& c = p->ctab;
  if (!c.cplx) {
    if (c.H == 1) return launch_tab_shape<1, false, 4>(p, src, d_out, d_partial, n_partial, s);
    if (c.H == 2) return launch_tab_shape<2, false, 2>(p, src, d_out, d_partial, n_partial, s);
    if (c.H == 4) return launch_tab_shape<4, false, 2>(p, src, d_out, d_partial, n_partial, s);
  } else {
    if (c.H == 1) return launch_tab_shape<1, true, 2>(p, src, d_out, d_partial, n_partial, s);
    if (c.H == 2) return launch_tab_shape<2, true, 2>(p, src, d_out, d_partial, n_partial, s);
  }
  set_error("table kernel: unsupported bond dimension");
  return TTN_ERR_UNSUPPORTED;
}

// Debug / test hook (not part of the ABI in include/ttneval.h; no CUDA calls): the table image of a
// description, so that CPU tests can walk it against an independent contraction.
//   meta[0..7] = {applicable, H, is_complex, bits0, n_groups, total_doubles, replicated layout, 0}; meta[8 + g] = gbits[g],
//   meta[8 + 16 + g] = goff[g]; site_bitpos[s] = stream bit of site index s (stride included).
int debug_table_image(const ttn_desc* d, int32_t budget_kb, int32_t* meta, double* image, int64_t image_cap,
                      int32_t* site_bitpos) {
  TableImage im;
  for (int i = 0; i < 8 + 2 * kTabMaxGroups; ++i) meta[i] = 0;
  bool base2 = true;
  for (int s = 0; s < d->n_sites; ++s) base2 = base2 && d->site_dim[s] == 2;
  const bool allow_rep = budget_kb >= 0;
  if (budget_kb < 0) budget_kb = -budget_kb; // negative budget: plain layout only
  if (!base2 || !make_table_image(d, (size_t)budget_kb * 1024, allow_rep, &im)) return TTN_OK;
  if ((int64_t)im.image.size() > image_cap) {
    set_error("debug_table_image: image buffer too small");
    return TTN_ERR_INVALID;
  }
  meta[0] = 1;
  meta[1] = im.H;
  meta[2] = im.cplx;
  meta[3] = im.bits0;
  meta[4] = (int)im.gbits.size();
  meta[5] = (int)im.image.size();
  meta[6] = im.rep;
  for (size_t g = 0; g < im.gbits.size(); ++g) {
    meta[8 + g] = im.gbits[g];
    meta[8 + kTabMaxGroups + g] = im.goff[g];
  }
  std::copy(im.image.begin(), im.image.end(), image);
  for (int v = 0; v < d->n_vertices; ++v) {
    int stride_bits = 0;
    for (int si = d->site_ptr[v + 1] - 1; si >= d->site_ptr[v]; --si) {
      site_bitpos[si] = im.pos_of[v] * im.bits0 + stride_bits;
      stride_bits += 1;
    }
  }
  return TTN_OK;
}

} // namespace ttn
