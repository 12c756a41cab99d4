// K4, team-sorted kernel — the bench kernel.  Host-side images (group merging, deep tables) and the
// common scheme are described in k_chain_mma.cu.
#include <algorithm>
#include <cstring>

#include "k_chain_common.cuh"

namespace ttn {

// =====================================================================================
// v6: team-sorted, B-stationary kernel for MERGED binary chains (one stream position per round,
// NCLS = 4, 8 or 16 slices per position).  A CTA holds NTEAM independent teams of 4 warps; a team
// owns a tile of 512 points.
//  * every warp is "home" to 128 points of the tile: K1, leaf rows, class counts, root;
//    with deep leaf / root tables (k_chain_mma.cu) the table rows are gathered cooperatively — CHI / 2 lanes per 128-byte
//    row: leaf rows by cp.async straight into the state tile, root rows chunk by chunk with a reduce-scatter over the
//    lanes of a row — 1 L1 wavefront per row instead of 4 (DESIGN.md "Cooperative gathers");
//  * per round the team counting-sorts its 512 points by class: per-warp match.any counts, one
//    word of 4 byte counters per class in shared memory, ONE team barrier per round (the counts of
//    round r+1 are published before the barrier of round r);
//  * warp w then owns classes w, w+4, ...: the class's site matrix sits in REGISTERS as DMMA B
//    fragments (read from L2 one class ahead) while the warp streams the class's rows through
//    gather -> DMMA -> scatter.  No B traffic through shared memory, class padding amortised over
//    512 points, and the teams drift apart so one team's sort hides under the other's DMMAs.
template <int CHI, int NBAT>
__device__ __forceinline__ void batch6(uint32_t state_base, const int (&rows)[4], int tq, int zrow,
                                       const double (&bf)[(CHI / 4) * (CHI / 8)]) {
  constexpr int NB = CHI / 8, KB = CHI / 4;
  double a[NBAT][KB], d[NBAT][KB];
#pragma unroll
  for (int b = 0; b < NBAT; ++b)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const double2 v = lds128(row_chunk<CHI>(state_base, rows[b], 4 * nb + tq));
      a[b][2 * nb] = v.x;
      a[b][2 * nb + 1] = v.y;
      d[b][2 * nb] = 0.0;
      d[b][2 * nb + 1] = 0.0;
    }
#pragma unroll
  for (int kb = 0; kb < KB; ++kb)
#pragma unroll
    for (int nbp = 0; nbp < NB; ++nbp)
#pragma unroll
      for (int b = 0; b < NBAT; ++b) dmma884(d[b][2 * nbp], d[b][2 * nbp + 1], a[b][kb], bf[kb * NB + nbp]);
#pragma unroll
  for (int b = 0; b < NBAT; ++b)
    if (rows[b] != zrow) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) sts128(row_chunk<CHI>(state_base, rows[b], 4 * nb + tq), d[b][2 * nb], d[b][2 * nb + 1]);
    }
}

#ifdef TTN_TEAM_CLOCKS
// debug build (make dbgteam, scripts/team_clocks.py): per-warp clock accounting of the phases of a tile
__device__ unsigned long long g_tphase[12];
#define TPH_DECL long long ph_t = clock64(); unsigned long long ph_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define TPH_MARK(i) { const long long t_ = clock64(); ph_acc[i] += (unsigned long long)(t_ - ph_t); ph_t = t_; }
#define TPH_FLUSH if (lane == 0) { for (int i_ = 0; i_ < 10; ++i_) atomicAdd(&g_tphase[i_], ph_acc[i_]); atomicAdd(&g_tphase[11], 1ull); }
#else
#define TPH_DECL
#define TPH_MARK(i)
#define TPH_FLUSH
#endif

template <int CHI, int NCLS, int PW_>
struct Team6 {
  static constexpr int TW = 4, PW = PW_, TP = TW * PW;
  static constexpr int LIST_CAP = TP + 8 * NCLS;
  static constexpr size_t STATE_BYTES = (size_t)(TP + 8) * CHI * 8;
  static constexpr size_t BYTES = (STATE_BYTES + 2 * LIST_CAP * 2 + 3 * 2 * NCLS * 4 + 127) / 128 * 128;
};

template <int CHI, int NCLS, int NTEAM, int PW_>
__global__ void __launch_bounds__(NTEAM * 128, 1)
    chain_mma6_kernel(ChainMmaDev ch, DigitTable dg, CoordSource src, double* __restrict__ out, int* err,
                      double* __restrict__ partial, int do_sum) {
  using T6 = Team6<CHI, NCLS, PW_>;
  constexpr int TW = T6::TW, PW = T6::PW, PPL = PW / 32, TP = T6::TP;
  constexpr int GB = (CHI >= 32) ? 2 : 4;   // 8-row groups per batch (register budget)
  constexpr int CPW = NCLS / TW;           // classes owned by a warp
  constexpr int BITS = slice_bits(NCLS);
  constexpr int NB = CHI / 8, KB = CHI / 4, CPR = CHI / 2;
  constexpr int NBF = KB * NB;             // B-fragment doubles per lane per class
  constexpr int LIST_CAP = T6::LIST_CAP;
  constexpr int ZROW = TP;                 // the team's all-zero row (class padding target)
  constexpr int NT = NTEAM * TW * 32;
  constexpr int PAR_BIT = (CHI >= 16) ? 2 : 0;  // row bit that selects the bank half of a 64-byte piece
  constexpr int HS = 2 * NCLS;             // counters per set: (parity, class)
  static_assert(NCLS % TW == 0 && NCLS <= 32 && (NCLS & (NCLS - 1)) == 0, "class count");

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ double red[2][NTEAM * TW];
  __shared__ Digit2 s_d2[kFeMaxSites]; // binary digits
  __shared__ Digit4 s_d4[kFeMaxSites]; // base 3 / 4 (two plain arrays: see k_chain_table.cu)
  __shared__ int s_cptr[TTN_MAX_COORDS + 1];

  const int tid = threadIdx.x, lane = tid & 31;
  const int team = tid / (TW * 32), warp = (tid >> 5) % TW;
  const int g = lane >> 2, tq = lane & 3;
  constexpr int RPI = 32 / CPR;            // deep tables: table rows per warp instruction (CPR lanes share a row)
  const int coj = lane % CPR, coq = lane / CPR;
  for (int i = tid; i <= dg.n_coords; i += NT) s_cptr[i] = dg.coord_ptr[i];
  for (int i = tid; i < (ch.k1_generic == 1 ? dg.n_sites : 0); i += NT) s_d4[i] = make_digit4(dg, i);
  for (int i = tid; i < (ch.k1_generic == 0 ? dg.n_sites : 0); i += NT) {
    const DigitEntry e = dg.entries[i];
    Digit2 d2;
    d2.thr1 = dg.thr[e.thr_off + 1];
    d2.sh = (uint32_t)e.shift;
    d2.wv = ((uint32_t)e.site << 16) | ((uint32_t)e.word << 8) | (uint32_t)e.stride;
    s_d2[i] = d2;
  }
  // leaf / root vectors of the CTA in shared memory (after the team regions), 16-byte chunks
  // XOR-swizzled by the slice number: lanes reading chunk j of different slices spread over banks
  double* s_leaf = reinterpret_cast<double*>(smem + (size_t)NTEAM * T6::BYTES);  // [NCLS][CHI]
  double* s_root = s_leaf + NCLS * CHI;                                           // [nout][NCLS][CHI]
  // deep leaf / root groups (build_chain_mma): tables of 2^leaf_bits / 2^root_bits vectors stay in
  // global memory and a point gathers one row of each
  const bool deep = ch.leaf_bits != BITS || ch.root_bits != BITS;
  for (int i = tid; i < (deep ? 0 : NCLS * CPR); i += NT) {
    const int sl = i / CPR, j = i % CPR, pj = j ^ (sl & (CPR - 1) & 7);
    s_leaf[sl * CHI + 2 * pj] = ch.leaf[sl * CHI + 2 * j];
    s_leaf[sl * CHI + 2 * pj + 1] = ch.leaf[sl * CHI + 2 * j + 1];
    for (int o = 0; o < ch.nout; ++o) {
      s_root[(o * NCLS + sl) * CHI + 2 * pj] = ch.root[(o * NCLS + sl) * CHI + 2 * j];
      s_root[(o * NCLS + sl) * CHI + 2 * pj + 1] = ch.root[(o * NCLS + sl) * CHI + 2 * j + 1];
    }
  }
  unsigned char* tbase = smem + (size_t)team * T6::BYTES;
  const uint32_t state_base = smem_u32(tbase);
  uint16_t* lists = reinterpret_cast<uint16_t*>(tbase + T6::STATE_BYTES);     // [2][LIST_CAP]
  const uint32_t s_leaf_u32 = smem_u32(s_leaf), s_root_u32 = smem_u32(s_root);
  uint32_t* hist = reinterpret_cast<uint32_t*>(lists + 2 * LIST_CAP);          // [3][2][16] (class, parity) counters
  for (int i = tid % (TW * 32); i < 3 * HS; i += TW * 32) hist[i] = 0u;
  for (int i = tid % (TW * 32); i < 8 * CHI; i += TW * 32)
    reinterpret_cast<double*>(tbase + (size_t)TP * CHI * 8)[i] = 0.0;
  __syncthreads();

  const int bar_id = 1 + team;
  const int64_t n_tiles = (src.npts + TP - 1) / TP;
  const int64_t tstride = (int64_t)gridDim.x * NTEAM;
  const int R = ch.n_rounds;
  const uint64_t MASK = (uint64_t)(NCLS - 1);
  const uint64_t LMASK = (1ull << ch.leaf_bits) - 1ull, RMASK = (1ull << ch.root_bits) - 1ull;
  const uint32_t lt = (1u << lane) - 1u;
  double sum_re = 0.0, sum_im = 0.0;
  int qh = 0; // global round counter of the team (rotates the counter sets)

  // B fragments of class c of round r (fragment order, see build_chain_mma): 32 consecutive doubles per (kb, nb)
  double bcur[NBF], bnxt[NBF];
  auto load_b = [&](double (&b)[NBF], int r, int c) {
    const double* F = ch.frags + ((size_t)r * NCLS + c) * (CHI * CHI) + lane;
#pragma unroll
    for (int i = 0; i < NBF; ++i) b[i] = __ldg(F + i * 32);
  };
  if (R > 0) load_b(bcur, 0, warp);

  // ---- K1: packed slice streams of the lane's PPL home points of one tile (interleaved for ILP)
  auto compute_words = [&](int64_t tile_, uint64_t (&w0)[PPL], uint64_t (&w1)[PPL]) {
    const int64_t p0 = tile_ * TP + warp * PW;
    double x[PPL], xn[PPL];
    auto load_x = [&](int c, double (&v)[PPL]) {
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        const int64_t p = p0 + k * 32 + lane;
        v[k] = p < src.npts ? load_coord(src, p, c) : 0.0;
      }
    };
#pragma unroll
    for (int k = 0; k < PPL; ++k) w0[k] = w1[k] = 0;
    if (!src.qcoords && dg.n_coords > 0) load_x(0, xn);
    for (int c = 0; c < dg.n_coords; ++c) {
      if (src.qcoords) {
        // coordinates quantised on the host while they were staged (ttn_api.cu, pack_coords): q = floor(x 2^L), x >= 1
        // saturated, domain checked there — every coordinate is on the run path (the host checked that too)
        const int L = ch.run_L[c], plow = ch.run_plow[c];
        const int rev = ch.run_rev[c]; // 1: bits reversed, 2: bit PAIRS reversed (base-4 digits)
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          const int64_t p = p0 + k * 32 + lane;
          unsigned long long q = p < src.npts ? (unsigned long long)__ldg(src.qcoords + p * dg.n_coords + c) : 0ull;
          if (rev) q = __brevll(q) >> (64 - L);
          if (rev == 2) q = ((q & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((q & 0x5555555555555555ull) << 1);
          if (plow < 64) {
            w0[k] += q << plow;
            if (plow + L > 64) w1[k] += q >> (64 - plow);
          } else {
            w1[k] += q << (plow - 64);
          }
        }
        continue;
      }
      // the loads of coordinate c + 1 are issued before coordinate c is converted, and nothing with a side effect sits
      // between the loads of one coordinate (the domain flag is raised once, afterwards): K1 waits for ONE memory
      // round trip per tile where a load -> check -> atomic sequence per point made it eight
#pragma unroll
      for (int k = 0; k < PPL; ++k) x[k] = xn[k];
      if (c + 1 < dg.n_coords) load_x(c + 1, xn);
      {
        bool bad = false;
#pragma unroll
        for (int k = 0; k < PPL; ++k)
          if (!coord_in_domain(x[k])) {
            bad = true;
            x[k] = 0.0;
          }
        if (bad) atomicOr(err, 1);
      }
      if (ch.run_L[c] > 0 && !src.digits) {
        const int L = ch.run_L[c], plow = ch.run_plow[c];
        const double scale = ch.run_scale[c];
        const int rev = ch.run_rev[c]; // 1: bits reversed, 2: bit PAIRS reversed (base-4 digits)
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          unsigned long long q = x[k] >= 1.0 ? ((1ull << L) - 1ull) : (unsigned long long)(x[k] * scale);
          if (rev) q = __brevll(q) >> (64 - L);
          if (rev == 2) q = ((q & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((q & 0x5555555555555555ull) << 1);
          if (plow < 64) {
            w0[k] += q << plow;
            if (plow + L > 64) w1[k] += q >> (64 - plow);
          } else {
            w1[k] += q << (plow - 64);
          }
        }
      } else if (ch.k1_generic == 1) {
        k1_digit4_run<PPL>(s_d4, s_cptr[c], s_cptr[c + 1], dg, src, p0 + lane, 32, x, w0, w1, err);
      } else if (ch.k1_generic) {
        for (int e_i = s_cptr[c]; e_i < s_cptr[c + 1]; ++e_i) k1_generic_site<PPL>(dg, src, e_i, p0 + lane, 32, x, w0, w1, err);
      } else {
        for (int e_i = s_cptr[c]; e_i < s_cptr[c + 1]; ++e_i) {
          const Digit2 e = s_d2[e_i];
          const uint32_t stride = e.wv & 0xffu;
          const bool hi = ((e.wv >> 8) & 0xffu) != 0;
#pragma unroll
          for (int k = 0; k < PPL; ++k) {
            const bool ge = src.digits ? (given_digit(src, p0 + k * 32 + lane, dg.n_sites, (int)(e.wv >> 16), 2, err) != 0)
                                       : (x[k] >= e.thr1);
            x[k] = __dsub_rn(x[k], ge ? e.thr1 : 0.0);
            const uint64_t bb = (uint64_t)(ge ? stride : 0u) << e.sh;
            if (hi) w1[k] += bb;
            else w0[k] += bb;
          }
        }
      }
    }
  };

  TPH_DECL
  for (int64_t tile = (int64_t)blockIdx.x * NTEAM + team; tile < n_tiles; tile += tstride) {
    const int64_t p0 = tile * TP + warp * PW; // first home point of this warp
    uint64_t w1[PPL], cw[PPL];
    {
      uint64_t w0[PPL];
      compute_words(tile, w0, w1);
      if (tile + tstride < n_tiles && (src.qcoords || (!src.grid && !src.digits))) {
        // the warp's coordinates of the team's NEXT tile -> L2 (a tile later K1's one round trip is an L2 hit)
        const int64_t pn = (tile + tstride) * TP + warp * PW;
        const int64_t np_ = min((int64_t)PW, src.npts - pn);
        auto pf = [&](const void* b, int64_t nbytes) {
          const char* e = reinterpret_cast<const char*>(b) + nbytes;
          for (const char* a = reinterpret_cast<const char*>((uintptr_t)b & ~(uintptr_t)127) + lane * 128; a < e; a += 32 * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        };
        if (np_ > 0) {
          if (src.qcoords) pf(src.qcoords + pn * dg.n_coords, np_ * dg.n_coords * 4);
          else if (src.layout == TTN_LAYOUT_AOS) pf(src.coords + pn * dg.n_coords, np_ * dg.n_coords * 8);
          else
            for (int c = 0; c < dg.n_coords; ++c) pf(src.coords + (int64_t)c * src.npts + pn, np_ * 8);
        }
      }
      TPH_MARK(0)
      if (deep && R > 0 && ch.root_bits > 16) {
        // the root-table row of every home point is known as soon as K1 is done: pull it into L2 now, the rounds
        // hide the HBM latency (tables of 2^20 rows do not stay L2-resident; smaller ones do, and with no rounds
        // in between the prefetch would only cost instructions)
        const int off = ch.leaf_bits + R * BITS;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          const uint64_t top = off == 0 ? w0[k] : (off < 64 ? ((w0[k] >> off) | (w1[k] << (64 - off))) : (w1[k] >> (off - 64)));
          const double* Rp = ch.root + (size_t)(top & RMASK) * CHI;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(Rp));
          if (CHI > 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(Rp + 16));
          if (ch.nout == 2) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Rp + ((size_t)CHI << ch.root_bits)));
            if (CHI > 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(Rp + ((size_t)CHI << ch.root_bits) + 16));
          }
        }
      }
      if (src.dbg_stream) { // test hook: the fused K1's digits, compared as integers by the tests
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          const int64_t p = p0 + k * 32 + lane;
          if (p < src.npts) {
            src.dbg_stream[2 * p] = w0[k];
            src.dbg_stream[2 * p + 1] = w1[k];
          }
        }
      }
      // ---- leaf rows
      if (deep) {
        // CPR lanes share one table row: a warp instruction copies 32 / CPR whole 128-byte-aligned rows global -> shared
        // (cp.async: 1 L1 wavefront per row where a lane walking its own row with 256-bit loads costs 4, no registers,
        // and every copy of the warp's 128 rows is in flight at once); the row index comes from the owner lane
        __syncwarp(); // the owners' root reads of the previous tile precede the other lanes' copies into those rows
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          cw[k] = w0[k];
          const uint32_t li = (uint32_t)(w0[k] & LMASK);
#pragma unroll
          for (int s = 0; s < CPR; ++s) {
            const int ol = s * RPI + coq;
            const uint32_t ri = __shfl_sync(0xffffffffu, li, ol);
            cp_async16_cg(row_chunk<CHI>(state_base, warp * PW + k * 32 + ol, coj), ch.leaf + (size_t)ri * CHI + 2 * coj);
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          cw[k] = w0[k];
          const int row = warp * PW + k * 32 + lane;
          const int sl = (int)(cw[k] & MASK);
          const uint32_t L = s_leaf_u32 + (uint32_t)sl * (CHI * 8);
#pragma unroll
          for (int j = 0; j < CPR; ++j) {
            const double2 v = lds128(L + (uint32_t)((j ^ (sl & (CPR - 1) & 7)) << 4));
            sts128(row_chunk<CHI>(state_base, row, j), v.x, v.y);
          }
        }
      }
    }
    auto shift_stream = [&]() {
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        cw[k] = (cw[k] >> BITS) | (w1[k] << (64 - BITS));
        w1[k] >>= BITS;
      }
    };
    {
      const int lb = ch.leaf_bits; // 1..21
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        cw[k] = (cw[k] >> lb) | (w1[k] << (64 - lb));
        w1[k] >>= lb;
      }
    }

    // classes of the next stream position: one shared-memory atomic per point on the team's
    // (class, parity) counters gives the point its position among the team's points of that class
    // and parity; info[k] = class | position << 8.  `parity` is the row bit that decides which half
    // of the bank line a 64-byte piece of the row occupies (row_chunk swizzle): it is a lane
    // constant for home rows.  Counter sets rotate over 3 buffers (qh = global round counter).
    uint32_t info[PPL];
    const int par = (lane >> PAR_BIT) & 1;
    auto count_round = [&](int hb) {
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        const int cls = (int)(cw[k] & MASK);
        const uint32_t pos = atomicAdd(hist + hb * HS + par * NCLS + cls, 1u);
        info[k] = (uint32_t)cls | (pos << 8);
      }
      shift_stream();
    };
    if (R > 0) count_round(qh % 3);
    TPH_MARK(1)
    if (deep) cp_async_wait_all();
    TPH_MARK(2)
    named_bar_sync(bar_id, TW * 32); // leaf rows + counts of round 0
    TPH_MARK(3)

    for (int r = 0; r < R; ++r, ++qh) {
      const int buf = r & 1, hb = qh % 3;
      uint16_t* list = lists + buf * LIST_CAP;
      // ---- list of round r.  Lane c holds class c's counts.  Inside a class the rows of opposite
      // parity are zipped into (even, odd) slot pairs — the two rows a quarter-warp gathers at once
      // then never collide — and the surplus of the larger parity follows.
      const int n0 = (lane < NCLS) ? (int)hist[hb * HS + lane] : 0;
      const int n1 = (lane < NCLS) ? (int)hist[hb * HS + NCLS + lane] : 0;
      const int total = n0 + n1, mzip = min(n0, n1);
      const int padded = (total + 7) & ~7;
      int incl = padded;
#pragma unroll
      for (int o = 1; o < NCLS; o <<= 1) {
        const int nn = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nn;
      }
      const int mystart = incl - padded;
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        const int c = (int)(info[k] & 255u), pos = (int)(info[k] >> 8);
        const int st = __shfl_sync(0xffffffffu, mystart, c), mm = __shfl_sync(0xffffffffu, mzip, c);
        list[st + (pos < mm ? 2 * pos + par : mm + pos)] = (uint16_t)(warp * PW + k * 32 + lane);
      }
      for (int c0 = 4 * warp; c0 < NCLS; c0 += 4 * TW) { // class padding -> zero row
        const int pc = c0 + (lane >> 3), pi = lane & 7;
        const int cn = __shfl_sync(0xffffffffu, total, pc & 31), cs = __shfl_sync(0xffffffffu, mystart, pc & 31);
        if (pc < NCLS && (cn & 7) && pi >= (cn & 7)) list[cs + (cn & ~7) + pi] = (uint16_t)ZROW;
      }
      if (r + 1 < R) count_round((qh + 1) % 3);
      if (warp == 0)
        for (int i = lane; i < HS; i += 32) hist[((qh + 2) % 3) * HS + i] = 0u; // last read before the previous barrier
      TPH_MARK(4)
      named_bar_sync(bar_id, TW * 32); // list r complete; rows of round r-1 written; counts r+1 final
      TPH_MARK(5)

      // ---- classes of this warp: B in registers, rows streamed through gather -> DMMA -> scatter.  Warp w owns
      // classes w, w + 4, ...  A BIG class (more than BIG_G groups of 8 rows) is split into four quarters, one per
      // warp, so that a skewed class distribution — a coordinate held constant along a cut, dyadic grid points whose
      // low digits are all zero: every point of the tile in ONE class — keeps all four warps busy instead of one
      // (measured on such inputs: rounds 3.5x slower with static ownership; now as fast as random points).  A round
      // without a big class — every round of random points at 16 classes — takes the static, fully unrolled schedule.
      // Either way the first class of a round is the warp's own class w, whose B is prefetched across the barrier.
      auto run_groups = [&](int st, int g_lo, int ng) {
        const uint16_t* Lc = list + st;
        int rows_nx[4] = {0, 0, 0, 0};
        if (ng > g_lo) {
#pragma unroll
          for (int b = 0; b < GB; ++b) rows_nx[b] = (int)Lc[(min(g_lo + b, ng - 1) << 3) + g];
        }
        for (int gi = g_lo; gi < ng; gi += GB) {
          const int nbat = min(GB, ng - gi);
          int rows[4] = {0, 0, 0, 0};
#pragma unroll
          for (int b = 0; b < GB; ++b) rows[b] = rows_nx[b];
          if (gi + GB < ng) {
#pragma unroll
            for (int b = 0; b < GB; ++b) rows_nx[b] = (int)Lc[(min(gi + GB + b, ng - 1) << 3) + g];
          }
          if (GB >= 4 && nbat == 4) batch6<CHI, (GB >= 4 ? 4 : 1)>(state_base, rows, tq, ZROW, bcur);
          else if (GB >= 4 && nbat == 3) batch6<CHI, (GB >= 4 ? 3 : 1)>(state_base, rows, tq, ZROW, bcur);
          else if (nbat == 2) batch6<CHI, 2>(state_base, rows, tq, ZROW, bcur);
          else batch6<CHI, 1>(state_base, rows, tq, ZROW, bcur);
        }
      };
      constexpr int BIG_G = 8;
      const int ngl = (total + 7) >> 3; // lane c: 8-row groups of class c
      if (__ballot_sync(0xffffffffu, lane < NCLS && ngl > BIG_G) == 0u) {
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const int c = warp + j * TW;
          {
            // prefetch the next class in the flattened (round, class) sequence; the sequence of the
            // next tile starts again at (0, warp)
            const int rn = (j + 1 < CPW) ? r : ((r + 1 < R) ? r + 1 : 0);
            const int cn_ = (j + 1 < CPW) ? c + TW : warp;
            load_b(bnxt, rn, cn_);
          }
          const int n_c = __shfl_sync(0xffffffffu, total, c), st = __shfl_sync(0xffffffffu, mystart, c);
          run_groups(st, 0, (n_c + 7) >> 3);
#pragma unroll
          for (int i = 0; i < NBF; ++i) bcur[i] = bnxt[i];
        }
      } else {
        uint32_t rest = __ballot_sync(0xffffffffu, lane < NCLS && ngl > 0 && (ngl > BIG_G || (lane % TW) == warp)) & ~(1u << warp);
        int c = warp;
        for (;;) {
          const int cn = rest ? __ffs(rest) - 1 : -1; // next class of this warp in this round
          load_b(bnxt, cn >= 0 ? r : ((r + 1 < R) ? r + 1 : 0), cn >= 0 ? cn : warp);
          const int n_c = __shfl_sync(0xffffffffu, total, c), st = __shfl_sync(0xffffffffu, mystart, c);
          const int ng_all = (n_c + 7) >> 3;
          const bool big = ng_all > BIG_G;
          run_groups(st, big ? (ng_all * warp) / TW : 0, big ? (ng_all * (warp + 1)) / TW : ng_all);
#pragma unroll
          for (int i = 0; i < NBF; ++i) bcur[i] = bnxt[i];
          if (cn < 0) break;
          c = cn;
          rest &= rest - 1;
        }
      }
    }
    TPH_MARK(6)
    if (R > 0) named_bar_sync(bar_id, TW * 32); // rows of the last round complete
    TPH_MARK(7)

    // ---- root: out = row . R[d_{n-1}] for the home rows
    if (deep) {
      // cooperative: the CPR lanes of a group read one state row and one root-table row chunk by chunk (coalesced:
      // 1 L1 wavefront per table row instead of 4), CPR rows per group and pass; a reduce-scatter over the group
      // leaves lane j with the dot product of the group's j-th row, and the 32 results of a pass are 32
      // consecutive points (owner lane = j * RPI + group)
      // KH passes share one memory round trip: their table chunks are all in flight before the first is used
      constexpr int KH = (CPR * 4 >= 64) ? 1 : ((64 / (CPR * 4) < PPL) ? 64 / (CPR * 4) : PPL);
      const int ol_out = coj * RPI + coq;
#pragma unroll
      for (int k0 = 0; k0 < PPL; k0 += KH) {
        double o[KH][2];
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) o[kk][0] = o[kk][1] = 0.0;
        for (int oi = 0; oi < ch.nout; ++oi) {
          const double* Rt = ch.root + ((size_t)oi * CHI << ch.root_bits);
          double2 q[KH][CPR];
#pragma unroll
          for (int kk = 0; kk < KH; ++kk) {
            if (k0 + kk < PPL) {
              const uint32_t ri_own = (uint32_t)(cw[(k0 + kk) % PPL] & RMASK);
#pragma unroll
              for (int s = 0; s < CPR; ++s) {
                const uint32_t ri = __shfl_sync(0xffffffffu, ri_own, s * RPI + coq);
                q[kk][s] = ldg128(Rt + (size_t)ri * CHI + 2 * coj);
              }
            }
          }
#pragma unroll
          for (int kk = 0; kk < KH; ++kk) {
            if (k0 + kk < PPL) {
              double part[CPR];
#pragma unroll
              for (int s = 0; s < CPR; ++s) {
                const double2 v = lds128(row_chunk<CHI>(state_base, warp * PW + (k0 + kk) * 32 + s * RPI + coq, coj));
                part[s] = fma(v.y, q[kk][s].y, v.x * q[kk][s].x);
              }
#pragma unroll
              for (int off = CPR / 2; off > 0; off >>= 1) {
                const bool up = (coj & off) != 0;
#pragma unroll
                for (int i = 0; i < off; ++i) {
                  const double send = up ? part[i] : part[i + off];
                  const double keep = up ? part[i + off] : part[i];
                  part[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
              if (oi == 0) o[kk][0] = part[0];
              else o[kk][1] = part[0];
            }
          }
        }
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) {
          const int64_t p = p0 + (k0 + kk) * 32 + ol_out;
          if (k0 + kk < PPL && p < src.npts) {
            if (out) {
              if (ch.nout == 2) reinterpret_cast<double2*>(out)[p] = make_double2(o[kk][0], o[kk][1]);
              else out[p] = o[kk][0];
            }
            accumulate_point(src, p, o[kk][0], o[kk][1], sum_re, sum_im);
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        const int row = warp * PW + k * 32 + lane;
        const int64_t p = p0 + k * 32 + lane;
        double o0 = 0.0, o1 = 0.0;
        const int sl = (int)(cw[k] & MASK);
        const uint32_t R0 = s_root_u32 + (uint32_t)sl * (CHI * 8), R1 = R0 + (uint32_t)(NCLS * CHI * 8);
#pragma unroll
        for (int j = 0; j < CPR; ++j) {
          const double2 v = lds128(row_chunk<CHI>(state_base, row, j));
          const uint32_t off = (uint32_t)((j ^ (sl & (CPR - 1) & 7)) << 4);
          const double2 q0 = lds128(R0 + off);
          o0 = fma(v.x, q0.x, o0);
          o0 = fma(v.y, q0.y, o0);
          if (ch.nout == 2) {
            const double2 q1 = lds128(R1 + off);
            o1 = fma(v.x, q1.x, o1);
            o1 = fma(v.y, q1.y, o1);
          }
        }
        if (p < src.npts) {
          if (out) {
            if (ch.nout == 2) reinterpret_cast<double2*>(out)[p] = make_double2(o0, o1);
            else out[p] = o0;
          }
          accumulate_point(src, p, o0, o1, sum_re, sum_im);
        }
      }
    }
    TPH_MARK(8)
    // the next tile's leaf rows overwrite home rows only: no barrier needed here, the first barrier
    // of the next tile orders them before any other warp's gather
  }

  TPH_FLUSH
  if (do_sum) {
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
      double xx = 0.0, yy = 0.0;
      for (int w = 0; w < NTEAM * TW; ++w) {
        xx += red[0][w];
        yy += red[1][w];
      }
      partial[2 * blockIdx.x] = xx;
      partial[2 * blockIdx.x + 1] = yy;
    }
  }
}
template <int CHI, int NCLS, int NTEAM, int PW_>
static int launch_mma6_inst(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                            cudaStream_t s) {
  using T6 = Team6<CHI, NCLS, PW_>;
  const ChainMmaDev& c = (src.pcie_bound && p->cmma_light_ok) ? p->cmma_light : p->cmma;
  const size_t smem = (size_t)NTEAM * T6::BYTES + (size_t)3 * NCLS * CHI * 8; // teams + leaf + root (re, im)
  auto kern = chain_mma6_kernel<CHI, NCLS, NTEAM, PW_>;
  TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (src.npts + T6::TP - 1) / T6::TP;
  const int grid = (int)std::min<int64_t>((n_tiles + NTEAM - 1) / NTEAM, p->sm_count);
  const int do_sum = d_partial != nullptr;
  kern<<<grid, NTEAM * 128, smem, s>>>(c, p->digits_mma, src, d_out, p->d_err, d_partial, do_sum);
  TTN_CUDA(cudaGetLastError());
  *n_partial = do_sum ? grid : 0;
  return TTN_OK;
}

// Applies to merged binary chains (build_chain_mma) of width 8, 16 or 32.  TTN_MMA_V6=0 switches the
// kernel off (the ring kernels then run the merged image), =2 runs two teams per CTA (experiments).
// The team count is a plan-time knob (ttn_plan::v6_teams, read from TTN_MMA_V6 by ttn_plan_create).

bool chain_team_applicable(const ttn_plan* p) {
  const ChainMmaDev& c = p->cmma;
  if (!p->v6_teams || !c.merged || c.spr != 1) return false;
  if (c.chi == 8 || c.chi == 16) return c.nsl == 4 || c.nsl == 8 || c.nsl == 16 || c.nsl == 32;
  return c.chi == 32 && (c.nsl == 4 || c.nsl == 8 || c.nsl == 16);
}

int launch_chain_team(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                      cudaStream_t s) {
  const ChainMmaDev& c = p->cmma;
  const int v6 = p->v6_teams;
  if (v6 && c.merged && c.spr == 1) {
#define TTN_V6_CASE(W, N)                                                                                   \
  if (c.chi == W && c.nsl == N)                                                                             \
    return v6 == 2 ? launch_mma6_inst<W, N, 2, 128>(p, src, d_out, d_partial, n_partial, s)                 \
         : v6 == 4 ? launch_mma6_inst<W, N, 4, 96>(p, src, d_out, d_partial, n_partial, s)                  \
                   : launch_mma6_inst<W, N, 3, 128>(p, src, d_out, d_partial, n_partial, s);
    TTN_V6_CASE(16, 4)
    TTN_V6_CASE(16, 8)
    TTN_V6_CASE(16, 16)
    TTN_V6_CASE(16, 32)
    TTN_V6_CASE(8, 4)
    TTN_V6_CASE(8, 8)
    TTN_V6_CASE(8, 16)
    TTN_V6_CASE(8, 32)
#undef TTN_V6_CASE
    // width 32: rows are 256 bytes and a class's B fragments take 64 registers -> 2 teams of 384 points
    if (c.chi == 32 && c.nsl == 4) return launch_mma6_inst<32, 4, 2, 96>(p, src, d_out, d_partial, n_partial, s);
    if (c.chi == 32 && c.nsl == 8) return launch_mma6_inst<32, 8, 2, 96>(p, src, d_out, d_partial, n_partial, s);
    if (c.chi == 32 && c.nsl == 16) return launch_mma6_inst<32, 16, 2, 96>(p, src, d_out, d_partial, n_partial, s);
  }
  set_error("team-sorted DMMA kernel: unsupported width / slice count");
  return TTN_ERR_UNSUPPORTED;
}

#ifdef TTN_TEAM_CLOCKS
int debug_team_clocks(unsigned long long* out12, int reset) {
  unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  TTN_CUDA(cudaMemcpyFromSymbol(out12, g_tphase, sizeof(z)));
  if (reset) TTN_CUDA(cudaMemcpyToSymbol(g_tphase, z, sizeof(z)));
  return TTN_OK;
}
#endif

} // namespace ttn
