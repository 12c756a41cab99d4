// K4 — chain (MPS-shaped) networks on the FP64 tensor pipe: points grouped by digit, DMMA tiles.
//
// Why: a thread-per-point kernel needs one 8-byte matrix element per DFMA per lane, i.e. 256 B
// delivered from shared memory per warp-DFMA; the LSU returns 128 B/clk/SM while the FP64 pipe
// retires a warp-DFMA every 0.5 clk/SM, so that design is capped at 25 % of FP64 peak (measured:
// 24 %, profiles/).  An m8n8k4 DMMA performs 8 FMAs per lane per (A,B) operand pair, and the B
// (matrix) fragment is reused by every 8-point group that selected the same slice.
//
// Common scheme:
//   * the state of a tile of points lives in shared memory, one row of CHI doubles per point (16-byte
//     chunks XOR-swizzled by row so that row gathers and owner accesses avoid bank conflicts);
//   * a ROUND covers one stream position (or `spr` consecutive sites).  Points are counting-sorted by
//     the slice they select there; each class is padded to a multiple of 8 rows (pad rows point at an
//     all-zero row); 8-row groups of one class are gathered as DMMA A fragments, multiplied by the
//     class's site matrix (B fragments, fragment order prepared on the host) and scattered back.
//   * the packed slice stream of a point (K1) is one bit string, bit position = chain position x
//     bits per vertex, consumed a few bits per round.
// Three kernels share it (launch_chain_mma picks):
//   chain_mma6_kernel  merged binary chains (build_chain_mma contracts k vertices per position, and
//                      optionally whole leaf / root groups into tables): independent 4-warp teams,
//                      team-wide sort with shared-memory atomics, site matrices in registers;
//   chain_mma5_kernel  width <= 16, anything else: warp-autonomous (128 points per warp, warp-local
//                      sort), B fragments through a TMA/mbarrier ring, D fragment of one site reused
//                      as A fragment of the next (B rows permuted on the host to match);
//   chain_mma3_kernel  width 32, anything else: CTA-wide sort by specialised front-end warps.
// Complex networks are embedded as real ones of twice the width: row = [re | im],
// M -> [[Re, Im], [-Im, Re]]: exactly the 8 flops per complex MAC of the flop rule.
#include <algorithm>
#include <cmath>
#include <cstring>

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

template <int CHI, int NBAT>
__device__ __forceinline__ void site_mma(double (&dst)[NBAT][CHI / 4], const double (&srcA)[NBAT][CHI / 4],
                                         uint32_t bb) {
  constexpr int NB = CHI / 8, KB = CHI / 4;
#pragma unroll
  for (int b = 0; b < NBAT; ++b)
#pragma unroll
    for (int j = 0; j < KB; ++j) dst[b][j] = 0.0;
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
    double bf[NB];
#pragma unroll
    for (int nbp = 0; nbp < NB; ++nbp) bf[nbp] = lds64(bb + (uint32_t)((kb * NB + nbp) * 32) * 8u);
#pragma unroll
    for (int nbp = 0; nbp < NB; ++nbp)
#pragma unroll
      for (int b = 0; b < NBAT; ++b) dmma884(dst[b][2 * nbp], dst[b][2 * nbp + 1], srcA[b][kb], bf[nbp]);
  }
}

// One batch: NBAT 8-row groups of one class.  The two register tiles ping-pong between A and D
// roles from site to site (the D fragment of one site is the A fragment of the next), so there
// are no register moves between sites.
template <int CHI, int NBAT>
__device__ __forceinline__ void process_batch5(uint32_t state_base, const int (&rows)[4], int tq,
                                               uint32_t stage_base, const int (&boff)[4], int sites, int pad_row) {
  constexpr int NB = CHI / 8, KB = CHI / 4;
  double t0[NBAT][KB], t1[NBAT][KB];
#pragma unroll
  for (int b = 0; b < NBAT; ++b) {
    // class padding slots read the scratch row (all zeros, never written: the stores below skip it)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const double2 v = lds128(row_chunk<CHI>(state_base, rows[b], 4 * nb + tq));
      t0[b][2 * nb] = v.x;
      t0[b][2 * nb + 1] = v.y;
    }
  }
  int s = 0;
  for (; s + 1 < sites; s += 2) {
    site_mma<CHI, NBAT>(t1, t0, stage_base + (uint32_t)boff[s]);
    site_mma<CHI, NBAT>(t0, t1, stage_base + (uint32_t)boff[s + 1]);
  }
  if (s < sites) {
    site_mma<CHI, NBAT>(t1, t0, stage_base + (uint32_t)boff[s]);
#pragma unroll
    for (int b = 0; b < NBAT; ++b)
      if (rows[b] != pad_row) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
          sts128(row_chunk<CHI>(state_base, rows[b], 4 * nb + tq), t1[b][2 * nb], t1[b][2 * nb + 1]);
      }
  } else {
#pragma unroll
    for (int b = 0; b < NBAT; ++b)
      if (rows[b] != pad_row) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
          sts128(row_chunk<CHI>(state_base, rows[b], 4 * nb + tq), t0[b][2 * nb], t0[b][2 * nb + 1]);
      }
  }
}

// compact base-2 digit entry for the branch-free fast path (16 bytes -> one LDS.128)
struct __align__(16) Digit2 {
  double thr1;   // |index_value_to_scalar(ind, 1)|
  uint32_t sh;   // shift inside the word
  uint32_t wv;   // (site << 16) | (word << 8) | stride
};

// =====================================================================================
// v3: warp-specialised version of the scheme above.  The MMA warps do nothing but
// gather -> DMMA -> scatter; everything else runs concurrently on front-end warps:
//   * digit warps   compute the packed slice streams of the NEXT tile (K1),
//   * the list warp counting-sorts the points of the CURRENT tile by class, one round ahead
//                   (double-buffered lists),
//   * B fragments are prefetched by MMA thread 0 into the ring slot the end-of-round barrier has
//     just freed (no producer warp, no "empty" barriers).
// Handshakes are mbarriers: tile_ready/tile_free (slice streams), list_full/list_empty, ring full.
constexpr int kFeMaxSites = 160; // static shared-memory copies of the digit tables (within the 12 KB the
constexpr int kFeMaxThr = 640;   // launchers reserve); larger networks take the chain / generic kernels

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

template <int CHI, int P, int NMW, int GB, bool B2, int NSLT, int SPRT>
__global__ void __launch_bounds__(NMW * 32 + 128, 1)
    chain_mma3_kernel(ChainMmaDev ch, DigitTable dg, CoordSource src, double* __restrict__ out, int* err,
                      double* __restrict__ partial, int do_sum, int n_stage, int resident,
                      uint32_t stage_stride) {
  constexpr int NTM = NMW * 32;    // MMA threads
  constexpr int NDT = 96;          // digit threads (3 warps); the 4th front-end warp builds lists
  constexpr int PPT = P / NTM;
  constexpr int CPR = CHI / 2;
  constexpr int LIST_CAP = P + 8 * kMaxClasses;
  static_assert(P % NTM == 0 && P % 32 == 0 && P / 32 <= 32, "tile shape");

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t ring_full[kMmaMaxStages];
  __shared__ __align__(8) uint64_t list_full[2], list_empty[2], tile_ready[2], tile_free[2];
  __shared__ double red[2][NMW];
  __shared__ int meta[2][kMaxClasses + 2];
  __shared__ DigitEntry s_ent[B2 ? 1 : kFeMaxSites];
  __shared__ double s_thr[B2 ? 1 : kFeMaxThr];
  __shared__ Digit2 s_d2[B2 ? kFeMaxSites : 1];
  __shared__ int s_cptr[TTN_MAX_COORDS + 1];

  unsigned char* state_p = smem;                                                    // (P + 8) rows
  ulonglong2* words = reinterpret_cast<ulonglong2*>(smem + (size_t)(P + 8) * CHI * 8); // [2][P]
  uint16_t* lists = reinterpret_cast<uint16_t*>(words + 2 * P);                     // [2][LIST_CAP]
  unsigned char* ring = reinterpret_cast<unsigned char*>(lists + 2 * LIST_CAP);
  ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ring) + 127) & ~(uintptr_t)127);

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < n_stage; ++s) mbar_init(smem_u32(&ring_full[s]), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&list_full[b]), 1);
      mbar_init(smem_u32(&list_empty[b]), NMW);
      mbar_init(smem_u32(&tile_ready[b]), NDT / 32);
      mbar_init(smem_u32(&tile_free[b]), NMW + 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = tid; i < 8 * CHI; i += NTM + 128) reinterpret_cast<double*>(state_p + (size_t)P * CHI * 8)[i] = 0.0;
  for (int i = tid; i <= dg.n_coords; i += NTM + 128) s_cptr[i] = dg.coord_ptr[i];
  if (B2) {
    for (int i = tid; i < dg.n_sites; i += NTM + 128) {
      const DigitEntry e = dg.entries[i];
      Digit2 d2;
      d2.thr1 = dg.thr[e.thr_off + 1];
      d2.sh = (uint32_t)e.shift;
      d2.wv = ((uint32_t)e.site << 16) | ((uint32_t)e.word << 8) | (uint32_t)e.stride;
      s_d2[i] = d2;
    }
  } else {
    for (int i = tid; i < dg.n_sites; i += NTM + 128) s_ent[i] = dg.entries[i];
    int nthr = 0; // thr[] length = max over entries of thr_off + base
    for (int i = 0; i < dg.n_sites; ++i) nthr = max(nthr, dg.entries[i].thr_off + dg.entries[i].base);
    for (int i = tid; i < nthr; i += NTM + 128) s_thr[i] = dg.thr[i];
  }
  __syncthreads();

  const int64_t n_tiles = (src.npts + P - 1) / P;
  const int n_rounds = ch.n_rounds, spr = SPRT ? SPRT : ch.spr, nsl = NSLT ? NSLT : ch.nsl, n_steps = ch.n_steps;
  const uint32_t site_bytes = (uint32_t)nsl * CHI * CHI * 8;
  const uint32_t ring_base = smem_u32(ring);
  const int bits = NSLT ? slice_bits(NSLT) : ch.bits;
  const uint64_t MASK = (1ull << bits) - 1ull;
  const int per_word = NSLT ? (NSLT <= 1 ? (1 << 30) : 64 / slice_bits(NSLT > 1 ? NSLT : 2)) : ch.per_word;
  const bool pow2 = (nsl & (nsl - 1)) == 0;
  const int lane = tid & 31;
  int64_t my_tiles = 0;
  if ((int64_t)blockIdx.x < n_tiles) my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  // slice of chain position `pos` of a packed stream
  auto slice_at = [&](const ulonglong2& w, int pos) -> int {
    if (bits == 0) return 0;
    const int wi = pos / per_word, sh = (pos - wi * per_word) * bits;
    return (int)(((wi ? w.y : w.x) >> sh) & MASK);
  };

  if (tid >= NTM + 32) {
    // ===== digit warps: K1 for the tiles of this CTA, one tile ahead of the MMA warps =====
    const int dtid = tid - (NTM + 32);
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int64_t tile = blockIdx.x + i * gridDim.x;
      const int b = (int)(i & 1);
      mbar_wait(smem_u32(&tile_free[b]), (uint32_t)(((i >> 1) & 1) ^ 1));
      for (int base = dtid; base < P; base += NDT * 4) {
        double x[4];
        uint64_t w0[4], w1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w0[q] = w1[q] = 0;
        for (int c = 0; c < dg.n_coords; ++c) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int pt = base + q * NDT;
            const int64_t p = tile * P + pt;
            x[q] = 0.0;
            if (pt < P && p < src.npts) {
              x[q] = load_coord(src, p, c);
              if (!coord_in_domain(x[q])) {
                atomicOr(err, 1);
                x[q] = 0.0;
              }
            }
          }
          if (B2) {
            for (int e_i = s_cptr[c]; e_i < s_cptr[c + 1]; ++e_i) {
              const Digit2 e = s_d2[e_i];
              const uint32_t stride = e.wv & 0xffu;
              const bool hi = ((e.wv >> 8) & 0xffu) != 0;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const bool ge = src.digits ? (given_digit(src, tile * P + base + q * NDT, dg.n_sites, (int)(e.wv >> 16), 2, err) != 0)
                                           : (x[q] >= e.thr1);
                x[q] = __dsub_rn(x[q], ge ? e.thr1 : 0.0);
                const uint64_t bb = (uint64_t)(ge ? stride : 0u) << e.sh;
                if (hi) w1[q] += bb;
                else w0[q] += bb;
              }
            }
          } else {
            for (int e_i = s_cptr[c]; e_i < s_cptr[c + 1]; ++e_i) {
              const DigitEntry e = s_ent[e_i];
              const double* thr = s_thr + e.thr_off;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int v = src.digits ? given_digit(src, tile * P + base + q * NDT, dg.n_sites, e.site, e.base, err)
                                         : greedy_digit_smem(x[q], thr, e.base);
                const uint64_t bb = (uint64_t)(v * e.stride) << e.shift;
                w0[q] += (e.word == 0) ? bb : 0ull;
                w1[q] += (e.word == 1) ? bb : 0ull;
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int pt = base + q * NDT;
          if (pt < P) words[b * P + pt] = make_ulonglong2(w0[q], w1[q]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tile_ready[b]));
    }
    return;
  }

  if (tid >= NTM) {
    // ===== list warp: counting sort by class, one round ahead (double-buffered lists) =====
    int64_t q = 0; // global round counter of this CTA
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int b = (int)(i & 1);
      mbar_wait(smem_u32(&tile_ready[b]), (uint32_t)((i >> 1) & 1));
      const ulonglong2* wt = words + b * P;
      for (int r = 0; r < n_rounds; ++r, ++q) {
        const int lb = (int)(q & 1);
        uint16_t* list = lists + lb * LIST_CAP;
        const int sites = SPRT ? SPRT : min(spr, n_steps - r * spr);
        int ncls = 1;
        for (int k = 0; k < sites; ++k) ncls *= nsl;
        const int pos0 = 1 + r * spr;
        mbar_wait(smem_u32(&list_empty[lb]), (uint32_t)(((q >> 1) & 1) ^ 1));
        // (word, shift) of the round's sites in the packed stream: uniform, hoisted
        int s_w[4] = {0, 0, 0, 0}, s_sh[4] = {0, 0, 0, 0}, s_mul[4] = {0, 0, 0, 0};
        {
          int mul = 1;
          for (int sI = 0; sI < sites && sI < 4; ++sI) {
            const int pos = pos0 + sI;
            const int wi = bits ? pos / per_word : 0;
            s_w[sI] = wi;
            s_sh[sI] = bits ? (pos - wi * per_word) * bits : 0;
            s_mul[sI] = mul;
            mul *= nsl;
          }
        }
        auto class_of = [&](const ulonglong2& w) -> int {
          int cls = 0;
#pragma unroll
          for (int sI = 0; sI < 4; ++sI)
            cls += (int)(((s_w[sI] ? w.y : w.x) >> s_sh[sI]) & MASK) * s_mul[sI]; // s_mul == 0 beyond `sites`
          return cls;
        };
        if (ncls <= 4 && sites <= 4) {
          // ---- bitmap rank: lane j keeps the membership mask of points 32j..32j+31 per class; all
          // ballots are independent across j, so the loop pipelines instead of serialising
          uint32_t mk0 = 0, mk1 = 0, mk2 = 0, mk3 = 0;
          uint64_t cache = 0; // 2 bits per j
#pragma unroll 8
          for (int j = 0; j < P / 32; ++j) {
            const int cls = class_of(wt[j * 32 + lane]);
            cache |= (uint64_t)cls << (2 * j);
            const uint32_t m0 = __ballot_sync(0xffffffffu, cls == 0);
            const uint32_t m1 = __ballot_sync(0xffffffffu, cls == 1);
            const uint32_t m2 = __ballot_sync(0xffffffffu, cls == 2);
            const uint32_t m3 = __ballot_sync(0xffffffffu, cls == 3);
            if (lane == j) {
              mk0 = m0; mk1 = m1; mk2 = m2; mk3 = m3;
            }
          }
          int i0 = __popc(mk0), i1 = __popc(mk1), i2 = __popc(mk2), i3 = __popc(mk3);
          const int n0 = i0, n1 = i1, n2 = i2, n3 = i3;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int a0 = __shfl_up_sync(0xffffffffu, i0, o), a1 = __shfl_up_sync(0xffffffffu, i1, o);
            const int a2 = __shfl_up_sync(0xffffffffu, i2, o), a3 = __shfl_up_sync(0xffffffffu, i3, o);
            if (lane >= o) {
              i0 += a0; i1 += a1; i2 += a2; i3 += a3;
            }
          }
          const int t0 = __shfl_sync(0xffffffffu, i0, 31), t1 = __shfl_sync(0xffffffffu, i1, 31);
          const int t2 = __shfl_sync(0xffffffffu, i2, 31), t3 = __shfl_sync(0xffffffffu, i3, 31);
          const int st0 = 0, st1 = st0 + ((t0 + 7) & ~7), st2 = st1 + ((t1 + 7) & ~7), st3 = st2 + ((t2 + 7) & ~7);
          const int total = st3 + ((t3 + 7) & ~7);
          if (lane <= kMaxClasses)
            meta[lb][lane] = lane == 0 ? st0 : (lane == 1 ? st1 : (lane == 2 ? st2 : (lane == 3 ? st3 : total)));
          if (lane < 8) { // class padding -> scratch row
            if ((t0 & 7) && lane >= (t0 & 7)) list[st0 + (t0 & ~7) + lane] = (uint16_t)P;
            if ((t1 & 7) && lane >= (t1 & 7)) list[st1 + (t1 & ~7) + lane] = (uint16_t)P;
            if ((t2 & 7) && lane >= (t2 & 7)) list[st2 + (t2 & ~7) + lane] = (uint16_t)P;
            if ((t3 & 7) && lane >= (t3 & 7)) list[st3 + (t3 & ~7) + lane] = (uint16_t)P;
          }
          // slot base of word j per class = class start + points of the class in earlier words
          const int b0 = st0 + i0 - n0, b1 = st1 + i1 - n1, b2 = st2 + i2 - n2, b3 = st3 + i3 - n3;
          const uint32_t lt = (1u << lane) - 1u;
#pragma unroll 8
          for (int j = 0; j < P / 32; ++j) {
            const int cls = (int)((cache >> (2 * j)) & 3);
            const int pb0 = __shfl_sync(0xffffffffu, b0, j), pb1 = __shfl_sync(0xffffffffu, b1, j);
            const int pb2 = __shfl_sync(0xffffffffu, b2, j), pb3 = __shfl_sync(0xffffffffu, b3, j);
            const uint32_t M0 = __shfl_sync(0xffffffffu, mk0, j), M1 = __shfl_sync(0xffffffffu, mk1, j);
            const uint32_t M2 = __shfl_sync(0xffffffffu, mk2, j), M3 = __shfl_sync(0xffffffffu, mk3, j);
            const int B = cls == 0 ? pb0 : (cls == 1 ? pb1 : (cls == 2 ? pb2 : pb3));
            const uint32_t M = cls == 0 ? M0 : (cls == 1 ? M1 : (cls == 2 ? M2 : M3));
            list[B + __popc(M & lt)] = (uint16_t)(j * 32 + lane);
          }
        } else {
          // pass 1: classes (cached 4 bits each) and per-class counts (lane c counts class c)
          uint64_t cache0 = 0, cache1 = 0;
          int mycnt = 0;
          for (int j = 0; j < P / 32; ++j) {
            int cls = 0;
            if (sites <= 4) {
              cls = class_of(wt[j * 32 + lane]);
            } else {
              const ulonglong2 w = wt[j * 32 + lane];
              int mul = 1;
              for (int s = 0; s < sites; ++s) {
                cls += slice_at(w, pos0 + s) * mul;
                mul *= nsl;
              }
            }
            if (j < 16) cache0 |= (uint64_t)cls << (4 * j);
            else cache1 |= (uint64_t)cls << (4 * (j - 16));
            for (int c = 0; c < ncls; ++c) {
              const uint32_t m = __ballot_sync(0xffffffffu, cls == c);
              if (lane == c) mycnt += __popc(m);
            }
          }
          // class start rows (each class padded to a multiple of 8 rows)
          const int padded = (lane < ncls) ? ((mycnt + 7) & ~7) : 0;
          int incl = padded;
  #pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
          }
          const int mystart = incl - padded;
          if (lane <= kMaxClasses) meta[lb][lane] = mystart; // lanes >= ncls hold the total
          if (lane < ncls)
            for (int k = mycnt; k < padded; ++k) list[mystart + k] = (uint16_t)P; // padding -> scratch row
          // pass 2: slots
          int run = mystart;
          for (int j = 0; j < P / 32; ++j) {
            const int cls = (int)(((j < 16) ? (cache0 >> (4 * j)) : (cache1 >> (4 * (j - 16)))) & 15);
            int slot = 0;
            for (int c = 0; c < ncls; ++c) {
              const uint32_t m = __ballot_sync(0xffffffffu, cls == c);
              const int base = __shfl_sync(0xffffffffu, run, c);
              if (cls == c) slot = base + __popc(m & ((1u << lane) - 1u));
              if (lane == c) run += __popc(m);
            }
            list[slot] = (uint16_t)(j * 32 + lane);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&list_full[lb]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tile_free[b]));
    }
    return;
  }

  // ===== MMA warps =====
  const int warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const uint32_t state_base = smem_u32(state_p);
  const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(ch.frags);
  const int64_t total_q = my_tiles * n_rounds;
  auto issue_round = [&](int64_t qq) { // thread 0 only: B fragments of global round qq -> its ring slot
    const int r = (int)(qq % n_rounds);
    const uint32_t s = resident ? (uint32_t)r : (uint32_t)(qq % n_stage);
    const uint32_t bytes = (uint32_t)(SPRT ? SPRT : min(spr, n_steps - r * spr)) * site_bytes;
    mbar_expect_tx(smem_u32(&ring_full[s]), bytes);
    bulk_g2s(ring_base + s * stage_stride, gsrc + (size_t)r * spr * site_bytes, bytes, smem_u32(&ring_full[s]));
  };
  if (tid == 0) {
    const int64_t first = resident ? min((int64_t)n_rounds, total_q) : min((int64_t)n_stage, total_q);
    for (int64_t qq = 0; qq < first; ++qq) issue_round(qq);
  }
  double sum_re = 0.0, sum_im = 0.0;
  int64_t q = 0;
  for (int64_t i = 0; i < my_tiles; ++i) {
    const int64_t tile = blockIdx.x + i * gridDim.x;
    const int b = (int)(i & 1);
    const ulonglong2* wt = words + b * P;
    mbar_wait(smem_u32(&tile_ready[b]), (uint32_t)((i >> 1) & 1));
    // ---- leaf: row(point) = L[d_0]
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int row = k * NTM + tid;
      const double* L = ch.leaf + (size_t)slice_at(wt[row], 0) * CHI;
#pragma unroll
      for (int j = 0; j < CPR; ++j) sts128(row_chunk<CHI>(state_base, row, j), __ldg(L + 2 * j), __ldg(L + 2 * j + 1));
    }
    named_bar_sync(1, NTM);
    // ---- rounds
    for (int r = 0; r < n_rounds; ++r, ++q) {
      const int lb = (int)(q & 1);
      const uint16_t* list = lists + lb * LIST_CAP;
      const int sites = SPRT ? SPRT : min(spr, n_steps - r * spr);
      int ncls = 1;
      for (int k = 0; k < sites; ++k) ncls *= nsl;
      mbar_wait(smem_u32(&list_full[lb]), (uint32_t)((q >> 1) & 1));
      const int mystart = meta[lb][min(lane, kMaxClasses)];
      const int total_rows = __shfl_sync(0xffffffffu, mystart, ncls);
      const uint32_t s_use = resident ? (uint32_t)r : (uint32_t)(q % n_stage);
      mbar_wait(smem_u32(&ring_full[s_use]), resident ? 0u : (uint32_t)((q / n_stage) & 1));
      const uint32_t stage_base = ring_base + s_use * stage_stride + (uint32_t)lane * 8u;

      const int n_groups = total_rows >> 3;
      const int gpw = (n_groups + NMW - 1) / NMW;
      const int gw0 = warp * gpw, gw1 = min(n_groups, gw0 + gpw);
      int rows_nx[4];
      for (int c = 0; c < ncls; ++c) { // classes outermost: B offsets once per class, no divisions
        const int cs = __shfl_sync(0xffffffffu, mystart, c) >> 3, ce = __shfl_sync(0xffffffffu, mystart, c + 1) >> 3;
        int gi = max(cs, gw0);
        const int gend = min(ce, gw1);
        if (gi >= gend) continue;
        int boff[4] = {0, 0, 0, 0};
        {
          int crem = c;
#pragma unroll
          for (int si = 0; si < 4; ++si) {
            if (si < sites) {
              const int dd = pow2 ? (crem & (int)MASK) : (crem % nsl);
              crem = pow2 ? (crem >> bits) : (crem / nsl);
              boff[si] = (si * nsl + dd) * (CHI * CHI * 8);
            }
          }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) rows_nx[b] = (int)list[(min(gi + b, gend - 1) << 3) + g];
        while (gi < gend) {
          const int nbat = min(GB, gend - gi);
          int rows[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) rows[b] = rows_nx[b];
          const int gnext = gi + nbat;
          if (gnext < gend) {
#pragma unroll
            for (int b = 0; b < 4; ++b) rows_nx[b] = (int)list[(min(gnext + b, gend - 1) << 3) + g];
          }
          if (nbat >= 4 && GB >= 4) process_batch5<CHI, (GB >= 4 ? 4 : 1)>(state_base, rows, tq, stage_base, boff, sites, P);
          else if (nbat == 3 && GB >= 3) process_batch5<CHI, (GB >= 3 ? 3 : 1)>(state_base, rows, tq, stage_base, boff, sites, P);
          else if (nbat == 2 && GB >= 2) process_batch5<CHI, (GB >= 2 ? 2 : 1)>(state_base, rows, tq, stage_base, boff, sites, P);
          else process_batch5<CHI, 1>(state_base, rows, tq, stage_base, boff, sites, P);
          gi = gnext;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&list_empty[lb]));
      named_bar_sync(1, NTM); // rows change hands between rounds; the ring slot is free again
      if (tid == 0 && !resident && q + n_stage < total_q) issue_round(q + n_stage);
    }
    // ---- root: out = row . R[d_{n-1}]
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int row = k * NTM + tid;
      const int64_t p = tile * P + row;
      double o0 = 0.0, o1 = 0.0;
      if (ch.n_vertices > 1) {
        const double* R0 = ch.root + (size_t)slice_at(wt[row], ch.root_pos) * CHI;
        const double* R1 = R0 + (size_t)nsl * CHI;
#pragma unroll
        for (int j = 0; j < CPR; ++j) {
          const double2 v = lds128(row_chunk<CHI>(state_base, row, j));
          o0 = fma(v.x, __ldg(R0 + 2 * j), o0);
          o0 = fma(v.y, __ldg(R0 + 2 * j + 1), o0);
          if (ch.nout == 2) {
            o1 = fma(v.x, __ldg(R1 + 2 * j), o1);
            o1 = fma(v.y, __ldg(R1 + 2 * j + 1), o1);
          }
        }
      } else {
        o0 = lds64(row_chunk<CHI>(state_base, row, 0));
        if (ch.nout == 2) o1 = lds64(row_chunk<CHI>(state_base, row, CPR / 2));
      }
      if (p < src.npts) {
        if (out) {
          if (ch.nout == 2) reinterpret_cast<double2*>(out)[p] = make_double2(o0, o1);
          else out[p] = o0;
        }
        accumulate_point(src, p, o0, o1, sum_re, sum_im);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&tile_free[b]));
  }

  if (do_sum) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum_re += __shfl_down_sync(0xffffffffu, sum_re, o);
      sum_im += __shfl_down_sync(0xffffffffu, sum_im, o);
    }
    if (lane == 0) {
      red[0][warp] = sum_re;
      red[1][warp] = sum_im;
    }
    named_bar_sync(1, NTM);
    if (tid == 0) {
      double x = 0.0, y = 0.0;
      for (int w = 0; w < NMW; ++w) {
        x += red[0][w];
        y += red[1][w];
      }
      partial[2 * blockIdx.x] = x;
      partial[2 * blockIdx.x + 1] = y;
    }
  }
}

// =====================================================================================
// v5: warp-autonomous version.  Every MMA warp owns PW = 128 points end to end: digits (K1), leaf
// rows, a warp-local counting sort per round (ballots only, no CTA barrier), gather -> DMMA ->
// scatter on its private state rows, root.  The warps of a CTA share nothing but the TMA-fed
// B-fragment ring (full/empty mbarriers), so they drift apart freely: while one warp sorts or
// gathers, the other warp on its SMSP keeps the DMMA pipe busy (one warp alone can saturate it:
// scripts/microbench/dmma_issue.cu).  Price: classes are padded to 8 rows per warp, not per CTA.
#ifdef TTN_PHASE_CLOCKS
__device__ unsigned long long g_phase[8];
#define PH_DECL long long ph_t = clock64(); unsigned long long ph_acc[6] = {0, 0, 0, 0, 0, 0};
#define PH_MARK(i) { const long long t_ = clock64(); ph_acc[i] += (unsigned long long)(t_ - ph_t); ph_t = t_; }
#define PH_FLUSH if (lane == 0) { for (int i_ = 0; i_ < 6; ++i_) atomicAdd(&g_phase[i_], ph_acc[i_]); atomicAdd(&g_phase[7], 1ull); }
#else
#define PH_DECL
#define PH_MARK(i)
#define PH_FLUSH
#endif

template <int CHI, int NMW, int GB, bool B2, int NSLT, int SPRT>
__global__ void __launch_bounds__(NMW * 32 + 32, 1)
    chain_mma5_kernel(ChainMmaDev ch, DigitTable dg, CoordSource src, double* __restrict__ out, int* err,
                      double* __restrict__ partial, int do_sum, int n_stage, int resident,
                      uint32_t stage_stride) {
  constexpr int PW = 128;          // points per warp sub-tile
  constexpr int PPL = PW / 32;     // points per lane
  constexpr int ROWS = PW + 8;     // + scratch rows (class padding target = row PW)
  constexpr int CPR = CHI / 2;
  constexpr int LIST_CAP = PW + 8 * kMaxClasses;
  constexpr int NT = NMW * 32;

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[kMmaMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMmaMaxStages];
  __shared__ double red[2][NMW];
  __shared__ DigitEntry s_ent[B2 ? 1 : kFeMaxSites];
  __shared__ double s_thr[B2 ? 1 : kFeMaxThr];
  __shared__ Digit2 s_d2[B2 ? kFeMaxSites : 1];
  __shared__ int s_cptr[TTN_MAX_COORDS + 1];

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < n_stage; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), NMW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = tid; i <= dg.n_coords; i += NT + 32) s_cptr[i] = dg.coord_ptr[i];
  if (B2) {
    for (int i = tid; i < dg.n_sites; i += NT + 32) {
      const DigitEntry e = dg.entries[i];
      Digit2 d2;
      d2.thr1 = dg.thr[e.thr_off + 1];
      d2.sh = (uint32_t)e.shift;
      d2.wv = ((uint32_t)e.site << 16) | ((uint32_t)e.word << 8) | (uint32_t)e.stride;
      s_d2[i] = d2;
    }
  } else {
    for (int i = tid; i < dg.n_sites; i += NT + 32) s_ent[i] = dg.entries[i];
    int nthr = 0;
    for (int i = 0; i < dg.n_sites; ++i) nthr = max(nthr, dg.entries[i].thr_off + dg.entries[i].base);
    for (int i = tid; i < nthr; i += NT + 32) s_thr[i] = dg.thr[i];
  }
  // per-warp regions: state rows, then lists, then the shared ring
  unsigned char* state_all = smem;
  uint8_t* list_all = reinterpret_cast<uint8_t*>(smem + (size_t)NMW * ROWS * CHI * 8);
  unsigned char* ring = reinterpret_cast<unsigned char*>(list_all) + NMW * 256;
  ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ring) + 127) & ~(uintptr_t)127);
  static_assert(LIST_CAP <= 256, "list capacity");
  for (int w = 0; w < NMW; ++w)
    for (int i = tid; i < 8 * CHI; i += NT + 32)
      reinterpret_cast<double*>(state_all + ((size_t)w * ROWS + PW) * CHI * 8)[i] = 0.0;
  __syncthreads();

  const int64_t n_sub = (src.npts + PW - 1) / PW;                 // sub-tiles in the launch
  const int64_t stride = (int64_t)gridDim.x * NMW;
  const int64_t n_iter = (n_sub + stride - 1) / stride;           // identical for every warp: ring lockstep
  // NSLT / SPRT != 0: slices per vertex and sites per round are compile-time constants (the host
  // pads the chain with identity sites so that every round is full), which lets the compiler
  // unroll the class and site loops and drop every division from the hot path
  const int n_rounds = ch.n_rounds, spr = SPRT ? SPRT : ch.spr, nsl = NSLT ? NSLT : ch.nsl, n_steps = ch.n_steps;
  const uint32_t site_bytes = (uint32_t)nsl * CHI * CHI * 8;
  const uint32_t ring_base = smem_u32(ring);

  if (tid >= NT) {
    // ===== producer warp: one elected lane streams the rounds' B fragments =====
    if (tid == NT && n_rounds > 0) {
      const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(ch.frags);
      if (resident) {
        for (int r = 0; r < n_rounds; ++r) {
          const uint32_t bytes = (uint32_t)(SPRT ? SPRT : min(spr, n_steps - r * spr)) * site_bytes;
          mbar_expect_tx(smem_u32(&full_bar[r]), bytes);
          bulk_g2s(ring_base + (uint32_t)r * stage_stride, gsrc + (size_t)r * spr * site_bytes, bytes,
                   smem_u32(&full_bar[r]));
        }
      } else {
        uint32_t slot = 0, phase = 0;
        for (int64_t it = 0; it < n_iter; ++it) {
          for (int r = 0; r < n_rounds; ++r) {
            const uint32_t bytes = (uint32_t)(SPRT ? SPRT : min(spr, n_steps - r * spr)) * site_bytes;
            mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1u);
            mbar_expect_tx(smem_u32(&full_bar[slot]), bytes);
            bulk_g2s(ring_base + slot * stage_stride, gsrc + (size_t)r * spr * site_bytes, bytes,
                     smem_u32(&full_bar[slot]));
            if (++slot == (uint32_t)n_stage) {
              slot = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
    return;
  }

  // ===== autonomous MMA warps =====
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const uint32_t state_base = smem_u32(state_all + (size_t)warp * ROWS * CHI * 8);
  uint8_t* list = list_all + warp * 256;
  const int bits = NSLT ? slice_bits(NSLT) : ch.bits;
  const uint64_t MASK = (1ull << bits) - 1ull;
  const uint32_t lt = (1u << lane) - 1u;
  const bool pow2 = (nsl & (nsl - 1)) == 0; // slice index == bit field of the stream
  double sum_re = 0.0, sum_im = 0.0;
  uint32_t slot = 0, phase = 0;
  PH_DECL

  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t sub = (int64_t)blockIdx.x * NMW + warp + it * stride;
    const bool live_sub = sub < n_sub;     // warp-uniform
    uint64_t w0[PPL], w1[PPL], cw[PPL];
    if (live_sub) {
      // ---- K1: digits of the lane's PPL points (interleaved for ILP)
      double x[PPL];
#pragma unroll
      for (int k = 0; k < PPL; ++k) w0[k] = w1[k] = 0;
      for (int c = 0; c < dg.n_coords; ++c) {
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          const int64_t p = sub * PW + k * 32 + lane;
          x[k] = 0.0;
          if (p < src.npts) {
            x[k] = load_coord(src, p, c);
            if (!coord_in_domain(x[k])) {
              atomicOr(err, 1);
              x[k] = 0.0;
            }
          }
        }
        if (B2 && ch.run_L[c] > 0 && !src.digits) {
          // whole coordinate at once: digits = bits of floor(x * 2^L) (exact; see build_chain_mma),
          // placed as one run of L consecutive stream positions
          const int L = ch.run_L[c], plow = ch.run_plow[c];
          const double scale = ch.run_scale[c];
          const bool rev = ch.run_rev[c] != 0;
#pragma unroll
          for (int k = 0; k < PPL; ++k) {
            unsigned long long q = x[k] >= 1.0 ? ((1ull << L) - 1ull) : (unsigned long long)(x[k] * scale);
            if (rev) q = __brevll(q) >> (64 - L);
            if (plow < 64) {
              w0[k] += q << plow;
              if (plow + L > 64) w1[k] += q >> (64 - plow);
            } else {
              w1[k] += q << (plow - 64);
            }
          }
        } else if (B2) {
          // base 2: the greedy loop is one compare + one subtract (no divergence), 4 points in flight
          for (int e_i = s_cptr[c]; e_i < s_cptr[c + 1]; ++e_i) {
            const Digit2 e = s_d2[e_i];
            const uint32_t stride = e.wv & 0xffu;
            const bool hi = ((e.wv >> 8) & 0xffu) != 0;
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
              const bool ge = src.digits ? (given_digit(src, sub * PW + k * 32 + lane, dg.n_sites, (int)(e.wv >> 16), 2, err) != 0)
                                         : (x[k] >= e.thr1);
              x[k] = __dsub_rn(x[k], ge ? e.thr1 : 0.0);
              const uint64_t bb = (uint64_t)(ge ? stride : 0u) << e.sh;
              if (hi) w1[k] += bb;
              else w0[k] += bb;
            }
          }
        } else {
          for (int e_i = s_cptr[c]; e_i < s_cptr[c + 1]; ++e_i) {
            const DigitEntry e = s_ent[e_i];
            const double* thr = s_thr + e.thr_off;
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
              const int v = src.digits ? given_digit(src, sub * PW + k * 32 + lane, dg.n_sites, e.site, e.base, err)
                                       : greedy_digit_smem(x[k], thr, e.base);
              const uint64_t bb = (uint64_t)(v * e.stride) << e.shift;
              w0[k] += (e.word == 0) ? bb : 0ull;
              w1[k] += (e.word == 1) ? bb : 0ull;
            }
          }
        }
      }
      // ---- leaf rows
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        cw[k] = w0[k];
        const int row = k * 32 + lane;
        const double* L = ch.leaf + (size_t)(cw[k] & MASK) * CHI;
#pragma unroll
        for (int j = 0; j < CPR; ++j) sts128(row_chunk<CHI>(state_base, row, j), __ldg(L + 2 * j), __ldg(L + 2 * j + 1));
      }
    }
    // the packed stream is consumed as a 128-bit shift register (cw = low word, w1 = high word):
    // position p of the stream is bits [p * bits, (p + 1) * bits), whatever the field width
    auto shift_stream = [&](int nb) {
      if (nb == 0) return;
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        cw[k] = (cw[k] >> nb) | (w1[k] << (64 - nb));
        w1[k] >>= nb;
      }
    };
    if (live_sub) shift_stream(bits);
    __syncwarp();
    PH_MARK(0)

    for (int r = 0; r < n_rounds; ++r) {
      const int sites = SPRT ? SPRT : min(spr, n_steps - r * spr);
      const uint32_t s_use = resident ? (uint32_t)r : slot;
      if (live_sub) {
        int ncls = 1;
        for (int k = 0; k < sites; ++k) ncls *= nsl;
        int cls[PPL];
        if (pow2) {
          // slices are bit fields: the class is the next bits*sites bits of the stream
          const int nb_ = bits * sites;
          const uint64_t rmask = (1ull << nb_) - 1ull;
#pragma unroll
          for (int k = 0; k < PPL; ++k) cls[k] = (int)(cw[k] & rmask);
          shift_stream(nb_);
        } else {
#pragma unroll
          for (int k = 0; k < PPL; ++k) cls[k] = 0;
          int mul = 1;
          for (int s = 0; s < sites; ++s) {
#pragma unroll
            for (int k = 0; k < PPL; ++k) cls[k] += (int)(cw[k] & MASK) * mul;
            mul *= nsl;
            shift_stream(bits);
          }
        }
        PH_MARK(5)
        // ---- warp-local counting sort by class (each class padded to 8 rows), branch-free:
        // match.any gives every lane the mask of its class-mates in a slice of 32 points; the
        // per-class slice counts are summed warp-wide as packed bytes with redux.add.
        // lane c ends up with the row count of class c.
        int mycnt = 0;
        {
          uint32_t rank[PPL];
          uint32_t pk[PPL][4]; // packed per-class counts of slice k: byte (c & 3) of word (c >> 2)
#pragma unroll
          for (int k = 0; k < PPL; ++k) {
            const uint32_t m = __match_any_sync(0xffffffffu, cls[k]);
            rank[k] = __popc(m & lt);
            const bool leader = (m & lt) == 0u;
            const uint32_t contrib = leader ? ((uint32_t)__popc(m) << (8 * (cls[k] & 3))) : 0u;
#pragma unroll
            for (int w = 0; w < 4; ++w)
              pk[k][w] = (w * 4 < ncls) ? __reduce_add_sync(0xffffffffu, ((cls[k] >> 2) == w) ? contrib : 0u) : 0u;
          }
          // lane c: totals and padded exclusive start of class c
          uint32_t tot_w[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            tot_w[w] = 0;
#pragma unroll
            for (int k = 0; k < PPL; ++k) tot_w[w] += pk[k][w]; // <= 128 per byte: no carry
          }
          const int lw = (lane >> 2) & 3, lsh = 8 * (lane & 3);
          mycnt = (lane < ncls) ? (int)((tot_w[lw] >> lsh) & 255u) : 0;
          const int padded = (mycnt + 7) & ~7;
          int incl = padded;
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) {
            const int nn = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nn;
          }
          const int mystart = incl - padded;
          for (int c0 = 0; c0 < ncls; c0 += 4) { // padding -> scratch row: lane = (class - c0) * 8 + i
            const int pc = c0 + (lane >> 3), pi = lane & 7;
            const int cn = __shfl_sync(0xffffffffu, mycnt, pc & 15), cs = __shfl_sync(0xffffffffu, mystart, pc & 15);
            if (pc < ncls && (cn & 7) && pi >= (cn & 7)) list[cs + (cn & ~7) + pi] = (uint8_t)PW;
          }
          // every point: class start + earlier slices of its class + rank inside its slice
          uint32_t before_w[4] = {0, 0, 0, 0};
#pragma unroll
          for (int k = 0; k < PPL; ++k) {
            const int st = __shfl_sync(0xffffffffu, mystart, cls[k]);
            const int cw_ = (cls[k] >> 2) & 3, csh = 8 * (cls[k] & 3);
            const uint32_t bw = cw_ == 0 ? before_w[0] : (cw_ == 1 ? before_w[1] : (cw_ == 2 ? before_w[2] : before_w[3]));
            list[st + (int)((bw >> csh) & 255u) + (int)rank[k]] = (uint8_t)(k * 32 + lane);
#pragma unroll
            for (int w = 0; w < 4; ++w) before_w[w] += pk[k][w];
          }
        }
        __syncwarp();
        PH_MARK(1)
        mbar_wait(smem_u32(&full_bar[s_use]), resident ? 0u : phase);
        PH_MARK(2)
        const uint32_t stage_base = ring_base + s_use * stage_stride + (uint32_t)lane * 8u;
        // ---- classes outermost: B offsets and group ranges are computed once per class; the row
        // indices of the next batch are fetched under the current batch's DMMAs
        int gi = 0;
        int rows_nx[4];
        for (int c = 0; c < ncls; ++c) {
          const int n_c = __shfl_sync(0xffffffffu, mycnt, c);
          const int gend = gi + ((n_c + 7) >> 3);
          if (gi == gend) continue;
          int boff[4] = {0, 0, 0, 0};
          {
            int crem = c;
#pragma unroll
            for (int si = 0; si < 4; ++si) {
              if (si < sites) {
                const int dd = pow2 ? (crem & (int)MASK) : (crem % nsl);
                crem = pow2 ? (crem >> bits) : (crem / nsl);
                boff[si] = (si * nsl + dd) * (CHI * CHI * 8);
              }
            }
          }
#pragma unroll
          for (int b = 0; b < 4; ++b) rows_nx[b] = (int)list[(min(gi + b, gend - 1) << 3) + g];
          while (gi < gend) {
            const int nbat = min(GB, gend - gi);
            int rows[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) rows[b] = rows_nx[b];
            const int gnext = gi + nbat;
            if (gnext < gend) {
#pragma unroll
              for (int b = 0; b < 4; ++b) rows_nx[b] = (int)list[(min(gnext + b, gend - 1) << 3) + g];
            }
            if (nbat >= 4 && GB >= 4) process_batch5<CHI, (GB >= 4 ? 4 : 1)>(state_base, rows, tq, stage_base, boff, sites, PW);
            else if (nbat == 3 && GB >= 3) process_batch5<CHI, (GB >= 3 ? 3 : 1)>(state_base, rows, tq, stage_base, boff, sites, PW);
            else if (nbat == 2 && GB >= 2) process_batch5<CHI, (GB >= 2 ? 2 : 1)>(state_base, rows, tq, stage_base, boff, sites, PW);
            else process_batch5<CHI, 1>(state_base, rows, tq, stage_base, boff, sites, PW);
            gi = gnext;
          }
        }
      
        PH_MARK(3)
      } else {
        mbar_wait(smem_u32(&full_bar[s_use]), resident ? 0u : phase); // keep the ring in lockstep
      }
      __syncwarp();
      if (!resident) {
        if (lane == 0) mbar_arrive(smem_u32(&empty_bar[slot]));
        if (++slot == (uint32_t)n_stage) {
          slot = 0;
          phase ^= 1u;
        }
      }
    }

    // ---- root: out = row . R[d_{n-1}]
    if (live_sub) {
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        const int row = k * 32 + lane;
        const int64_t p = sub * PW + row;
        double o0 = 0.0, o1 = 0.0;
        if (ch.n_vertices > 1) {
          const double* R0 = ch.root + (size_t)(cw[k] & MASK) * CHI;
          const double* R1 = R0 + (size_t)nsl * CHI;
#pragma unroll
          for (int j = 0; j < CPR; ++j) {
            const double2 v = lds128(row_chunk<CHI>(state_base, row, j));
            o0 = fma(v.x, __ldg(R0 + 2 * j), o0);
            o0 = fma(v.y, __ldg(R0 + 2 * j + 1), o0);
            if (ch.nout == 2) {
              o1 = fma(v.x, __ldg(R1 + 2 * j), o1);
              o1 = fma(v.y, __ldg(R1 + 2 * j + 1), o1);
            }
          }
        } else {
          o0 = lds64(row_chunk<CHI>(state_base, row, 0));
          if (ch.nout == 2) o1 = lds64(row_chunk<CHI>(state_base, row, CPR / 2));
        }
        if (p < src.npts) {
          if (out) {
            if (ch.nout == 2) reinterpret_cast<double2*>(out)[p] = make_double2(o0, o1);
            else out[p] = o0;
          }
          accumulate_point(src, p, o0, o1, sum_re, sum_im);
        }
      }
      __syncwarp();
      PH_MARK(4)
    }
  }
  PH_FLUSH

  if (do_sum) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum_re += __shfl_down_sync(0xffffffffu, sum_re, o);
      sum_im += __shfl_down_sync(0xffffffffu, sum_im, o);
    }
    if (lane == 0) {
      red[0][warp] = sum_re;
      red[1][warp] = sum_im;
    }
    named_bar_sync(1, NT);
    if (tid == 0) {
      double xx = 0.0, yy = 0.0;
      for (int w = 0; w < NMW; ++w) {
        xx += red[0][w];
        yy += red[1][w];
      }
      partial[2 * blockIdx.x] = xx;
      partial[2 * blockIdx.x + 1] = yy;
    }
  }
}

// =====================================================================================
// v6: team-sorted, B-stationary kernel for MERGED binary chains (one stream position per round,
// NCLS = 4, 8 or 16 slices per position).  A CTA holds NTEAM independent teams of 4 warps; a team
// owns a tile of 512 points.
//  * every warp is "home" to 128 points of the tile: K1, leaf rows, class counts, root;
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
  __shared__ Digit2 s_d2[kFeMaxSites];
  __shared__ int s_cptr[TTN_MAX_COORDS + 1];

  const int tid = threadIdx.x, lane = tid & 31;
  const int team = tid / (TW * 32), warp = (tid >> 5) % TW;
  const int g = lane >> 2, tq = lane & 3;
  for (int i = tid; i <= dg.n_coords; i += NT) s_cptr[i] = dg.coord_ptr[i];
  for (int i = tid; i < dg.n_sites; i += NT) {
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
    double x[PPL];
#pragma unroll
    for (int k = 0; k < PPL; ++k) w0[k] = w1[k] = 0;
    for (int c = 0; c < dg.n_coords; ++c) {
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        const int64_t p = p0 + k * 32 + lane;
        x[k] = 0.0;
        if (p < src.npts) {
          x[k] = load_coord(src, p, c);
          if (!coord_in_domain(x[k])) {
            atomicOr(err, 1);
            x[k] = 0.0;
          }
        }
      }
      if (ch.run_L[c] > 0 && !src.digits) {
        const int L = ch.run_L[c], plow = ch.run_plow[c];
        const double scale = ch.run_scale[c];
        const bool rev = ch.run_rev[c] != 0;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          unsigned long long q = x[k] >= 1.0 ? ((1ull << L) - 1ull) : (unsigned long long)(x[k] * scale);
          if (rev) q = __brevll(q) >> (64 - L);
          if (plow < 64) {
            w0[k] += q << plow;
            if (plow + L > 64) w1[k] += q >> (64 - plow);
          } else {
            w1[k] += q << (plow - 64);
          }
        }
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

  for (int64_t tile = (int64_t)blockIdx.x * NTEAM + team; tile < n_tiles; tile += tstride) {
    const int64_t p0 = tile * TP + warp * PW; // first home point of this warp
    uint64_t w1[PPL], cw[PPL];
    {
      uint64_t w0[PPL];
      compute_words(tile, w0, w1);
      // ---- leaf rows
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
        cw[k] = w0[k];
        const int row = warp * PW + k * 32 + lane;
        if (deep) {
          const double* L = ch.leaf + (size_t)(cw[k] & LMASK) * CHI;
#pragma unroll
          for (int j = 0; j < CPR; j += 2) {
            double a0, a1, a2, a3;
            ldg256(L + 2 * j, a0, a1, a2, a3);
            sts128(row_chunk<CHI>(state_base, row, j), a0, a1);
            sts128(row_chunk<CHI>(state_base, row, j + 1), a2, a3);
          }
        } else {
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
    named_bar_sync(bar_id, TW * 32); // leaf rows + counts of round 0

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
      named_bar_sync(bar_id, TW * 32); // list r complete; rows of round r-1 written; counts r+1 final

      // ---- owned classes: B in registers, rows streamed through gather -> DMMA -> scatter
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
        const int ng = (n_c + 7) >> 3;
        const uint16_t* Lc = list + st;
        int rows_nx[4] = {0, 0, 0, 0};
        if (ng > 0) {
#pragma unroll
          for (int b = 0; b < GB; ++b) rows_nx[b] = (int)Lc[(min(b, ng - 1) << 3) + g];
        }
        for (int gi = 0; gi < ng; gi += GB) {
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
#pragma unroll
        for (int i = 0; i < NBF; ++i) bcur[i] = bnxt[i];
      }
    }
    if (R > 0) named_bar_sync(bar_id, TW * 32); // rows of the last round complete

    // ---- root: out = row . R[d_{n-1}] for the home rows
#pragma unroll
    for (int k = 0; k < PPL; ++k) {
      const int row = warp * PW + k * 32 + lane;
      const int64_t p = p0 + k * 32 + lane;
      double o0 = 0.0, o1 = 0.0;
      if (deep) {
        const double* R0 = ch.root + (size_t)(cw[k] & RMASK) * CHI;
        const double* R1 = R0 + ((size_t)CHI << ch.root_bits);
#pragma unroll
        for (int j = 0; j < CPR; j += 2) {
          const double2 v = lds128(row_chunk<CHI>(state_base, row, j));
          const double2 w = lds128(row_chunk<CHI>(state_base, row, j + 1));
          double q0, q1, q2, q3;
          ldg256(R0 + 2 * j, q0, q1, q2, q3);
          o0 = fma(v.x, q0, o0);
          o0 = fma(v.y, q1, o0);
          o0 = fma(w.x, q2, o0);
          o0 = fma(w.y, q3, o0);
          if (ch.nout == 2) {
            ldg256(R1 + 2 * j, q0, q1, q2, q3);
            o1 = fma(v.x, q0, o1);
            o1 = fma(v.y, q1, o1);
            o1 = fma(w.x, q2, o1);
            o1 = fma(w.y, q3, o1);
          }
        }
      } else {
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
      }
      if (p < src.npts) {
        if (out) {
          if (ch.nout == 2) reinterpret_cast<double2*>(out)[p] = make_double2(o0, o1);
          else out[p] = o0;
        }
        accumulate_point(src, p, o0, o1, sum_re, sum_im);
      }
    }
    // the next tile's leaf rows overwrite home rows only: no barrier needed here, the first barrier
    // of the next tile orders them before any other warp's gather
  }

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

// ------------------------------------------------------------------------------ host side

static int mma_width(int w) {
  for (int o : {8, 16, 32})
    if (w <= o) return o;
  return 0;
}

// Host image of a chain before upload: position 0 = leaf, steps, last = root; every position has
// the same (padded) number of slices; matrices are CHI x CHI row-major in the real embedding.
struct ChainImage {
  int nsl = 1;
  std::vector<double> leaf;               // [nsl][CHI]
  std::vector<double> root;               // [nout][nsl][CHI]
  std::vector<std::vector<double>> steps; // [t][nsl * CHI * CHI]
};

// Group merging (plan-time pre-contraction): k consecutive positions of a chain whose slice
// indices are bit fields (bits0 bits each) are contracted into one position with 2^(k*bits0)
// slices, slice index = sum_i s_i << (bits0 * i) with i = 0 the position nearest the leaf — i.e.
// exactly the same packed bit stream read k*bits0 bits at a time — so K1 and the run fast path are
// untouched while a point needs 1/k as many matrix-vector products.  Products are accumulated in
// long double and rounded once.  Needs (#positions) % k == 0 and >= 2k positions.
static ChainImage merge_groups(const ChainImage& a, int CHI, int nout, int k, int bits0, int kL, int kR) {
  // positions: leaf group = leaf + steps[0, kL-1), middle groups of k steps, root group =
  // steps[T-(kR-1), T) + root.  kL and kR may be much larger than k ("deep" leaf / root tables of
  // 2^(bits0 kL) vectors, kept in global memory): those are built with vector products only, one
  // member at a time (rounded to double per member, dot products accumulated in long double).
  typedef long double ld;
  const size_t M = (size_t)CHI * CHI;
  const int T = (int)a.steps.size(), S0 = a.nsl, SK = 1 << (bits0 * k);
  ChainImage m;
  m.nsl = SK;
  // ---- leaf table: T_i[s + (b << bits0 i)] = T_{i-1}[s] . E_{i-1}[b]
  {
    std::vector<double> cur((size_t)S0 * CHI);
    for (size_t i = 0; i < cur.size(); ++i) cur[i] = a.leaf[i];
    size_t n_cur = (size_t)1 << bits0;
    for (int i = 1; i < kL; ++i) {
      std::vector<double> nxt(n_cur * ((size_t)1 << bits0) * CHI, 0.0);
      const std::vector<double>& E = a.steps[i - 1];
      for (int bsl = 0; bsl < S0; ++bsl) {
        const double* B = E.data() + (size_t)bsl * M;
        for (size_t sidx = 0; sidx < n_cur; ++sidx) {
          const double* A = cur.data() + sidx * CHI;
          double* C = nxt.data() + (sidx + ((size_t)bsl << (bits0 * i))) * CHI;
          ld acc[32];
          for (int j = 0; j < CHI; ++j) acc[j] = 0.0L;
          for (int kk = 0; kk < CHI; ++kk) {
            const ld av = A[kk];
            if (av == 0.0L) continue;
            for (int j = 0; j < CHI; ++j) acc[j] += av * (ld)B[(size_t)kk * CHI + j];
          }
          for (int j = 0; j < CHI; ++j) C[j] = (double)acc[j];
        }
      }
      cur.swap(nxt);
      n_cur <<= bits0;
    }
    m.leaf.swap(cur);
  }
  // ---- middle groups: matrix products in long double, rounded once
  auto extend = [&](const std::vector<ld>& cur, int n_cur, const std::vector<double>& E, int member) {
    std::vector<ld> nxt((size_t)n_cur * (1 << bits0) * M, 0.0L);
    for (int sidx = 0; sidx < n_cur; ++sidx)
      for (int bsl = 0; bsl < S0; ++bsl) {
        const ld* A = cur.data() + (size_t)sidx * M;
        const double* B = E.data() + (size_t)bsl * M;
        ld* C = nxt.data() + (size_t)(sidx + (bsl << (bits0 * member))) * M;
        for (int i = 0; i < CHI; ++i)
          for (int kk = 0; kk < CHI; ++kk) {
            const ld av = A[(size_t)i * CHI + kk];
            if (av == 0.0L) continue;
            for (int j = 0; j < CHI; ++j) C[(size_t)i * CHI + j] += av * (ld)B[(size_t)kk * CHI + j];
          }
      }
    return nxt;
  };
  const int G = (T - (kL - 1) - (kR - 1)) / k;
  for (int g = 0; g < G; ++g) {
    const int t0 = kL - 1 + g * k;
    std::vector<ld> cur((size_t)(1 << bits0) * M, 0.0L);
    for (int sl = 0; sl < S0; ++sl)
      for (size_t i = 0; i < M; ++i) cur[(size_t)sl * M + i] = a.steps[t0][(size_t)sl * M + i];
    int n_cur = 1 << bits0;
    for (int i = 1; i < k; ++i) {
      cur = extend(cur, n_cur, a.steps[t0 + i], i);
      n_cur <<= bits0;
    }
    std::vector<double> E((size_t)SK * M);
    for (size_t i = 0; i < E.size(); ++i) E[i] = (double)cur[i];
    m.steps.push_back(std::move(E));
  }
  // ---- root table, built from the root backwards (the member nearest the leaf owns the LOW bits):
  // R_j[s + (idx << bits0)][i] = sum_l E_j[s][i][l] R_{j+1}[idx][l]
  {
    const size_t SR = (size_t)1 << (bits0 * kR);
    m.root.assign((size_t)nout * SR * CHI, 0.0);
    for (int o = 0; o < nout; ++o) {
      std::vector<double> cur((size_t)S0 * CHI);
      for (size_t i = 0; i < cur.size(); ++i) cur[i] = a.root[(size_t)o * S0 * CHI + i];
      size_t n_cur = (size_t)1 << bits0;
      for (int j = kR - 2; j >= 0; --j) { // step index T - (kR - 1) + j
        const std::vector<double>& E = a.steps[T - (kR - 1) + j];
        std::vector<double> nxt(n_cur * ((size_t)1 << bits0) * CHI, 0.0);
        for (int sl = 0; sl < S0; ++sl) {
          const double* A = E.data() + (size_t)sl * M;
          for (size_t idx = 0; idx < n_cur; ++idx) {
            const double* Rv = cur.data() + idx * CHI;
            double* C = nxt.data() + ((size_t)sl + (idx << bits0)) * CHI;
            for (int i = 0; i < CHI; ++i) {
              ld acc = 0.0L;
              for (int l = 0; l < CHI; ++l) acc += (ld)A[(size_t)i * CHI + l] * (ld)Rv[l];
              C[i] = (double)acc;
            }
          }
        }
        cur.swap(nxt);
        n_cur <<= bits0;
      }
      std::copy(cur.begin(), cur.end(), m.root.begin() + (size_t)o * SR * CHI);
    }
  }
  return m;
}

static int upload_chain_image(ttn_plan* p, const ChainImage& im, int CHI, ChainMmaDev& c) {
  const int NSL = im.nsl, T = (int)im.steps.size();
  const size_t M = (size_t)CHI * CHI;
  std::vector<double> frags((size_t)std::max(T, 1) * NSL * M, 0.0);
  const int NB = CHI / 8, KB = CHI / 4;
  for (int t = 0; t < T; ++t)
    for (int s = 0; s < NSL; ++s) {
      // B-fragment order: [kb][nb'][lane = 4 g' + t'] = E[8 (kb/2) + 2 t' + (kb%2)][8 nb' + g']
      const double* E = im.steps[t].data() + (size_t)s * M;
      double* F = frags.data() + ((size_t)t * NSL + s) * M;
      for (int kb = 0; kb < KB; ++kb)
        for (int nbp = 0; nbp < NB; ++nbp)
          for (int ln = 0; ln < 32; ++ln) {
            const int gp = ln >> 2, tp = ln & 3;
            F[((size_t)kb * NB + nbp) * 32 + ln] = E[(size_t)(8 * (kb / 2) + 2 * tp + (kb % 2)) * CHI + 8 * nbp + gp];
          }
    }
  double *d_leaf, *d_root, *d_frags;
  TTN_CUDA(cudaMalloc(&d_leaf, im.leaf.size() * 8));
  p->allocs.push_back(d_leaf);
  TTN_CUDA(cudaMalloc(&d_root, im.root.size() * 8));
  p->allocs.push_back(d_root);
  TTN_CUDA(cudaMalloc(&d_frags, frags.size() * 8));
  p->allocs.push_back(d_frags);
  TTN_CUDA(cudaMemcpy(d_leaf, im.leaf.data(), im.leaf.size() * 8, cudaMemcpyHostToDevice));
  TTN_CUDA(cudaMemcpy(d_root, im.root.data(), im.root.size() * 8, cudaMemcpyHostToDevice));
  TTN_CUDA(cudaMemcpy(d_frags, frags.data(), frags.size() * 8, cudaMemcpyHostToDevice));
  c.leaf = d_leaf;
  c.root = d_root;
  c.frags = d_frags;
  return TTN_OK;
}

int build_chain_mma(ttn_plan* p, const ttn_desc* d) {
  p->cmma_ok = false;
  p->cmma_plain_ok = false;
  if (!p->is_chain) return TTN_OK;
  const int n = d->n_vertices;
  const bool cplx = d->is_complex != 0;
  const int NC = cplx ? 2 : 1;
  int maxchi = 1, maxsl = 1;
  for (int v = 0; v < n; ++v) {
    maxchi = std::max(maxchi, d->link_dim[v]);
    maxsl = std::max(maxsl, p->nslices[v]);
  }
  const int CHI = cplx ? mma_width(2 * maxchi) : mma_width(maxchi);
  const int H = CHI / 2; // complex: row = [re(H) | im(H)]
  if (CHI == 0 || maxsl > 4) return TTN_OK;
  const int NSL0 = maxsl; // slices per vertex of the network as given
  const int bits0 = NSL0 <= 1 ? 0 : (NSL0 <= 2 ? 1 : 2); // stream bits per VERTEX
  // group merging (merge_groups): k vertices per stream position when the slice indices are bit
  // fields, the merged position has <= 16 slices and one position's matrices stay <= 32 KB (one
  // ring stage).  TTN_MMA_MERGE caps k (0 or 1: one vertex per position).
  int kmerge = 1;
  {
    // measured on B200 (scripts/merge_probe.py): 4 > 3 > 2 > 1 for binary chains; TTN_MMA_MERGE=k
    // asks for exactly k (falling back to smaller groups when k does not apply)
    std::vector<int> cand = {4, 3, 2};
    if (const char* e = getenv("TTN_MMA_MERGE")) {
      cand.clear();
      for (int kk = std::min(atoi(e), 5); kk >= 2; --kk) cand.push_back(kk);
    }
    // the team-sorted kernel (v6) reads site matrices straight from L2 into registers: no ring-stage
    // limit and up to 32 slices per position; the ring kernels take <= 16 slices and <= 32 KB per position
    const bool v6_on = !(getenv("TTN_MMA_V6") && atoi(getenv("TTN_MMA_V6")) == 0) && p->all_base2;
    if (NSL0 == 2 || NSL0 == 4)
      for (int kk : cand) {
        const int sb = bits0 * kk;
        const bool fits = v6_on ? sb <= (CHI <= 16 ? 5 : 4) : (sb <= 4 && ((size_t)1 << sb) * CHI * CHI * 8 <= 32 * 1024);
        if (fits && n >= 2 * kk && (kk != 3 || CHI <= 16 || v6_on)) {
          kmerge = kk; // 3- and 5-bit fields straddle words: only the kernels with the 128-bit stream read those
          break;
        }
      }
  }
  const bool merge = kmerge > 1;
  // Deep leaf / root groups (team-sorted kernel only): the first kL and the last kR vertices are
  // contracted into tables of 2^(bits0 kL) leaf vectors / 2^(bits0 kR) root vectors in global memory
  // (<= 128 MB each); a point gathers one row of each and runs only the middle groups as rounds.
  int kL = kmerge, kR = kmerge;
  {
    const bool v6_on = !(getenv("TTN_MMA_V6") && atoi(getenv("TTN_MMA_V6")) == 0) && p->all_base2;
    // default: only when the two tables absorb the WHOLE chain with <= 2^16 rows each (L2-resident,
    // built in a fraction of a second): the evaluation is then two row gathers and a dot product.
    // TTN_MMA_DEEP=b sets the budget to 2^b rows of 16 doubles for any chain (0 disables); measured
    // on config 2 with b = 20: 5.7 G points/s device-resident instead of 3.6 G, 2 x 128 MB of
    // tables, 2 s of plan time, the kernel then bound by random 128-byte HBM gathers
    int deep_bits = (n * bits0 <= 32) ? 16 : 0;
    if (const char* e = getenv("TTN_MMA_DEEP")) deep_bits = std::min(atoi(e), 22);
    if (merge && v6_on && deep_bits > 0) {
      const int lb = deep_bits - (CHI >= 32 ? 1 : 0) + (CHI <= 8 ? 1 : 0), rb = lb - (cplx ? 1 : 0);
      kL = std::max(kmerge, std::min(lb / bits0, n / 2));
      kR = std::max(kmerge, std::min(rb / bits0, n - kL));
      const int mid = n - kL - kR;
      const int delta = mid >= 0 ? (kmerge - mid % kmerge) % kmerge : 0;
      if (mid >= 0 && kR - delta >= kmerge) kR -= delta; // whole middle groups without identity padding
      if (n - kL - kR < 0) kL = kR = kmerge;
    }
  }

  // sites per round: as many as keep the class count <= 4 (binary digits: 2 sites -> 4 classes)
  int spr = 1;
  if (!merge) {
    while (true) {
      int cls = 1;
      for (int k = 0; k < spr + 1; ++k) cls *= NSL0;
      if (NSL0 <= 1 || cls > 4 || spr >= 4) break;
      ++spr;
    }
    if (NSL0 <= 1) spr = 2;
    if (const char* e = getenv("TTN_MMA_SPR")) spr = std::max(1, atoi(e));
    int cls = 1;
    for (int k = 0; k < spr; ++k) cls *= std::max(NSL0, 1);
    while (cls > kMaxClasses && spr > 1) {
      --spr;
      cls /= NSL0;
    }
  }
  // the chain is padded with identity sites (slice bits always 0) up to a whole number of rounds
  // (merged: up to an even number of positions), so that every round has exactly `spr` sites
  const int n_steps = n >= 2 ? n - 2 : 0;
  const int n_steps_p = merge ? kL + kR + (n - kL - kR + kmerge - 1) / kmerge * kmerge - 2 : (n_steps + spr - 1) / spr * spr;
  const int root_pos = n >= 2 ? 1 + n_steps_p : 0; // in vertices
  const int n_pos = root_pos + 1;
  const int n_words = bits0 == 0 ? 0 : (n_pos * bits0 + 63) / 64;
  if (n_words > 2) return TTN_OK;

  std::vector<int> order(n), pos_of(n);
  {
    int v = d->root;
    for (int pos = n - 1; pos >= 0; --pos) {
      order[pos] = v;
      if (pos > 0) v = p->child[p->child_ptr[v]];
    }
    for (int pos = 0; pos < n; ++pos) pos_of[order[pos]] = pos;
    if (n >= 2) pos_of[order[n - 1]] = root_pos;
  }
  const double* T = reinterpret_cast<const double*>(d->tensors);
  auto elem = [&](int v, int64_t idx, double* re, double* im) {
    *re = T[(d->tensor_ptr[v] + idx) * NC];
    *im = cplx ? T[(d->tensor_ptr[v] + idx) * NC + 1] : 0.0;
  };
  const int nout = cplx ? 2 : 1;
  const size_t M = (size_t)CHI * CHI;
  ChainImage im;
  im.nsl = NSL0;
  im.leaf.assign((size_t)NSL0 * CHI, 0.0);
  im.root.assign((size_t)nout * NSL0 * CHI, 0.0);
  {
    const int v = order[0], b = d->link_dim[v];
    for (int s = 0; s < p->nslices[v]; ++s)
      for (int j = 0; j < b; ++j) {
        double re, imv;
        elem(v, (int64_t)s * b + j, &re, &imv);
        im.leaf[(size_t)s * CHI + j] = re;
        if (cplx) im.leaf[(size_t)s * CHI + H + j] = imv;
      }
  }
  if (n >= 2) {
    const int v = order[n - 1], a = d->link_dim[order[n - 2]];
    for (int s = 0; s < p->nslices[v]; ++s)
      for (int i = 0; i < a; ++i) {
        double re, imv;
        elem(v, (int64_t)s * a + i, &re, &imv);
        im.root[(size_t)s * CHI + i] = re; // out_re = v_re.R_re - v_im.R_im
        if (cplx) {
          im.root[(size_t)s * CHI + H + i] = -imv;
          im.root[((size_t)NSL0 + s) * CHI + i] = imv; // out_im = v_re.R_im + v_im.R_re
          im.root[((size_t)NSL0 + s) * CHI + H + i] = re;
        }
      }
  }
  double flops_exec = 0.0;
  std::vector<int> dim_in, dim_out;
  dim_out.reserve(n_steps_p + 1);
  for (int t = 0; t < n_steps_p; ++t) {
    const bool dummy = t >= n_steps;
    const int v = dummy ? -1 : order[t + 1];
    const int a = dummy ? 0 : d->link_dim[order[t]], b = dummy ? 0 : d->link_dim[v];
    std::vector<double> Es((size_t)NSL0 * M, 0.0);
    for (int s = 0; s < (dummy ? NSL0 : p->nslices[v]); ++s) {
      double* E = Es.data() + (size_t)s * M;
      if (dummy) {
        for (int i = 0; i < CHI; ++i) E[(size_t)i * CHI + i] = 1.0;
      }
      for (int i = 0; i < a; ++i)
        for (int j = 0; j < b; ++j) {
          double re, imv;
          elem(v, ((int64_t)s * a + i) * b + j, &re, &imv);
          E[(size_t)i * CHI + j] = re;
          if (cplx) {
            E[(size_t)i * CHI + H + j] = imv;
            E[(size_t)(H + i) * CHI + j] = -imv;
            E[(size_t)(H + i) * CHI + H + j] = re;
          }
        }
    }
    im.steps.push_back(std::move(Es));
    flops_exec += (cplx ? 8.0 : 2.0) * a * b;
    dim_in.push_back(dummy ? dim_out.back() : a);
    dim_out.push_back(dummy ? dim_out.back() : b);
  }
  const double root_flops = n >= 2 ? (cplx ? 8.0 : 2.0) * d->link_dim[order[n - 2]] : 0.0;
  p->cmma_flops_exec = flops_exec + root_flops;
  if (merge) {
    // one product per middle group: steps [kL - 1 + g k, kL - 1 + g k + k - 1]
    double fm = 0.0;
    const int G = (n_steps_p + 2 - kL - kR) / kmerge;
    for (int g = 0; g < G; ++g) fm += (cplx ? 8.0 : 2.0) * dim_in[kL - 1 + g * kmerge] * dim_out[kL - 1 + g * kmerge + kmerge - 1];
    p->cmma_flops_exec = fm + (cplx ? 8.0 : 2.0) * dim_in[n_steps_p - (kR - 1)];
  }

  // the unmerged image stays available for the prefix-shared grid kernel (k_grid_share.cu)
  ChainMmaDev& c = p->cmma;
  c = ChainMmaDev{};
  c.n_vertices = n;
  c.chi = CHI;
  c.nout = nout;
  c.root_pos = merge ? 1 + (n_steps_p + 2 - kL - kR) / kmerge : root_pos; // in stream positions (uniform groups only)
  int rc;
  if (NSL0 == 2) {
    ChainMmaDev& q = p->cmma_plain;
    q = ChainMmaDev{};
    q.n_vertices = n;
    q.chi = CHI;
    q.nout = nout;
    q.nsl = NSL0;
    q.n_steps = n_steps_p;
    if (!merge) {
      if ((rc = upload_chain_image(p, im, CHI, q))) return rc;
      p->cmma_plain_ok = true;
    } else {
      // identity padding is not part of the plain image the grid kernel walks
      ChainImage plain = im;
      plain.steps.resize(n_steps);
      q.n_steps = n_steps;
      if ((rc = upload_chain_image(p, plain, CHI, q))) return rc;
      p->cmma_plain_ok = true;
    }
  }
  int NSL = NSL0, bits = bits0;
  if (merge) {
    const ChainImage mg = merge_groups(im, CHI, nout, kmerge, bits0, kL, kR);
    if ((rc = upload_chain_image(p, mg, CHI, c))) return rc;
    NSL = mg.nsl;
    bits = bits0 * kmerge;
    spr = 1;
    c.n_steps = (int)mg.steps.size();
  } else if (p->cmma_plain_ok) {
    c.leaf = p->cmma_plain.leaf;
    c.root = p->cmma_plain.root;
    c.frags = p->cmma_plain.frags;
    c.n_steps = n_steps_p;
  } else {
    if ((rc = upload_chain_image(p, im, CHI, c))) return rc;
    c.n_steps = n_steps_p;
  }
  c.merged = merge ? kmerge : 0;
  c.leaf_bits = merge ? bits0 * kL : bits0; // stream bits the leaf / root group consumes
  c.root_bits = merge ? bits0 * kR : bits0;
  c.nsl = NSL;
  c.bits = bits;
  c.per_word = bits == 0 ? (1 << 30) : 64 / bits;
  c.n_words = n_words;
  c.spr = spr;
  c.n_rounds = c.n_steps / spr;

  // the DMMA kernels get their own copy of the digit table: (word, shift) are the BIT position of
  // the vertex in THEIR packed stream (identity padding shifts the root; a merged position reads
  // the bits of its two vertices as one 2-bit slice index)
  p->digits_mma = p->digits;
  std::vector<DigitEntry> ent(std::max(d->n_sites, 1));
  if (d->n_sites > 0) {
    TTN_CUDA(cudaMemcpy(ent.data(), p->digits.entries, sizeof(DigitEntry) * d->n_sites, cudaMemcpyDeviceToHost));
    for (int i = 0; i < d->n_sites; ++i) {
      const int bitpos = pos_of[ent[i].vertex] * bits0;
      ent[i].word = bits0 ? bitpos / 64 : 0;
      ent[i].shift = bits0 ? bitpos % 64 : 0;
    }
    DigitEntry* d_ent;
    TTN_CUDA(cudaMalloc(&d_ent, sizeof(DigitEntry) * d->n_sites));
    p->allocs.push_back(d_ent);
    TTN_CUDA(cudaMemcpy(d_ent, ent.data(), sizeof(DigitEntry) * d->n_sites, cudaMemcpyHostToDevice));
    p->digits_mma.entries = d_ent;
  }
  // K1 "run" fast path per coordinate slot.  Conditions (all checked here, bitwise):
  //   every site index of the slot is binary, its vertex carries no other site index (1 bit per
  //   vertex), the digit numbers are exactly 1..L with thresholds exactly 2^-k, L <= 63, and
  //   the digits sit on CONSECUTIVE stream bits in increasing or decreasing order.
  // Then the greedy loop (abstractindexmap.jl:121-138) yields digit k = bit (L-k) of floor(x * 2^L):
  // x >= 2^-k ? subtract : keep is exact in binary floating point (the subtraction clears the leading
  // bit), the scaling by 2^L is exact, and x >= 1 saturates to all ones exactly as the loop does.
  for (int cidx = 0; cidx < TTN_MAX_COORDS; ++cidx) c.run_L[cidx] = 0;
  if (bits0 == 1) {
    std::vector<int32_t> cptr(d->n_coords + 1);
    TTN_CUDA(cudaMemcpy(cptr.data(), p->digits_mma.coord_ptr, sizeof(int32_t) * (d->n_coords + 1), cudaMemcpyDeviceToHost));
    for (int cidx = 0; cidx < d->n_coords; ++cidx) {
      const int L = cptr[cidx + 1] - cptr[cidx];
      if (L < 1 || L > 63) continue;
      bool ok = true;
      int step = 0, first_pos = -1;
      for (int k = 0; k < L && ok; ++k) {
        const DigitEntry& e = ent[cptr[cidx] + k];
        const int pos = e.word * 64 + e.shift;
        ok = ok && e.base == 2 && e.stride == 1 && p->nslices[e.vertex] == 2;
        ok = ok && d->site_digit[e.site] == k + 1 && d->thr[e.thr_off + 1] == std::ldexp(1.0, -(k + 1));
        if (k == 0) first_pos = pos;
        else if (k == 1) step = pos - first_pos;
        if (k >= 1) ok = ok && (pos - first_pos == step * k);
      }
      if (L == 1) step = 1;
      if (!ok || (step != 1 && step != -1)) continue;
      c.run_L[cidx] = L;
      c.run_rev[cidx] = step == 1 ? 1 : 0;
      c.run_plow[cidx] = step == 1 ? first_pos : first_pos - (L - 1);
      c.run_scale[cidx] = std::ldexp(1.0, L);
    }
  }
  // both kernels keep the digit and threshold tables in static shared memory
  p->cmma_ok = p->digits.n_sites <= kFeMaxSites && d->n_sites <= kFeMaxSites && d->thr_ptr[d->n_sites] <= kFeMaxThr;
  return TTN_OK;
}

template <int CHI, int P, int NMW, int GB, bool B2, int NSLT, int SPRT>
static int launch_mma3_inst(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial,
                            int* n_partial, cudaStream_t s) {
  const ChainMmaDev& c = p->cmma;
  constexpr int LIST_CAP = P + 8 * kMaxClasses;
  const size_t fixed = (size_t)(P + 8) * CHI * 8 + (size_t)2 * P * 16 + (size_t)2 * LIST_CAP * 2 + 128;
  const size_t smem_max = 227 * 1024 - 12 * 1024; // static shared: barriers, digit tables, meta
  const size_t stage = (size_t)c.spr * c.nsl * CHI * CHI * 8;
  if (fixed + stage > smem_max) {
    set_error("chain DMMA kernel: one round of site matrices does not fit in shared memory");
    return TTN_ERR_UNSUPPORTED;
  }
  int n_stage = (int)std::min<size_t>((smem_max - fixed) / stage, (size_t)kMmaMaxStages);
  int resident = 0;
  if (c.n_rounds <= n_stage) {
    n_stage = std::max(c.n_rounds, 1);
    resident = 1;
  }
  const size_t smem = fixed + (size_t)n_stage * stage;
  auto kern = chain_mma3_kernel<CHI, P, NMW, GB, B2, NSLT, SPRT>;
  TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  const int64_t n_tiles = (src.npts + P - 1) / P;
  const int grid = (int)std::min<int64_t>(n_tiles, p->sm_count);
  const int do_sum = d_partial != nullptr;
  kern<<<grid, NMW * 32 + 128, smem, s>>>(c, p->digits_mma, src, d_out, p->d_err, d_partial, do_sum, n_stage, resident,
                                          (uint32_t)stage);
  TTN_CUDA(cudaGetLastError());
  *n_partial = do_sum ? grid : 0;
  return TTN_OK;
}

template <int CHI, int NMW, int GB, bool B2, int NSLT, int SPRT>
static int launch_mma5_inst(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial,
                            int* n_partial, cudaStream_t s) {
  const ChainMmaDev& c = p->cmma;
  constexpr int PW = 128, ROWS = PW + 8;
  const size_t fixed = (size_t)NMW * ROWS * CHI * 8 + (size_t)NMW * 256 + 128;
  const size_t smem_max = 227 * 1024 - 12 * 1024; // static shared: barriers, digit tables
  const size_t stage = (size_t)c.spr * c.nsl * CHI * CHI * 8;
  if (fixed + stage > smem_max) {
    set_error("chain DMMA kernel: one round of site matrices does not fit in shared memory");
    return TTN_ERR_UNSUPPORTED;
  }
  int n_stage = (int)std::min<size_t>((smem_max - fixed) / stage, (size_t)kMmaMaxStages);
  int resident = 0;
  if (c.n_rounds <= n_stage) {
    n_stage = std::max(c.n_rounds, 1);
    resident = 1;
  }
  const size_t smem = fixed + (size_t)n_stage * stage;
  auto kern = chain_mma5_kernel<CHI, NMW, GB, B2, NSLT, SPRT>;
  TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  const int64_t n_sub = (src.npts + PW - 1) / PW;
  const int grid = (int)std::min<int64_t>((n_sub + NMW - 1) / NMW, p->sm_count);
  const int do_sum = d_partial != nullptr;
  kern<<<grid, NMW * 32 + 32, smem, s>>>(c, p->digits_mma, src, d_out, p->d_err, d_partial, do_sum, n_stage, resident,
                                         (uint32_t)stage);
  TTN_CUDA(cudaGetLastError());
  *n_partial = do_sum ? grid : 0;
  return TTN_OK;
}

template <int CHI, int NCLS, int NTEAM, int PW_>
static int launch_mma6_inst(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                            cudaStream_t s) {
  using T6 = Team6<CHI, NCLS, PW_>;
  const ChainMmaDev& c = p->cmma;
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

int launch_chain_mma(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double* d_partial,
                     int* n_partial, cudaStream_t s) {
  (void)st;
  *n_partial = 0;
  if (src.npts == 0) return TTN_OK;
  if (!p->cmma_ok) {
    set_error("DMMA chain kernel requested but the network is not a supported chain");
    return TTN_ERR_UNSUPPORTED;
  }
  const ChainMmaDev& c = p->cmma;
  // fast instances: binary digits with 2 sites per round, or two binary digits per vertex (4 slices)
  // with 1 site per round; everything else takes the runtime-generic instance
  const bool f22 = p->all_base2 && c.nsl == 2 && c.spr == 2;
  const bool f41 = p->all_base2 && c.nsl == 4 && c.spr == 1;
  const bool f161 = p->all_base2 && c.nsl == 16 && c.spr == 1; // 4 merged binary vertices per position
  const bool f81 = p->all_base2 && c.nsl == 8 && c.spr == 1;   // 3 merged binary vertices per position
  // the warp-autonomous kernel's 2-sites-per-round instance has no partial last round
  const bool f22w = f22 && c.n_steps % 2 == 0;
  if (p->digits.n_sites > kFeMaxSites || p->info.n_sites > kFeMaxSites || p->fe_thr_len > kFeMaxThr) {
    set_error("DMMA chain kernel: digit tables exceed the kernel's shared-memory copies (160 sites / 640 thresholds)");
    return TTN_ERR_UNSUPPORTED;
  }
  // merged binary chains of width <= 16: team-sorted, B-stationary kernel (v6), 3 teams per CTA
  // (TTN_MMA_V6=0 falls back to the warp-autonomous kernel, =2 runs two teams: experiments)
  static const int v6 = getenv("TTN_MMA_V6") ? atoi(getenv("TTN_MMA_V6")) : 3;
  if (v6 && c.merged && p->all_base2 && c.spr == 1) {
#define TTN_V6_CASE(W, N)                                                                                   \
  if (c.chi == W && c.nsl == N)                                                                             \
    return v6 == 2 ? launch_mma6_inst<W, N, 2, 128>(p, src, d_out, d_partial, n_partial, s)                 \
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
  switch (c.chi) {
    case 8: // warp-autonomous kernel (v5)
      if (f22w) return launch_mma5_inst<8, 8, 4, true, 2, 2>(p, src, d_out, d_partial, n_partial, s);
      if (f41) return launch_mma5_inst<8, 8, 4, true, 4, 1>(p, src, d_out, d_partial, n_partial, s);
      if (f81) return launch_mma5_inst<8, 8, 4, true, 8, 1>(p, src, d_out, d_partial, n_partial, s);
      if (f161) return launch_mma5_inst<8, 8, 4, true, 16, 1>(p, src, d_out, d_partial, n_partial, s);
      return launch_mma5_inst<8, 8, 4, false, 0, 0>(p, src, d_out, d_partial, n_partial, s);
    case 16:
      if (f22w) return launch_mma5_inst<16, 8, 4, true, 2, 2>(p, src, d_out, d_partial, n_partial, s);
      if (f41) return launch_mma5_inst<16, 8, 4, true, 4, 1>(p, src, d_out, d_partial, n_partial, s);
      if (f81) return launch_mma5_inst<16, 8, 4, true, 8, 1>(p, src, d_out, d_partial, n_partial, s);
      if (f161) return launch_mma5_inst<16, 8, 4, true, 16, 1>(p, src, d_out, d_partial, n_partial, s);
      return launch_mma5_inst<16, 8, 4, false, 0, 0>(p, src, d_out, d_partial, n_partial, s);
    case 32: // CTA-sorted, warp-specialised kernel (v3): 128 rows per warp would not fit at this width
      if (f22) return launch_mma3_inst<32, 512, 8, 3, true, 2, 2>(p, src, d_out, d_partial, n_partial, s);
      if (f41) return launch_mma3_inst<32, 512, 8, 3, true, 4, 1>(p, src, d_out, d_partial, n_partial, s);
      return launch_mma3_inst<32, 512, 8, 2, false, 0, 0>(p, src, d_out, d_partial, n_partial, s);
  }
  set_error("DMMA chain kernel: unsupported width");
  return TTN_ERR_UNSUPPORTED;
}

#ifdef TTN_PHASE_CLOCKS
int debug_phase_clocks(unsigned long long* out8, int reset) {
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  TTN_CUDA(cudaMemcpyFromSymbol(out8, g_phase, sizeof(z)));
  if (reset) TTN_CUDA(cudaMemcpyToSymbol(g_phase, z, sizeof(z)));
  return TTN_OK;
}
#endif

} // namespace ttn
