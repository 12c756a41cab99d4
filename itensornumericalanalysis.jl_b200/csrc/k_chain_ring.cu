// K4, ring kernels — chains that are NOT merged (base 3, ...) or when the team-sorted kernel is switched
// off: the warp-autonomous kernel (width <= 16) and the CTA-sorted, warp-specialised kernel (width 32).
// Site matrices (B fragments) reach shared memory through a ring of TMA bulk copies signalled by
// mbarriers; see k_chain_mma.cu for the common scheme and the host-side images.
#include <algorithm>
#include <cstring>

#include "k_chain_common.cuh"

namespace ttn {

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

// =====================================================================================
// v3: warp-specialised version of the scheme above.  The MMA warps do nothing but
// gather -> DMMA -> scatter; everything else runs concurrently on front-end warps:
//   * digit warps   compute the packed slice streams of the NEXT tile (K1),
//   * the list warp counting-sorts the points of the CURRENT tile by class, one round ahead
//                   (double-buffered lists),
//   * B fragments are prefetched by MMA thread 0 into the ring slot the end-of-round barrier has
//     just freed (no producer warp, no "empty" barriers).
// Handshakes are mbarriers: tile_ready/tile_free (slice streams), list_full/list_empty, ring full.
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

// Dispatch to the compile-time instances.  Fast instances: binary digits with 2 sites per round, or slice
// counts of 4 / 8 / 16 with 1 site per round; everything else takes the runtime-generic instance.
int launch_chain_ring(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                      cudaStream_t s) {
  const ChainMmaDev& c = p->cmma;
  const bool f22 = p->all_base2 && c.nsl == 2 && c.spr == 2;
  const bool f41 = p->all_base2 && c.nsl == 4 && c.spr == 1;
  const bool f161 = p->all_base2 && c.nsl == 16 && c.spr == 1; // 4 merged binary vertices per position
  const bool f81 = p->all_base2 && c.nsl == 8 && c.spr == 1;   // 3 merged binary vertices per position
  // the warp-autonomous kernel's 2-sites-per-round instance has no partial last round
  const bool f22w = f22 && c.n_steps % 2 == 0;
  if (c.nsl > kMaxClasses) {
    set_error("DMMA ring kernels: more than 16 slices per position (merged for the team-sorted kernel)");
    return TTN_ERR_UNSUPPORTED;
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
