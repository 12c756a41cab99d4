// K2+K3 — chain (MPS-shaped) networks: one thread per point, per-point state in registers,
// site matrices streamed through a shared-memory ring by bulk async copies (TMA, UBLKCP) that
// signal mbarriers; a dedicated producer warp feeds NT consumer threads.
//
// The chain is rooted at one end (the packer does that), so the whole contraction
// (project + scalar, src/itensornetworkfunction.jl:84-106) is
//     v   = L[d_0]                      leaf slice, a row vector of length chi
//     v  <- v * M_t[d_t]   t = 1..n-2   digit-selected chi x chi slice of site t
//     out = v . R[d_{n-1}]              root slice, a column vector
// Every point picks its own slice d_t per site, so there is no GEMM across points at small chi:
// each thread keeps v in registers and reads M_t[d_t][i][j] from shared memory.  The stage image
// interleaves the slices of one site at 16-byte granularity,
//     real   : chunk q = (i*(CHI/2) + j/2)*NSL + d  holds (M[d][i][j], M[d][i][j+1]),  j even
//     complex: chunk q = (i*CHI + j)*NSL + d        holds (re, im) of M[d][i][j]
// so the lanes of a warp — which differ only in d — read NSL adjacent 16-byte chunks with one
// LDS.128: at most NSL*16 contiguous bytes, no bank conflict, 2 (real) or 4 (complex) DFMAs per
// load.  Bond dimensions are zero-padded to CHI and slice counts to NSL (exact: padded entries
// are never selected / contribute +0.0).
#include <algorithm>
#include <cstring>

#include "k_async.cuh"
#include "k_digits.cuh"

namespace ttn {

constexpr int kMaxStages = 64;

template <int CHI, int NSL, bool CPLX, int NT>
__global__ void __launch_bounds__(NT + 32, 1)
    chain_kernel(ChainDev ch, DigitTable dg, CoordSource src, double* __restrict__ out, int* err,
                 double* __restrict__ partial, int do_sum, int n_stage, int resident) {
  constexpr int NC = CPLX ? 2 : 1;
  constexpr uint32_t STAGE = (uint32_t)CHI * CHI * NSL * 8 * NC;
  constexpr int NWARP = NT / 32;
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ double red[2][NWARP];

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < n_stage; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), NWARP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int64_t n_tiles = (src.npts + NT - 1) / NT;
  const int n_steps = ch.n_steps;
  const uint32_t ring_base = smem_u32(ring);

  if (tid >= NT) {
    // ===== producer warp: one elected lane issues the bulk copies =====
    if (tid == NT && n_steps > 0) {
      const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(ch.steps);
      if (resident) {
        for (int t = 0; t < n_steps; ++t) {
          mbar_expect_tx(smem_u32(&full_bar[t]), STAGE);
          bulk_g2s(ring_base + (uint32_t)t * STAGE, gsrc + (size_t)t * STAGE, STAGE, smem_u32(&full_bar[t]));
        }
      } else {
        uint32_t slot = 0, phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
          for (int t = 0; t < n_steps; ++t) {
            mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1u);
            mbar_expect_tx(smem_u32(&full_bar[slot]), STAGE);
            bulk_g2s(ring_base + slot * STAGE, gsrc + (size_t)t * STAGE, STAGE, smem_u32(&full_bar[slot]));
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

  // ===== consumers: one point per thread per tile =====
  const int lane = tid & 31;
  constexpr uint64_t MASK = (NSL <= 1) ? 0ull : (NSL <= 2 ? 1ull : 3ull);
  const int bits = ch.bits, per_word = ch.per_word;
  double sum_re = 0.0, sum_im = 0.0;
  uint32_t slot = 0, phase = 0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p = tile * NT + tid;
    const bool live = p < src.npts;
    // ---- K1: digits -> packed slice stream (position 0 = leaf ... n-1 = root)
    uint64_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    if (live) {
      for (int c = 0; c < dg.n_coords; ++c) {
        double x = load_coord(src, p, c);
        if (!coord_in_domain(x)) {
          atomicOr(err, 1);
          x = 0.0;
        }
        for (int k = dg.coord_ptr[c]; k < dg.coord_ptr[c + 1]; ++k) {
          const DigitEntry e = dg.entries[k];
          const int v = src.digits ? given_digit(src, p, dg.n_sites, e.site, e.base, err) : greedy_digit(x, dg.thr + e.thr_off, e.base);
          const uint64_t b = (uint64_t)(v * e.stride) << e.shift;
          w0 += (e.word == 0) ? b : 0ull;
          w1 += (e.word == 1) ? b : 0ull;
          w2 += (e.word == 2) ? b : 0ull;
          w3 += (e.word == 3) ? b : 0ull;
        }
      }
    }
    uint64_t cw = w0;
    int in_word = 0, widx = 0;
    auto next_slice = [&]() -> uint32_t {
      const uint32_t d = (uint32_t)(cw & MASK);
      cw >>= bits;
      if (++in_word == per_word) {
        in_word = 0;
        ++widx;
        cw = (widx == 1) ? w1 : (widx == 2 ? w2 : w3);
      }
      return d;
    };

    // ---- leaf
    double v[CHI * NC];
    {
      const uint32_t d = next_slice();
      const double* L = ch.leaf + (size_t)d * CHI * NC;
#pragma unroll
      for (int j = 0; j < CHI * NC; ++j) v[j] = __ldg(L + j);
    }
    // ---- middle sites through the ring
    for (int t = 0; t < n_steps; ++t) {
      const uint32_t d = next_slice();
      const uint32_t s_use = resident ? (uint32_t)t : slot;
      mbar_wait(smem_u32(&full_bar[s_use]), resident ? 0u : phase);
      const uint32_t sb = ring_base + s_use * STAGE + d * 16u;
      double acc[CHI * NC];
#pragma unroll
      for (int j = 0; j < CHI * NC; ++j) acc[j] = 0.0;
      if (!CPLX) {
#pragma unroll
        for (int i = 0; i < CHI; ++i) {
#pragma unroll
          for (int jp = 0; jp < CHI / 2; ++jp) {
            const double2 m = lds128(sb + (uint32_t)((i * (CHI / 2) + jp) * NSL) * 16u);
            acc[2 * jp] = fma(v[i], m.x, acc[2 * jp]);
            acc[2 * jp + 1] = fma(v[i], m.y, acc[2 * jp + 1]);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < CHI; ++i) {
#pragma unroll
          for (int j = 0; j < CHI; ++j) {
            const double2 m = lds128(sb + (uint32_t)((i * CHI + j) * NSL) * 16u);
            acc[2 * j] = fma(v[2 * i], m.x, acc[2 * j]);
            acc[2 * j] = fma(-v[2 * i + 1], m.y, acc[2 * j]);
            acc[2 * j + 1] = fma(v[2 * i], m.y, acc[2 * j + 1]);
            acc[2 * j + 1] = fma(v[2 * i + 1], m.x, acc[2 * j + 1]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CHI * NC; ++j) v[j] = acc[j];
      if (!resident) {
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty_bar[slot]));
        if (++slot == (uint32_t)n_stage) {
          slot = 0;
          phase ^= 1u;
        }
      }
    }
    // ---- root
    double vr, vi = 0.0;
    if (ch.n_vertices > 1) {
      const uint32_t d = next_slice();
      const double* R = ch.root + (size_t)d * CHI * NC;
      double ar = 0.0, ai = 0.0;
#pragma unroll
      for (int i = 0; i < CHI; ++i) {
        if (!CPLX) {
          ar = fma(v[i], __ldg(R + i), ar);
        } else {
          const double rr = __ldg(R + 2 * i), ri = __ldg(R + 2 * i + 1);
          ar = fma(v[2 * i], rr, ar);
          ar = fma(-v[2 * i + 1], ri, ar);
          ai = fma(v[2 * i], ri, ai);
          ai = fma(v[2 * i + 1], rr, ai);
        }
      }
      vr = ar;
      vi = ai;
    } else {
      vr = v[0];
      if (CPLX) vi = v[1];
    }
    if (live) {
      if (out) {
        if (CPLX) {
          reinterpret_cast<double2*>(out)[p] = make_double2(vr, vi);
        } else {
          out[p] = vr;
        }
      }
      accumulate_point(src, p, vr, vi, sum_re, sum_im);
    }
  }

  if (do_sum) {
    // deterministic: fixed-order shuffle tree per warp, then warp 0 adds the NWARP partials in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum_re += __shfl_down_sync(0xffffffffu, sum_re, o);
      sum_im += __shfl_down_sync(0xffffffffu, sum_im, o);
    }
    if (lane == 0) {
      red[0][tid >> 5] = sum_re;
      red[1][tid >> 5] = sum_im;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(NT) : "memory");
    if (tid == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < NWARP; ++w) {
        a += red[0][w];
        b += red[1][w];
      }
      partial[2 * blockIdx.x] = a;
      partial[2 * blockIdx.x + 1] = b;
    }
  }
}

// ------------------------------------------------------------------------------ host side

static int pad_chi(int chi, bool cplx) {
  const int opts_r[] = {2, 4, 8, 16, 32};
  const int opts_c[] = {2, 4, 8, 16};
  if (cplx) {
    for (int o : opts_c)
      if (chi <= o) return o;
  } else {
    for (int o : opts_r)
      if (chi <= o) return o;
  }
  return 0;
}

bool chain_supported(int chi, int nsl, bool cplx) { return pad_chi(chi, cplx) != 0 && nsl >= 1 && nsl <= 4; }

// Pack the chain into the device images the kernel streams.  `d` uses the layout documented in
// include/ttneval.h: tensor of v = [slice][child][parent].
int build_chain(ttn_plan* p, const ttn_desc* d) {
  p->chain_ok = false;
  if (!p->is_chain) return TTN_OK;
  const int n = d->n_vertices;
  const bool cplx = d->is_complex != 0;
  const int NC = cplx ? 2 : 1;
  int maxchi = 1, maxsl = 1;
  for (int v = 0; v < n; ++v) {
    maxchi = std::max(maxchi, d->link_dim[v]);
    maxsl = std::max(maxsl, p->nslices[v]);
  }
  const int CHI = pad_chi(maxchi, cplx);
  if (CHI == 0 || maxsl > 4) return TTN_OK; // not supported by this kernel; other kernels take over
  const int NSL = maxsl;
  const int bits = NSL <= 1 ? 0 : (NSL <= 2 ? 1 : 2);
  const int per_word = bits == 0 ? (1 << 30) : 64 / bits;
  const int n_words = bits == 0 ? 0 : (n + per_word - 1) / per_word;
  if (n_words > 4) return TTN_OK;

  // positions: 0 = leaf ... n-1 = root
  std::vector<int> order(n);
  {
    int v = d->root;
    for (int pos = n - 1; pos >= 0; --pos) {
      order[pos] = v;
      if (pos > 0) v = p->child[p->child_ptr[v]];
    }
  }
  p->chain_order = order;
  std::vector<int> pos_of(n);
  for (int pos = 0; pos < n; ++pos) pos_of[order[pos]] = pos;

  const double* T = reinterpret_cast<const double*>(d->tensors);
  const int n_steps = n >= 2 ? n - 2 : 0;
  const size_t stage_elems = (size_t)CHI * CHI * NSL * NC;
  std::vector<double> leaf((size_t)NSL * CHI * NC, 0.0), root((size_t)NSL * CHI * NC, 0.0);
  std::vector<double> steps(stage_elems * (size_t)std::max(n_steps, 1), 0.0);
  {
    const int v = order[0];
    const int b = d->link_dim[v];
    for (int s = 0; s < p->nslices[v]; ++s)
      for (int j = 0; j < b; ++j)
        for (int k = 0; k < NC; ++k)
          leaf[((size_t)s * CHI + j) * NC + k] = T[(d->tensor_ptr[v] + (int64_t)s * b + j) * NC + k];
  }
  if (n >= 2) {
    const int v = order[n - 1];
    const int a = d->link_dim[order[n - 2]];
    for (int s = 0; s < p->nslices[v]; ++s)
      for (int i = 0; i < a; ++i)
        for (int k = 0; k < NC; ++k)
          root[((size_t)s * CHI + i) * NC + k] = T[(d->tensor_ptr[v] + (int64_t)s * a + i) * NC + k];
  }
  for (int t = 0; t < n_steps; ++t) {
    const int v = order[t + 1];
    const int a = d->link_dim[order[t]];
    const int b = d->link_dim[v];
    double* img = steps.data() + stage_elems * t;
    for (int s = 0; s < p->nslices[v]; ++s)
      for (int i = 0; i < a; ++i)
        for (int j = 0; j < b; ++j) {
          const int64_t src = (d->tensor_ptr[v] + ((int64_t)s * a + i) * b + j) * NC;
          if (!cplx) {
            img[(((size_t)i * (CHI / 2) + j / 2) * NSL + s) * 2 + (j & 1)] = T[src];
          } else {
            img[(((size_t)i * CHI + j) * NSL + s) * 2 + 0] = T[src];
            img[(((size_t)i * CHI + j) * NSL + s) * 2 + 1] = T[src + 1];
          }
        }
  }
  double *d_leaf, *d_root, *d_steps;
  TTN_CUDA(cudaMalloc(&d_leaf, leaf.size() * 8));
  p->allocs.push_back(d_leaf);
  TTN_CUDA(cudaMalloc(&d_root, root.size() * 8));
  p->allocs.push_back(d_root);
  TTN_CUDA(cudaMalloc(&d_steps, steps.size() * 8));
  p->allocs.push_back(d_steps);
  TTN_CUDA(cudaMemcpy(d_leaf, leaf.data(), leaf.size() * 8, cudaMemcpyHostToDevice));
  TTN_CUDA(cudaMemcpy(d_root, root.data(), root.size() * 8, cudaMemcpyHostToDevice));
  TTN_CUDA(cudaMemcpy(d_steps, steps.data(), steps.size() * 8, cudaMemcpyHostToDevice));

  ChainDev& c = p->chain;
  c.n_steps = n_steps;
  c.n_vertices = n;
  c.chi = CHI;
  c.nsl = NSL;
  c.bits = bits;
  c.per_word = per_word;
  c.n_words = n_words;
  c.leaf = d_leaf;
  c.root = d_root;
  c.steps = d_steps;
  c.stage_bytes = (int64_t)stage_elems * 8;

  // (word, shift) of every digit entry: the vertex's position in the packed slice stream
  std::vector<DigitEntry> ent(d->n_sites);
  if (d->n_sites > 0) {
    TTN_CUDA(cudaMemcpy(ent.data(), p->digits.entries, sizeof(DigitEntry) * d->n_sites, cudaMemcpyDeviceToHost));
    for (auto& e : ent) {
      const int pos = pos_of[e.vertex];
      e.word = bits ? pos / per_word : 0;
      e.shift = bits ? (pos % per_word) * bits : 0;
    }
    TTN_CUDA(cudaMemcpy(const_cast<DigitEntry*>(p->digits.entries), ent.data(), sizeof(DigitEntry) * d->n_sites,
                        cudaMemcpyHostToDevice));
  }
  p->chain_ok = true;
  return TTN_OK;
}

template <int CHI, int NSL, bool CPLX, int NT>
static int launch_inst(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                       cudaStream_t s) {
  constexpr size_t STAGE = (size_t)CHI * CHI * NSL * 8 * (CPLX ? 2 : 1);
  const size_t smem_budget = 200 * 1024;
  int n_stage = (int)std::min<size_t>(smem_budget / STAGE, (size_t)kMaxStages);
  const int n_steps = p->chain.n_steps;
  int resident = 0;
  if (n_steps <= n_stage) {
    n_stage = std::max(n_steps, 1);
    resident = 1;
  }
  const size_t smem = (size_t)n_stage * STAGE;
  auto kern = chain_kernel<CHI, NSL, CPLX, NT>;
  TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_budget)));
  const int64_t n_tiles = (src.npts + NT - 1) / NT;
  const int grid = (int)std::min<int64_t>(n_tiles, p->sm_count);
  const int do_sum = d_partial != nullptr;
  kern<<<grid, NT + 32, smem, s>>>(p->chain, p->digits, src, d_out, p->d_err, d_partial, do_sum, n_stage, resident);
  TTN_CUDA(cudaGetLastError());
  *n_partial = do_sum ? grid : 0;
  return TTN_OK;
}

template <int CHI, bool CPLX, int NT>
static int launch_nsl(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                      cudaStream_t s) {
  switch (p->chain.nsl) {
    case 1: return launch_inst<CHI, 1, CPLX, NT>(p, src, d_out, d_partial, n_partial, s);
    case 2: return launch_inst<CHI, 2, CPLX, NT>(p, src, d_out, d_partial, n_partial, s);
    case 3: return launch_inst<CHI, 3, CPLX, NT>(p, src, d_out, d_partial, n_partial, s);
    case 4: return launch_inst<CHI, 4, CPLX, NT>(p, src, d_out, d_partial, n_partial, s);
  }
  set_error("chain kernel: unsupported slice count");
  return TTN_ERR_UNSUPPORTED;
}

int launch_chain(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double* d_partial,
                 int* n_partial, cudaStream_t s) {
  (void)st;
  *n_partial = 0;
  if (src.npts == 0) return TTN_OK;
  if (!p->chain_ok) {
    set_error("chain kernel requested but the network is not a supported chain");
    return TTN_ERR_UNSUPPORTED;
  }
  const bool cplx = p->info.is_complex != 0;
  const int chi = p->chain.chi;
  if (!cplx) {
    switch (chi) {
      case 2: return launch_nsl<2, false, 512>(p, src, d_out, d_partial, n_partial, s);
      case 4: return launch_nsl<4, false, 512>(p, src, d_out, d_partial, n_partial, s);
      case 8: return launch_nsl<8, false, 512>(p, src, d_out, d_partial, n_partial, s);
      case 16: return launch_nsl<16, false, 512>(p, src, d_out, d_partial, n_partial, s);
      case 32: return launch_nsl<32, false, 256>(p, src, d_out, d_partial, n_partial, s);
    }
  } else {
    switch (chi) {
      case 2: return launch_nsl<2, true, 512>(p, src, d_out, d_partial, n_partial, s);
      case 4: return launch_nsl<4, true, 512>(p, src, d_out, d_partial, n_partial, s);
      case 8: return launch_nsl<8, true, 512>(p, src, d_out, d_partial, n_partial, s);
      case 16: return launch_nsl<16, true, 256>(p, src, d_out, d_partial, n_partial, s);
    }
  }
  set_error("chain kernel: unsupported bond dimension");
  return TTN_ERR_UNSUPPORTED;
}

} // namespace ttn
