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

#include "k_chain_common.cuh"

namespace ttn {

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
static ChainImage merge_groups(const ChainImage& a, int CHI, int nout, int k, int kL, int kR, int kLh, int kRh) {
  // positions: leaf group = leaf + steps[0, kL-1), middle groups of k steps, root group =
  // steps[T-(kR-1), T) + root.  kL and kR may be much larger than k ("deep" leaf / root tables of
  // 2^(bits0 kL) vectors, kept in global memory): those are built with vector products only, one
  // member at a time (rounded to double per member, dot products accumulated in long double).
  typedef long double ld;
  const size_t M = (size_t)CHI * CHI;
  // Slice indices are mixed-radix numbers in radix S0 = a.nsl (member i of a group has weight S0^i): for S0 = 2 / 4
  // that is the bit-field form, for S0 = 3 (base-3 digits, one site index per vertex) the fields are dense — 3^k
  // classes per position in a ceil(log2 3^k)-bit field, 3^12 rows in a 2^20-row table budget.
  const int T = (int)a.steps.size(), S0 = a.nsl;
  int SK = 1, pk = 1;
  for (int i = 0; i < k; ++i) pk *= S0;
  while (SK < pk) SK <<= 1; // classes per merged position, padded to a power of two (rows >= S0^k stay zero)
  ChainImage m;
  m.nsl = SK;
  // ---- leaf table: T_i[s + (b << bits0 i)] = T_{i-1}[s] . E_{i-1}[b]
  {
    std::vector<double> cur((size_t)S0 * CHI);
    for (size_t i = 0; i < cur.size(); ++i) cur[i] = a.leaf[i];
    size_t n_cur = (size_t)S0;
    for (int i = 1; i < kLh; ++i) { // levels kLh .. kL - 1 are added on the device (extend_deep_tables)
      std::vector<double> nxt(n_cur * (size_t)S0 * CHI, 0.0);
      const std::vector<double>& E = a.steps[i - 1];
      for (int bsl = 0; bsl < S0; ++bsl) {
        const double* B = E.data() + (size_t)bsl * M;
        for (size_t sidx = 0; sidx < n_cur; ++sidx) {
          const double* A = cur.data() + sidx * CHI;
          double* C = nxt.data() + (sidx + (size_t)bsl * n_cur) * CHI;
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
      n_cur *= S0;
    }
    if (cur.size() < (size_t)SK * CHI) cur.resize((size_t)SK * CHI, 0.0); // the kernels stage SK rows of a uniform group
    m.leaf.swap(cur);
  }
  // ---- middle groups: matrix products in long double, rounded once
  auto extend = [&](const std::vector<ld>& cur, int n_cur, const std::vector<double>& E) {
    std::vector<ld> nxt((size_t)n_cur * S0 * M, 0.0L);
    for (int sidx = 0; sidx < n_cur; ++sidx)
      for (int bsl = 0; bsl < S0; ++bsl) {
        const ld* A = cur.data() + (size_t)sidx * M;
        const double* B = E.data() + (size_t)bsl * M;
        ld* C = nxt.data() + (size_t)(sidx + bsl * n_cur) * M;
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
    std::vector<ld> cur((size_t)S0 * M, 0.0L);
    for (int sl = 0; sl < S0; ++sl)
      for (size_t i = 0; i < M; ++i) cur[(size_t)sl * M + i] = a.steps[t0][(size_t)sl * M + i];
    int n_cur = S0;
    for (int i = 1; i < k; ++i) {
      cur = extend(cur, n_cur, a.steps[t0 + i]);
      n_cur *= S0;
    }
    std::vector<double> E((size_t)SK * M, 0.0);
    for (size_t i = 0; i < cur.size(); ++i) E[i] = (double)cur[i];
    m.steps.push_back(std::move(E));
  }
  // ---- root table, built from the root backwards (the member nearest the leaf owns the LOW bits):
  // R_j[s + (idx << bits0)][i] = sum_l E_j[s][i][l] R_{j+1}[idx][l]
  {
    size_t SRr = 1; // members kR - kRh - 1 .. 0 are added on the device
    for (int i = 0; i < kRh; ++i) SRr *= S0;
    size_t SRp = 1;
    while (SRp < SRr) SRp <<= 1;
    // table o starts at row o * SR with SR a power of two: the kernels address the second (imaginary-part) table of a
    // complex network as root + (CHI << root_bits); a uniform group is staged with SK rows
    const size_t SR = std::max(SRp, (size_t)SK);
    m.root.assign((size_t)nout * SR * CHI, 0.0);
    for (int o = 0; o < nout; ++o) {
      std::vector<double> cur((size_t)S0 * CHI);
      for (size_t i = 0; i < cur.size(); ++i) cur[i] = a.root[(size_t)o * S0 * CHI + i];
      size_t n_cur = (size_t)S0;
      for (int j = kR - 2; j >= kR - kRh; --j) { // step index T - (kR - 1) + j
        const std::vector<double>& E = a.steps[T - (kR - 1) + j];
        std::vector<double> nxt(n_cur * (size_t)S0 * CHI, 0.0);
        for (int sl = 0; sl < S0; ++sl) {
          const double* A = E.data() + (size_t)sl * M;
          for (size_t idx = 0; idx < n_cur; ++idx) {
            const double* Rv = cur.data() + idx * CHI;
            double* C = nxt.data() + ((size_t)sl + idx * S0) * CHI;
            for (int i = 0; i < CHI; ++i) {
              ld acc = 0.0L;
              for (int l = 0; l < CHI; ++l) acc += (ld)A[(size_t)i * CHI + l] * (ld)Rv[l];
              C[i] = (double)acc;
            }
          }
        }
        cur.swap(nxt);
        n_cur *= S0;
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

// ---- deep leaf / root tables, upper levels on the device.  The host (merge_groups) builds the tables of the
// first kLh / last kRh vertices; every further member doubles (x 2^bits0) the table with one vector-matrix
// product per row.  Dot products are accumulated in double-double (error-free two_prod via FMA + two_sum) and
// rounded once per member — the same "long accumulation, one rounding per member" the host loop does in long
// double, at a few ms for 2^20 rows instead of seconds.
struct dd2 {
  double hi, lo;
};
__device__ __forceinline__ void dd2_mac(dd2& acc, double a, double b) {
  const double pr = __dmul_rn(a, b);
  const double e = __fma_rn(a, b, -pr);
  const double t = __dadd_rn(acc.hi, pr);
  const double bb = __dsub_rn(t, acc.hi);
  const double err = __dadd_rn(__dsub_rn(acc.hi, __dsub_rn(t, bb)), __dsub_rn(pr, bb));
  acc.lo = __dadd_rn(acc.lo, __dadd_rn(err, e));
  acc.hi = t;
}
// nxt[s + b * n_cur][j] = sum_k cur[s][k] E[b][k][j]
__global__ void deep_leaf_extend_kernel(const double* __restrict__ cur, size_t n_cur, const double* __restrict__ E, int S0,
                                        int CHI, double* __restrict__ nxt) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cur * S0 * CHI) return;
  const int j = (int)(idx % CHI), b = (int)((idx / CHI) % S0);
  const size_t sidx = idx / ((size_t)CHI * S0);
  const double* A = cur + sidx * CHI;
  const double* B = E + (size_t)b * CHI * CHI + j;
  dd2 acc{0.0, 0.0};
  for (int k = 0; k < CHI; ++k) dd2_mac(acc, A[k], __ldg(B + (size_t)k * CHI));
  nxt[(sidx + (size_t)b * n_cur) * CHI + j] = __dadd_rn(acc.hi, acc.lo);
}
// nxt[sl + idx * S0][i] = sum_l E[sl][i][l] cur[idx][l]
__global__ void deep_root_extend_kernel(const double* __restrict__ cur, size_t n_cur, const double* __restrict__ E, int S0,
                                        int CHI, double* __restrict__ nxt) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cur * S0 * CHI) return;
  const int i = (int)(idx % CHI), sl = (int)((idx / CHI) % S0);
  const size_t ridx = idx / ((size_t)CHI * S0);
  const double* A = E + ((size_t)sl * CHI + i) * CHI;
  const double* Rv = cur + ridx * CHI;
  dd2 acc{0.0, 0.0};
  for (int l = 0; l < CHI; ++l) dd2_mac(acc, __ldg(A + l), Rv[l]);
  nxt[((size_t)sl + ridx * (size_t)S0) * CHI + i] = __dadd_rn(acc.hi, acc.lo);
}

static int extend_deep_tables(ttn_plan* p, const ChainImage& a, int CHI, int nout, int kLh, int kL, int kRh,
                              int kR, int SK, ChainMmaDev& c) {
  auto ipow = [](size_t b, int e) {
    size_t r = 1;
    for (int i = 0; i < e; ++i) r *= b;
    return r;
  };
  const int S0 = a.nsl, T = (int)a.steps.size();
  const size_t M = (size_t)CHI * CHI;
  double* d_E = nullptr;
  TTN_CUDA(cudaMalloc(&d_E, (size_t)S0 * M * 8));
  int rc = TTN_OK;
  auto run = [&](bool leaf, const double* start, size_t n_start, int lv0, int lv1, double** result) -> int {
    // levels lv0 .. lv1 - 1; ping-pong between two buffers of the final size, the last level lands in `fin`
    const size_t rows_fin = n_start * ipow((size_t)S0, lv1 - lv0);
    double *fin = nullptr, *tmp = nullptr;
    TTN_CUDA(cudaMalloc(&fin, rows_fin * CHI * 8));
    p->allocs.push_back(fin);
    if (cudaMalloc(&tmp, std::max<size_t>(rows_fin / S0, 1) * CHI * 8) != cudaSuccess) {
      set_error("deep tables: out of device memory");
      return TTN_ERR_NOMEM;
    }
    const int nlev = lv1 - lv0;
    const double* cur = start;
    size_t n_cur = n_start;
    for (int q = 0; q < nlev; ++q) {
      double* dst = ((nlev - 1 - q) % 2 == 0) ? fin : tmp;
      // leaf level i uses step i - 1; root member j (descending) uses step T - (kR - 1) + j
      const int step = leaf ? (lv0 + q) - 1 : T - (kR - 1) + (kR - 1 - (lv0 + q));
      cudaMemcpy(d_E, a.steps[step].data(), (size_t)S0 * M * 8, cudaMemcpyHostToDevice);
      const size_t total = n_cur * S0 * CHI;
      const unsigned grid = (unsigned)((total + 255) / 256);
      if (leaf) deep_leaf_extend_kernel<<<grid, 256>>>(cur, n_cur, d_E, S0, CHI, dst);
      else deep_root_extend_kernel<<<grid, 256>>>(cur, n_cur, d_E, S0, CHI, dst);
      cudaDeviceSynchronize(); // d_E is reused by the next level
      cur = dst;
      n_cur *= S0;
    }
    cudaFree(tmp);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error(std::string("deep tables: ") + cudaGetErrorString(e));
      return TTN_ERR_CUDA;
    }
    *result = fin;
    return TTN_OK;
  };
  if (kL > kLh) {
    double* fin = nullptr;
    rc = run(true, c.leaf, ipow((size_t)S0, kLh), kLh, kL, &fin);
    if (rc == TTN_OK) c.leaf = fin;
  }
  if (rc == TTN_OK && kR > kRh) {
    // nout tables back to back: [o][2^(bits0 kR)][CHI]
    const size_t rows_h = ipow((size_t)S0, kRh), rows_f = ipow((size_t)S0, kR);
    auto pow2_ceil = [](size_t x) {
      size_t r = 1;
      while (r < x) r <<= 1;
      return r;
    };
    const size_t stride_h = std::max(pow2_ceil(rows_h), (size_t)SK), stride_f = pow2_ceil(rows_f); // rows between two tables
    double* all = nullptr;
    if (nout == 1) {
      rc = run(false, c.root, rows_h, kRh, kR, &all);
    } else {
      if (cudaMalloc(&all, (size_t)nout * stride_f * CHI * 8) != cudaSuccess) {
        set_error("deep tables: out of device memory");
        rc = TTN_ERR_NOMEM;
      } else {
        p->allocs.push_back(all);
        for (int o = 0; o < nout && rc == TTN_OK; ++o) {
          double* one = nullptr;
          rc = run(false, c.root + (size_t)o * stride_h * CHI, rows_h, kRh, kR, &one); // host stride: merge_groups' SR
          if (rc == TTN_OK) cudaMemcpy(all + (size_t)o * stride_f * CHI, one, rows_f * CHI * 8, cudaMemcpyDeviceToDevice);
          if (rc == TTN_OK) { // `one` was pushed to p->allocs by run(): release it now
            p->allocs.pop_back();
            cudaFree(one);
          }
        }
      }
    }
    if (rc == TTN_OK) c.root = all;
  }
  cudaFree(d_E);
  return rc;
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
  p->v6_teams = getenv("TTN_MMA_V6") ? atoi(getenv("TTN_MMA_V6")) : 3;
  // Base-3 chains (every site-carrying vertex has ONE site index of dimension 3): RADIX mode — the slice indices of a
  // group of vertices are packed as one radix-3 number (3^3 = 27 classes in a 5-bit field per round, 3^12 rows in a
  // 2^20-row table) instead of one 2-bit field per vertex (16 classes per round of which 9 are populated, 4^10 rows
  // of which 3^10 are reachable): a 60-site chain runs 12 rounds per point instead of 20.  Needs the team-sorted
  // kernel (the only one that reads odd field widths from the 128-bit stream) and a chain that splits into whole
  // groups; anything else keeps the 2-bit fields.  TTN_MMA_RADIX=0 switches it off.
  bool radix = maxsl == 3 && p->v6_teams != 0 && !(getenv("TTN_MMA_RADIX") && atoi(getenv("TTN_MMA_RADIX")) == 0) && !getenv("TTN_MMA_MERGE");
  for (int v = 0; v < n && radix; ++v) radix = p->nslices[v] == 3 || p->nslices[v] == 1;
  int rk = 0, rkL = 0, rkR = 0; // radix: vertices per round, in the leaf table, in the root table
  if (radix) {
    auto max_k = [](int bitsb) { // largest k with 3^k <= 2^bitsb
      int k = 0;
      double pw = 1.0;
      while (pw * 3.0 <= std::ldexp(1.0, bitsb)) pw *= 3.0, ++k;
      return k;
    };
    rk = CHI <= 16 ? 3 : 2;
    int deep_bits = 20;
    if (const char* e = getenv("TTN_MMA_DEEP")) deep_bits = std::min(atoi(e), 22);
    const int lb = deep_bits - (CHI >= 32 ? 1 : 0) + (CHI <= 8 ? 1 : 0), rb = lb - (cplx ? 1 : 0);
    rkL = deep_bits > 0 ? std::max(rk, std::min(max_k(lb), n / 2)) : rk;
    rkR = deep_bits > 0 ? std::max(rk, std::min(max_k(rb), n - rkL)) : rk;
    const int mid = n - rkL - rkR;
    const int delta = mid >= 0 ? (rk - mid % rk) % rk : 0;
    if (mid >= 0 && rkR - delta >= rk) rkR -= delta;
    radix = n >= 2 * rk && n - rkL - rkR >= 0 && (n - rkL - rkR) % rk == 0;
  }
  const int NSL0 = radix ? 3 : (maxsl == 3 ? 4 : maxsl); // slices per vertex; without radix mode a base-3 index takes a
                                                         // 2-bit field (slice 3: zero matrix, never selected)
  const int bits0 = NSL0 <= 1 ? 0 : (NSL0 <= 2 ? 1 : 2); // stream bits per VERTEX (bit-field mode)
  auto ceil_log2 = [](double x) {
    int b = 0;
    while (std::ldexp(1.0, b) < x) ++b;
    return b;
  };
  // stream bits of a field holding k vertices
  auto fbits = [&](int k) { return radix ? ceil_log2(std::pow(3.0, k)) : bits0 * k; };
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
    const bool v6_on = p->v6_teams != 0;
    if (radix) kmerge = rk;
    else if (NSL0 == 2 || NSL0 == 4)
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
    const bool v6_on = p->v6_teams != 0;
    // default budget: 2^16 rows of 16 doubles when the two tables then absorb the WHOLE chain (L2-resident:
    // the evaluation is two row gathers and a dot product), else 2^20 rows (2 x 128 MB).  Measured on config 2
    // (scripts/deep_bits_sweep.py, 1e8 points, G points/s device-resident): no deep tables 3.58, 2^12 4.37,
    // 2^16 5.20, 2^18 5.55, 2^20 5.98 — fewer DMMA rounds per point (13 -> 5) against two random 128-byte row
    // gathers.  The upper table levels are built on the device (extend_deep_tables: 2^20 rows in milliseconds;
    // the host loop took 2 s).  TTN_MMA_DEEP=b overrides (0 disables).
    int deep_bits = (n * bits0 <= 32) ? 16 : 20;
    if (const char* e = getenv("TTN_MMA_DEEP")) deep_bits = std::min(atoi(e), 22);
    if (radix) {
      kL = rkL;
      kR = rkR;
    } else if (merge && v6_on && deep_bits > 0) {
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
  const int grp_bits = radix ? slice_bits(32 >> (kmerge == 2 ? 1 : 0)) : bits0 * kmerge; // radix: 27 -> 5 bits, 9 -> 4 bits
  const int n_groups_mid = merge ? (n_steps_p + 2 - kL - kR) / kmerge : 0;
  const int n_words = radix ? (fbits(kL) + n_groups_mid * grp_bits + fbits(kR) + 63) / 64 : (bits0 == 0 ? 0 : (n_pos * bits0 + 63) / 64);
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
    // host levels: up to 2^10 rows per table; the rest of a deep table is built on the device
    const int khost = radix ? 6 : 10 / bits0; // 3^6 = 729 rows
    const int kLh = std::min(kL, std::max(kmerge, khost)), kRh = std::min(kR, std::max(kmerge, khost));
    const ChainImage mg = merge_groups(im, CHI, nout, kmerge, kL, kR, kLh, kRh);
    if ((rc = upload_chain_image(p, mg, CHI, c))) return rc;
    if ((kL > kLh || kR > kRh) && (rc = extend_deep_tables(p, im, CHI, nout, kLh, kL, kRh, kR, mg.nsl, c))) return rc;
    NSL = mg.nsl;
    bits = grp_bits;
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
  c.k1_generic = k1_generic_mode(d);
  c.leaf_bits = merge ? fbits(kL) : bits0; // stream bits the leaf / root group consumes
  c.root_bits = merge ? fbits(kR) : bits0;
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
    p->cmma_site_fbits.assign(d->n_sites, std::max(bits0, 1));
    for (int i = 0; i < d->n_sites; ++i) {
      int bitpos = pos_of[ent[i].vertex] * bits0;
      if (radix) {
        // field of the vertex's group and its weight inside it: leaf table | middle groups | root table
        const int q = pos_of[ent[i].vertex]; // chain position (radix mode has no identity padding)
        int field, idx, fb;
        if (q < kL) field = 0, idx = q, fb = fbits(kL);
        else if (q >= n - kR) field = fbits(kL) + n_groups_mid * grp_bits, idx = q - (n - kR), fb = fbits(kR);
        else field = fbits(kL) + ((q - kL) / kmerge) * grp_bits, idx = (q - kL) % kmerge, fb = grp_bits;
        int mult = 1;
        for (int t = 0; t < idx; ++t) mult *= 3;
        ent[i].stride *= mult;
        bitpos = field;
        p->cmma_site_fbits[ent[i].site] = fb;
      }
      ent[i].word = (bits0 || radix) ? bitpos / 64 : 0;
      ent[i].shift = (bits0 || radix) ? bitpos % 64 : 0;
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
  // Base-4 digits (one site index of dimension 4 per vertex, thresholds exactly v 4^-k): digit k of the greedy loop is the
  // bit PAIR (2(L-k)+1, 2(L-k)) of floor(x 4^L) — the same argument as for base 2 with two bits per step (x_rn stays an
  // exact multiple of 4^-L's grid, the compare picks floor(x_rn 4^k)) — so a coordinate whose digits sit on consecutive
  // 2-bit fields takes the run path as a run of 2L bits; run_rev = 2: digit 1 on the LOWEST field (pairs reversed).
  if (bits0 == 2 && !radix) {
    std::vector<int32_t> cptr(d->n_coords + 1);
    TTN_CUDA(cudaMemcpy(cptr.data(), p->digits_mma.coord_ptr, sizeof(int32_t) * (d->n_coords + 1), cudaMemcpyDeviceToHost));
    for (int cidx = 0; cidx < d->n_coords; ++cidx) {
      const int L = cptr[cidx + 1] - cptr[cidx];
      if (L < 1 || 2 * L > 62) continue;
      bool ok = true;
      int step = 0, first_pos = -1;
      for (int k = 0; k < L && ok; ++k) {
        const DigitEntry& e = ent[cptr[cidx] + k];
        const int pos = e.word * 64 + e.shift;
        ok = ok && e.base == 4 && e.stride == 1 && p->nslices[e.vertex] == 4 && d->site_digit[e.site] == k + 1;
        for (int v = 1; v < 4 && ok; ++v) ok = d->thr[e.thr_off + v] == v * std::ldexp(1.0, -2 * (k + 1));
        if (k == 0) first_pos = pos;
        else if (k == 1) step = pos - first_pos;
        if (k >= 1) ok = ok && (pos - first_pos == step * k);
      }
      if (L == 1) step = 2;
      if (!ok || (step != 2 && step != -2)) continue;
      c.run_L[cidx] = 2 * L;
      c.run_rev[cidx] = step == 2 ? 2 : 0;
      c.run_plow[cidx] = step == 2 ? first_pos : first_pos - 2 * (L - 1);
      c.run_scale[cidx] = std::ldexp(1.0, 2 * L);
    }
  }
  // Light variant for host-buffer calls: measured on config 2 (bench.py e2e, pinned buffers, 2 Mi-point chunks) the
  // deep-table kernel — two random 128-byte row gathers per point at > 5 G points/s — slows the H2D / D2H copies
  // that run beside it (e2e 2.79 G points/s for ANY table size, L2-resident ones included, against 3.19 = the
  // copy ceiling with uniform groups): such calls are PCIe-bound, so they take the image without deep tables.
  // Same packed stream (built only when both images pad the chain identically), same digits, same kernel instance.
  p->cmma_light_ok = false;
  const bool light_on = !(getenv("TTN_MMA_LIGHT") && atoi(getenv("TTN_MMA_LIGHT")) == 0); // 0: host-buffer calls run the deep image too
  if (light_on && merge && !radix && (kL > kmerge || kR > kmerge)) {
    const int n_steps_u = kmerge + kmerge + (n - 2 * kmerge + kmerge - 1) / kmerge * kmerge - 2;
    if (n_steps_u == n_steps_p) {
      ChainMmaDev& q = p->cmma_light;
      q = c; // layout, run fast path, k1 mode are shared; (leaf, root, frags, rounds) replaced below
      const ChainImage mu = merge_groups(im, CHI, nout, kmerge, kmerge, kmerge, kmerge, kmerge);
      if ((rc = upload_chain_image(p, mu, CHI, q))) return rc;
      q.n_steps = (int)mu.steps.size();
      q.n_rounds = q.n_steps;
      q.leaf_bits = q.root_bits = bits0 * kmerge;
      double fm = 0.0;
      for (int gI = 0; gI < q.n_steps; ++gI) fm += (cplx ? 8.0 : 2.0) * dim_in[kmerge - 1 + gI * kmerge] * dim_out[kmerge - 1 + gI * kmerge + kmerge - 1];
      p->cmma_light_flops = fm + (cplx ? 8.0 : 2.0) * dim_in[n_steps_p - (kmerge - 1)];
      // ... where the light image keeps up with the copies: its rate falls with rounds x width^2 (config 2: 13 x 16^2 ->
      // 3.6 G points/s against a 3.2 G copy ceiling; the config-4 shape: 5 x 32^2 -> 2.1 G, and host-buffer calls of that
      // shape ran at 2.0 G end to end where the deep image — two row gathers from L2-resident tables — is PCIe-bound)
      p->cmma_light_ok = (double)q.n_rounds * CHI * CHI <= 3600.0;
    }
  }
  // both kernels keep the digit and threshold tables in static shared memory
  p->cmma_ok = p->digits.n_sites <= kFeMaxSites && d->n_sites <= kFeMaxSites && d->thr_ptr[d->n_sites] <= kFeMaxThr;
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
  if (p->digits.n_sites > kFeMaxSites || p->info.n_sites > kFeMaxSites || p->fe_thr_len > kFeMaxThr) {
    set_error("DMMA chain kernel: digit tables exceed the kernel's shared-memory copies (160 sites / 640 thresholds)");
    return TTN_ERR_UNSUPPORTED;
  }
  if (chain_team_applicable(p)) return launch_chain_team(p, src, d_out, d_partial, n_partial, s);
  return launch_chain_ring(p, src, d_out, d_partial, n_partial, s);
}

} // namespace ttn
