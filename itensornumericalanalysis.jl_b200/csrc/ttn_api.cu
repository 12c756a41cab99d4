// C ABI of libttneval.so (include/ttneval.h): plan construction, chunked/pipelined evaluation,
// error reporting.  Everything here is host code around the kernels in k_*.cu.
#include <immintrin.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <new>
#include <numeric>
#include <thread>

#include "ttn_internal.h"

namespace ttn {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

template <typename T>
static int upload(ttn_plan* p, const std::vector<T>& h, const T** out) {
  void* d = nullptr;
  const size_t bytes = std::max<size_t>(h.size() * sizeof(T), 16);
  TTN_CUDA(cudaMalloc(&d, bytes));
  p->allocs.push_back(d);
  if (!h.empty()) TTN_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const T*>(d);
  return TTN_OK;
}

static int fail(int code, const std::string& msg) {
  set_error(msg);
  return code;
}

// ---- vertices with more than two children ------------------------------------------------------------------
// The per-vertex GEMM tree kernel contracts at most two children per vertex (a Khatri-Rao GEMM).  A vertex with
// children c_1 < ... < c_k (k > 2; uniform_tree, test/test_realitensorfunction.jl:129) is BINARISED at plan time:
// its last two children are hung under a new site-less vertex u whose tensor is the identity
// T_u[a][b][(a, b)] = 1 — u's message is the Kronecker product of theirs, dimension chi_a chi_b — and v keeps u as
// its last child; v's own tensor [site..][c_1]..[c_{k-1}][c_k][parent] is untouched, the merged axis (c_{k-1}, c_k)
// is already contiguous in it.  Repeated until two children remain.  Done only when every new link fits the tree
// kernel (<= 64 real / 32 complex) and every vertex has <= 8 slices; otherwise the network is left as it is (generic
// kernel).  The function evaluated is the same; flops_per_point stays the rule of the ORIGINAL network.
struct BinDesc {
  ttn_desc d{};
  std::vector<int32_t> parent, link_dim, site_ptr;
  std::vector<int64_t> tensor_ptr;
  std::vector<double> tensors;
};
static bool binarize_desc(const ttn_desc* o, BinDesc* b) {
  const int n = o->n_vertices;
  if (getenv("TTN_TREE_BINARIZE") && atoi(getenv("TTN_TREE_BINARIZE")) == 0) return false;
  if (n <= 0 || o->root < 0 || o->root >= n || !o->parent || !o->link_dim || !o->site_ptr || !o->tensor_ptr || !o->tensors) return false;
  const int NC = o->is_complex ? 2 : 1, WL = 64 / NC;
  std::vector<std::vector<int>> ch((size_t)n);
  int maxk = 0;
  for (int v = 0; v < n; ++v) {
    const int q = o->parent[v];
    if (v == o->root) continue;
    if (q < 0 || q >= n || q == v || o->link_dim[v] < 1) return false;
    ch[q].push_back(v);
    maxk = std::max(maxk, (int)ch[q].size());
  }
  if (maxk <= 2) return false;
  for (int v = 0; v < n; ++v) {
    int64_t ns = 1;
    for (int si = o->site_ptr[v]; si < o->site_ptr[v + 1]; ++si) ns *= std::max(o->site_dim[si], 1);
    if (ns > 8 || o->link_dim[v] > WL) return false;
  }
  b->parent.assign(o->parent, o->parent + n);
  b->link_dim.assign(o->link_dim, o->link_dim + n);
  b->site_ptr.assign(o->site_ptr, o->site_ptr + n + 1);
  b->tensor_ptr.assign(o->tensor_ptr, o->tensor_ptr + n + 1);
  const double* T = reinterpret_cast<const double*>(o->tensors);
  b->tensors.assign(T, T + o->tensor_ptr[n] * NC);
  for (int v = 0; v < n; ++v) {
    std::vector<int> cur = ch[v];
    while (cur.size() > 2) {
      const int c2 = cur.back(), c1 = cur[cur.size() - 2];
      const int64_t dim = (int64_t)b->link_dim[c1] * b->link_dim[c2];
      if (dim > WL) return false;
      const int u = (int)b->parent.size();
      b->parent.push_back(v);
      b->link_dim.push_back((int32_t)dim);
      b->parent[c1] = b->parent[c2] = u;
      b->site_ptr.push_back(b->site_ptr.back());
      const size_t off = b->tensors.size();
      b->tensors.resize(off + (size_t)dim * dim * NC, 0.0);
      for (int64_t i = 0; i < dim; ++i) b->tensors[off + (size_t)(i * dim + i) * NC] = 1.0;
      b->tensor_ptr.push_back(b->tensor_ptr.back() + dim * dim);
      cur.pop_back();
      cur.pop_back();
      cur.push_back(u);
    }
  }
  b->d = *o;
  b->d.n_vertices = (int32_t)b->parent.size();
  b->d.parent = b->parent.data();
  b->d.link_dim = b->link_dim.data();
  b->d.site_ptr = b->site_ptr.data();
  b->d.tensor_ptr = b->tensor_ptr.data();
  b->d.tensors = b->tensors.data();
  return true;
}

static int build_plan_impl(ttn_plan* p, const ttn_desc* d);

static int build_plan(ttn_plan* p, const ttn_desc* o) {
  BinDesc bin;
  if (!o->parent || o->n_vertices <= 0 || !binarize_desc(o, &bin)) return build_plan_impl(p, o);
  const int rc = build_plan_impl(p, &bin.d);
  if (rc != TTN_OK) return rc;
  // report the network the caller described: vertex count, link dimensions and the flop rule of the ORIGINAL tree
  const int n = o->n_vertices;
  std::vector<std::vector<int>> ch((size_t)n);
  for (int v = 0; v < n; ++v)
    if (v != o->root) ch[o->parent[v]].push_back(v);
  double macs = 0.0;
  int max_link = 1;
  for (int v = 0; v < n; ++v) {
    max_link = std::max(max_link, o->link_dim[v]);
    double rest = o->link_dim[v];
    for (int c : ch[v]) rest *= o->link_dim[c];
    for (int c : ch[v]) {
      macs += rest; // sum_j c_j ... c_k p
      rest /= o->link_dim[c];
    }
  }
  p->info.n_vertices = n;
  p->info.max_link_dim = max_link;
  p->info.flops_per_point = (o->is_complex ? 8.0 : 2.0) * macs;
  p->info.tensor_bytes = o->tensor_ptr[n] * (o->is_complex ? 2 : 1) * 8;
  p->binarized = true;
  return TTN_OK;
}

static int build_plan_impl(ttn_plan* p, const ttn_desc* d) {
  const int n = d->n_vertices;
  if (d->abi_version != TTN_ABI_VERSION) return fail(TTN_ERR_INVALID, "ttn_desc.abi_version mismatch");
  if (n <= 0) return fail(TTN_ERR_INVALID, "n_vertices must be positive");
  if (d->n_coords < 0 || d->n_coords > TTN_MAX_COORDS)
    return fail(TTN_ERR_UNSUPPORTED, "n_coords must be in [0, 16]");
  if (d->root < 0 || d->root >= n || d->parent[d->root] != -1)
    return fail(TTN_ERR_INVALID, "root must have parent -1");
  if (d->site_ptr[0] != 0 || d->site_ptr[n] != d->n_sites || d->thr_ptr[0] != 0)
    return fail(TTN_ERR_INVALID, "site_ptr / thr_ptr are inconsistent");
  p->parent.assign(d->parent, d->parent + n);
  p->link_dim.assign(d->link_dim, d->link_dim + n);
  p->child_ptr.assign(n + 1, 0);
  for (int v = 0; v < n; ++v) {
    if (v == d->root) continue;
    const int q = d->parent[v];
    if (q < 0 || q >= n || q == v) return fail(TTN_ERR_INVALID, "parent[] out of range");
    p->child_ptr[q + 1]++;
  }
  for (int v = 0; v < n; ++v) p->child_ptr[v + 1] += p->child_ptr[v];
  p->child.assign(std::max(n - 1, 1), 0);
  {
    std::vector<int> fill(n, 0);
    for (int v = 0; v < n; ++v) {
      const int q = d->parent[v];
      if (q >= 0) p->child[p->child_ptr[q] + fill[q]++] = v;
    }
  }
  // post order (children before parents); detects cycles / disconnected input
  p->post.clear();
  {
    std::vector<int> stack{d->root}, it(n, 0);
    while (!stack.empty()) {
      const int v = stack.back();
      const int k = p->child_ptr[v] + it[v];
      if (k < p->child_ptr[v + 1]) {
        it[v]++;
        stack.push_back(p->child[k]);
      } else {
        p->post.push_back(v);
        stack.pop_back();
      }
    }
  }
  if ((int)p->post.size() != n) return fail(TTN_ERR_INVALID, "parent[] does not describe a tree rooted at root");
  if (d->link_dim[d->root] != 1) return fail(TTN_ERR_INVALID, "link_dim[root] must be 1");

  p->slice_size.assign(n, 0);
  p->tensor_off.assign(d->tensor_ptr, d->tensor_ptr + n);
  p->msg_off.assign(n, 0);
  p->nslices.assign(n, 1);
  int64_t msg_total = 0, max_inter = 1;
  int max_link = 1;
  double macs = 0;
  p->is_chain = true;
  for (int v = 0; v < n; ++v) {
    if (d->link_dim[v] < 1) return fail(TTN_ERR_INVALID, "link_dim must be >= 1");
    max_link = std::max(max_link, d->link_dim[v]);
    int64_t s = d->link_dim[v];
    const int nchild = p->child_ptr[v + 1] - p->child_ptr[v];
    if (nchild > 1) p->is_chain = false;
    for (int ci = p->child_ptr[v]; ci < p->child_ptr[v + 1]; ++ci) s *= d->link_dim[p->child[ci]];
    p->slice_size[v] = s;
    int64_t rest = s;
    for (int ci = p->child_ptr[v]; ci < p->child_ptr[v + 1]; ++ci) {
      macs += (double)rest; // SURVEY §8(d) flop rule
      rest /= d->link_dim[p->child[ci]];
      if (ci != p->child_ptr[v + 1] - 1) max_inter = std::max(max_inter, rest);
    }
    int64_t ns = 1;
    for (int si = d->site_ptr[v]; si < d->site_ptr[v + 1]; ++si) {
      if (d->site_dim[si] < 1 || d->site_dim[si] > 255) return fail(TTN_ERR_UNSUPPORTED, "site dimension must be in [1, 255]");
      ns *= d->site_dim[si];
      if (ns > (1 << 20)) return fail(TTN_ERR_UNSUPPORTED, "too many slices on one vertex");
    }
    p->nslices[v] = (int)ns;
    if (d->tensor_ptr[v + 1] - d->tensor_ptr[v] != ns * s)
      return fail(TTN_ERR_INVALID, "tensor_ptr does not match prod(site dims) * prod(link dims) at vertex " + std::to_string(v));
    p->msg_off[v] = msg_total;
    msg_total += d->link_dim[v];
  }

  // ---- digit table: entries grouped by coordinate slot, ascending digit number (stable)
  std::vector<DigitEntry> entries;
  std::vector<int32_t> coord_ptr(d->n_coords + 1, 0);
  {
    std::vector<int> site_vertex(d->n_sites), site_stride(d->n_sites);
    for (int v = 0; v < n; ++v) {
      int stride = 1;
      for (int si = d->site_ptr[v + 1] - 1; si >= d->site_ptr[v]; --si) {
        site_vertex[si] = v;
        site_stride[si] = stride;
        stride *= d->site_dim[si];
      }
    }
    for (int c = 0; c < d->n_coords; ++c) {
      std::vector<int> sites;
      for (int s = 0; s < d->n_sites; ++s) {
        if (d->site_coord[s] < 0 || d->site_coord[s] >= d->n_coords)
          return fail(TTN_ERR_INVALID, "site_coord out of range");
        if (d->site_coord[s] == c) sites.push_back(s);
      }
      std::stable_sort(sites.begin(), sites.end(), [&](int a, int b) { return d->site_digit[a] < d->site_digit[b]; });
      for (int s : sites) {
        if (d->thr_ptr[s + 1] - d->thr_ptr[s] != d->site_dim[s]) return fail(TTN_ERR_INVALID, "thr_ptr does not match site_dim");
        if (d->thr[d->thr_ptr[s]] != 0.0) return fail(TTN_ERR_INVALID, "thr[0] of every site index must be 0");
        DigitEntry e{};
        e.site = s;
        e.base = d->site_dim[s];
        e.thr_off = d->thr_ptr[s];
        e.vertex = site_vertex[s];
        e.stride = site_stride[s];
        e.pad_ = (d->site_ptr[site_vertex[s] + 1] - d->site_ptr[site_vertex[s]] == 1) ? 1 : 0; // the vertex's only site index
        entries.push_back(e);
      }
      coord_ptr[c + 1] = (int)entries.size();
    }
  }

  // ---- device upload
  int rc;
  std::vector<double> thr(d->thr, d->thr + d->thr_ptr[d->n_sites]);
  const int NC = d->is_complex ? 2 : 1;
  std::vector<double> tensors(reinterpret_cast<const double*>(d->tensors),
                              reinterpret_cast<const double*>(d->tensors) + d->tensor_ptr[n] * NC);
  p->fe_thr_len = d->thr_ptr[d->n_sites];
  p->all_base2 = true;
  for (int s = 0; s < d->n_sites; ++s) p->all_base2 = p->all_base2 && d->site_dim[s] == 2;
  p->digits.n_coords = d->n_coords;
  p->digits.n_sites = d->n_sites;
  if ((rc = upload(p, coord_ptr, &p->digits.coord_ptr))) return rc;
  if ((rc = upload(p, entries, &p->digits.entries))) return rc;
  if ((rc = upload(p, thr, &p->digits.thr))) return rc;
  TreeDev& t = p->tree;
  t.n_vertices = n;
  t.root = d->root;
  t.is_complex = d->is_complex;
  t.msg_total = msg_total;
  t.max_inter = max_inter;
  if ((rc = upload(p, p->post, &t.post))) return rc;
  if ((rc = upload(p, p->child_ptr, &t.child_ptr))) return rc;
  if ((rc = upload(p, p->child, &t.child))) return rc;
  if ((rc = upload(p, p->link_dim, &t.link_dim))) return rc;
  if ((rc = upload(p, p->slice_size, &t.slice_size))) return rc;
  if ((rc = upload(p, p->tensor_off, &t.tensor_off))) return rc;
  if ((rc = upload(p, p->msg_off, &t.msg_off))) return rc;
  if ((rc = upload(p, tensors, &t.tensors))) return rc;

  if ((rc = build_chain(p, d))) return rc;
  if ((rc = build_chain_mma(p, d))) return rc;
  if ((rc = build_chain_gemm(p, d))) return rc;
  if ((rc = build_chain_table(p, d))) return rc;
  if ((rc = build_grid_share(p, d))) return rc;
  if (!p->is_chain && (rc = build_tree_gemm(p, d))) return rc;

  ttn_info& I = p->info;
  I.n_vertices = n;
  I.n_coords = d->n_coords;
  I.is_complex = d->is_complex;
  I.n_sites = d->n_sites;
  I.max_link_dim = max_link;
  I.is_chain = p->is_chain;
  // planner: DMMA tiles once the (real-embedded) row is wide enough to fill them, the register
  // kernel for narrow chains, the generic kernel for everything that is not a chain
  const int width = (d->is_complex ? 2 : 1) * max_link;
  if (p->ctab_ok && width <= 4) I.auto_kernel = TTN_KERNEL_TABLE; // HBM-bound regime: group tables in shared memory
  else if (p->cmma_ok && width >= 6) I.auto_kernel = TTN_KERNEL_DMMA;
  else if (p->chain_ok) I.auto_kernel = TTN_KERNEL_CHAIN;
  else if (p->cmma_ok) I.auto_kernel = TTN_KERNEL_DMMA;
  else if (p->cgemm_ok) I.auto_kernel = TTN_KERNEL_GEMM;
  else if (p->tgemm_ok && max_link >= 3) I.auto_kernel = TTN_KERNEL_TREE; // measured cross-over (scripts/tree_small_chi.py, round 2: comb chi = 3
                                                                          // 748 vs 753 M points/s generic, chi = 4 748 vs 592, binary tree chi = 3 2.9 G vs 0.98 G)
  else I.auto_kernel = TTN_KERNEL_GENERIC;
  I.device = p->device;
  I.kernels_available = (1 << TTN_KERNEL_GENERIC) | (p->chain_ok ? (1 << TTN_KERNEL_CHAIN) : 0) |
                        (p->cmma_ok ? (1 << TTN_KERNEL_DMMA) : 0) | (p->cgemm_ok ? (1 << TTN_KERNEL_GEMM) : 0) |
                        (p->tgemm_ok ? (1 << TTN_KERNEL_TREE) : 0) | (p->gshare_ok ? (1 << TTN_KERNEL_GRID) : 0) |
                        (p->ctab_ok ? (1 << TTN_KERNEL_TABLE) : 0);
  I.flops_per_point = (d->is_complex ? 8.0 : 2.0) * macs;
  p->exec_rule_flops = I.flops_per_point;
  I.bytes_per_point = 8.0 * d->n_coords + (d->is_complex ? 16.0 : 8.0);
  I.tensor_bytes = d->tensor_ptr[n] * NC * 8;
  return TTN_OK;
}

// Every entry point runs on the plan's device and restores the caller's current device on exit (a GC-driven
// ttn_plan_destroy must not move a torch / CUDA.jl caller to another GPU).
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static void destroy_plan(ttn_plan* p) {
  if (!p) return;
  for (ttn_plan* r : p->replicas) destroy_plan(r);
  p->replicas.clear();
  DeviceGuard guard(p->device);
  for (auto& st : p->streams) {
    if (st.s) cudaStreamSynchronize(st.s);
    if (st.d_coords) cudaFree(st.d_coords);
    if (st.d_out) cudaFree(st.d_out);
    if (st.d_weights) cudaFree(st.d_weights);
    if (st.d_digits) cudaFree(st.d_digits);
    if (st.d_work) cudaFree(st.d_work);
    if (st.d_partial) cudaFree(st.d_partial);
    if (st.d_gemm) cudaFree(st.d_gemm);
    if (st.d_partial2) cudaFree(st.d_partial2);
    if (st.d_sel) cudaFree(st.d_sel);
    if (st.d_refine) cudaFree(st.d_refine);
    if (st.h_coords) cudaFreeHost(st.h_coords);
    if (st.h_out) cudaFreeHost(st.h_out);
    if (st.h_weights) cudaFreeHost(st.h_weights);
    if (st.h_digits) cudaFreeHost(st.h_digits);
    if (st.h_q) cudaFreeHost(st.h_q);
    if (st.d_q) cudaFree(st.d_q);
    if (st.k0) cudaEventDestroy(st.k0);
    if (st.k1) cudaEventDestroy(st.k1);
    if (st.ev_h2d) cudaEventDestroy(st.ev_h2d);
    if (st.ev_d2h) cudaEventDestroy(st.ev_d2h);
    if (st.s) cudaStreamDestroy(st.s);
  }
  for (void* a : p->allocs) cudaFree(a);
  if (p->d_grid_out) cudaFree(p->d_grid_out);
  if (p->d_err) cudaFree(p->d_err);
  if (p->d_sum) cudaFree(p->d_sum);
  if (p->t0) cudaEventDestroy(p->t0);
  if (p->t1) cudaEventDestroy(p->t1);
  delete p;
}

static int run_kernel(ttn_plan* p, int kernel, Stream& st, const CoordSource& src, double* d_out,
                      double* d_partial, int* n_partial, int* extra_launches) {
  switch (kernel) {
    case TTN_KERNEL_TREE: {
      int nl = 0;
      const int rc = launch_tree_gemm(p, st, src, d_out, d_partial, n_partial, st.s, &nl);
      *extra_launches += nl - 1;
      return rc;
    }
    case TTN_KERNEL_GEMM: {
      int nl = 0;
      const int rc = launch_chain_gemm(p, st, src, d_out, d_partial, n_partial, st.s, &nl);
      *extra_launches += nl - 1;
      return rc;
    }
    case TTN_KERNEL_GENERIC: return launch_generic(p, st, src, d_out, d_partial, n_partial, st.s);
    case TTN_KERNEL_CHAIN: return launch_chain(p, st, src, d_out, d_partial, n_partial, st.s);
    case TTN_KERNEL_DMMA: return launch_chain_mma(p, st, src, d_out, d_partial, n_partial, st.s);
    case TTN_KERNEL_TABLE: return launch_chain_table(p, st, src, d_out, d_partial, n_partial, st.s);
    default: break;
  }
  return fail(TTN_ERR_UNSUPPORTED, "requested kernel is not available in this build");
}

constexpr int kMaxChunks = 4096;

// ---- host staging pool ------------------------------------------------------------------------------------
// cudaMemcpyAsync on PAGEABLE memory (a Julia Array, a numpy array) is staged by the driver on the calling
// thread at 10-20 GB/s and serialises the H2D / kernel / D2H pipeline.  Pageable chunks are therefore copied
// to / from a pinned ring (Stream::h_*) by a small pool of host threads, slice by slice.  The pool is a
// process-wide leaked singleton (worker threads must not be joined from a library destructor).
// Streaming copy for the staging ring: the destination (the pinned ring on the way in, the caller's array on the
// way out) is not read by the CPU soon, so non-temporal stores skip the read-for-ownership of every destination
// line — a third less DRAM traffic than memcpy, on a path whose host side is memory-bound (the DMA engines read /
// write the same DRAM at PCIe rate meanwhile).
__attribute__((target("avx2"))) static void stream_copy_avx2(char* dst, const char* src, size_t n) {
  const size_t head = std::min(n, (size_t)((32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31));
  if (head) memcpy(dst, src, head);
  dst += head, src += head, n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i a1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i a2 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i a3 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a0);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), a1);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), a2);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), a3);
  }
  _mm_sfence();
  if (i < n) memcpy(dst + i, src + i, n - i);
}
static void stream_copy(void* dst, const void* src, size_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && n >= 4096) stream_copy_avx2(static_cast<char*>(dst), static_cast<const char*>(src), n);
  else memcpy(dst, src, n);
}

// Host-side quantisation of run-path coordinates (TTN_STAGE_AUTO): q = floor(x 2^L) with x >= 1 saturating to
// 2^L - 1 — exactly what the kernels' K1 run path computes from the double (k_chain_team.cu), so the digits are
// bit-identical — written as uint32 into the pinned ring with streaming stores.  Returns true if a coordinate is
// negative or NaN (TTN_ERR_DOMAIN, as on the device).
struct PackParams {
  int nc = 0;
  double scale[TTN_MAX_COORDS];
  uint32_t qmax[TTN_MAX_COORDS];
};
// AVX2 body for 1, 2 or 4 coordinates per point and L <= 31 (signed 32-bit conversion): four doubles per step,
// q = trunc(min(x 2^L, 2^L - 1)) — the min IS the saturation of x >= 1, and for x < 1 it never bites
// (floor(x 2^L) <= 2^L - 1) — one compare for the domain flag, one streaming 16-byte store.
__attribute__((target("avx2"))) static bool pack_coords_avx2(uint32_t* dst, const double* src, size_t n_doubles, const PackParams& pp) {
  double sc[4], qm[4];
  for (int i = 0; i < 4; ++i) {
    sc[i] = pp.scale[i % pp.nc];
    qm[i] = (double)pp.qmax[i % pp.nc];
  }
  const __m256d vs = _mm256_loadu_pd(sc), vq = _mm256_loadu_pd(qm), zero = _mm256_setzero_pd();
  int ok = 0xF;
  size_t i = 0;
  for (; i + 4 <= n_doubles; i += 4) {
    const __m256d x = _mm256_loadu_pd(src + i);
    ok &= _mm256_movemask_pd(_mm256_cmp_pd(x, zero, _CMP_GE_OQ));
    // max(t, 0) with t first: a NaN / negative lane (flagged above) becomes 0, never an out-of-range field
    const __m128i q = _mm256_cvttpd_epi32(_mm256_max_pd(_mm256_min_pd(_mm256_mul_pd(x, vs), vq), zero));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), q);
  }
  _mm_sfence();
  bool bad = ok != 0xF;
  for (; i < n_doubles; ++i) { // tail (n_doubles is a multiple of nc; i % 4 == 0 keeps the coordinate phase)
    const double x = src[i];
    const int c = (int)(i % pp.nc);
    if (!(x >= 0.0)) {
      bad = true;
      dst[i] = 0u;
    } else {
      dst[i] = x >= 1.0 ? pp.qmax[c] : (uint32_t)(unsigned long long)(x * pp.scale[c]);
    }
  }
  return bad;
}

static bool pack_coords(uint32_t* dst, const double* src, size_t npts, const PackParams& pp) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && (pp.nc == 1 || pp.nc == 2 || pp.nc == 4) && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    bool l31 = true;
    for (int c = 0; c < pp.nc; ++c) l31 = l31 && pp.qmax[c] <= 0x7fffffffu;
    if (l31) return pack_coords_avx2(dst, src, npts * pp.nc, pp);
  }
  bool bad = false;
  auto one = [&](double x, int c) -> uint32_t {
    if (!(x >= 0.0)) {
      bad = true;
      return 0u;
    }
    if (x >= 1.0) return pp.qmax[c];
    return (uint32_t)(unsigned long long)(x * pp.scale[c]);
  };
  if (pp.nc == 2 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    for (size_t p = 0; p < npts; ++p) {
      const unsigned long long lo = one(src[2 * p], 0), hi = one(src[2 * p + 1], 1);
      _mm_stream_si64(reinterpret_cast<long long*>(dst) + p, (long long)(lo | (hi << 32)));
    }
    _mm_sfence();
  } else {
    const int nc = pp.nc;
    for (size_t p = 0; p < npts; ++p)
      for (int c = 0; c < nc; ++c) dst[p * nc + c] = one(src[p * nc + c], c);
  }
  return bad;
}

class CopyPool {
 public:
  static CopyPool& get() {
    static CopyPool* pool = new CopyPool();
    return *pool;
  }
  // dst[0:bytes) = src[0:bytes), in parallel; returns when done.  The caller takes slices too.
  void copy(void* dst, const void* src, size_t bytes) {
    constexpr size_t kSlice = (size_t)1 << 20;
    if (bytes <= 2 * kSlice || workers_ == 0) {
      stream_copy(dst, src, bytes);
      return;
    }
    auto job = std::make_shared<Job>();
    job->dst = static_cast<char*>(dst);
    job->src = static_cast<const char*>(src);
    job->bytes = bytes;
    job->n_slices = (bytes + kSlice - 1) / kSlice;
    {
      std::lock_guard<std::mutex> lk(mu_);
      jobs_.push_back(job);
    }
    cv_.notify_all();
    work_on(*job);
    std::unique_lock<std::mutex> lk(job->mu);
    job->cv.wait(lk, [&] { return job->done.load() == job->n_slices; });
  }
  // dst[0 : npts * nc) = quantised src[0 : npts * nc), in parallel (slices of 64 Ki points); returns the domain flag
  bool pack(uint32_t* dst, const double* src, size_t npts, const PackParams& pp) {
    constexpr size_t kPts = (size_t)1 << 16;
    if (npts <= 2 * kPts || workers_ == 0) return pack_coords(dst, src, npts, pp);
    auto job = std::make_shared<Job>();
    job->dst = reinterpret_cast<char*>(dst);
    job->src = reinterpret_cast<const char*>(src);
    job->bytes = npts; // points
    job->n_slices = (npts + kPts - 1) / kPts;
    job->kind = 1;
    job->pp = pp;
    {
      std::lock_guard<std::mutex> lk(mu_);
      jobs_.push_back(job);
    }
    cv_.notify_all();
    work_on(*job);
    std::unique_lock<std::mutex> lk(job->mu);
    job->cv.wait(lk, [&] { return job->done.load() == job->n_slices; });
    return job->bad.load() != 0;
  }
  int threads() const { return workers_ + 1; }

 private:
  struct Job {
    char* dst;
    const char* src;
    size_t bytes, n_slices;
    int kind = 0; // 0: streaming copy of `bytes` bytes; 1: pack_coords of `bytes` POINTS
    PackParams pp;
    std::atomic<int> bad{0};
    std::atomic<size_t> next{0}, done{0};
    std::mutex mu;
    std::condition_variable cv;
  };
  CopyPool() {
    int n = 8;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::min(16, std::max(1, CPU_COUNT(&set)));
    if (const char* e = getenv("TTN_HOST_THREADS")) n = std::max(1, std::min(atoi(e), 64));
    workers_ = n - 1;
    for (int i = 0; i < workers_; ++i) std::thread([this] { worker(); }).detach();
  }
  static void work_on(Job& j) {
    constexpr size_t kSlice = (size_t)1 << 20;
    for (;;) {
      const size_t i = j.next.fetch_add(1);
      if (i >= j.n_slices) return;
      if (j.kind == 1) {
        constexpr size_t kPts = (size_t)1 << 16;
        const size_t p0 = i * kPts, np = std::min(kPts, j.bytes - p0);
        if (pack_coords(reinterpret_cast<uint32_t*>(j.dst) + p0 * j.pp.nc, reinterpret_cast<const double*>(j.src) + p0 * j.pp.nc, np, j.pp))
          j.bad.store(1);
      } else {
        const size_t off = i * kSlice, len = std::min(kSlice, j.bytes - off);
        stream_copy(j.dst + off, j.src + off, len);
      }
      if (j.done.fetch_add(1) + 1 == j.n_slices) {
        std::lock_guard<std::mutex> lk(j.mu);
        j.cv.notify_all();
      }
    }
  }
  void worker() {
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] {
          while (!jobs_.empty() && jobs_.front()->next.load() >= jobs_.front()->n_slices) jobs_.pop_front();
          return !jobs_.empty();
        });
        job = jobs_.front();
      }
      work_on(*job);
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::shared_ptr<Job>> jobs_;
  int workers_ = 0;
};

// How a caller buffer is reached from the plan's device.
enum Space {
  SPACE_NONE = 0,
  SPACE_DIRECT,   // device memory of the plan's device (or managed): kernels use the pointer
  SPACE_ASYNC,    // pinned host memory: cudaMemcpyAsync in place (PCIe)
  SPACE_PAGEABLE, // pageable host memory: through the pinned staging ring
  SPACE_PEER      // another GPU's memory: cudaMemcpyAsync in place (NVLink peer copy)
};

static Space classify(const void* ptr, int declared_mem, int device, int host_staging, size_t bytes) {
  if (!ptr) return SPACE_NONE;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return declared_mem == TTN_MEM_DEVICE ? SPACE_DIRECT : SPACE_ASYNC;
  }
  if (a.type == cudaMemoryTypeDevice) return a.device == device ? SPACE_DIRECT : SPACE_PEER;
  if (a.type == cudaMemoryTypeManaged) return SPACE_DIRECT;
  if (a.type == cudaMemoryTypeHost) return SPACE_ASYNC;
  // unregistered host memory; small buffers are not worth the ring
  if (host_staging == TTN_STAGE_OFF || bytes < ((size_t)4 << 20)) return SPACE_ASYNC;
  return SPACE_PAGEABLE;
}

static int ensure_stream_buffers(ttn_plan* p, Stream& st, int64_t chunk, bool need_coords, bool need_out) {
  const int NC = p->info.is_complex ? 2 : 1;
  if (st.cap_points < chunk || (need_coords && !st.d_coords) || (need_out && !st.d_out)) {
    if (st.d_coords) cudaFree(st.d_coords);
    if (st.d_out) cudaFree(st.d_out);
    st.d_coords = st.d_out = nullptr;
    st.cap_points = 0;
    TTN_CUDA(cudaMalloc(&st.d_coords, std::max<size_t>(16, sizeof(double) * (size_t)chunk * std::max(p->info.n_coords, 1))));
    TTN_CUDA(cudaMalloc(&st.d_out, sizeof(double) * (size_t)chunk * NC));
    if (st.d_weights) cudaFree(st.d_weights);
    st.d_weights = nullptr;
    TTN_CUDA(cudaMalloc(&st.d_weights, sizeof(double) * (size_t)chunk));
    st.cap_points = chunk;
  }
  const size_t pc = (size_t)3 * (p->sm_count * 8 + 8);
  if (st.partial_cap < pc) {
    if (st.d_partial) cudaFree(st.d_partial);
    st.d_partial = nullptr;
    st.partial_cap = 0;
    TTN_CUDA(cudaMalloc(&st.d_partial, pc * sizeof(double)));
    st.partial_cap = pc;
  }
  return TTN_OK;
}

// pinned staging ring of one stream (allocated on the first pageable call, at chunk capacity)
static int ensure_host_ring(ttn_plan* p, Stream& st, int64_t chunk, bool coords, bool out, bool weights) {
  const int NC = p->info.is_complex ? 2 : 1;
  if (st.h_cap_points < chunk) {
    if (st.h_coords) cudaFreeHost(st.h_coords);
    if (st.h_out) cudaFreeHost(st.h_out);
    if (st.h_weights) cudaFreeHost(st.h_weights);
    st.h_coords = st.h_out = st.h_weights = nullptr;
    st.h_cap_points = chunk;
  }
  if (coords && !st.h_coords)
    TTN_CUDA(cudaHostAlloc(&st.h_coords, sizeof(double) * (size_t)st.h_cap_points * std::max(p->info.n_coords, 1), cudaHostAllocPortable));
  if (out && !st.h_out) TTN_CUDA(cudaHostAlloc(&st.h_out, sizeof(double) * (size_t)st.h_cap_points * NC, cudaHostAllocPortable));
  if (weights && !st.h_weights) TTN_CUDA(cudaHostAlloc(&st.h_weights, sizeof(double) * (size_t)st.h_cap_points, cudaHostAllocPortable));
  if (!st.ev_h2d) TTN_CUDA(cudaEventCreateWithFlags(&st.ev_h2d, cudaEventDisableTiming));
  if (!st.ev_d2h) TTN_CUDA(cudaEventCreateWithFlags(&st.ev_d2h, cudaEventDisableTiming));
  return TTN_OK;
}

static int validate_opts(const ttn_opts* opts, const void* out) {
  if (opts->reduce_sum < 0 || opts->reduce_sum > TTN_REDUCE_WEIGHTED) return fail(TTN_ERR_INVALID, "bad reduce mode");
  if (opts->reduce_sum == TTN_REDUCE_WEIGHTED && !opts->weights) return fail(TTN_ERR_INVALID, "TTN_REDUCE_WEIGHTED needs opts->weights");
  if (!out && opts->reduce_sum == TTN_REDUCE_NONE) return fail(TTN_ERR_INVALID, "out is NULL and reduce_sum is 0: nothing to compute");
  if (opts->accuracy != TTN_ACCURACY_FP64 && opts->accuracy != TTN_ACCURACY_REFINED) return fail(TTN_ERR_INVALID, "bad accuracy mode");
  if (opts->host_staging < TTN_STAGE_AUTO || opts->host_staging > TTN_STAGE_COPY) return fail(TTN_ERR_INVALID, "bad host_staging mode");
  if (!(opts->refine_tau >= 0.0) || opts->refine_tau > 1e3) return fail(TTN_ERR_INVALID, "refine_tau must be in [0, 1000]");
  return TTN_OK;
}

// Shared driver of ttn_evaluate / ttn_evaluate_grid / ttn_evaluate_indices on ONE device.
// soa_stride: distance (in points) between two coordinate slots of a host / peer SOA array (the whole call's
// npts; a multi-device block is a sub-range of it).
static int evaluate_impl(ttn_plan* p, CoordSource base, const double* coords, void* out, ttn_opts* opts,
                         const uint8_t* digits = nullptr, int64_t soa_stride = 0) {
  std::lock_guard<std::mutex> lock(p->mu);
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(TTN_ERR_CUDA, "cudaSetDevice failed");
  const auto wall0 = std::chrono::steady_clock::now();
  const int NC = p->info.is_complex ? 2 : 1;
  const int64_t npts = base.npts;
  if (soa_stride == 0) soa_stride = npts;
  opts->sum_out[0] = opts->sum_out[1] = 0.0;
  opts->kernel_ms = opts->total_ms = 0.f;
  opts->n_launches = 0;
  opts->n_devices_used = 1;
  opts->staged = 0;
  opts->n_refined = 0;
  opts->flops_executed = 0.0;
  opts->h2d_bytes = opts->d2h_bytes = 0;
  int rc = validate_opts(opts, out);
  if (rc) return rc;
  if (npts < 0) return fail(TTN_ERR_INVALID, "npts < 0");
  const bool refine = opts->accuracy == TTN_ACCURACY_REFINED;
  const double tau = opts->refine_tau > 0.0 ? opts->refine_tau : 0.02;
  const bool do_sum = opts->reduce_sum != TTN_REDUCE_NONE;
  const int n_sites = std::max(p->info.n_sites, 1);

  const Space sp_out = classify(out, opts->out_mem, p->device, opts->host_staging, sizeof(double) * (size_t)npts * NC);
  const Space sp_w = opts->reduce_sum == TTN_REDUCE_WEIGHTED
                         ? classify(opts->weights, opts->weights_mem, p->device, opts->host_staging, sizeof(double) * (size_t)npts)
                         : SPACE_NONE;

  // full dyadic grid of a binary chain: prefix-shared expansion (one pass over the whole grid)
  if (base.grid && !refine && (opts->kernel == TTN_KERNEL_AUTO || opts->kernel == TTN_KERNEL_GRID) && grid_share_applicable(p, base) &&
      (sp_w == SPACE_NONE || sp_w == SPACE_DIRECT)) {
    opts->kernel_used = TTN_KERNEL_GRID;
    if (npts == 0) return TTN_OK;
    Stream& st = p->streams[0];
    rc = ensure_stream_buffers(p, st, 1, false, false);
    if (rc) return rc;
    double* d_out = reinterpret_cast<double*>(out);
    bool scratch_ok = true;
    if (out && sp_out != SPACE_DIRECT) {
      // grow-only scratch for the whole grid (the kernel writes every level-L row once); when the device cannot
      // hold it the per-point kernels below evaluate the grid chunk by chunk instead
      const size_t need = sizeof(double) * (size_t)npts * NC;
      if (p->grid_out_bytes < need) {
        if (p->d_grid_out) cudaFree(p->d_grid_out);
        p->d_grid_out = nullptr;
        p->grid_out_bytes = 0;
        if (cudaMalloc(&p->d_grid_out, need) == cudaSuccess) p->grid_out_bytes = need;
        else {
          cudaGetLastError();
          scratch_ok = false;
        }
      }
      d_out = p->d_grid_out;
    }
    if (scratch_ok) {
      CoordSource src = base;
      src.reduce_mode = opts->reduce_sum;
      src.weights = opts->reduce_sum == TTN_REDUCE_WEIGHTED ? opts->weights : nullptr;
      cudaEventRecord(st.k0, st.s);
      int n_partial = 0;
      double flops = 0.0;
      rc = launch_grid_share(p, st, src, d_out, do_sum ? st.d_partial : nullptr, &n_partial, st.s, &opts->n_launches, &flops);
      if (rc == TTN_OK && do_sum) {
        rc = launch_sum_partials(p, st.d_partial, n_partial, NC, p->d_sum, st.s);
        opts->n_launches += 1;
      }
      cudaEventRecord(st.k1, st.s);
      if (rc == TTN_OK && out && sp_out != SPACE_DIRECT) {
        // D2H in slices so that a pageable destination goes through the staging ring as well
        const int64_t slice = (int64_t)1 << 22;
        for (int64_t f0 = 0; f0 < npts && rc == TTN_OK; f0 += slice) {
          const int64_t m = std::min(slice, npts - f0);
          opts->d2h_bytes += (int64_t)(sizeof(double) * m * NC);
          if (sp_out == SPACE_PAGEABLE) {
            rc = ensure_host_ring(p, st, slice, false, true, false);
            if (rc) break;
            cudaMemcpyAsync(st.h_out, d_out + f0 * NC, sizeof(double) * m * NC, cudaMemcpyDeviceToHost, st.s);
            cudaStreamSynchronize(st.s);
            CopyPool::get().copy(reinterpret_cast<double*>(out) + f0 * NC, st.h_out, sizeof(double) * m * NC);
            opts->staged |= 2;
          } else {
            cudaMemcpyAsync(reinterpret_cast<double*>(out) + f0 * NC, d_out + f0 * NC, sizeof(double) * m * NC, cudaMemcpyDefault, st.s);
          }
        }
      }
      const cudaError_t e = cudaStreamSynchronize(st.s);
      if (rc == TTN_OK && e != cudaSuccess) rc = fail(TTN_ERR_CUDA, std::string("grid kernel: ") + cudaGetErrorString(e));
      if (rc == TTN_OK) {
        cudaEventElapsedTime(&opts->kernel_ms, st.k0, st.k1);
        opts->flops_executed = flops;
        if (do_sum) {
          double hs[2];
          cudaMemcpy(hs, p->d_sum, sizeof(hs), cudaMemcpyDeviceToHost);
          opts->sum_out[0] = hs[0];
          opts->sum_out[1] = hs[1];
        }
      }
      opts->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
      return rc;
    }
  }
  if (opts->kernel == TTN_KERNEL_GRID)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_GRID needs ttn_evaluate_grid on the FULL dyadic grid (count = 2^L, step = 2^-L, first = 0) of a chain with one binary site index per vertex and width <= 32, in TTN_ACCURACY_FP64");
  int kernel = opts->kernel == TTN_KERNEL_AUTO ? p->info.auto_kernel : opts->kernel;
  if (kernel == TTN_KERNEL_CHAIN && !p->chain_ok)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_CHAIN: network is not a chain with chi <= 32 (real) / 16 (complex) and <= 4 slices per vertex");
  if (kernel == TTN_KERNEL_TABLE && !p->ctab_ok)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_TABLE: network is not a chain of binary site indices with chi <= 4 (real) / 2 (complex), <= 2 site indices per vertex and <= 128 slice bits");
  if (kernel == TTN_KERNEL_TREE && !p->tgemm_ok)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_TREE: network is not a tree the per-vertex GEMM kernel covers (<= 2 children per vertex, chi <= 64 real / 32 complex, <= 8 slices per vertex)");
  if (kernel == TTN_KERNEL_GEMM && !p->cgemm_ok)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_GEMM: network is not a chain with 32 < (real-embedded) width <= 256 and <= 8 slices per vertex");
  if (kernel == TTN_KERNEL_DMMA && !p->cmma_ok)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_DMMA: network is not a chain with chi <= 32 (real) / 16 (complex), <= 4 slices per vertex and <= 128 slice bits");
  opts->kernel_used = kernel;
  if (npts == 0) return TTN_OK;
  if (!base.grid && !coords && !digits) return fail(TTN_ERR_INVALID, "coords is NULL");

  const void* in_ptr = digits ? static_cast<const void*>(digits) : static_cast<const void*>(coords);
  const size_t in_bytes = digits ? (size_t)npts * n_sites : sizeof(double) * (size_t)npts * std::max(base.n_coords, 1);
  Space sp_in = base.grid ? SPACE_NONE : classify(in_ptr, opts->coords_mem, p->device, opts->host_staging, in_bytes);
  // a sub-range of an SOA array is strided: only the staged copies can address it
  if (sp_in == SPACE_DIRECT && !digits && base.layout == TTN_LAYOUT_SOA && soa_stride != npts) sp_in = SPACE_PEER; // device-to-device copies
  const bool in_staged = sp_in == SPACE_ASYNC || sp_in == SPACE_PAGEABLE || sp_in == SPACE_PEER;
  const bool out_staged = sp_out == SPACE_ASYNC || sp_out == SPACE_PAGEABLE || sp_out == SPACE_PEER;
  const bool w_staged = sp_w == SPACE_ASYNC || sp_w == SPACE_PAGEABLE || sp_w == SPACE_PEER;
  // host buffers: the call is bound by the PCIe copies that run beside the kernels (launch_chain_team)
  const bool pcie_bound = sp_in == SPACE_ASYNC || sp_in == SPACE_PAGEABLE || sp_out == SPACE_ASYNC || sp_out == SPACE_PAGEABLE;
  // Host-side quantisation of the coordinates (include/ttneval.h, TTN_STAGE_AUTO): every coordinate on the K1 run path
  // of the team-sorted kernel with L <= 32, AoS host doubles in PAGEABLE memory — the staging threads touch every byte
  // of such an array anyway, and writing 4 bytes per coordinate into the ring instead of 8 takes a quarter off their
  // DRAM traffic and half off the H2D bytes (config 2, 1 GPU: 2.0 -> 2.5 G points/s).  Pinned arrays are read by the
  // copy engines directly, which is the cheapest path for the HOST memory system (16 B per point, no CPU pass): with
  // TTN_HOST_QUANT=2 they are quantised too, +8 % on one GPU (3.19 -> 3.46 G points/s, then bound by host DRAM at
  // ~140 GB/s instead of PCIe) but a loss as soon as several GPUs share that DRAM, so it is not the default.
  // HYBRID (mode 5, the default where this GPU has the host to itself: ttn_plan::host_peers == 1): two of every three chunks
  // of a PINNED array are quantised by the host threads while the copy engines read the third in place — 10.7 B per point
  // over PCIe instead of 16, the host pass and the DMA sharing the host's DRAM bandwidth — and the call then runs the
  // deep-table image (with a third less copy traffic beside it the gathers cost the copies less than the 2x faster kernel
  // gains).  Config 2, one GPU, end to end (scripts/microbench/quant_probe.sh): doubles + light image 3.15 G points/s,
  // 1 of 2 chunks quantised + deep image 3.76, 2 of 3: 4.03, every chunk: 3.44 (the 16 host threads then set the pace),
  // 1 of 2 + light image 3.37 (the light kernel's 3.6 G sets it).  With several GPUs on one host the host's DRAM
  // bandwidth is the ceiling and every extra host pass costs: mode 1.
  PackParams pp;
  const int quant_mode = getenv("TTN_HOST_QUANT") ? atoi(getenv("TTN_HOST_QUANT")) : (p->host_peers <= 1 ? 5 : 1);
  bool quant = opts->host_staging == TTN_STAGE_AUTO && !base.grid && !digits && coords && !refine && base.layout == TTN_LAYOUT_AOS &&
               (sp_in == SPACE_PAGEABLE || (sp_in == SPACE_ASYNC && quant_mode >= 2)) && quant_mode != 0 &&
               kernel == TTN_KERNEL_DMMA && chain_team_applicable(p) && base.n_coords >= 1 && npts >= ((int64_t)1 << 20);
  if (quant) {
    pp.nc = base.n_coords;
    for (int c = 0; c < base.n_coords && quant; ++c) {
      const int L = p->cmma.run_L[c];
      quant = L >= 1 && L <= 32;
      pp.scale[c] = p->cmma.run_scale[c];
      pp.qmax[c] = L >= 32 ? 0xffffffffu : ((1u << (L & 31)) - 1u);
    }
  }
  // which chunks are quantised: all of a pageable array; of a PINNED one all (mode 2) or a fraction — the copy engines
  // read the other chunks in place while the host threads pack, so the call uses the PCIe link (16 B per direct point,
  // 8 B per packed one) and the host's DRAM bandwidth (24 B against 40 B) together
  auto quant_chunk = [&](int ci) {
    if (!quant) return false;
    if (sp_in == SPACE_PAGEABLE || quant_mode == 2) return true;
    switch (quant_mode) {
      case 3: return (ci & 1) == 1;  // 1 of 2
      case 4: return (ci % 3) == 2;  // 1 of 3
      case 6: return (ci & 3) != 0;  // 3 of 4
      default: return (ci % 3) != 0; // 5: 2 of 3
    }
  };
  // the image the team-sorted kernel runs: pinned arrays in a quantising mode take the deep-table image (above)
  const bool light_image = pcie_bound && !(quant && sp_in == SPACE_ASYNC);
  bool host_bad = false;
  // the refine pass and its functionals need the values of a chunk in device memory
  const bool need_dout = out_staged || (refine && !out);
  const bool chunked = in_staged || need_dout || w_staged;

  TTN_CUDA(cudaMemsetAsync(p->d_err, 0, sizeof(int), p->streams[0].s));
  TTN_CUDA(cudaStreamSynchronize(p->streams[0].s));

  int64_t chunk = npts;
  int n_chunks = 1;
  if (chunked) {
    chunk = opts->chunk_points > 0 ? opts->chunk_points : (int64_t)1 << 21; // measured best on PCIe 5 x16 (scripts/e2e_probe.py)
    chunk = std::min(chunk, npts);
    n_chunks = (int)((npts + chunk - 1) / chunk);
    if (n_chunks > kMaxChunks) {
      chunk = (npts + kMaxChunks - 1) / kMaxChunks;
      n_chunks = (int)((npts + chunk - 1) / chunk);
    }
  }
  if (!chunked && npts > (int64_t)0x7fffff00 && refine) return fail(TTN_ERR_UNSUPPORTED, "TTN_ACCURACY_REFINED: more than 2^31 points in one un-chunked call");
  if (sp_in == SPACE_DIRECT && !digits && base.layout == TTN_LAYOUT_SOA && n_chunks > 1)
    return fail(TTN_ERR_UNSUPPORTED, "SOA device coordinates with a host output buffer are not supported; use AOS");
  // per-chunk timing events, destroyed on EVERY exit path (TTN_CUDA returns early on allocation failures)
  struct EventSet {
    std::vector<cudaEvent_t> v;
    ~EventSet() {
      for (auto e : v)
        if (e) cudaEventDestroy(e);
    }
  } evs;
  evs.v.assign(2 * (size_t)n_chunks, nullptr);
  std::vector<cudaEvent_t>& ev = evs.v;
  const int n_streams = chunked ? 3 : 1;
  // pending D2H of a staged-pageable output: (destination, bytes) per stream, drained before the slot is reused
  struct Pending {
    double* dst = nullptr;
    size_t bytes = 0;
  } pend[3];
  auto drain = [&](int si) {
    if (!pend[si].dst) return;
    cudaEventSynchronize(p->streams[si].ev_d2h);
    CopyPool::get().copy(pend[si].dst, p->streams[si].h_out, pend[si].bytes);
    pend[si].dst = nullptr;
  };
  const bool any_pageable = sp_in == SPACE_PAGEABLE || sp_out == SPACE_PAGEABLE || sp_w == SPACE_PAGEABLE || quant;
  for (int ci = 0; ci < n_chunks && rc == TTN_OK; ++ci) {
    const int si = ci % n_streams;
    Stream& st = p->streams[si];
    const int64_t first = (int64_t)ci * chunk;
    const int64_t m = std::min(chunk, npts - first);
    rc = ensure_stream_buffers(p, st, chunked ? chunk : 1, in_staged && !digits, need_dout);
    if (rc) break;
    if (any_pageable) {
      rc = ensure_host_ring(p, st, chunk, sp_in == SPACE_PAGEABLE && !digits && !quant, sp_out == SPACE_PAGEABLE, sp_w == SPACE_PAGEABLE);
      if (rc) break;
      drain(si);                                                  // values of chunk ci - 3 -> caller's array
      if (sp_in == SPACE_PAGEABLE || sp_w == SPACE_PAGEABLE || quant) cudaEventSynchronize(st.ev_h2d); // ring slot free again
    }
    CoordSource src = base;
    src.npts = m;
    src.reduce_mode = opts->reduce_sum;
    src.weights = nullptr;
    src.pcie_bound = light_image ? 1 : 0;
    if (opts->reduce_sum == TTN_REDUCE_WEIGHTED) {
      if (w_staged) opts->h2d_bytes += (int64_t)(sizeof(double) * m);
      if (sp_w == SPACE_PAGEABLE) {
        CopyPool::get().copy(st.h_weights, opts->weights + first, sizeof(double) * m);
        cudaMemcpyAsync(st.d_weights, st.h_weights, sizeof(double) * m, cudaMemcpyHostToDevice, st.s);
        src.weights = st.d_weights;
      } else if (sp_w == SPACE_ASYNC || sp_w == SPACE_PEER) {
        cudaMemcpyAsync(st.d_weights, opts->weights + first, sizeof(double) * m, cudaMemcpyDefault, st.s);
        src.weights = st.d_weights;
      } else {
        src.weights = opts->weights + first;
      }
    }
    if (digits) {
      if (in_staged) {
        if (!st.d_digits || st.digits_cap < (size_t)chunk * n_sites) {
          if (st.d_digits) cudaFree(st.d_digits);
          st.d_digits = nullptr;
          st.digits_cap = 0;
          TTN_CUDA(cudaMalloc(&st.d_digits, (size_t)chunk * n_sites));
          st.digits_cap = (size_t)chunk * n_sites;
        }
        const uint8_t* hsrc = digits + (size_t)first * p->info.n_sites;
        const size_t nbytes = (size_t)m * p->info.n_sites;
        if (sp_in == SPACE_PAGEABLE) {
          if (st.h_digits_cap < (size_t)chunk * n_sites) {
            if (st.h_digits) cudaFreeHost(st.h_digits);
            st.h_digits = nullptr;
            st.h_digits_cap = 0;
            TTN_CUDA(cudaHostAlloc(&st.h_digits, (size_t)chunk * n_sites, cudaHostAllocPortable));
            st.h_digits_cap = (size_t)chunk * n_sites;
          }
          CopyPool::get().copy(st.h_digits, hsrc, nbytes);
          hsrc = st.h_digits;
          opts->staged |= 1;
        }
        cudaMemcpyAsync(st.d_digits, hsrc, nbytes, cudaMemcpyDefault, st.s);
        opts->h2d_bytes += (int64_t)nbytes;
        src.digits = st.d_digits;
      } else {
        src.digits = digits + (size_t)first * p->info.n_sites;
      }
    } else if (base.grid) {
      src.first = base.first + first;
    } else if (quant_chunk(ci)) {
      if (st.q_cap_points < chunk) {
        if (st.h_q) cudaFreeHost(st.h_q);
        if (st.d_q) cudaFree(st.d_q);
        st.h_q = st.d_q = nullptr;
        st.q_cap_points = 0;
        TTN_CUDA(cudaHostAlloc(&st.h_q, sizeof(uint32_t) * (size_t)chunk * base.n_coords, cudaHostAllocPortable));
        TTN_CUDA(cudaMalloc(&st.d_q, sizeof(uint32_t) * (size_t)chunk * base.n_coords));
        st.q_cap_points = chunk;
      }
      host_bad = CopyPool::get().pack(st.h_q, coords + first * base.n_coords, (size_t)m, pp) || host_bad;
      const size_t qb = sizeof(uint32_t) * (size_t)m * base.n_coords;
      cudaMemcpyAsync(st.d_q, st.h_q, qb, cudaMemcpyHostToDevice, st.s);
      opts->h2d_bytes += (int64_t)qb;
      opts->staged |= 4 | (sp_in == SPACE_PAGEABLE ? 1 : 0);
      src.qcoords = st.d_q;
      src.coords = nullptr;
    } else if (in_staged) {
      const size_t row = sizeof(double) * (size_t)m;
      opts->h2d_bytes += (int64_t)(row * base.n_coords);
      if (base.layout == TTN_LAYOUT_AOS) {
        const double* hsrc = coords + first * base.n_coords;
        if (sp_in == SPACE_PAGEABLE) {
          CopyPool::get().copy(st.h_coords, hsrc, row * base.n_coords);
          hsrc = st.h_coords;
          opts->staged |= 1;
        }
        cudaMemcpyAsync(st.d_coords, hsrc, row * base.n_coords, cudaMemcpyDefault, st.s);
      } else {
        for (int c = 0; c < base.n_coords; ++c) {
          const double* hsrc = coords + (int64_t)c * soa_stride + first;
          if (sp_in == SPACE_PAGEABLE) {
            CopyPool::get().copy(st.h_coords + (int64_t)c * m, hsrc, row);
            hsrc = st.h_coords + (int64_t)c * m;
            opts->staged |= 1;
          }
          cudaMemcpyAsync(st.d_coords + (int64_t)c * m, hsrc, row, cudaMemcpyDefault, st.s);
        }
      }
      src.coords = st.d_coords;
    } else {
      // device coordinates: a chunk is a sub-range of the caller's array
      src.coords = base.layout == TTN_LAYOUT_AOS ? coords + first * base.n_coords : coords;
    }
    if (sp_in == SPACE_PAGEABLE || sp_w == SPACE_PAGEABLE || quant) cudaEventRecord(st.ev_h2d, st.s);
    double* d_out = nullptr;
    if (need_dout) d_out = st.d_out;
    else if (out) d_out = reinterpret_cast<double*>(out) + first * NC;
    cudaEventCreate(&ev[2 * ci]);
    cudaEventCreate(&ev[2 * ci + 1]);
    cudaEventRecord(ev[2 * ci], st.s);
    int n_partial = 0;
    rc = run_kernel(p, kernel, st, src, d_out, (do_sum && !refine) ? st.d_partial : nullptr, &n_partial, &opts->n_launches);
    if (rc) break;
    opts->n_launches += 1;
    if (refine) {
      if (st.sel_cap < (size_t)(chunked ? chunk : npts)) {
        if (st.d_sel) cudaFree(st.d_sel);
        st.d_sel = nullptr;
        st.sel_cap = 0;
        const size_t cap = (size_t)(chunked ? chunk : npts);
        TTN_CUDA(cudaMalloc(&st.d_sel, sizeof(int32_t) * (cap + 1)));
        st.sel_cap = cap;
      }
      rc = launch_refine(p, st, src, d_out, tau, st.d_sel, do_sum ? st.d_partial : nullptr, &n_partial, st.s, &opts->n_launches);
      if (rc) break;
      // the count of this chunk, summed on the host after the streams have drained
      cudaMemcpyAsync(p->d_nsel + ci, st.d_sel, sizeof(int32_t), cudaMemcpyDeviceToDevice, st.s);
    }
    if (do_sum) {
      rc = launch_sum_partials(p, st.d_partial, n_partial, NC, p->d_sum + 2 * ci, st.s);
      if (rc) break;
      opts->n_launches += 1;
    }
    cudaEventRecord(ev[2 * ci + 1], st.s);
    if (out && need_dout) {
      double* dst = reinterpret_cast<double*>(out) + first * NC;
      const size_t nbytes = sizeof(double) * (size_t)m * NC;
      opts->d2h_bytes += (int64_t)nbytes;
      if (sp_out == SPACE_PAGEABLE) {
        cudaMemcpyAsync(st.h_out, st.d_out, nbytes, cudaMemcpyDeviceToHost, st.s);
        cudaEventRecord(st.ev_d2h, st.s);
        pend[si].dst = dst;
        pend[si].bytes = nbytes;
        opts->staged |= 2;
      } else {
        cudaMemcpyAsync(dst, st.d_out, nbytes, cudaMemcpyDefault, st.s);
      }
    }
  }
  for (int s = 0; s < n_streams; ++s) drain((n_chunks + s) % n_streams); // oldest first
  cudaError_t e = cudaSuccess;
  for (int s = 0; s < n_streams; ++s) {
    cudaError_t es = cudaStreamSynchronize(p->streams[s].s);
    if (es != cudaSuccess) e = es;
  }
  if (rc == TTN_OK && e != cudaSuccess) rc = fail(TTN_ERR_CUDA, std::string("kernel execution: ") + cudaGetErrorString(e));
  if (rc == TTN_OK) {
    float total = 0.f;
    for (int ci = 0; ci < n_chunks; ++ci) {
      float ms = 0.f;
      if (ev[2 * ci] && ev[2 * ci + 1] && cudaEventElapsedTime(&ms, ev[2 * ci], ev[2 * ci + 1]) == cudaSuccess) total += ms;
    }
    opts->kernel_ms = total;
    opts->flops_executed = ((kernel == TTN_KERNEL_DMMA && p->cmma.merged)   ? ((light_image && p->cmma_light_ok && chain_team_applicable(p)) ? p->cmma_light_flops : p->cmma_flops_exec)
                            : (kernel == TTN_KERNEL_GEMM && p->cgemm.merged) ? p->cgemm_flops_exec
                            : kernel == TTN_KERNEL_TREE ? p->tgemm_flops_exec
                            : kernel == TTN_KERNEL_TABLE ? p->ctab_flops_exec
                                                                             : p->exec_rule_flops) *
                           (double)npts;
    int herr = 0;
    cudaMemcpy(&herr, p->d_err, sizeof(int), cudaMemcpyDeviceToHost);
    if (herr & 2)
      rc = fail(TTN_ERR_INVALID, "an index value is out of range for its site index");
    else if (herr || host_bad)
      rc = fail(TTN_ERR_DOMAIN,
                "a coordinate is negative or NaN (the reference's digit loop, abstractindexmap.jl:121-138, does not terminate on such input)");
    if (do_sum && rc == TTN_OK) {
      std::vector<double> hs(2 * (size_t)n_chunks);
      cudaMemcpy(hs.data(), p->d_sum, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost);
      double a = 0.0, b = 0.0;
      for (int ci = 0; ci < n_chunks; ++ci) { // fixed order => deterministic
        a += hs[2 * ci];
        b += hs[2 * ci + 1];
      }
      opts->sum_out[0] = a;
      opts->sum_out[1] = b;
    }
    if (refine && rc == TTN_OK) {
      std::vector<int32_t> hn((size_t)n_chunks);
      cudaMemcpy(hn.data(), p->d_nsel, sizeof(int32_t) * hn.size(), cudaMemcpyDeviceToHost);
      for (int32_t v : hn) opts->n_refined += v;
    }
  }
  opts->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  return rc;
}

// Contiguous block of device g: [g * ceil(n / G), min(n, (g + 1) * ceil(n / G)))  (SURVEY 8 e)
static void shard_bounds(int64_t n, int g, int G, int64_t* lo, int64_t* hi) {
  const int64_t per = (n + G - 1) / G;
  *lo = std::min(n, (int64_t)g * per);
  *hi = std::min(n, *lo + per);
}

// Multi-device plans: every replica evaluates its block on its own host thread (own streams, own pipeline)
// straight from / into the caller's arrays; sums are added in device order.
static int evaluate_any(ttn_plan* p, CoordSource base, const double* coords, void* out, ttn_opts* opts,
                        const uint8_t* digits = nullptr) {
  if (p->replicas.empty()) return evaluate_impl(p, base, coords, out, opts, digits);
  const int G = (int)p->replicas.size();
  const int NC = p->info.is_complex ? 2 : 1;
  const int64_t npts = base.npts;
  int rc = validate_opts(opts, out);
  if (rc) return rc;
  if (npts < 0) return fail(TTN_ERR_INVALID, "npts < 0");
  const auto wall0 = std::chrono::steady_clock::now();
  std::vector<ttn_opts> ro((size_t)G, *opts);
  std::vector<int> rrc((size_t)G, TTN_OK);
  std::vector<std::string> rerr((size_t)G);
  auto work = [&](int g) {
    int64_t lo, hi;
    shard_bounds(npts, g, G, &lo, &hi);
    CoordSource b = base;
    b.npts = hi - lo;
    if (base.grid) b.first = base.first + lo;
    ttn_opts& o = ro[g];
    if (o.weights) o.weights = o.weights + lo;
    const double* c = coords;
    if (c) c = base.layout == TTN_LAYOUT_AOS ? c + lo * base.n_coords : c + lo;
    const uint8_t* dgt = digits ? digits + (size_t)lo * p->info.n_sites : nullptr;
    void* o_ptr = out ? static_cast<void*>(reinterpret_cast<double*>(out) + lo * NC) : nullptr;
    if (hi == lo) {
      o.kernel_used = p->replicas[g]->info.auto_kernel;
      o.kernel_ms = o.total_ms = 0.f;
      o.n_launches = 0;
      o.n_refined = 0;
      o.staged = 0;
      o.flops_executed = 0.0;
      o.sum_out[0] = o.sum_out[1] = 0.0;
      return;
    }
    rrc[g] = evaluate_impl(p->replicas[g], b, c, o_ptr, &o, dgt, npts);
    if (rrc[g] != TTN_OK) rerr[g] = g_err;
  };
  std::vector<std::thread> th;
  for (int g = 1; g < G; ++g) th.emplace_back(work, g);
  work(0);
  for (auto& t : th) t.join();
  opts->sum_out[0] = opts->sum_out[1] = 0.0;
  opts->kernel_ms = 0.f;
  opts->n_launches = 0;
  opts->n_devices_used = 0;
  opts->staged = 0;
  opts->n_refined = 0;
  opts->flops_executed = 0.0;
  opts->h2d_bytes = opts->d2h_bytes = 0;
  opts->kernel_used = ro[0].kernel_used;
  for (int g = 0; g < G; ++g) {
    if (rrc[g] != TTN_OK && rc == TTN_OK) {
      rc = rrc[g];
      set_error("device " + std::to_string(p->replicas[g]->device) + ": " + rerr[g]);
    }
    opts->sum_out[0] += ro[g].sum_out[0]; // device order => deterministic
    opts->sum_out[1] += ro[g].sum_out[1];
    opts->kernel_ms = std::max(opts->kernel_ms, ro[g].kernel_ms); // the devices run concurrently
    opts->n_launches += ro[g].n_launches;
    opts->n_devices_used += ro[g].n_launches > 0 ? 1 : 0;
    opts->staged |= ro[g].staged;
    opts->n_refined += ro[g].n_refined;
    opts->flops_executed += ro[g].flops_executed;
    opts->h2d_bytes += ro[g].h2d_bytes;
    opts->d2h_bytes += ro[g].d2h_bytes;
  }
  opts->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  return rc;
}

} // namespace ttn

#ifdef TTN_PHASE_CLOCKS
namespace ttn { int debug_phase_clocks(unsigned long long* out8, int reset); }
#endif
#ifdef TTN_TEAM_CLOCKS
namespace ttn { int debug_team_clocks(unsigned long long* out12, int reset); }
#endif
using namespace ttn;

extern "C" {

int ttn_abi_version(void) { return TTN_ABI_VERSION; }

const char* ttn_last_error(void) { return g_err.c_str(); }

int ttn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static int create_single(const ttn_desc* desc, int32_t device, ttn_plan** out) {
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(TTN_ERR_CUDA, "no CUDA device available (libttneval has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return fail(TTN_ERR_INVALID, "device index out of range");
  ttn_plan* p = new (std::nothrow) ttn_plan();
  if (!p) return fail(TTN_ERR_NOMEM, "out of host memory");
  p->device = device;
  if (const char* lws = getenv("LOCAL_WORLD_SIZE")) p->host_peers = std::max(1, atoi(lws));
  int rc = TTN_OK;
  {
    DeviceGuard guard(device);
    do {
      if (!guard.ok) { rc = fail(TTN_ERR_CUDA, "cudaSetDevice failed"); break; }
      cudaDeviceProp prop;
      if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { rc = fail(TTN_ERR_CUDA, "cudaGetDeviceProperties failed"); break; }
      if (prop.major != 10) { rc = fail(TTN_ERR_CUDA, std::string("libttneval is built for sm_100a only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor)); break; }
      p->sm_count = prop.multiProcessorCount;
      bool ok = true;
      for (auto& st : p->streams) {
        ok = ok && cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreate(&st.k0) == cudaSuccess && cudaEventCreate(&st.k1) == cudaSuccess;
      }
      ok = ok && cudaMalloc(&p->d_err, sizeof(int)) == cudaSuccess;
      ok = ok && cudaMalloc(&p->d_sum, sizeof(double) * 2 * kMaxChunks) == cudaSuccess;
      ok = ok && cudaMalloc(&p->d_nsel, sizeof(int32_t) * kMaxChunks) == cudaSuccess;
      ok = ok && cudaEventCreate(&p->t0) == cudaSuccess && cudaEventCreate(&p->t1) == cudaSuccess;
      if (!ok) { rc = fail(TTN_ERR_CUDA, std::string("plan resources: ") + cudaGetErrorString(cudaGetLastError())); break; }
      rc = build_plan(p, desc);
    } while (0);
  }
  if (rc != TTN_OK) {
    const std::string keep = g_err;
    destroy_plan(p);
    g_err = keep;
    return rc;
  }
  p->info.n_devices = 1;
  *out = p;
  return TTN_OK;
}

int ttn_plan_create(const ttn_desc* desc, int32_t device, ttn_plan** out) {
  if (!desc || !out) return fail(TTN_ERR_INVALID, "null argument");
  return create_single(desc, device, out);
}

int ttn_plan_create_multi(const ttn_desc* desc, int32_t n_devices, const int32_t* devices, ttn_plan** out) {
  if (!desc || !out) return fail(TTN_ERR_INVALID, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(TTN_ERR_CUDA, "no CUDA device available (libttneval has no CPU fallback)");
  }
  if (n_devices < 1 || n_devices > ndev) return fail(TTN_ERR_INVALID, "n_devices must be in [1, ttn_device_count()]");
  std::vector<int> devs((size_t)n_devices);
  for (int g = 0; g < n_devices; ++g) {
    devs[g] = devices ? devices[g] : g;
    if (devs[g] < 0 || devs[g] >= ndev) return fail(TTN_ERR_INVALID, "device index out of range");
    for (int h = 0; h < g; ++h)
      if (devs[h] == devs[g]) return fail(TTN_ERR_INVALID, "devices[] lists a device twice");
  }
  ttn_plan* mp = new (std::nothrow) ttn_plan();
  if (!mp) return fail(TTN_ERR_NOMEM, "out of host memory");
  mp->device = devs[0];
  mp->replicas.assign((size_t)n_devices, nullptr);
  // the replicas are built concurrently (a plan with tables takes up to ~1.5 s per device)
  std::vector<int> rrc((size_t)n_devices, TTN_OK);
  std::vector<std::string> rerr((size_t)n_devices);
  auto build = [&](int g) {
    rrc[g] = create_single(desc, devs[g], &mp->replicas[g]);
    if (rrc[g] != TTN_OK) rerr[g] = g_err;
  };
  {
    std::vector<std::thread> th;
    for (int g = 1; g < n_devices; ++g) th.emplace_back(build, g);
    build(0);
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < n_devices; ++g)
    if (rrc[g] != TTN_OK) {
      const int rc = rrc[g];
      const std::string msg = "device " + std::to_string(devs[g]) + ": " + rerr[g];
      destroy_plan(mp);
      return fail(rc, msg);
    }
  // peer access both ways, so that blocks of a device-resident array on another GPU move over NVLink
  for (int g = 0; g < n_devices; ++g) {
    DeviceGuard guard(devs[g]);
    for (int h = 0; h < n_devices; ++h) {
      int can = 0;
      if (h != g && cudaDeviceCanAccessPeer(&can, devs[g], devs[h]) == cudaSuccess && can) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(devs[h], 0);
        if (e != cudaSuccess) cudaGetLastError(); // already enabled (by the caller's framework) is fine
      }
    }
  }
  for (ttn_plan* r : mp->replicas) r->host_peers = std::max(r->host_peers, n_devices);
  mp->info = mp->replicas[0]->info;
  mp->info.n_devices = n_devices;
  mp->sm_count = mp->replicas[0]->sm_count;
  *out = mp;
  return TTN_OK;
}

void ttn_plan_destroy(ttn_plan* plan) { destroy_plan(plan); }

int ttn_plan_info(const ttn_plan* plan, ttn_info* info) {
  if (!plan || !info) return fail(TTN_ERR_INVALID, "null argument");
  *info = plan->info;
  return TTN_OK;
}

int ttn_evaluate(ttn_plan* plan, const double* coords, int64_t npts, int32_t n_coords, int32_t layout,
                 void* out, ttn_opts* opts) {
  if (!plan || !opts) return fail(TTN_ERR_INVALID, "null argument");
  if (n_coords != plan->info.n_coords)
    return fail(TTN_ERR_INVALID, "n_coords does not match the plan (all dimensions must be supplied; the reference throws a missing-key error in project, itensornetworkfunction.jl:90)");
  if (layout != TTN_LAYOUT_AOS && layout != TTN_LAYOUT_SOA) return fail(TTN_ERR_INVALID, "bad layout");
  CoordSource src{};
  src.coords = nullptr;
  src.npts = npts;
  src.n_coords = n_coords;
  src.layout = layout;
  src.grid = 0;
  return evaluate_any(plan, src, coords, out, opts);
}

int ttn_evaluate_grid(ttn_plan* plan, const ttn_grid* grid, void* out_or_null, ttn_opts* opts) {
  if (!plan || !opts || !grid) return fail(TTN_ERR_INVALID, "null argument");
  if (grid->n_coords != plan->info.n_coords) return fail(TTN_ERR_INVALID, "grid.n_coords does not match the plan");
  CoordSource src{};
  src.npts = grid->npts;
  src.n_coords = grid->n_coords;
  src.layout = TTN_LAYOUT_AOS;
  src.grid = 1;
  src.first = grid->first;
  int64_t total = 1;
  for (int c = 0; c < grid->n_coords; ++c) {
    if (grid->count[c] < 1) return fail(TTN_ERR_INVALID, "grid.count must be >= 1");
    if (!(grid->step[c] >= 0.0)) return fail(TTN_ERR_DOMAIN, "grid.step must be >= 0");
    src.step[c] = grid->step[c];
    src.count[c] = grid->count[c];
    total *= grid->count[c];
  }
  if (grid->first < 0 || grid->npts < 0 || grid->first + grid->npts > total)
    return fail(TTN_ERR_INVALID, "grid range exceeds the number of grid points");
  return evaluate_any(plan, src, nullptr, out_or_null, opts);
}

int ttn_evaluate_indices(ttn_plan* plan, const uint8_t* index_values, int64_t npts, void* out, ttn_opts* opts) {
  if (!plan || !opts) return fail(TTN_ERR_INVALID, "null argument");
  if (npts > 0 && !index_values) return fail(TTN_ERR_INVALID, "index_values is NULL");
  if (plan->info.n_sites == 0) return fail(TTN_ERR_INVALID, "the network has no site indices");
  CoordSource src{};
  src.npts = npts;
  src.n_coords = plan->info.n_coords;
  src.layout = TTN_LAYOUT_AOS;
  return evaluate_any(plan, src, nullptr, out, opts, index_values);
}

int ttn_digits(ttn_plan* plan, const double* coords, int64_t npts, int32_t n_coords, int32_t layout,
               uint8_t* digits_out, ttn_opts* opts) {
  if (!plan || !opts || !digits_out) return fail(TTN_ERR_INVALID, "null argument");
  if (!plan->replicas.empty()) plan = plan->replicas[0]; // integer work of a few bytes per point: one device
  if (n_coords != plan->info.n_coords) return fail(TTN_ERR_INVALID, "n_coords does not match the plan");
  if (layout != TTN_LAYOUT_AOS && layout != TTN_LAYOUT_SOA) return fail(TTN_ERR_INVALID, "bad layout");
  if (npts < 0) return fail(TTN_ERR_INVALID, "npts < 0");
  opts->n_launches = 0;
  if (npts == 0) return TTN_OK;
  if (!coords) return fail(TTN_ERR_INVALID, "coords is NULL");
  std::lock_guard<std::mutex> lock(plan->mu);
  DeviceGuard guard(plan->device);
  if (!guard.ok) return fail(TTN_ERR_CUDA, "cudaSetDevice failed");
  Stream& st = plan->streams[0];
  const int ns = std::max(plan->info.n_sites, 1);
  const bool direct_c = classify(coords, opts->coords_mem, plan->device, TTN_STAGE_OFF, 0) == SPACE_DIRECT;
  const bool direct_o = classify(digits_out, opts->out_mem, plan->device, TTN_STAGE_OFF, 0) == SPACE_DIRECT;
  TTN_CUDA(cudaMemsetAsync(plan->d_err, 0, sizeof(int), st.s));
  // chunks through the stream's grow-only buffers (no per-call allocation once warm)
  const int64_t chunk = (direct_c && direct_o) ? npts : std::min<int64_t>(npts, (int64_t)1 << 21);
  if (direct_c && layout == TTN_LAYOUT_SOA && chunk != npts)
    return fail(TTN_ERR_UNSUPPORTED, "SOA device coordinates with a host digit buffer are not supported; use AOS");
  int rc = TTN_OK;
  if (!direct_c) rc = ensure_stream_buffers(plan, st, chunk, true, false);
  if (rc == TTN_OK && !direct_o && st.digits_cap < (size_t)chunk * ns) {
    if (st.d_digits) cudaFree(st.d_digits);
    st.d_digits = nullptr;
    st.digits_cap = 0;
    if (cudaMalloc(&st.d_digits, (size_t)chunk * ns) != cudaSuccess) rc = fail(TTN_ERR_NOMEM, "out of device memory (digit buffer)");
    else st.digits_cap = (size_t)chunk * ns;
  }
  for (int64_t first = 0; first < npts && rc == TTN_OK; first += chunk) {
    const int64_t m = std::min(chunk, npts - first);
    CoordSource src{};
    src.npts = m;
    src.n_coords = n_coords;
    src.layout = layout;
    if (direct_c) {
      src.coords = layout == TTN_LAYOUT_AOS ? coords + first * n_coords : coords;
    } else {
      if (layout == TTN_LAYOUT_AOS) {
        cudaMemcpyAsync(st.d_coords, coords + first * n_coords, sizeof(double) * (size_t)m * n_coords, cudaMemcpyDefault, st.s);
      } else {
        for (int c = 0; c < n_coords; ++c)
          cudaMemcpyAsync(st.d_coords + (int64_t)c * m, coords + (int64_t)c * npts + first, sizeof(double) * (size_t)m,
                          cudaMemcpyDefault, st.s);
      }
      src.coords = st.d_coords;
    }
    uint8_t* d_o = direct_o ? digits_out + (size_t)first * plan->info.n_sites : st.d_digits;
    rc = launch_digits(plan, src, d_o, st.s);
    opts->n_launches += 1;
    if (rc == TTN_OK && !direct_o)
      cudaMemcpyAsync(digits_out + (size_t)first * plan->info.n_sites, st.d_digits, (size_t)m * plan->info.n_sites,
                      cudaMemcpyDefault, st.s);
    if (!direct_c || !direct_o) cudaStreamSynchronize(st.s); // the single buffer pair is reused by the next chunk
  }
  const cudaError_t e = cudaStreamSynchronize(st.s);
  int herr = 0;
  cudaMemcpy(&herr, plan->d_err, sizeof(int), cudaMemcpyDeviceToHost);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(TTN_ERR_CUDA, std::string("digits kernel: ") + cudaGetErrorString(e));
  if (herr) return fail(TTN_ERR_DOMAIN, "a coordinate is negative or NaN");
  return TTN_OK;
}

int ttn_host_register(void* ptr, uint64_t bytes) {
  if (!ptr || bytes == 0) return fail(TTN_ERR_INVALID, "null / empty range");
  const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return TTN_OK;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(TTN_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
  }
  return TTN_OK;
}

int ttn_host_unregister(void* ptr) {
  if (!ptr) return fail(TTN_ERR_INVALID, "null pointer");
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(TTN_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
  }
  return TTN_OK;
}

#ifdef TTN_PHASE_CLOCKS
int ttn_debug_phase_clocks(unsigned long long* out8, int reset) { return ttn::debug_phase_clocks(out8, reset); }
#endif
#ifdef TTN_TEAM_CLOCKS
int ttn_debug_team_clocks(unsigned long long* out12, int reset) { return ttn::debug_team_clocks(out12, reset); }
#endif

/* Test hook, not part of include/ttneval.h (no CUDA calls): the table kernel's plan-time image of a
 * description, see debug_table_image in k_chain_table.cu. */
int ttn_debug_table_image(const ttn_desc* desc, int32_t budget_kb, int32_t* meta, double* image, int64_t image_cap,
                          int32_t* site_bitpos) {
  if (!desc || !meta || !image || !site_bitpos) return fail(TTN_ERR_INVALID, "null argument");
  return debug_table_image(desc, budget_kb, meta, image, image_cap, site_bitpos);
}

/* Test hook, not part of include/ttneval.h: the packed slice stream — i.e. the DIGITS — that the kernels with a
 * fused K1 really use (team-sorted DMMA kernel: run fast path floor(x 2^L); table kernel: 32-bit saturating
 * conversion, bit-deposit network).  One launch of `kernel` (TTN_KERNEL_DMMA / TTN_KERNEL_TABLE) over npts AoS host
 * points; words_out[2p], [2p+1] = the 128-bit stream of point p; site_bit[s] = stream bit of site index s of the

 * description (its digit = the field of ceil(log2(dim)) bits starting there).  tests/test_gpu_round2.py compares these integers with the CPU greedy loop. */
int ttn_debug_slice_stream(ttn_plan* plan, const double* coords, int64_t npts, int32_t kernel, uint64_t* words_out,
                           int32_t* site_bit, int32_t* site_stride, int32_t* site_fbits) {
  if (!plan || !coords || !words_out || !site_bit || !site_stride || !site_fbits || npts <= 0)
    return fail(TTN_ERR_INVALID, "null / empty argument");
  if (!plan->replicas.empty()) plan = plan->replicas[0];
  const DigitTable* dt = nullptr;
  if (kernel == TTN_KERNEL_DMMA && plan->cmma_ok && chain_team_applicable(plan)) dt = &plan->digits_mma;
  if (kernel == TTN_KERNEL_TABLE && plan->ctab_ok) dt = &plan->digits_tab;
  if (!dt) return fail(TTN_ERR_UNSUPPORTED, "slice-stream dump: the team-sorted DMMA kernel or the table kernel must apply");
  std::lock_guard<std::mutex> lock(plan->mu);
  DeviceGuard guard(plan->device);
  if (!guard.ok) return fail(TTN_ERR_CUDA, "cudaSetDevice failed");
  const int nc = plan->info.n_coords, ns = plan->info.n_sites, NC = plan->info.is_complex ? 2 : 1;
  Stream& st = plan->streams[0];
  int rc = ensure_stream_buffers(plan, st, npts, true, true);
  if (rc) return rc;
  unsigned long long* d_words = nullptr;
  TTN_CUDA(cudaMalloc(&d_words, sizeof(unsigned long long) * 2 * (size_t)npts));
  cudaMemsetAsync(plan->d_err, 0, sizeof(int), st.s);
  cudaMemcpyAsync(st.d_coords, coords, sizeof(double) * (size_t)npts * nc, cudaMemcpyHostToDevice, st.s);
  CoordSource src{};
  src.coords = st.d_coords;
  src.npts = npts;
  src.n_coords = nc;
  src.layout = TTN_LAYOUT_AOS;
  src.dbg_stream = d_words;
  int n_partial = 0, extra = 0;
  rc = run_kernel(plan, kernel, st, src, st.d_out, nullptr, &n_partial, &extra);
  (void)NC;
  if (rc == TTN_OK) {
    cudaMemcpyAsync(words_out, d_words, sizeof(unsigned long long) * 2 * (size_t)npts, cudaMemcpyDeviceToHost, st.s);
    const cudaError_t e = cudaStreamSynchronize(st.s);
    if (e != cudaSuccess) rc = fail(TTN_ERR_CUDA, std::string("slice-stream dump: ") + cudaGetErrorString(e));
  }
  cudaFree(d_words);
  if (rc) return rc;
  std::vector<DigitEntry> ent((size_t)std::max(ns, 1));
  TTN_CUDA(cudaMemcpy(ent.data(), dt->entries, sizeof(DigitEntry) * (size_t)ns, cudaMemcpyDeviceToHost));
  // digit of site s = (field / stride) % dim(s), field = site_fbits[s] stream bits starting at bit site_bit[s]
  for (int i = 0; i < ns; ++i) {
    site_bit[ent[i].site] = ent[i].word * 64 + ent[i].shift;
    site_stride[ent[i].site] = ent[i].stride;
    site_fbits[ent[i].site] = kernel == TTN_KERNEL_DMMA ? plan->cmma_site_fbits[ent[i].site] : plan->ctab_bits0;
  }
  return TTN_OK;
}

/* Test hooks without CUDA calls (CPU test suite), not part of include/ttneval.h.
 * ttn_debug_pack_coords: the host-side quantisation of run-path coordinates exactly as the staging threads do it
 *   (pack_coords): q[p * nc + c] = min(floor(x 2^L[c]), 2^L[c] - 1); returns 1 if a coordinate is negative / NaN, else 0.
 * ttn_debug_binarize: the plan-time splitting of vertices with more than two children (binarize_desc).  Returns the
 *   new vertex count (0: the description is left as it is); fills parent / link_dim / tensor_ptr (cap entries, + 1 for
 *   tensor_ptr) and copies the new tensor blob (tensor_cap doubles) when the buffers are large enough, else returns -1. */
int ttn_debug_pack_coords(const double* coords, int64_t npts, int32_t nc, const int32_t* L, uint32_t* q_out) {
  if (!coords || !L || !q_out || nc < 1 || nc > TTN_MAX_COORDS || npts < 0) return -1;
  PackParams pp;
  pp.nc = nc;
  for (int c = 0; c < nc; ++c) {
    if (L[c] < 1 || L[c] > 32) return -1;
    pp.scale[c] = std::ldexp(1.0, L[c]);
    pp.qmax[c] = L[c] >= 32 ? 0xffffffffu : ((1u << L[c]) - 1u);
  }
  return pack_coords(q_out, coords, (size_t)npts, pp) ? 1 : 0;
}

int ttn_debug_binarize(const ttn_desc* desc, int32_t cap, int32_t* parent, int32_t* link_dim, int64_t* tensor_ptr,
                       double* tensors, int64_t tensor_cap) {
  if (!desc || !parent || !link_dim || !tensor_ptr || !tensors) return -1;
  BinDesc b;
  if (!binarize_desc(desc, &b)) return 0;
  const int n = b.d.n_vertices;
  if (n > cap || (int64_t)b.tensors.size() > tensor_cap) return -1;
  std::copy(b.parent.begin(), b.parent.end(), parent);
  std::copy(b.link_dim.begin(), b.link_dim.end(), link_dim);
  std::copy(b.tensor_ptr.begin(), b.tensor_ptr.end(), tensor_ptr);
  std::copy(b.tensors.begin(), b.tensors.end(), tensors);
  return n;
}

int ttn_measure_fp64_peak(int32_t device, double* dfma_tflops, double* dmma_tflops) {
  if (!dfma_tflops || !dmma_tflops) return fail(TTN_ERR_INVALID, "null argument");
  return measure_fp64_peak(device, dfma_tflops, dmma_tflops);
}

} // extern "C"
