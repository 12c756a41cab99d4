// C ABI of libttneval.so (include/ttneval.h): plan construction, chunked/pipelined evaluation,
// error reporting.  Everything here is host code around the kernels in k_*.cu.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <new>
#include <numeric>

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

static int build_plan(ttn_plan* p, const ttn_desc* d) {
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
  else if (p->tgemm_ok && max_link >= 8) I.auto_kernel = TTN_KERNEL_TREE; // measured cross-over (scripts/tree_small_chi.py)
  else I.auto_kernel = TTN_KERNEL_GENERIC;
  I.device = p->device;
  I.kernels_available = (1 << TTN_KERNEL_GENERIC) | (p->chain_ok ? (1 << TTN_KERNEL_CHAIN) : 0) |
                        (p->cmma_ok ? (1 << TTN_KERNEL_DMMA) : 0) | (p->cgemm_ok ? (1 << TTN_KERNEL_GEMM) : 0) |
                        (p->tgemm_ok ? (1 << TTN_KERNEL_TREE) : 0) | (p->gshare_ok ? (1 << TTN_KERNEL_GRID) : 0) |
                        (p->ctab_ok ? (1 << TTN_KERNEL_TABLE) : 0);
  I.flops_per_point = (d->is_complex ? 8.0 : 2.0) * macs;
  I.bytes_per_point = 8.0 * d->n_coords + (d->is_complex ? 16.0 : 8.0);
  I.tensor_bytes = d->tensor_ptr[n] * NC * 8;
  return TTN_OK;
}

static void destroy_plan(ttn_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
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
    if (st.k0) cudaEventDestroy(st.k0);
    if (st.k1) cudaEventDestroy(st.k1);
    if (st.s) cudaStreamDestroy(st.s);
  }
  for (void* a : p->allocs) cudaFree(a);
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
  const size_t pc = (size_t)2 * (p->sm_count * 8 + 8);
  if (st.partial_cap < pc) {
    if (st.d_partial) cudaFree(st.d_partial);
    TTN_CUDA(cudaMalloc(&st.d_partial, pc * sizeof(double)));
    st.partial_cap = pc;
  }
  return TTN_OK;
}

// Shared driver of ttn_evaluate / ttn_evaluate_grid.
static int evaluate_impl(ttn_plan* p, CoordSource base, const double* coords, void* out, ttn_opts* opts,
                         const uint8_t* digits = nullptr) {
  std::lock_guard<std::mutex> lock(p->mu);
  TTN_CUDA(cudaSetDevice(p->device));
  const auto wall0 = std::chrono::steady_clock::now();
  const int NC = p->info.is_complex ? 2 : 1;
  const int64_t npts = base.npts;
  // full dyadic grid of a binary chain: prefix-shared expansion (one pass over the whole grid)
  if (base.grid && (opts->kernel == TTN_KERNEL_AUTO || opts->kernel == TTN_KERNEL_GRID) && grid_share_applicable(p, base) &&
      !(opts->reduce_sum == TTN_REDUCE_WEIGHTED && opts->weights_mem == TTN_MEM_HOST)) {
    opts->sum_out[0] = opts->sum_out[1] = 0.0;
    opts->kernel_used = TTN_KERNEL_GRID;
    opts->n_launches = 0;
    const bool do_sum_g = opts->reduce_sum != TTN_REDUCE_NONE;
    if (!out && !do_sum_g) return fail(TTN_ERR_INVALID, "out is NULL and reduce_sum is 0: nothing to compute");
    Stream& st = p->streams[0];
    int rc = ensure_stream_buffers(p, st, 1, false, false);
    if (rc) return rc;
    double* d_out = reinterpret_cast<double*>(out);
    double* tmp = nullptr;
    const bool out_host_g = out && opts->out_mem == TTN_MEM_HOST;
    if (out_host_g) {
      TTN_CUDA(cudaMalloc(&tmp, sizeof(double) * (size_t)base.npts * NC));
      d_out = tmp;
    }
    CoordSource src = base;
    src.reduce_mode = opts->reduce_sum;
    src.weights = opts->reduce_sum == TTN_REDUCE_WEIGHTED ? opts->weights : nullptr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st.s);
    int n_partial = 0;
    double flops = 0.0;
    rc = launch_grid_share(p, st, src, d_out, do_sum_g ? st.d_partial : nullptr, &n_partial, st.s, &opts->n_launches, &flops);
    if (rc == TTN_OK && do_sum_g) {
      rc = launch_sum_partials(p, st.d_partial, n_partial, NC, p->d_sum, st.s);
      opts->n_launches += 1;
    }
    cudaEventRecord(e1, st.s);
    if (rc == TTN_OK && out_host_g)
      cudaMemcpyAsync(out, tmp, sizeof(double) * (size_t)base.npts * NC, cudaMemcpyDeviceToHost, st.s);
    const cudaError_t e = cudaStreamSynchronize(st.s);
    if (rc == TTN_OK && e != cudaSuccess) rc = fail(TTN_ERR_CUDA, std::string("grid kernel: ") + cudaGetErrorString(e));
    if (rc == TTN_OK) {
      cudaEventElapsedTime(&opts->kernel_ms, e0, e1);
      opts->flops_executed = flops;
      if (do_sum_g) {
        double hs[2];
        cudaMemcpy(hs, p->d_sum, sizeof(hs), cudaMemcpyDeviceToHost);
        opts->sum_out[0] = hs[0];
        opts->sum_out[1] = hs[1];
      }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (tmp) cudaFree(tmp);
    opts->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    return rc;
  }
  if (opts->kernel == TTN_KERNEL_GRID)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_GRID needs ttn_evaluate_grid on the FULL dyadic grid (count = 2^L, step = 2^-L, first = 0) of a chain with one binary site index per vertex and width <= 32");
  int kernel = opts->kernel == TTN_KERNEL_AUTO ? p->info.auto_kernel : opts->kernel;
  if (kernel == TTN_KERNEL_CHAIN && !p->chain_ok)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_CHAIN: network is not a chain with chi <= 32 (real) / 16 (complex) and <= 4 slices per vertex");
  if (kernel == TTN_KERNEL_TABLE && !p->ctab_ok)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_TABLE: network is not a chain of binary site indices with chi <= 4 (real) / 2 (complex), <= 2 site indices per vertex and <= 128 slice bits");
  if (kernel == TTN_KERNEL_TREE && !p->tgemm_ok)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_TREE: network is not a real tree with <= 2 children per vertex, chi <= 64 and <= 8 slices per vertex");
  if (kernel == TTN_KERNEL_GEMM && !p->cgemm_ok)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_GEMM: network is not a chain with 32 < (real-embedded) width <= 256 and <= 8 slices per vertex");
  if (kernel == TTN_KERNEL_DMMA && !p->cmma_ok)
    return fail(TTN_ERR_UNSUPPORTED, "TTN_KERNEL_DMMA: network is not a chain with chi <= 32 (real) / 16 (complex), <= 4 slices per vertex and <= 128 slice bits");
  const bool coords_host = !base.grid && opts->coords_mem == TTN_MEM_HOST;
  const int n_sites = std::max(p->info.n_sites, 1);
  const bool out_host = out != nullptr && opts->out_mem == TTN_MEM_HOST;
  const bool do_sum = opts->reduce_sum != TTN_REDUCE_NONE;
  if (opts->reduce_sum < 0 || opts->reduce_sum > TTN_REDUCE_WEIGHTED) return fail(TTN_ERR_INVALID, "bad reduce mode");
  if (opts->reduce_sum == TTN_REDUCE_WEIGHTED && !opts->weights) return fail(TTN_ERR_INVALID, "TTN_REDUCE_WEIGHTED needs opts->weights");
  const bool weights_host = opts->reduce_sum == TTN_REDUCE_WEIGHTED && opts->weights_mem == TTN_MEM_HOST;
  opts->sum_out[0] = opts->sum_out[1] = 0.0;
  opts->kernel_ms = opts->total_ms = 0.f;
  opts->kernel_used = kernel;
  opts->n_launches = 0;
  if (npts < 0) return fail(TTN_ERR_INVALID, "npts < 0");
  if (!out && !do_sum) return fail(TTN_ERR_INVALID, "out is NULL and reduce_sum is 0: nothing to compute");
  if (npts == 0) return TTN_OK;
  if (!base.grid && !coords && !digits) return fail(TTN_ERR_INVALID, "coords is NULL");

  TTN_CUDA(cudaMemsetAsync(p->d_err, 0, sizeof(int), p->streams[0].s));
  TTN_CUDA(cudaStreamSynchronize(p->streams[0].s));

  int64_t chunk = npts;
  int n_chunks = 1;
  if (coords_host || out_host || weights_host) {
    chunk = opts->chunk_points > 0 ? opts->chunk_points : (int64_t)1 << 21; // measured best on PCIe 5 x16 (scripts/e2e_probe.py)
    chunk = std::min(chunk, npts);
    n_chunks = (int)((npts + chunk - 1) / chunk);
    if (n_chunks > kMaxChunks) {
      chunk = (npts + kMaxChunks - 1) / kMaxChunks;
      n_chunks = (int)((npts + chunk - 1) / chunk);
    }
  }
  std::vector<cudaEvent_t> ev(2 * (size_t)n_chunks, nullptr);
  auto cleanup_events = [&]() {
    for (auto e : ev)
      if (e) cudaEventDestroy(e);
  };
  int rc = TTN_OK;
  const int n_streams = (coords_host || out_host || weights_host) ? 3 : 1;
  for (int ci = 0; ci < n_chunks && rc == TTN_OK; ++ci) {
    Stream& st = p->streams[ci % n_streams];
    const int64_t first = (int64_t)ci * chunk;
    const int64_t m = std::min(chunk, npts - first);
    rc = ensure_stream_buffers(p, st, (coords_host || out_host || weights_host) ? chunk : 1, coords_host, out_host);
    if (rc) break;
    CoordSource src = base;
    src.npts = m;
    src.reduce_mode = opts->reduce_sum;
    src.weights = nullptr;
    if (opts->reduce_sum == TTN_REDUCE_WEIGHTED) {
      if (weights_host) {
        cudaMemcpyAsync(st.d_weights, opts->weights + first, sizeof(double) * m, cudaMemcpyHostToDevice, st.s);
        src.weights = st.d_weights;
      } else {
        src.weights = opts->weights + first;
      }
    }
    if (digits) {
      if (coords_host) {
        if (!st.d_digits || st.digits_cap < (size_t)chunk * n_sites) {
          if (st.d_digits) cudaFree(st.d_digits);
          st.d_digits = nullptr;
          TTN_CUDA(cudaMalloc(&st.d_digits, (size_t)chunk * n_sites));
          st.digits_cap = (size_t)chunk * n_sites;
        }
        cudaMemcpyAsync(st.d_digits, digits + (size_t)first * p->info.n_sites, (size_t)m * p->info.n_sites,
                        cudaMemcpyHostToDevice, st.s);
        src.digits = st.d_digits;
      } else {
        src.digits = digits + (size_t)first * p->info.n_sites;
      }
    } else if (base.grid) {
      src.first = base.first + first;
    } else if (coords_host) {
      if (base.layout == TTN_LAYOUT_AOS) {
        cudaMemcpyAsync(st.d_coords, coords + first * base.n_coords, sizeof(double) * m * base.n_coords,
                        cudaMemcpyHostToDevice, st.s);
      } else {
        for (int c = 0; c < base.n_coords; ++c)
          cudaMemcpyAsync(st.d_coords + (int64_t)c * m, coords + (int64_t)c * npts + first, sizeof(double) * m,
                          cudaMemcpyHostToDevice, st.s);
      }
      src.coords = st.d_coords;
    } else {
      // device coordinates: a chunk is a sub-range of the caller's array
      if (base.layout == TTN_LAYOUT_AOS) {
        src.coords = coords + first * base.n_coords;
      } else if (n_chunks == 1) {
        src.coords = coords;
      } else {
        cleanup_events();
        return fail(TTN_ERR_UNSUPPORTED, "SOA device coordinates with a host output buffer are not supported; use AOS");
      }
    }
    double* d_out = nullptr;
    if (out) d_out = out_host ? st.d_out : reinterpret_cast<double*>(out) + first * NC;
    cudaEventCreate(&ev[2 * ci]);
    cudaEventCreate(&ev[2 * ci + 1]);
    cudaEventRecord(ev[2 * ci], st.s);
    int n_partial = 0;
    rc = run_kernel(p, kernel, st, src, d_out, do_sum ? st.d_partial : nullptr, &n_partial, &opts->n_launches);
    if (rc) break;
    opts->n_launches += 1;
    if (do_sum) {
      rc = launch_sum_partials(p, st.d_partial, n_partial, NC, p->d_sum + 2 * ci, st.s);
      if (rc) break;
      opts->n_launches += 1;
    }
    cudaEventRecord(ev[2 * ci + 1], st.s);
    if (out_host)
      cudaMemcpyAsync(reinterpret_cast<double*>(out) + first * NC, st.d_out, sizeof(double) * m * NC,
                      cudaMemcpyDeviceToHost, st.s);
  }
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
    opts->flops_executed = ((kernel == TTN_KERNEL_DMMA && p->cmma.merged)   ? p->cmma_flops_exec
                            : (kernel == TTN_KERNEL_GEMM && p->cgemm.merged) ? p->cgemm_flops_exec
                            : (kernel == TTN_KERNEL_TREE && !p->tg_tab.empty()) ? p->tgemm_flops_exec
                            : kernel == TTN_KERNEL_TABLE ? p->ctab_flops_exec
                                                                             : p->info.flops_per_point) *
                           (double)npts;
    int herr = 0;
    cudaMemcpy(&herr, p->d_err, sizeof(int), cudaMemcpyDeviceToHost);
    if (herr & 2)
      rc = fail(TTN_ERR_INVALID, "an index value is out of range for its site index");
    else if (herr)
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
  }
  cleanup_events();
  opts->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  return rc;
}

} // namespace ttn

#ifdef TTN_PHASE_CLOCKS
namespace ttn { int debug_phase_clocks(unsigned long long* out8, int reset); }
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

int ttn_plan_create(const ttn_desc* desc, int32_t device, ttn_plan** out) {
  if (!desc || !out) return fail(TTN_ERR_INVALID, "null argument");
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
  int rc = TTN_OK;
  do {
    if (cudaSetDevice(device) != cudaSuccess) { rc = fail(TTN_ERR_CUDA, "cudaSetDevice failed"); break; }
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
    ok = ok && cudaEventCreate(&p->t0) == cudaSuccess && cudaEventCreate(&p->t1) == cudaSuccess;
    if (!ok) { rc = fail(TTN_ERR_CUDA, std::string("plan resources: ") + cudaGetErrorString(cudaGetLastError())); break; }
    rc = build_plan(p, desc);
  } while (0);
  if (rc != TTN_OK) {
    const std::string keep = g_err;
    destroy_plan(p);
    g_err = keep;
    return rc;
  }
  *out = p;
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
  return evaluate_impl(plan, src, coords, out, opts);
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
  return evaluate_impl(plan, src, nullptr, out_or_null, opts);
}

int ttn_evaluate_indices(ttn_plan* plan, const uint8_t* index_values, int64_t npts, void* out, ttn_opts* opts) {
  if (!plan || !opts) return fail(TTN_ERR_INVALID, "null argument");
  if (npts > 0 && !index_values) return fail(TTN_ERR_INVALID, "index_values is NULL");
  if (plan->info.n_sites == 0) return fail(TTN_ERR_INVALID, "the network has no site indices");
  CoordSource src{};
  src.npts = npts;
  src.n_coords = plan->info.n_coords;
  src.layout = TTN_LAYOUT_AOS;
  return evaluate_impl(plan, src, nullptr, out, opts, index_values);
}

int ttn_digits(ttn_plan* plan, const double* coords, int64_t npts, int32_t n_coords, int32_t layout,
               uint8_t* digits_out, ttn_opts* opts) {
  if (!plan || !opts || !digits_out) return fail(TTN_ERR_INVALID, "null argument");
  if (n_coords != plan->info.n_coords) return fail(TTN_ERR_INVALID, "n_coords does not match the plan");
  if (npts == 0) return TTN_OK;
  if (!coords) return fail(TTN_ERR_INVALID, "coords is NULL");
  std::lock_guard<std::mutex> lock(plan->mu);
  TTN_CUDA(cudaSetDevice(plan->device));
  Stream& st = plan->streams[0];
  const int ns = std::max(plan->info.n_sites, 1);
  const bool host_c = opts->coords_mem == TTN_MEM_HOST, host_o = opts->out_mem == TTN_MEM_HOST;
  double* d_c = nullptr;
  uint8_t* d_o = nullptr;
  TTN_CUDA(cudaMemsetAsync(plan->d_err, 0, sizeof(int), st.s));
  if (host_c) {
    TTN_CUDA(cudaMalloc(&d_c, sizeof(double) * (size_t)npts * std::max(n_coords, 1)));
    TTN_CUDA(cudaMemcpyAsync(d_c, coords, sizeof(double) * (size_t)npts * n_coords, cudaMemcpyHostToDevice, st.s));
  }
  if (host_o) TTN_CUDA(cudaMalloc(&d_o, (size_t)npts * ns));
  CoordSource src{};
  src.coords = host_c ? d_c : coords;
  src.npts = npts;
  src.n_coords = n_coords;
  src.layout = layout;
  int rc = launch_digits(plan, src, host_o ? d_o : digits_out, st.s);
  opts->n_launches = 1;
  if (rc == TTN_OK && host_o)
    cudaMemcpyAsync(digits_out, d_o, (size_t)npts * plan->info.n_sites, cudaMemcpyDeviceToHost, st.s);
  cudaError_t e = cudaStreamSynchronize(st.s);
  int herr = 0;
  cudaMemcpy(&herr, plan->d_err, sizeof(int), cudaMemcpyDeviceToHost);
  if (d_c) cudaFree(d_c);
  if (d_o) cudaFree(d_o);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(TTN_ERR_CUDA, std::string("digits kernel: ") + cudaGetErrorString(e));
  if (herr) return fail(TTN_ERR_DOMAIN, "a coordinate is negative or NaN");
  return TTN_OK;
}

#ifdef TTN_PHASE_CLOCKS
int ttn_debug_phase_clocks(unsigned long long* out8, int reset) { return ttn::debug_phase_clocks(out8, reset); }
#endif

/* Test hook, not part of include/ttneval.h (no CUDA calls): the table kernel's plan-time image of a
 * description, see debug_table_image in k_chain_table.cu. */
int ttn_debug_table_image(const ttn_desc* desc, int32_t budget_kb, int32_t* meta, double* image, int64_t image_cap,
                          int32_t* site_bitpos) {
  if (!desc || !meta || !image || !site_bitpos) return fail(TTN_ERR_INVALID, "null argument");
  return debug_table_image(desc, budget_kb, meta, image, image_cap, site_bitpos);
}

int ttn_measure_fp64_peak(int32_t device, double* dfma_tflops, double* dmma_tflops) {
  if (!dfma_tflops || !dmma_tflops) return fail(TTN_ERR_INVALID, "null argument");
  return measure_fp64_peak(device, dfma_tflops, dmma_tflops);
}

} // extern "C"
