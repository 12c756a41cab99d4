/*
 * ttneval.h — C ABI of libttneval.so: batched evaluation of a quantics tree
 * tensor-network function on B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE path of ITensorNumericalAnalysis.jl
 * (reference at /root/reference, v0.2.1): the caller loop around
 *     evaluate(fitn::ITensorNetworkFunction, xs::Vector, dims)   src/itensornetworkfunction.jl:96-106
 * i.e.  calculate_ind_values  (src/IndexMaps/realindexmap.jl:67-76,
 *                              src/IndexMaps/complexindexmap.jl:116-132,
 *                              greedy loop src/IndexMaps/abstractindexmap.jl:121-138)
 *    -> project               (src/itensornetworkfunction.jl:84-94)
 *    -> scalar(tn; alg="bp")  (src/itensornetworkfunction.jl:105; ITensorNetworks.jl, un-vendored)
 * executed for MANY points per call.  The reference has no FFI for this path
 * (it is 100% Julia); the Julia method that binds these symbols with `ccall`
 * is shown in INTEGRATION.md and shipped in
 * itensornumericalanalysis.jl_b200/julia/TTNEvalB200.jl.
 *
 * Conventions
 *  - plain C, no C++ types, no torch types; all pointers are caller-owned for
 *    the duration of the (synchronous) call; the plan copies everything it
 *    needs at ttn_plan_create time.
 *  - every function returning int returns 0 (TTN_OK) on success, else a
 *    TTN_ERR_* code; the message is available from ttn_last_error() (thread
 *    local).  No C++ exception crosses this boundary.
 *  - there is NO CPU fallback behind this ABI: if no CUDA device is usable,
 *    calls fail with TTN_ERR_CUDA.
 */
#ifndef TTNEVAL_H
#define TTNEVAL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTN_ABI_VERSION 3

/* ---- error codes ------------------------------------------------------ */
enum {
  TTN_OK = 0,
  TTN_ERR_INVALID = 1, /* malformed description / arguments */
  TTN_ERR_DOMAIN = 2,  /* a coordinate is negative or NaN: the reference's greedy loop
                          (abstractindexmap.jl:121-138) never terminates on such input; we
                          return an error instead (documented deviation) */
  TTN_ERR_CUDA = 3,    /* CUDA runtime failure, or no device */
  TTN_ERR_UNSUPPORTED = 4,
  TTN_ERR_NOMEM = 5
};

/* ---- coordinate layouts ------------------------------------------------ */
enum {
  TTN_LAYOUT_AOS = 0, /* coords[p * n_coords + c]  (Julia D x Npts Matrix, column major;
                         Vector{<:Vector} after reduce(hcat, ...)) */
  TTN_LAYOUT_SOA = 1  /* coords[c * npts + p] */
};

/* ---- memory spaces for coords / out ------------------------------------ */
enum {
  TTN_MEM_HOST = 0,  /* host pointers; H2D / D2H copies happen inside the call */
  TTN_MEM_DEVICE = 1 /* device pointers on the plan's device; no copies */
};

/* ---- pageable host buffers (ABI 3).  cudaMemcpyAsync from pageable memory is staged by the driver on one thread
 * and collapses the H2D / kernel / D2H pipeline, so ttn_evaluate copies pageable buffers through an internal
 * pinned ring with several host threads (TTN_HOST_THREADS, default min(16, cores)).  Buffers the caller pinned
 * (cudaHostAlloc / ttn_host_register) are used in place.  A registration is never cached across calls: a stale
 * one would outlive a freed Julia / numpy array. */
enum {
  TTN_STAGE_AUTO = 0, /* pinned buffers in place, pageable buffers through the staging ring; and, where the digits of
                         every coordinate are the bits of floor(x 2^L) (binary digits 1..L on consecutive chain
                         vertices, L <= 32: the K1 run path), PAGEABLE coordinates are QUANTISED to that L-bit grid
                         index by the staging threads on their way into the ring — 4 bytes per coordinate cross PCIe
                         instead of 8, bit-exact (the evaluation depends on nothing else); opts->staged bit 2 reports it.
                         PINNED coordinates of such a chain (>= 2^20 points) go HYBRID when this GPU has the host to itself
                         (a single-device plan in a process without LOCAL_WORLD_SIZE > 1): two of every three chunks are
                         quantised by the host threads while the copy engines read the third in place (10.7 B per 2-D point
                         over PCIe instead of 16); the environment variable TTN_HOST_QUANT overrides (0 never, 1 pageable
                         arrays only, 2 every chunk, 3 / 4 / 5 / 6: 1 of 2, 1 of 3, 2 of 3, 3 of 4 chunks of a pinned array) */
  TTN_STAGE_OFF = 1,  /* hand every host pointer straight to cudaMemcpyAsync */
  TTN_STAGE_COPY = 2  /* like AUTO, but coordinates always travel as doubles */
};

/* ---- accuracy modes (ABI 3) */
enum {
  TTN_ACCURACY_FP64 = 0,    /* plain FP64 kernels (every BASELINE number) */
  TTN_ACCURACY_REFINED = 1  /* FP64 kernels, then the points whose value is small against the RMS of the batch
                               (cancellation: the only points whose floored relative error can exceed 1e-12) are
                               re-evaluated by a double-double kernel and overwritten */
};

/* ---- fused quadrature functionals: what is accumulated over the evaluated points.
 * With the grid generator these generalise integrate() (src/integration.jl:6-32) to functionals a
 * pure contraction cannot express. */
enum {
  TTN_REDUCE_NONE = 0,
  TTN_REDUCE_SUM = 1,      /* sum_p f(p)                      (integrate(...; take_sum=true)) */
  TTN_REDUCE_ABS2 = 2,     /* sum_p |f(p)|^2                  -> sum_out[0] */
  TTN_REDUCE_WEIGHTED = 3  /* sum_p w[p] * f(p), w real, one weight per point (opts->weights) */
};

/* ---- kernel selection --------------------------------------------------- */
enum {
  TTN_KERNEL_AUTO = 0,
  TTN_KERNEL_GENERIC = 1, /* any tree, any dims: thread-per-point, messages in HBM scratch */
  TTN_KERNEL_CHAIN = 2,   /* MPS-shaped networks: per-point state in registers, site matrices
                             streamed through a shared-memory ring by bulk async copies (TMA) */
  TTN_KERNEL_DMMA = 3,    /* chains up to width 32: points grouped by digit, FP64 DMMA tiles, state in
                             shared memory */
  TTN_KERNEL_TREE = 5,    /* trees with <= 2 children per vertex, chi <= 64 real / 32 complex (binary trees, combs):
                             vertex by vertex over a chunk, degree-3 vertices as Khatri-Rao FP64 DMMA GEMMs, subtree
                             message tables, runs of single-child vertices merged into one GEMM */
  TTN_KERNEL_GRID = 6,    /* ttn_evaluate_grid on a FULL dyadic grid of a binary chain: prefix-shared level-by-level
                             expansion (~L/2 times fewer flops than independent points) */
  TTN_KERNEL_TABLE = 7,   /* narrow chains of binary site indices (chi <= 4 real / 2 complex): groups of vertices
                             pre-contracted at plan time into shared-memory tables, one lookup per group; the
                             HBM-bound point-streaming kernel */
  TTN_KERNEL_GEMM = 4     /* wide chains (width 33..256, e.g. complex chi = 128): per-site class-grouped FP64
                             DMMA GEMM with gathered rows, state in HBM/L2 */
};

/*
 * Flat description of one ITensorNetworkFunction on a TREE, rooted at `root`.
 *
 * Vertices are numbered 0..n_vertices-1.  `parent[v]` is the vertex towards the
 * root (-1 for the root).  The children of v are the vertices u with
 * parent[u]==v, taken in ascending u.  `link_dim[v]` is the dimension of the
 * link index between v and parent[v] (must be 1 for the root).
 *
 * Site indices (the "Digit" indices of src/digit_inds.jl:72-115) are listed in
 * CSR form per vertex: vertex v owns entries site_ptr[v] .. site_ptr[v+1]-1.
 * A vertex may own zero, one or several site indices
 * (src/digit_inds.jl:78-83, test/test_complexitensorfunction.jl:183-191).
 * For site index s:
 *   site_dim[s]    its dimension b (the base; 2 for binary digits)
 *   site_coord[s]  which REAL coordinate slot drives it, 0..n_coords-1.  For a
 *                  RealIndexMap slot i is xs[i]; for a ComplexIndexMap slot 2i is
 *                  real(xs[i]) and slot 2i+1 is imag(xs[i])
 *                  (complexindexmap.jl:116-132).
 *   site_digit[s]  its digit number (1 = most significant; realindexmap.jl:48-58).
 *                  Within one coordinate slot the greedy loop visits site
 *                  indices in ascending digit number (realindexmap.jl:72).
 *   thr[thr_ptr[s] + v], v = 0..b-1 :  abs(index_value_to_scalar(imap, ind, v))
 *                  (realindexmap.jl:13-15, complexindexmap.jl:25-32), computed
 *                  by the caller with the reference's own function so that the
 *                  thresholds are bit-identical for every base.
 *
 * Tensor of vertex v: element type double (is_complex==0) or interleaved
 * (re,im) double pairs (is_complex==1), starting at element offset
 * tensor_ptr[v] of `tensors`, C order (last axis fastest), axes
 *   [site_0] ... [site_{m-1}] [child_0] ... [child_{k-1}] [parent]
 * so that one setting of the site indices selects one contiguous slice
 * (what `project` does with onehot, src/itensornetworkfunction.jl:84-94).
 */
typedef struct ttn_desc {
  int32_t abi_version; /* TTN_ABI_VERSION */
  int32_t n_vertices;
  int32_t n_coords;
  int32_t is_complex;
  int32_t root;
  int32_t n_sites; /* == site_ptr[n_vertices] */
  const int32_t* parent;     /* [n_vertices] */
  const int32_t* link_dim;   /* [n_vertices] */
  const int32_t* site_ptr;   /* [n_vertices + 1] */
  const int32_t* site_dim;   /* [n_sites] */
  const int32_t* site_coord; /* [n_sites] */
  const int32_t* site_digit; /* [n_sites] */
  const int32_t* thr_ptr;    /* [n_sites + 1] */
  const double* thr;         /* [thr_ptr[n_sites]] */
  const int64_t* tensor_ptr; /* [n_vertices + 1], in elements (complex counts as one) */
  const void* tensors;
} ttn_desc;

typedef struct ttn_opts {
  int32_t coords_mem;   /* TTN_MEM_* */
  int32_t out_mem;      /* TTN_MEM_* */
  int32_t kernel;       /* TTN_KERNEL_* (AUTO = planner's choice) */
  int32_t reduce_sum;   /* TTN_REDUCE_*: 0 = write one value per point to out; otherwise also/only accumulate
                           the chosen functional into sum_out (out may then be NULL) */
  int64_t chunk_points; /* 0 = auto; host-memory calls are pipelined in chunks of this many points */
  /* outputs */
  double sum_out[2];    /* (re, im) of the sum when reduce_sum != 0 */
  float kernel_ms;      /* device time of the contraction kernel launches of this call, measured
                           with CUDA events on the launching stream */
  float total_ms;       /* device time of the whole call incl. copies (events) */
  int32_t kernel_used;  /* TTN_KERNEL_* actually run */
  int32_t n_launches;   /* kernels launched by this call */
  /* inputs (appended in ABI 2) */
  const double* weights; /* TTN_REDUCE_WEIGHTED: npts weights, in the memory space weights_mem */
  int32_t weights_mem;   /* TTN_MEM_* */
  int32_t reserved_;
  /* output */
  double flops_executed; /* FP64 flops the kernels of this call executed.  <= flops_per_point * npts (the
                            SURVEY 8(d) rule): plan-time contraction (merged chain positions, leaf/root and
                            subtree tables) and the grid kernel's prefix sharing remove work */
  /* inputs (appended in ABI 3) */
  int32_t host_staging;  /* TTN_STAGE_*: what to do with PAGEABLE host buffers (a Julia Array, a numpy array) */
  int32_t accuracy;      /* TTN_ACCURACY_* */
  double refine_tau;     /* TTN_ACCURACY_REFINED: points with |f| < refine_tau * rms(f) are re-evaluated in
                            double-double arithmetic, rms taken over the chunk of the call the point belongs to (the
                            whole call for device-resident buffers, <= chunk_points for host buffers, one block per
                            GPU of a multi-device plan); 0 = default (0.02) */
  /* outputs (ABI 3) */
  int32_t n_devices_used; /* GPUs that took part in this call (multi-device plans shard the points) */
  int32_t staged;         /* bit 0: coords went through the pinned staging ring, bit 1: out did, bit 2: coords were
                             quantised on the host (TTN_STAGE_AUTO) */
  int64_t n_refined;      /* points re-evaluated by TTN_ACCURACY_REFINED */
  int64_t h2d_bytes;      /* bytes this call copied host -> device (coordinates / index settings / weights) */
  int64_t d2h_bytes;      /* bytes this call copied device -> host (values) */
} ttn_opts;

/* Uniform grid generator — grid_points(imap, N, d), src/IndexMaps/realindexmap.jl:78-86:
 * coordinate slot c takes the values i * step[c], i = 0..count[c]-1, and the evaluation set is
 * the Cartesian product of all slots (slot 0 slowest).  Points are generated on the device; no
 * coordinate bytes are read. */
typedef struct ttn_grid {
  int32_t n_coords;
  const double* step;   /* [n_coords] a / b^L */
  const int64_t* count; /* [n_coords] number of kept points (those < 1) */
  int64_t first;        /* linear index of the first grid point of this call (sharding) */
  int64_t npts;         /* how many consecutive grid points to evaluate */
} ttn_grid;

typedef struct ttn_info {
  int32_t n_vertices, n_coords, is_complex, n_sites;
  int32_t max_link_dim;
  int32_t is_chain;          /* every vertex has <= 1 child once rooted */
  int32_t auto_kernel;       /* TTN_KERNEL_* chosen by the planner */
  int32_t device;
  int32_t kernels_available; /* bit k set: TTN_KERNEL_k can run this network */
  int32_t n_devices;         /* 1, or the number of GPUs of a ttn_plan_create_multi plan */
  double flops_per_point;    /* SURVEY §8(d) flop rule: 2 (real) / 8 (complex) * sum of MACs */
  double bytes_per_point;    /* 8 * n_coords read + 8/16 written */
  int64_t tensor_bytes;
} ttn_info;

typedef struct ttn_plan ttn_plan; /* opaque; owns all device memory and streams */

/* Build a plan on CUDA device `device` (>= 0).  Replaces the per-point `copy(fitn)` +
 * dictionary construction of the reference (itensornetworkfunction.jl:85, realindexmap.jl:69). */
int ttn_plan_create(const ttn_desc* desc, int32_t device, ttn_plan** out);
/* The same plan replicated on n_devices GPUs of this process (devices[i], or 0..n_devices-1 when devices is
 * NULL) — SURVEY 8(b),(e): one process drives all GPUs of the box.  Every ttn_evaluate* call on such a plan
 * splits its points into n_devices contiguous blocks (GPU g takes [g*ceil(n/G), ...)), each GPU copies its
 * block in and its values out on its own streams (host buffers: straight into / out of the caller's arrays;
 * device buffers on another GPU: peer copies over NVLink), there is no exchange step, and for reduce_sum != 0
 * the per-GPU sums are added on the host in device order (deterministic).  `device` of ttn_info is devices[0]. */
int ttn_plan_create_multi(const ttn_desc* desc, int32_t n_devices, const int32_t* devices, ttn_plan** out);
void ttn_plan_destroy(ttn_plan* plan);
int ttn_plan_info(const ttn_plan* plan, ttn_info* info);

/* out[p] = f(point p).  `out` holds npts doubles (real network) or npts (re,im) pairs.
 * Replaces the caller loop over evaluate(), e.g. examples/2d_laplace_solver.jl:49-53. */
int ttn_evaluate(ttn_plan* plan, const double* coords, int64_t npts, int32_t n_coords,
                 int32_t layout, void* out, ttn_opts* opts);

/* Same on a generated grid (grid_points), optionally with the summed-grid quadrature
 * (identity: sum over all b^L points == integrate(fitn; take_sum=true), src/integration.jl:6-17). */
int ttn_evaluate_grid(ttn_plan* plan, const ttn_grid* grid, void* out_or_null, ttn_opts* opts);

/* Evaluation at given index settings — the inner loop of TCI (ext/ITensorNumericalAnalysisTCIExt/tci_util.jl:21-55
 * evaluates candidate pivots / fibres one index setting at a time) and the batched form of
 * project() + scalar() (src/itensornetworkfunction.jl:84-106) without the coordinate -> digit step.
 * index_values[p * n_sites + s] (uint8) = value of site index s of the description, 0 <= value < site_dim[s];
 * memory space = opts->coords_mem.  Out-of-range values return TTN_ERR_INVALID. */
int ttn_evaluate_indices(ttn_plan* plan, const uint8_t* index_values, int64_t npts, void* out, ttn_opts* opts);

/* Digit decomposition only — batched calculate_ind_values (realindexmap.jl:67-76).
 * digits_out[p * n_sites + s] (uint8) = value chosen for site index s of the description. */
int ttn_digits(ttn_plan* plan, const double* coords, int64_t npts, int32_t n_coords,
               int32_t layout, uint8_t* digits_out, ttn_opts* opts);

/* FP64 roofline denominators measured on the plan's device: a dependent-free DFMA loop
 * (vector pipe) and an mma.sync m8n8k4 f64 loop (DMMA pipe); TFLOP/s. */
int ttn_measure_fp64_peak(int32_t device, double* dfma_tflops, double* dmma_tflops);

/* Pin / unpin a caller-owned host range (cudaHostRegister, portable) so that ttn_evaluate uses it in place at
 * the full PCIe rate.  The caller must unregister before freeing the memory. */
int ttn_host_register(void* ptr, uint64_t bytes);
int ttn_host_unregister(void* ptr);

const char* ttn_last_error(void);
int ttn_device_count(void);
int ttn_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TTNEVAL_H */
