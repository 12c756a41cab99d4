// Internal structures of libttneval.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/ttneval.h"

namespace ttn {

void set_error(const std::string& msg);
#define TTN_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      ttn::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));              \
      return TTN_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// ---- digit extraction tables (device) ------------------------------------------------
// One entry per site index, grouped by coordinate slot and sorted by digit number ascending
// (sort(indices; by=digit), src/IndexMaps/realindexmap.jl:72).
struct DigitEntry {
  int32_t site;    // site id in the description (for ttn_digits output)
  int32_t base;    // dim(ind)
  int32_t thr_off; // offset into thr[]: thr[thr_off + v] = abs(index_value_to_scalar(ind, v))
  int32_t vertex;  // owner vertex
  int32_t stride;  // stride of this digit inside the vertex's mixed-radix slice index
  int32_t word;    // chain kernel: which 64-bit word of the packed slice stream
  int32_t shift;   // chain kernel: bit offset inside that word
  int32_t pad_;    // 1: the owner vertex carries no other site index (its slice index IS this digit times stride)
};

struct DigitTable {
  int32_t n_coords;
  int32_t n_sites;
  const int32_t* coord_ptr;  // [n_coords + 1] into entries
  const DigitEntry* entries; // [n_sites]
  const double* thr;
};

// Where coordinates come from: a caller array or the grid generator (grid_points,
// src/IndexMaps/realindexmap.jl:78-86).
#define TTN_MAX_COORDS 16
struct CoordSource {
  const double* coords; // device pointer, or nullptr in grid mode
  int64_t npts;         // points in this launch
  int32_t n_coords;
  int32_t layout;       // TTN_LAYOUT_*
  int32_t grid;         // 1: generate
  int64_t first;        // grid: linear index of point 0 of this launch
  double step[TTN_MAX_COORDS];
  int64_t count[TTN_MAX_COORDS];
  const uint8_t* digits; // index-setting mode (ttn_evaluate_indices): digits[p * n_sites + site], else nullptr
  int32_t reduce_mode;   // TTN_REDUCE_* (what the kernels accumulate per point)
  const double* weights; // TTN_REDUCE_WEIGHTED: device pointer, one weight per point of this launch
  const uint32_t* qcoords; // host-quantised coordinates (TTN_STAGE_AUTO): qcoords[p * n_coords + c] = min(floor(x 2^L_c), 2^L_c - 1);
                           // nullptr otherwise.  Only kernels whose K1 runs every coordinate on the run path read it
  int32_t pcie_bound;    // 1: this launch is one chunk of a host-buffer call (H2D / D2H copies run beside it): launchers may
                         // pick the variant that leaves the memory system to the copy engines (launch_chain_team)
  unsigned long long* dbg_stream; // test hook (ttn_debug_slice_stream): the kernels with a FUSED K1 (team-sorted DMMA
                                  // kernel, table kernel) store the packed slice stream of point p — the digits they
                                  // really use — at dbg_stream[2p], [2p + 1]; nullptr in every product call
};

// ---- generic tree program (device) ---------------------------------------------------
struct TreeDev {
  int32_t n_vertices, root, is_complex;
  const int32_t* post;      // post order
  const int32_t* child_ptr; // CSR children
  const int32_t* child;
  const int32_t* link_dim;
  const int64_t* slice_size; // elements per slice
  const int64_t* tensor_off; // element offset of the vertex tensor
  const int64_t* msg_off;    // per-point message offsets (elements)
  const double* tensors;
  int64_t msg_total;  // elements per point in the message area
  int64_t max_inter;  // elements of the largest intermediate
};

// ---- chain (MPS) program --------------------------------------------------------------
struct ChainDev {
  int32_t n_steps;     // middle vertices streamed through the shared-memory ring
  int32_t n_vertices;  // chain length (leaf .. root)
  int32_t chi;         // padded bond dimension (template instance)
  int32_t nsl;         // padded slices per vertex (template instance)
  int32_t bits;        // bits per vertex in the packed slice stream
  int32_t per_word;    // vertices per 64-bit word
  int32_t n_words;
  const double* leaf;  // [nsl][chi] (x2 if complex)
  const double* root;  // [nsl][chi]
  const double* steps; // [n_steps] stage images, see k_chain.cu for the layout
  int64_t stage_bytes;
};

// chain networks on FP64 tensor cores (k_chain_mma.cu)
struct ChainMmaDev {
  int32_t n_vertices, n_steps, n_rounds;
  int32_t spr;      // sites per round (one shared-memory round trip of the state per round)
  int32_t nsl;      // padded slices per vertex
  int32_t chi;      // real row width (2*chi for complex networks, embedded), padded: 8, 16 or 32
  int32_t nout;     // 1 (real) or 2 (re, im)
  int32_t bits, per_word, n_words;
  int32_t root_pos; // position of the root in the packed slice stream (after identity padding)
  int32_t merged;   // k > 0: k vertices pre-contracted into one stream position (build_chain_mma)
  int32_t leaf_bits, root_bits; // stream bits of the leaf / root group (deep tables: up to 21)
  int32_t k1_generic;           // 1: some site index is not binary -> K1 runs the tabulated greedy loop (base 3 / 4)
  // per coordinate slot: "run" fast path of K1 (see build_chain_mma): L == 0 -> use the table loop
  int32_t run_L[TTN_MAX_COORDS];     // number of binary digits
  int32_t run_plow[TTN_MAX_COORDS];  // lowest stream position of the run
  int32_t run_rev[TTN_MAX_COORDS];   // 1: digit 1 at the LOWEST position (bit-reversed placement)
  double run_scale[TTN_MAX_COORDS];  // 2^L
  const double* leaf;  // [nsl][chi]
  const double* root;  // [nout][nsl][chi]
  const double* frags; // [n_steps][nsl][chi*chi], B-fragment order
};

// wide chains as per-site grouped GEMMs (k_chain_gemm.cu)
struct ChainGemmDev {
  int32_t n_vertices, n_steps, nsl, W, nout;
  int32_t n_pos;                 // stream positions (= vertices, or vertex pairs when merged)
  int32_t merged;                // 2: vertex pairs pre-contracted at plan time (build_chain_gemm)
  const double* leaf;            // [nsl][W]
  const double* root;            // [nout][nsl][W]
  const double* frags;           // [n_steps][nsl][W*W], B-fragment order
  const int32_t* pos_of_vertex;  // stream position of every vertex | (bit shift inside the position) << 16
  // leaf / root tables (build_gemm_tables): the first 1 + tab_L and the last 1 + tab_R positions are
  // tabulated (nsl^(1+tab) rows each); only the middle steps [tab_L, n_steps - tab_R) run as GEMMs
  int32_t tab_L, tab_R;
  const double* leaf_tab;        // [nsl^(1+tab_L)][W] or null
  const double* root_tab;        // [nout][nsl^(1+tab_R)][W] or null
};

// narrow chains as shared-memory group tables (k_chain_table.cu)
constexpr int kTabMaxGroups = 16;
constexpr int kTabMaskCoords = 8; // coordinate slots that can use the masked K1 run path
struct ChainTabDev {
  int32_t n_groups;      // >= 2: group 0 = leaf vectors, last = root vectors, between: H x H matrices
  int32_t H;             // padded bond dimension (a complex entry counts once): 1, 2 or 4
  int32_t cplx;
  int32_t rep;           // 1: chi = 1 tables replicated per lane (one 128-byte line per entry; conflict-free)
  int32_t k1_generic;    // 1: some site index is not binary (base 3 / 4): tabulated greedy loop in K1
  int32_t total_doubles; // image size (even)
  int32_t gbits[kTabMaxGroups]; // stream bits consumed by the group (<= 16)
  int32_t goff[kTabMaxGroups];  // offset of the group's table in the image, in doubles (even)
  int32_t run_L[TTN_MAX_COORDS], run_plow[TTN_MAX_COORDS], run_rev[TTN_MAX_COORDS]; // K1 run fast path
  double run_scale[TTN_MAX_COORDS];
  // run_kind: 0 = tabulated greedy loop, 1 = digits on consecutive stream bits (shift), 2 = digits on any
  // monotone set of stream bits (interleaved dimensions, two site indices per vertex): the bits of
  // floor(x 2^L) are deposited into the mask exp_m by a 6-step shift/select network (masks exp_mv)
  int32_t run_kind[TTN_MAX_COORDS];
  int32_t exp_nlo[kTabMaskCoords];        // popcount(exp_m[c][0])
  uint64_t exp_m[kTabMaskCoords][2];      // stream bits of the coordinate's digits (word 0, word 1)
  uint64_t exp_mv[kTabMaskCoords][2][6];
  const double* image;
};

// trees with <= 2 children per vertex as per-vertex (Khatri-Rao) GEMMs (k_tree_gemm.cu)
struct TreeGemmDev {
  int32_t n_vertices, W, root;
  const double* blob;       // per-vertex images: leaf rows / B-fragment tensors / padded root tensor
  const int32_t* nslices;   // [n_vertices]
};

struct Stream {
  cudaStream_t s = nullptr;
  cudaEvent_t k0 = nullptr, k1 = nullptr;
  double* d_coords = nullptr;
  double* d_out = nullptr;
  double* d_weights = nullptr;
  uint8_t* d_digits = nullptr;
  size_t digits_cap = 0;
  int64_t cap_points = 0;
  double* d_work = nullptr; // generic-kernel workspace
  size_t work_bytes = 0;
  double* d_partial = nullptr; // per-CTA partial sums (reduce_sum)
  size_t partial_cap = 0;
  void* d_gemm = nullptr;      // GEMM-path workspace (state ping-pong, slices, lists)
  size_t gemm_bytes = 0;
  double* d_partial2 = nullptr;
  size_t partial2_bytes = 0;
  // pinned staging ring for PAGEABLE caller buffers (allocated on first use, chunk capacity)
  double* h_coords = nullptr;
  double* h_out = nullptr;
  double* h_weights = nullptr;
  uint8_t* h_digits = nullptr;
  int64_t h_cap_points = 0;
  size_t h_digits_cap = 0;
  cudaEvent_t ev_h2d = nullptr, ev_d2h = nullptr;
  uint32_t* h_q = nullptr;    // quantised coordinates: pinned ring slot and device buffer (chunk capacity)
  uint32_t* d_q = nullptr;
  int64_t q_cap_points = 0;
  // refine pass (TTN_ACCURACY_REFINED): compacted point list + values
  int32_t* d_sel = nullptr;   // [1 + points]: count, then the selected point indices
  size_t sel_cap = 0;         // in points
  void* d_refine = nullptr;   // double-double workspace
  size_t refine_bytes = 0;
};

} // namespace ttn

struct ttn_plan {
  int device = 0;
  int sm_count = 0;
  ttn_info info{};
  // host copies
  std::vector<int32_t> parent, link_dim, post, child_ptr, child, chain_order;
  std::vector<int64_t> slice_size, tensor_off, msg_off;
  std::vector<int32_t> nslices;
  bool is_chain = false;
  double exec_rule_flops = 0.0; // flop rule of the tree the kernels really walk (== info.flops_per_point unless binarized)
  bool binarized = false; // vertices with > 2 children were split with Kronecker pair vertices (ttn_api.cu, binarize_desc)
  // device memory
  std::vector<void*> allocs;
  ttn::DigitTable digits{};
  ttn::DigitTable digits_mma{}; // same table, (word, shift) for the DMMA kernels' stream layout
  std::vector<int32_t> cmma_site_fbits; // per site index: width of the stream field its digit lives in (test hook)
  ttn::TreeDev tree{};
  ttn::ChainDev chain{};
  bool chain_ok = false;
  ttn::ChainMmaDev cmma{};
  ttn::ChainMmaDev cmma_plain{}; // one vertex per position (leaf/root/frags only): grid-share kernel
  ttn::ChainMmaDev cmma_light{}; // the merged image WITHOUT deep leaf / root tables (same stream layout): host-buffer calls
  bool cmma_light_ok = false;
  double cmma_light_flops = 0.0;
  bool cmma_plain_ok = false;
  double cgemm_flops_exec = 0.0; // flops per point the GEMM chain kernel executes
  double cmma_flops_exec = 0.0;  // flops per point the DMMA chain kernel executes (merged: about half the rule)
  bool cmma_ok = false;
  ttn::ChainGemmDev cgemm{};
  bool cgemm_ok = false;
  bool gshare_ok = false;            // prefix-shared full-grid evaluation (k_grid_share.cu)
  std::vector<int> gs_coord, gs_digit, gs_L;
  ttn::ChainTabDev ctab{};
  bool ctab_ok = false;
  ttn::DigitTable digits_tab{}; // same table, (word, shift) for the table kernel's stream layout
  double ctab_flops_exec = 0.0;
  int ctab_bits0 = 1;           // stream bits per vertex of the table kernel's image (test hook)
  ttn::TreeGemmDev tgemm{};
  bool tgemm_ok = false;
  std::vector<int64_t> tg_frag_off;
  // subtree message tables of the tree kernel (build_tree_tables): per vertex -1 = computed, -2 = inside a
  // tabulated subtree, >= 0 = index of its table
  std::vector<int32_t> tg_tab_of, tg_tab_ns, tg_tab_bits;
  std::vector<double*> tg_tab;
  std::vector<int32_t*> tg_tab_vs;
  double tgemm_flops_exec = 0.0;
  // evaluation plan of the tree kernel (build_tree_merge): runs of single-child vertices merged into one GEMM
  std::vector<int32_t> tg_role;       // per vertex: 0 = as is, 1 = absorbed into a merged run, 2 = top of a merged run
  std::vector<int32_t> tg_in, tg_mnsl; // merged top: vertex whose message feeds the run; number of classes
  std::vector<int64_t> tg_mfrag_off;  // merged top: offset of its class matrices in tg_mblob
  double* tg_mblob = nullptr;         // (in allocs)
  void* tg_gv = nullptr;              // device array of TgClass: the vertices that run as GEMMs (in allocs)
  int tg_ngv = 0;
  int32_t* tg_zero_v = nullptr;       // vertices whose slice index is accumulated (0 or >= 2 site indices): zero-filled per chunk
  int tg_n_zero = 0;
  void* tg_tabrefs = nullptr;         // device array of (vs, ns) per subtree table (in allocs)
  bool all_base2 = false; // every site index has dimension 2 (branch-free digit path)
  int fe_thr_len = 0; // length of the threshold table (front-end shared-memory copy)
  int v6_teams = 3;        // teams per CTA of the team-sorted kernel (TTN_MMA_V6 at ttn_plan_create; 0 = ring kernels)
  int host_peers = 1;      // GPUs that stream through this host's memory system at the same time as far as the library can
                           // tell: the replicas of a multi-device plan, else LOCAL_WORLD_SIZE (one process per GPU under
                           // torchrun / mpirun wrappers that export it).  Decides the host-side quantisation default (ttn_api.cu)
  double* d_grid_out = nullptr; // grow-only scratch of the grid kernel's host-output path
  size_t grid_out_bytes = 0;
  std::vector<ttn_plan*> replicas; // non-empty: a multi-device plan (ttn_plan_create_multi); this object owns them
  int* d_err = nullptr;    // domain-error flag
  double* d_sum = nullptr; // (re, im) per chunk
  int32_t* d_nsel = nullptr; // refined points per chunk
  ttn::Stream streams[3];
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  std::mutex mu;
};

namespace ttn {
// kernel launchers (defined in the .cu files).  All return a TTN_* code.
int launch_digits(ttn_plan* p, const CoordSource& src, uint8_t* d_digits, cudaStream_t s);
int launch_generic(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out,
                   double* d_partial, int* n_partial, cudaStream_t s);
int launch_chain(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out,
                 double* d_partial, int* n_partial, cudaStream_t s);
int launch_sum_partials(ttn_plan* p, const double* d_partial, int n_partial, int nc,
                        double* d_sum, cudaStream_t s);
int build_chain(ttn_plan* p, const ttn_desc* d);
int build_chain_mma(ttn_plan* p, const ttn_desc* d);
int build_chain_gemm(ttn_plan* p, const ttn_desc* d);
int build_tree_gemm(ttn_plan* p, const ttn_desc* d);
int build_grid_share(ttn_plan* p, const ttn_desc* d);
bool grid_share_applicable(const ttn_plan* p, const CoordSource& src);
int launch_grid_share(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double* d_partial,
                      int* n_partial, cudaStream_t s, int* n_launches, double* flops_executed);
int launch_tree_gemm(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double* d_partial,
                     int* n_partial, cudaStream_t s, int* n_launches);
int launch_chain_gemm(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double* d_partial,
                      int* n_partial, cudaStream_t s, int* n_launches);
int launch_chain_mma(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out,
                     double* d_partial, int* n_partial, cudaStream_t s);
// DMMA chain kernels behind launch_chain_mma: the team-sorted kernel (k_chain_team.cu) for merged binary
// chains, the ring kernels (k_chain_ring.cu) for everything else
bool chain_team_applicable(const ttn_plan* p);
int launch_chain_team(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                      cudaStream_t s);
int launch_chain_ring(ttn_plan* p, const CoordSource& src, double* d_out, double* d_partial, int* n_partial,
                      cudaStream_t s);
int build_chain_table(ttn_plan* p, const ttn_desc* d);
int launch_chain_table(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double* d_partial,
                       int* n_partial, cudaStream_t s);
int debug_table_image(const ttn_desc* d, int32_t budget_kb, int32_t* meta, double* image, int64_t image_cap,
                      int32_t* site_bitpos);
bool chain_supported(int chi, int nsl, bool cplx);
int measure_fp64_peak(int device, double* dfma, double* dmma);
// TTN_ACCURACY_REFINED (k_refine.cu): select the points with |f| < thr, re-evaluate them in double-double
int launch_refine(ttn_plan* p, Stream& st, const CoordSource& src, double* d_out, double tau, int32_t* d_sel,
                  double* d_partial, int* n_partial, cudaStream_t s, int* n_launches);
} // namespace ttn
