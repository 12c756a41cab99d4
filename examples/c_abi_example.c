/* Minimal C caller of libttneval.so (include/ttneval.h): a 3-vertex binary MPS with chi = 2 evaluated at four
 * points.  It shows what any FFI (Julia ccall, ctypes, cgo, JNI) has to provide: the flat description, host
 * buffers, an options struct; and what it gets back: values or an error code + message.
 *
 *   gcc -std=c11 -Iinclude examples/c_abi_example.c -Litensornumericalanalysis.jl_b200/csrc -lttneval \
 *       -Wl,-rpath,$PWD/itensornumericalanalysis.jl_b200/csrc -o c_abi_example && ./c_abi_example
 *
 * Without a CUDA device ttn_plan_create fails with TTN_ERR_CUDA (there is no CPU fallback) and the program says so. */
#include <stdio.h>
#include <string.h>

#include "ttneval.h"

int main(void) {
  printf("libttneval ABI %d, %d CUDA device(s)\n", ttn_abi_version(), ttn_device_count());

  /* chain 0 - 1 - 2 rooted at vertex 2; one binary site index per vertex; 1-D function, digits 1, 2, 3 */
  const int32_t parent[3] = {1, 2, -1};
  const int32_t link_dim[3] = {2, 2, 1};
  const int32_t site_ptr[4] = {0, 1, 2, 3};
  const int32_t site_dim[3] = {2, 2, 2};
  const int32_t site_coord[3] = {0, 0, 0};
  const int32_t site_digit[3] = {1, 2, 3};
  const int32_t thr_ptr[4] = {0, 2, 4, 6};
  const double thr[6] = {0.0, 0.5, 0.0, 0.25, 0.0, 0.125};       /* |index_value_to_scalar(ind, v)| = v 2^-digit */
  /* tensors [site][child][parent]: f(x) = 1 + x as the chain (1, x_acc) -> ... : leaf (1, d/2), middle
   * [[1, d/4], [0, 1]], root (1 + ... ) */
  const int64_t tensor_ptr[4] = {0, 4, 12, 16};
  const double tensors[16] = {
      /* vertex 0 (leaf): [s][parent]          */ 1.0, 0.0, 1.0, 0.5,
      /* vertex 1:        [s][child][parent]   */ 1.0, 0.0, 0.0, 1.0, 1.0, 0.25, 0.0, 1.0,
      /* vertex 2 (root): [s][child]           */ 1.0, 1.0, 1.125, 1.0};
  ttn_desc d;
  memset(&d, 0, sizeof d);
  d.abi_version = TTN_ABI_VERSION;
  d.n_vertices = 3;
  d.n_coords = 1;
  d.is_complex = 0;
  d.root = 2;
  d.n_sites = 3;
  d.parent = parent;
  d.link_dim = link_dim;
  d.site_ptr = site_ptr;
  d.site_dim = site_dim;
  d.site_coord = site_coord;
  d.site_digit = site_digit;
  d.thr_ptr = thr_ptr;
  d.thr = thr;
  d.tensor_ptr = tensor_ptr;
  d.tensors = tensors;

  ttn_plan* plan = NULL;
  int rc = ttn_plan_create(&d, 0, &plan);
  if (rc != TTN_OK) {
    printf("ttn_plan_create: error %d: %s\n", rc, ttn_last_error());
    return rc == TTN_ERR_CUDA ? 0 : 1; /* no GPU here: expected */
  }
  const double xs[4] = {0.0, 0.375, 0.5, 0.999};
  double out[4];
  ttn_opts o;
  memset(&o, 0, sizeof o);
  o.coords_mem = TTN_MEM_HOST;
  o.out_mem = TTN_MEM_HOST;
  o.kernel = TTN_KERNEL_AUTO;
  rc = ttn_evaluate(plan, xs, 4, 1, TTN_LAYOUT_AOS, out, &o);
  if (rc != TTN_OK) {
    printf("ttn_evaluate: error %d: %s\n", rc, ttn_last_error());
    ttn_plan_destroy(plan);
    return 1;
  }
  for (int i = 0; i < 4; ++i) printf("f(%.3f) = %.6f   (1 + floor(8 x) / 8 = %.6f)\n", xs[i], out[i], 1.0 + (double)(int)(xs[i] * 8) / 8);
  ttn_plan_destroy(plan);
  return 0;
}
