/* C99 consumer of include/neko_top_b200.h: proves that the header is plain C (no C++/torch types in any
 * signature), that every declared entry point resolves at link time against libneko_top_b200.so, and that
 * the error convention works without a GPU (status code + message instead of abort when asked).
 * Built and run by tests/test_abi.py (CPU) -- it makes no compute call. */
#include <stdio.h>
#include <string.h>

#include "../../../include/neko_top_b200.h"

struct sym { const char* name; void (*fn)(void); };
static const struct sym table[] = {
  {"b200_set_abort_on_error", (void (*)(void))&b200_set_abort_on_error},
  {"b200_version", (void (*)(void))&b200_version},
  {"b200_last_error", (void (*)(void))&b200_last_error},
  {"b200_launch_count", (void (*)(void))&b200_launch_count},
  {"b200_adjrhs_create", (void (*)(void))&b200_adjrhs_create},
  {"b200_adjrhs_free", (void (*)(void))&b200_adjrhs_free},
  {"b200_adjrhs_set_stream", (void (*)(void))&b200_adjrhs_set_stream},
  {"b200_adjrhs_set_space", (void (*)(void))&b200_adjrhs_set_space},
  {"b200_adjrhs_set_geometry", (void (*)(void))&b200_adjrhs_set_geometry},
  {"b200_adjrhs_set_params", (void (*)(void))&b200_adjrhs_set_params},
  {"b200_adjrhs_set_lube_mask", (void (*)(void))&b200_adjrhs_set_lube_mask},
  {"b200_adjrhs_compute", (void (*)(void))&b200_adjrhs_compute},
  {"b200_adjrhs_step", (void (*)(void))&b200_adjrhs_step},
  {"b200_adjrhs_step_host", (void (*)(void))&b200_adjrhs_step_host},
  {"b200_adv_adjoint_compute", (void (*)(void))&b200_adv_adjoint_compute},
  {"b200_adv_linear_compute", (void (*)(void))&b200_adv_linear_compute},
  {"b200_adv_dealias_init", (void (*)(void))&b200_adv_dealias_init},
  {"b200_adv_adjoint_dealias_compute", (void (*)(void))&b200_adv_adjoint_dealias_compute},
  {"b200_adv_linear_dealias_compute", (void (*)(void))&b200_adv_linear_dealias_compute},
  {"b200_adjrhs_set_dealias", (void (*)(void))&b200_adjrhs_set_dealias},
  {"b200_brinkman_compute", (void (*)(void))&b200_brinkman_compute},
  {"b200_lube_compute", (void (*)(void))&b200_lube_compute},
  {"b200_opcolv", (void (*)(void))&b200_opcolv},
  {"b200_ramp_forward", (void (*)(void))&b200_ramp_forward},
  {"b200_ramp_backward", (void (*)(void))&b200_ramp_backward},
  {"b200_sensitivity", (void (*)(void))&b200_sensitivity},
  {"b200_steady_field_update", (void (*)(void))&b200_steady_field_update},
  {"b200_curl", (void (*)(void))&b200_curl},
  {"b200_curlcurl_forcing", (void (*)(void))&b200_curlcurl_forcing},
  {"b200_min_dissipation_objective", (void (*)(void))&b200_min_dissipation_objective},
  {"b200_mask_exterior_const", (void (*)(void))&b200_mask_exterior_const},
  {"b200_sumab", (void (*)(void))&b200_sumab},
  {"b200_makeabf", (void (*)(void))&b200_makeabf},
  {"b200_makebdf", (void (*)(void))&b200_makebdf},
  {"b200_makeabf_bdf", (void (*)(void))&b200_makeabf_bdf},
  {"b200_gs_init", (void (*)(void))&b200_gs_init},
  {"b200_gs_get_classes", (void (*)(void))&b200_gs_get_classes},
  {"b200_gs_op", (void (*)(void))&b200_gs_op},
  {"b200_gs_op3", (void (*)(void))&b200_gs_op3},
  {"b200_comm_unique_id", (void (*)(void))&b200_comm_unique_id},
  {"b200_comm_init", (void (*)(void))&b200_comm_init},
  {"b200_gs_init_shared", (void (*)(void))&b200_gs_init_shared},
  {"b200_adjrhs_set_boundary_elements", (void (*)(void))&b200_adjrhs_set_boundary_elements},
  {"b200_adjrhs_set_element_order", (void (*)(void))&b200_adjrhs_set_element_order},
  {"b200_adjrhs_set_gs_fused", (void (*)(void))&b200_adjrhs_set_gs_fused},
  {"b200_adjrhs_gs_info", (void (*)(void))&b200_adjrhs_gs_info},
  {"b200_adjrhs_enable_timing", (void (*)(void))&b200_adjrhs_enable_timing},
  {"b200_adjrhs_get_timing", (void (*)(void))&b200_adjrhs_get_timing},
  {"b200_adjrhs_get_phase_timing", (void (*)(void))&b200_adjrhs_get_phase_timing},
};

int main(void) {
  const int n = (int)(sizeof table / sizeof table[0]);
  int i, zero = 0, lx = 8, nelv = 4, dev = 0, rc;
  void* h = NULL;
  for (i = 0; i < n; i++)
    if (!table[i].fn) { fprintf(stderr, "unresolved %s\n", table[i].name); return 1; }
  if (b200_version() <= 0) return 2;
  b200_set_abort_on_error(&zero);                 /* report errors instead of aborting */
  lx = 3;                                         /* outside 4..10: must fail with a message, no device needed */
  rc = b200_adjrhs_create(&h, &lx, &nelv, &dev);
  if (rc == B200_OK || h != NULL || strlen(b200_last_error()) == 0) return 3;
  printf("abi_check: %d entry points, version %d, error path ok (%s)\n", n, b200_version(), b200_last_error());
  return 0;
}
