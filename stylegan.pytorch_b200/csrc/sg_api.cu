// Error reporting and library metadata for libsgb200.so.
#include "sg_common.cuh"
#include "../../include/sgb200.h"
#include <string.h>
#include <stdio.h>

static thread_local char g_err[512] = {0};

void sg_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

int sg_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

SG_EXPORT const char* sg_last_error(void) { return g_err; }
SG_EXPORT int sg_abi_version(void) { return SGB200_ABI_VERSION; }
SG_EXPORT int sg_geom_size(void) { return (int)sizeof(SgConvGeom); }
SG_EXPORT int sg_entry_size(void) { return (int)sizeof(SgTensorEntry); }

// Fast-path switch (csrc/sg_fast.cuh): 1 = specialised kernels where the shape allows (default), 0 = general kernels
// only.  Returns the previous setting.  Both give the same results; the switch exists for the tests that say so.
int g_sg_fast_paths = 1;
SG_EXPORT int sg_fast_paths(int enable) {
  int old = g_sg_fast_paths;
  if (enable >= 0) g_sg_fast_paths = enable ? 1 : 0;
  return old;
}
