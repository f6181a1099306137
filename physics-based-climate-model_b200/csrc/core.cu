// Error plumbing + version for the pcm_b200 C ABI.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace pcm {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const int on = [] {
    const char* e = getenv("PCM_PDL");
    return (e != nullptr && e[0] == '1') ? 1 : 0;          // opt-in: measured neutral on B200 (see common.cuh)
  }();
  return on != 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PCM_ERR_CUDA;
  }
  return PCM_OK;
}
}  // namespace pcm

extern "C" const char* pcm_last_error(void) { return pcm::g_err; }
extern "C" int pcm_version(void) { return 100; }
