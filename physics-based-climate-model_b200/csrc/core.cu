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
    return (e != nullptr && e[0] == '0') ? 0 : 1;          // default on (PCM_PDL=0: classic launches; see common.cuh)
  }();
  return on != 0;
}

unsigned long long* dropout_epoch_cell() {
  static unsigned long long* ptr = nullptr;
  if (ptr == nullptr) {
    if (cudaMalloc(&ptr, sizeof(unsigned long long)) != cudaSuccess) { ptr = nullptr; return nullptr; }
    cudaMemset(ptr, 0, sizeof(unsigned long long));
  }
  return ptr;
}

__global__ void dropout_epoch_advance_kernel(unsigned long long* cell) {
  PCM_PDL_ENTRY();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cell += 1ull;
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
extern "C" int pcm_version(void) { return 200; }

extern "C" int pcm_dropout_epoch_advance(pcm_stream_t s) {
  unsigned long long* cell = pcm::dropout_epoch_cell();
  PCM_REQUIRE(cell != nullptr, "dropout_epoch_advance: could not allocate the epoch cell");
  pcm::launch(pcm::dropout_epoch_advance_kernel, 1, 32, 0, (cudaStream_t)s, cell);
  return pcm::check_launch("dropout_epoch_advance");
}

extern "C" long long pcm_dropout_epoch(void) {
  unsigned long long v = 0;
  unsigned long long* cell = pcm::dropout_epoch_cell();
  if (cell == nullptr || cudaMemcpy(&v, cell, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (long long)v;
}
