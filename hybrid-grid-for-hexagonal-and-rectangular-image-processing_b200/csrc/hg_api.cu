// hg_api.cu -- version, thread-local error string, launch counter.
#include "hg_common.cuh"
#include "hg_ptx.cuh"
#include <string.h>
#include <atomic>

namespace hg {
static thread_local char g_err[512] = "";
static thread_local const char* g_last_launch = "";
static std::atomic<int64_t> g_launches{0};   // process-wide: the autograd engine launches the backward kernels from its own thread

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
int finish_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  g_launches += 1;
  g_last_launch = what;
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return HG_OK;
}

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}
}  // namespace hg

extern "C" {
int hg_version(void) { return HG_VERSION; }
const char* hg_last_error(void) { return hg::g_err; }
const char* hg_last_launch(void) { return hg::g_last_launch; }
int64_t hg_launch_count(void) { return hg::g_launches.load(); }
void hg_reset_launch_count(void) { hg::g_launches.store(0); }
}
