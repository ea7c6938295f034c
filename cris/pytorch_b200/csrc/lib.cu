// lib.cu — library-wide state of libcris_b200: error string, ABI version, device check.
#include <stdarg.h>
#include <stddef.h>
#include <string.h>

#include "common.cuh"

namespace cris {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace cris

extern "C" {

const char* cris_last_error(void) { return cris::g_err; }

int cris_abi_version(void) { return 2; }

int cris_gemm_args_size(void) { return (int)sizeof(cris_gemm_args); }
int cris_gemm_args_last_offset(void) { return (int)offsetof(cris_gemm_args, d_col_stride); }

uint64_t cris_launch_count(void) { return cris::g_launches.load(); }
void cris_add_launch_count(uint64_t n) { cris::g_launches.fetch_add(n); }

int cris_device_check(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cris::set_error("no CUDA device");
    return -1;
  }
  if (prop.major != 10) {
    cris::set_error("libcris_b200 needs an sm_100 (B200) device, found sm_%d%d", prop.major, prop.minor);
    return -1;
  }
  return 0;
}
}
