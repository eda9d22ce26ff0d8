// Library-level entry points of libmvregfus_b200: ABI version, thread-local error text, launch counter.
#include "common.cuh"

namespace mvf {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int64_t& launch_counter() {
  static thread_local int64_t n = 0;
  return n;
}

int sm_count() {
  static thread_local int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

}  // namespace mvf

extern "C" int mvf_abi_version(void) { return MVF_ABI_VERSION; }

extern "C" const char* mvf_last_error(void) { return mvf::error_buffer(); }

extern "C" int64_t mvf_launch_count(int reset) {
  int64_t n = mvf::launch_counter();
  if (reset) mvf::launch_counter() = 0;
  return n;
}
