// Library-level entry points: version, error string, device query.
#include "tg_common.cuh"

#include <cstdlib>
#include <mutex>

static thread_local char g_err[512] = "";

void tg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool tg_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TECOGAN_B200_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

extern "C" {

int tg_version(void) { return TG_ABI_VERSION; }

const char* tg_last_error_string(void) { return g_err; }

int tg_device_sm_count(int* out_sm_count) {
  TG_REQUIRE(out_sm_count != nullptr, TG_E_INVALID, "device_sm_count: null pointer");
  static int cached[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { tg_set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
  if (dev >= 0 && dev < 64 && cached[dev] > 0) { *out_sm_count = cached[dev]; return TG_OK; }
  int sms = 0;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { tg_set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  if (dev >= 0 && dev < 64) cached[dev] = sms;
  *out_sm_count = sms;
  return TG_OK;
}

}  // extern "C"
