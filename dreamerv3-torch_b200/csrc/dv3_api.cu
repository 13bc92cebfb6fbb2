// Library-level entry points: version, thread-local error text, device probe, and the
// single-step wrappers (obs_step / img_step) over the sequence kernels.
#include <stdarg.h>
#include "dv3_common.cuh"

namespace dv3 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return DV3_ERR_CUDA;
}

}  // namespace dv3

extern "C" int dv3_version(void) { return DV3_ABI_VERSION; }

extern "C" const char* dv3_last_error(void) { return dv3::g_err; }

extern "C" int dv3_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  DV3_CHECK_CUDA(cudaGetDevice(&dev));
  DV3_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DV3_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

// RSSM.obs_step (networks.py:174-206) == observe over a length-1 sequence from the caller's state.
extern "C" int dv3_obs_step_fwd(const dv3_rssm_dims* d, const dv3_rssm_params* p,
                                const dv3_observe_io* io, void* stream) {
  DV3_REQUIRE(io, DV3_ERR_NULL, "obs_step_fwd: io is NULL");
  DV3_REQUIRE(io->T == 1, DV3_ERR_BAD_SHAPE, "obs_step_fwd: T=%d, a step has T == 1", io->T);
  return dv3_observe_fwd(d, p, io, stream);
}

// RSSM.img_step (networks.py:208-233) == one given-action transition: H == 2, no actor.
extern "C" int dv3_img_step_fwd(const dv3_rssm_dims* d, const dv3_rssm_params* p,
                                const dv3_imagine_io* io, void* stream) {
  DV3_REQUIRE(io, DV3_ERR_NULL, "img_step_fwd: io is NULL");
  DV3_REQUIRE(io->H == 2, DV3_ERR_BAD_SHAPE, "img_step_fwd: H=%d, a step has H == 2", io->H);
  return dv3_imagine_fwd(d, p, nullptr, io, stream);
}
