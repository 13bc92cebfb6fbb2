// Library-level entry points: version, thread-local error text, device probe, and the
// single-step wrappers (obs_step / img_step) over the sequence kernels.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>
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

static std::atomic<unsigned> g_env_epoch{1};

const char* env_cached(const char* name, EnvSlot& slot) {
  const unsigned ep = g_env_epoch.load(std::memory_order_relaxed);
  if (slot.epoch != ep) {
    const char* e = getenv(name);
    slot.has = e != nullptr;
    slot.val[0] = 0;
    if (e) { strncpy(slot.val, e, sizeof(slot.val) - 1); slot.val[sizeof(slot.val) - 1] = 0; }
    slot.epoch = ep;
  }
  return slot.has ? slot.val : nullptr;
}

// 0 = off (default), 1 = every launch_pdl() launch, 2 = only launches with <= 64 KB of dynamic shared
// memory (the row kernels: resident early they do not take an SM away from a 200 KB GEMM CTA)
int pdl_mode() {
  const char* e = DV3_ENV("DV3_PDL");
  return (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
}
bool pdl_enabled() { return pdl_mode() == 1; }   // measured: 16.76 ms/step with it, 16.33 without -> off

int sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = sms[dev & 63];
  if (!v) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

// ---- launch counter + optional GEMM timing --------------------------------------------
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { cudaEvent_t a, b; int kind; double flops; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t g_cur = nullptr;

static cudaEvent_t take_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
bool prof_on() { return g_prof_on; }
void prof_begin(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_cur = take_event();
  cudaEventRecord(g_cur, st);
}
void prof_end(cudaStream_t st, int kind, double flops) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t b = take_event();
  cudaEventRecord(b, st);
  g_prof.push_back({g_cur, b, kind, flops});
  g_cur = nullptr;
}

}  // namespace dv3

extern "C" long long dv3_launch_count(void) { return dv3::g_launches.load(); }

extern "C" void dv3_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(dv3::g_prof_mu);
  dv3::g_prof_on = on != 0;
}

// Sums the CUDA-event durations recorded since the last read, per GEMM kind (0 skinny, 1 tiled),
// waits for the recorded events, and clears the record.  Arrays have 2 entries.
extern "C" int dv3_prof_read(double* ms, double* flops, long long* launches) {
  std::lock_guard<std::mutex> lk(dv3::g_prof_mu);
  for (int k = 0; k < 2; ++k) { ms[k] = 0; flops[k] = 0; launches[k] = 0; }
  for (auto& r : dv3::g_prof) {
    DV3_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    DV3_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.kind] += t; flops[r.kind] += r.flops; launches[r.kind] += 1;
    dv3::g_pool.push_back(r.a); dv3::g_pool.push_back(r.b);
  }
  dv3::g_prof.clear();
  return 0;
}

extern "C" int dv3_version(void) { return DV3_ABI_VERSION; }

extern "C" void dv3_reload_env(void) { dv3::g_env_epoch.fetch_add(1); }

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
