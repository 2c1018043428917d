// Error plumbing and ABI version of libmvptr_b200.so.
#include <stdarg.h>

#include <vector>

#include "common.cuh"

namespace mvptr {
unsigned long long g_launch_count = 0;
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mvptr

namespace mvptr {
// Optional per-launch CUDA-event profiler (bench.py's roofline leg).  Disabled by default; when
// enabled every entry point brackets its launches with events on the launching stream.
struct ProfRec {
  const char* name;
  double work;
  cudaEvent_t e0, e1;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
ProfScope::ProfScope(const char* name, double work, cudaStream_t s) : active_(g_prof_on), stream_(s) {
  if (!active_) return;
  ProfRec r{name, work, nullptr, nullptr};
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, s);
  g_prof.push_back(r);
  index_ = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (active_) cudaEventRecord(g_prof[index_].e1, stream_);
}
}  // namespace mvptr

extern "C" int mvptr_profile_enable(int on) {
  mvptr::g_prof_on = on != 0;
  return 0;
}
// Synchronises, then writes up to `cap` records (name pointer, work, milliseconds); returns the count.
extern "C" int mvptr_profile_collect(const char** names, double* work, float* ms, int cap) {
  using namespace mvptr;
  cudaDeviceSynchronize();
  int n = 0;
  for (auto& r : g_prof) {
    if (n < cap) {
      float t = 0.f;
      cudaEventElapsedTime(&t, r.e0, r.e1);
      names[n] = r.name;
      work[n] = r.work;
      ms[n] = t;
      ++n;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return n;
}

extern "C" unsigned long long mvptr_launch_count(void) { return mvptr::g_launch_count; }
extern "C" int mvptr_abi_version(void) { return MVPTR_ABI_VERSION; }
extern "C" const char* mvptr_last_error(void) { return mvptr::g_err; }

extern "C" int mvptr_set_epoch_gemm(const uint32_t*, void*);
extern "C" int mvptr_set_epoch_rows(const uint32_t*, void*);
extern "C" int mvptr_set_epoch_attn(const uint32_t*, void*);
extern "C" int mvptr_set_epoch_wra(const uint32_t*, void*);
extern "C" int mvptr_set_epoch_attn_tc(const uint32_t*, void*);
extern "C" int mvptr_set_epoch_gemm_addr(uint32_t**);
extern "C" int mvptr_set_epoch_rows_addr(uint32_t**);
extern "C" int mvptr_set_epoch_attn_addr(uint32_t**);
extern "C" int mvptr_set_epoch_wra_addr(uint32_t**);
extern "C" int mvptr_set_epoch_attn_tc_addr(uint32_t**);
// src: device or PINNED host pointer (read when the copy executes, i.e. at graph replay time)
extern "C" int mvptr_set_dropout_epoch(const uint32_t* src, void* stream) {
  if (int rc = mvptr_set_epoch_gemm(src, stream)) return rc;
  if (int rc = mvptr_set_epoch_rows(src, stream)) return rc;
  if (int rc = mvptr_set_epoch_wra(src, stream)) return rc;
  if (int rc = mvptr_set_epoch_attn_tc(src, stream)) return rc;  // every translation unit that hashes dropout masks
  return mvptr_set_epoch_attn(src, stream);
}

namespace mvptr {
// One thread: the per-replay parameters of a captured training step.  `ring` is a ring of 16-byte slots
// {float lr, float step, uint32 epoch, pad} in PINNED HOST memory (read over PCIe, system-scope loads) that the
// host fills for replay n at slot n % slots BEFORE launching it; `counter` is a device word counting the
// replays that have executed.  The host may run any number (< slots) of replays ahead: replay n always reads
// slot n, never "whatever the host wrote last" (which is what per-replay memcpy nodes from one host word did).
__global__ void step_params_kernel(const uint32_t* __restrict__ ring, uint32_t slots, uint32_t* __restrict__ counter,
                                   float* __restrict__ dyn, uint32_t* e0, uint32_t* e1, uint32_t* e2, uint32_t* e3,
                                   uint32_t* e4) {
  const uint32_t n = *counter;
  const uint32_t* s = ring + (size_t)(n % slots) * 4;
  uint32_t lr, st, ep;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(lr) : "l"(s) : "memory");
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(st) : "l"(s + 1) : "memory");
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(ep) : "l"(s + 2) : "memory");
  if (dyn != nullptr) {
    dyn[0] = __uint_as_float(lr);
    dyn[1] = __uint_as_float(st);
  }
  *e0 = ep;
  *e1 = ep;
  *e2 = ep;
  *e3 = ep;
  *e4 = ep;
  *counter = n + 1;
}
}  // namespace mvptr

extern "C" int mvptr_step_params(const void* ring, int slots, uint32_t* counter, float* dyn_lr_step, void* stream) {
  using namespace mvptr;
  if (!ring || !counter || slots <= 0) MVPTR_FAIL(MVPTR_ERR_ARG, "step_params: null ring / counter or no slots");
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, ring);
  if (e != cudaSuccess) MVPTR_FAIL(MVPTR_ERR_CUDA, "step_params: %s", cudaGetErrorString(e));
  const void* dev_ring = ring;
  if (attr.type == cudaMemoryTypeHost) dev_ring = attr.devicePointer;  // pinned: the device-visible alias
  else if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)
    MVPTR_FAIL(MVPTR_ERR_ARG, "step_params: the ring must be pinned host memory or device memory");
  if (!dev_ring) MVPTR_FAIL(MVPTR_ERR_ARG, "step_params: pinned ring has no device mapping");
  static uint32_t* addr[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  if (!addr[0]) {
    uint32_t* a[5];
    if (int rc = mvptr_set_epoch_gemm_addr(&a[0])) return rc;
    if (int rc = mvptr_set_epoch_rows_addr(&a[1])) return rc;
    if (int rc = mvptr_set_epoch_attn_addr(&a[2])) return rc;
    if (int rc = mvptr_set_epoch_wra_addr(&a[3])) return rc;
    if (int rc = mvptr_set_epoch_attn_tc_addr(&a[4])) return rc;
    for (int i = 4; i >= 0; --i) addr[i] = a[i];
  }
  step_params_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(static_cast<const uint32_t*>(dev_ring), (uint32_t)slots, counter,
                                                        dyn_lr_step, addr[0], addr[1], addr[2], addr[3], addr[4]);
  MVPTR_CHECK_LAUNCH("step_params");
  return 0;
}
