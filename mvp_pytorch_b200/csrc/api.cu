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
// src: device or PINNED host pointer (read when the copy executes, i.e. at graph replay time)
extern "C" int mvptr_set_dropout_epoch(const uint32_t* src, void* stream) {
  if (int rc = mvptr_set_epoch_gemm(src, stream)) return rc;
  if (int rc = mvptr_set_epoch_rows(src, stream)) return rc;
  return mvptr_set_epoch_attn(src, stream);
}
