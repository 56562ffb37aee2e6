// vi_version / vi_last_error and the error sink shared by all translation units.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void vi_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* vi_last_error(void) { return g_err; }
extern "C" const char* vi_version(void) { return "volinterp_b200 0.1.0 (sm_100a)"; }

// SMs of the current device (per device ordinal: a process may drive several, possibly different, GPUs)
int vi_sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = (dev >= 0 && dev < 64) ? dev : 0;
  if (cache[slot] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[slot] = n;
  }
  return cache[slot];
}

// ---- launch accounting ---------------------------------------------------------------------
#include <vector>
namespace {
struct Span { cudaEvent_t a, b; int kind; };
bool g_prof_on = false;
int64_t g_launches[VI_K_COUNT] = {0};
double g_ms[VI_K_COUNT] = {0};
std::vector<Span> g_spans;
cudaEvent_t g_cur;
int64_t g_rotations = 0, g_systems = 0;
}  // namespace

void vi_prof_count_rotations(int64_t rotations, int64_t systems) { g_rotations += rotations; g_systems += systems; }

extern "C" int vi_profile_counters(int64_t* rotations, int64_t* systems) {
  if (rotations) *rotations = g_rotations;
  if (systems) *systems = g_systems;
  return VI_OK;
}

void vi_prof_launch_begin(int kind, cudaStream_t s) {
  g_launches[kind] += 1;
  if (!g_prof_on) return;
  cudaEventCreate(&g_cur);
  cudaEventRecord(g_cur, s);
}

void vi_prof_launch_end(int kind, cudaStream_t s) {
  if (!g_prof_on) return;
  Span sp;
  sp.a = g_cur;
  sp.kind = kind;
  cudaEventCreate(&sp.b);
  cudaEventRecord(sp.b, s);
  g_spans.push_back(sp);
}

extern "C" int vi_profile_enable(int32_t on) {
  g_prof_on = on != 0;
  return VI_OK;
}

extern "C" int vi_profile_reset(void) {
  for (auto& sp : g_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  g_spans.clear();
  for (int k = 0; k < VI_K_COUNT; ++k) { g_launches[k] = 0; g_ms[k] = 0.0; }
  g_rotations = 0; g_systems = 0;
  return VI_OK;
}

extern "C" int vi_profile_read(double* ms_by_kind, int64_t* launches_by_kind, int32_t nkinds) {
  if (nkinds < VI_K_COUNT) { vi_set_error("need room for %d kinds", (int)VI_K_COUNT); return VI_EINVAL; }
  for (auto& sp : g_spans) {
    cudaError_t e = cudaEventSynchronize(sp.b);
    if (e != cudaSuccess) { vi_set_error("cudaEventSynchronize -> %s", cudaGetErrorString(e)); return VI_ECUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, sp.a, sp.b);
    g_ms[sp.kind] += ms;
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  g_spans.clear();
  for (int k = 0; k < VI_K_COUNT; ++k) {
    if (ms_by_kind) ms_by_kind[k] = g_ms[k];
    if (launches_by_kind) launches_by_kind[k] = g_launches[k];
  }
  return VI_OK;
}
