// Error reporting for the C ABI (include/cirs_b200.h): thread-local message + ABI version.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "../../include/cirs_b200.h"

static thread_local char g_err[512] = "";

void cirs_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

static long long g_launches = 0;
void cirs_note_launch(void) { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }
extern "C" int64_t cirs_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// ---- optional per-kernel timing (bench.py: roofline.achieved is measured with these events, live)
#include <map>
#include <string>
#include <vector>
#include <mutex>
struct ProfRec { const char* name; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_pool;
static std::mutex g_prof_mu;
static cudaEvent_t prof_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
bool cirs_profile_begin(const char* name, cudaStream_t st) {
  if (!g_prof_on) return false;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{name, prof_event(), prof_event()};
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  return true;
}
void cirs_profile_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, st);
}
extern "C" void cirs_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
}
// Synchronises the device, writes "name count total_ms\n" lines into buf (truncated at n), clears the records.
extern "C" int cirs_profile_report(char* buf, int n) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<long long, double>> agg;
  if (const char* path = getenv("CIRS_PROFILE_TIMELINE")) {   // debugging aid: "name start_us dur_us" per launch, in issue order
    if (FILE* f = fopen(path, "a")) {
      for (auto& r : g_prof) {
        float t0 = 0.f, dt = 0.f;
        cudaEventElapsedTime(&t0, g_prof.front().a, r.a);
        cudaEventElapsedTime(&dt, r.a, r.b);
        fprintf(f, "%s %.2f %.2f\n", r.name, t0 * 1e3, dt * 1e3);
      }
      fprintf(f, "--\n");
      fclose(f);
    }
  }
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { auto& x = agg[r.name]; x.first++; x.second += ms; }
    g_pool.push_back(r.a); g_pool.push_back(r.b);
  }
  g_prof.clear();
  std::string out;
  for (auto& kv : agg) out += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
  if (buf && n > 0) { strncpy(buf, out.c_str(), n - 1); buf[n - 1] = 0; }
  return (int)out.size();
}

extern "C" const char* cirs_last_error(void) { return g_err; }
extern "C" int cirs_abi_version(void) { return CIRS_ABI_VERSION; }
