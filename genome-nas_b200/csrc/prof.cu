// Launch counter and optional per-kernel CUDA-event timing (used by bench.py for the roofline line).
#include "common.cuh"
#include "../../include/gnas.h"
#include <algorithm>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace gnas {
namespace {
struct Rec { int name; 
#ifndef GNAS_EMUL
  cudaEvent_t a, b;
#endif
};
long long g_launches = 0;
bool g_enabled = false;
std::vector<std::string> g_names;
std::map<std::string, int> g_index;
std::vector<Rec> g_recs;

std::string tidy(const char* kernel, const char* where) {
  std::string k(kernel);
  while (!k.empty() && (k.front() == '(' || k.front() == ' ')) k.erase(k.begin());
  while (!k.empty() && (k.back() == ')' || k.back() == ' ')) k.pop_back();
  // template launchers: append the "[with ...]" part of __PRETTY_FUNCTION__
  std::string w(where);
  const size_t p = w.find("[with ");
  if (p != std::string::npos && k.find('<') != std::string::npos) {
    const size_t q = w.find(']', p);
    const size_t lt = k.find('<');
    k = k.substr(0, lt) + " " + w.substr(p, q == std::string::npos ? std::string::npos : q - p + 1);
  }
  return k;
}
}  // namespace

static int g_side_on = 1;

SideStream side_stream(int which) {
  SideStream none; none.st = nullptr; none.fork = nullptr; none.join[0] = none.join[1] = nullptr;
#ifdef GNAS_EMUL
  (void)which;
  return none;
#else
  constexpr int kWhich = 4;
  static SideStream table[kWhich][kMaxDevices];
  static bool made[kWhich][kMaxDevices] = {{false}};
  static int enabled[kWhich] = {-1, -1, -1, -1};
  if (which < 0 || which >= kWhich || !g_side_on) return none;
  if (enabled[which] < 0) { const char* e = getenv(which == 0 ? "GNAS_RHN_SIDE" : "GNAS_WG_SIDE"); enabled[which] = e ? atoi(e) : 1; }
  int dev = 0;
  if (!enabled[which] || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return none;
  if (!made[which][dev]) {
    SideStream s;
    if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) return none;
    cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.join[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.join[1], cudaEventDisableTiming);
    table[which][dev] = s;
    made[which][dev] = true;
  }
  return table[which][dev];
#endif
}

void prof_before(const char* kernel, const char* where, cudaStream_t stream) {
  ++g_launches;
  if (!g_enabled) return;
#ifndef GNAS_EMUL
  const std::string name = tidy(kernel, where);
  auto it = g_index.find(name);
  int id;
  if (it == g_index.end()) { id = (int)g_names.size(); g_names.push_back(name); g_index[name] = id; } else id = it->second;
  Rec r; r.name = id;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, stream);
  g_recs.push_back(r);
#else
  (void)kernel; (void)where; (void)stream;
#endif
}

void prof_after(cudaStream_t stream) {
  if (!g_enabled) return;
#ifndef GNAS_EMUL
  if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, stream);
#else
  (void)stream;
#endif
}
}  // namespace gnas

using namespace gnas;

extern "C" int64_t gnas_kernel_launches(void) { return g_launches; }

extern "C" int gnas_profile_enable(int on) { g_enabled = on != 0; return 0; }

// Median elapsed time (ms) of `n` EMPTY event pairs recorded back to back on `stream`: what an event pair reads
// when nothing runs between its two records.  bench.py subtracts it, measured live in the same process, from
// every bracketed launch (it replaces the round-1 calibration constant).
extern "C" double gnas_profile_event_overhead_ms(int n, void* stream) {
#ifndef GNAS_EMUL
  if (n < 1) n = 1;
  if (n > 4096) n = 4096;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<cudaEvent_t> ev(2 * (size_t)n);
  for (auto& e : ev) cudaEventCreate(&e);
  for (int i = 0; i < n; ++i) { cudaEventRecord(ev[2 * i], st); cudaEventRecord(ev[2 * i + 1], st); }
  cudaStreamSynchronize(st);
  std::vector<float> t((size_t)n, 0.f);
  for (int i = 0; i < n; ++i) cudaEventElapsedTime(&t[i], ev[2 * i], ev[2 * i + 1]);
  for (auto& e : ev) cudaEventDestroy(e);
  std::sort(t.begin(), t.end());
  return (double)t[(size_t)n / 2];
#else
  (void)n; (void)stream;
  return 0.0;
#endif
}

extern "C" int64_t gnas_profile_report(char* buf, int64_t cap) {
  std::string out;
#ifndef GNAS_EMUL
  cudaDeviceSynchronize();
  std::vector<double> ms(g_names.size(), 0.0);
  std::vector<long long> cnt(g_names.size(), 0);
  for (Rec& r : g_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.name] += t; cnt[r.name] += 1; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  for (size_t i = 0; i < g_names.size(); ++i) {
    if (!cnt[i]) continue;
    char line[768];
    snprintf(line, sizeof(line), "%s\t%lld\t%.6f\n", g_names[i].c_str(), cnt[i], ms[i]);
    out += line;
  }
#endif
  g_recs.clear();
  if ((int64_t)out.size() + 1 > cap) return -1;
  if (buf) { memcpy(buf, out.c_str(), out.size() + 1); }
  return (int64_t)out.size();
}

extern "C" int gnas_side_streams(int on) {
  const int prev = gnas::g_side_on;
  gnas::g_side_on = on ? 1 : 0;
  return prev;
}
