// Per-kernel-class event profiler of the B200 backend (see B200Defs.h).
#include <map>
#include <mutex>
#include <sstream>
#include <cstdlib>
#include "B200Defs.h"

namespace BaSpaCho {
namespace b200 {
namespace {

struct Rec {
  int cls;
  double flops, bytes;
  cudaEvent_t e0, e1;
  cudaStream_t st;
};
struct State {
  bool enabled = false;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  std::mutex mu;
  cudaEvent_t getEvent() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    B200_CUDA(cudaEventCreate(&e));
    return e;
  }
};
State& state() {
  static State s;
  return s;
}
const char* kNames[KC_COUNT] = {"gemm",     "potrf_block", "trsm_block",  "elim_factor", "elim_gather",
                                "assemble", "solve_elim",  "solve_dense", "other",       "lump_chol"};

}  // namespace

void ensureDynSmem(const void* kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return;
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = done[{kernel, dev}];
  if (have >= bytes) return;
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  have = bytes;
}

void profileEnable(bool on) { state().enabled = on; }
bool profileEnabled() { return state().enabled; }

void profileBegin(cudaStream_t st, int cls, double flops, double bytes) {
  State& s = state();
  std::lock_guard<std::mutex> lock(s.mu);
  Rec r{cls, flops, bytes, s.getEvent(), s.getEvent(), st};
  B200_CUDA(cudaEventRecord(r.e0, st));
  s.recs.push_back(r);
}

void profileEnd(cudaStream_t st) {
  State& s = state();
  std::lock_guard<std::mutex> lock(s.mu);
  B200_CUDA(cudaEventRecord(s.recs.back().e1, st));
}

std::string profileReportJson() {
  State& s = state();
  std::lock_guard<std::mutex> lock(s.mu);
  B200_CUDA(cudaDeviceSynchronize());
  double ms[KC_COUNT] = {0}, flops[KC_COUNT] = {0}, bytes[KC_COUNT] = {0};
  int64_t n[KC_COUNT] = {0};
  // BSPB200_PROFILE_TIMELINE=1: every record as [class, stream index, start ms, end ms, flops] relative to the first
  // record (events of different streams share the device's time base): where the lanes of a tree level overlap and
  // where they wait
  const char* tl = getenv("BSPB200_PROFILE_TIMELINE");
  const bool timeline = tl && atoi(tl) != 0 && !s.recs.empty();
  std::stringstream tss;
  tss.precision(7);
  std::map<cudaStream_t, int> streamIdx;
  for (Rec& r : s.recs) {
    float t = 0;
    B200_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms[r.cls] += t, flops[r.cls] += r.flops, bytes[r.cls] += r.bytes, n[r.cls]++;
    if (timeline) {
      float t0 = 0;
      B200_CUDA(cudaEventElapsedTime(&t0, s.recs.front().e0, r.e0));
      const int si = streamIdx.emplace(r.st, (int)streamIdx.size()).first->second;
      tss << (&r == &s.recs.front() ? "" : ", ") << "[" << r.cls << ", " << si << ", " << t0 << ", " << t0 + t << ", " << r.flops << "]";
    }
  }
  for (Rec& r : s.recs) {
    s.pool.push_back(r.e0);
    s.pool.push_back(r.e1);
  }
  s.recs.clear();
  std::stringstream ss;
  ss.precision(10);
  ss << "{";
  if (timeline) ss << "\"timeline\": [" << tss.str() << "], ";
  for (int c = 0; c < KC_COUNT; c++)
    ss << (c ? ", " : "") << "\"" << kNames[c] << "\": {\"launches\": " << n[c] << ", \"ms\": " << ms[c]
       << ", \"flops\": " << flops[c] << ", \"bytes\": " << bytes[c] << "}";
  ss << "}";
  return ss.str();
}

}  // namespace b200
}  // namespace BaSpaCho
