// Small host helpers shared by the symbolic layer, the Solver driver and the backends.
// Role of reference baspacho/baspacho/Utils.h (OpStat :49-121, bisect :154-166, permutation helpers :169-196).
#pragma once

#include <chrono>
#include <cstdint>
#include <functional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#ifdef __CUDACC__
#define BSP_HD __host__ __device__
#else
#define BSP_HD
#endif

namespace BaSpaCho {

[[noreturn]] void throwError(const char* file, int line, const std::string& msg);

std::string secondsToString(double secs, int precision = 2);

struct DefaultSyncOps {
  static void sync() {}
};

// Accumulating timer with an optional per-call callback (args are the op sizes).
// Same observable fields as the reference's OpStat: enabled/numRuns/totTime/maxTime/lastTime/callBack.
template <typename... Args>
struct OpStat {
  using Clock = std::chrono::steady_clock;

  template <typename SyncOps>
  class Instance {
   public:
    Instance() = default;
    Instance(OpStat* s, const Args&... a) : stat_(s), t0_(Clock::now()), args_(a...) {}
    Instance(Instance&& o) noexcept : stat_(o.stat_), t0_(o.t0_), args_(std::move(o.args_)) {
      o.stat_ = nullptr;
    }
    Instance(const Instance&) = delete;
    Instance& operator=(const Instance&) = delete;
    ~Instance() {
      if (!stat_) return;
      SyncOps::sync();
      double dt = std::chrono::duration<double>(Clock::now() - t0_).count();
      stat_->numRuns++;
      stat_->lastTime = dt;
      stat_->totTime += dt;
      if (dt > stat_->maxTime) stat_->maxTime = dt;
      if (stat_->callBack) {
        std::apply([&](const Args&... a) { stat_->callBack(dt, a...); }, args_);
      }
    }

   private:
    OpStat* stat_ = nullptr;
    typename Clock::time_point t0_{};
    std::tuple<Args...> args_;
  };

  template <typename SyncOps = DefaultSyncOps>
  Instance<SyncOps> instance(const Args&... args) {
    if (!enabled) return Instance<SyncOps>();
    return Instance<SyncOps>(this, args...);
  }

  void reset() { numRuns = 0, totTime = maxTime = lastTime = 0.0; }

  std::string toString() const {
    std::stringstream ss;
    ss << "#=" << numRuns << ", time=" << secondsToString(totTime)
       << ", last=" << secondsToString(lastTime) << ", max=" << secondsToString(maxTime);
    return ss.str();
  }

  bool enabled = true;
  int64_t numRuns = 0;
  double totTime = 0, maxTime = 0, lastTime = 0;
  std::function<void(double, const Args&...)> callBack;
};

template <typename T>
bool isStrictlyIncreasing(const std::vector<T>& v, size_t b, size_t e) {
  for (size_t i = b + 1; i < e; i++)
    if (!(v[i - 1] < v[i])) return false;
  return true;
}

template <typename T>
bool isWeaklyIncreasing(const std::vector<T>& v, size_t b, size_t e) {
  for (size_t i = b + 1; i < e; i++)
    if (v[i] < v[i - 1]) return false;
  return true;
}

// largest index a in [0,size) with array[a] <= needle (0 if none); array sorted ascending
BSP_HD inline int64_t bisect(const int64_t* array, int64_t size, int64_t needle) {
  int64_t lo = 0, hi = size;
  while (hi - lo > 1) {
    int64_t mid = lo + ((hi - lo) >> 1);
    if (array[mid] <= needle) lo = mid; else hi = mid;
  }
  return lo;
}

template <class It>
inline void shiftConcat(std::vector<int64_t>& target, int64_t shift, It first, It last) {
  for (; first != last; ++first) target.push_back(*first + shift);
}

// out[perm[i]] = w[i]
template <typename T, typename It>
void leftPermute(It out, const std::vector<int64_t>& perm, const std::vector<T>& w) {
  for (size_t i = 0; i < perm.size(); i++) out[perm[i]] = w[i];
}

std::vector<int64_t> composePermutations(const std::vector<int64_t>& v, const std::vector<int64_t>& w);
std::vector<int64_t> inversePermutation(const std::vector<int64_t>& v);

// exclusive prefix sum in place over v[0..n-1], total stored in the last element and returned
int64_t cumSumVec(std::vector<int64_t>& v);
// undo the "advance the pointers while filling" idiom: shift right by one, v[downTo]=value
void rewindVec(std::vector<int64_t>& v, int64_t downTo = 0, int64_t value = 0);

}  // namespace BaSpaCho
