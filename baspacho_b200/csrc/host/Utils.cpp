#include "Utils.h"
#include <cmath>
#include <cstdio>
#include "DebugMacros.h"

namespace BaSpaCho {

void throwError(const char* file, int line, const std::string& msg) {
  std::stringstream ss;
  ss << "[" << file << ":" << line << "] Check failed: " << msg;
  throw std::runtime_error(ss.str());
}

std::string secondsToString(double secs, int precision) {
  char buf[64];
  double us = secs * 1e6;
  if (us < 1000.0) {
    snprintf(buf, sizeof buf, "%.0fus", us);
  } else if (us < 1e5) {
    snprintf(buf, sizeof buf, "%.*fms", precision, us * 1e-3);
  } else if (secs < 60.0) {
    snprintf(buf, sizeof buf, "%.*fs", precision, secs);
  } else if (secs < 3600.0) {
    snprintf(buf, sizeof buf, "%dm%ds", int(secs / 60), int(std::lround(secs)) % 60);
  } else {
    snprintf(buf, sizeof buf, "%dh%dm", int(secs / 3600), (int(secs) % 3600) / 60);
  }
  return buf;
}

std::vector<int64_t> composePermutations(const std::vector<int64_t>& v, const std::vector<int64_t>& w) {
  BASPACHO_CHECK_EQ(v.size(), w.size());
  std::vector<int64_t> out(w.size());
  for (size_t i = 0; i < w.size(); i++) out[i] = v[w[i]];
  return out;
}

std::vector<int64_t> inversePermutation(const std::vector<int64_t>& p) {
  std::vector<int64_t> inv(p.size());
  for (size_t i = 0; i < p.size(); i++) inv[p[i]] = (int64_t)i;
  return inv;
}

int64_t cumSumVec(std::vector<int64_t>& v) {
  int64_t acc = 0;
  for (size_t i = 0; i + 1 < v.size(); i++) {
    int64_t x = v[i];
    v[i] = acc;
    acc += x;
  }
  v.back() = acc;
  return acc;
}

void rewindVec(std::vector<int64_t>& v, int64_t downTo, int64_t value) {
  for (int64_t i = (int64_t)v.size() - 1; i > downTo; i--) v[i] = v[i - 1];
  v[downTo] = value;
}

}  // namespace BaSpaCho
