// See TestingUtils.h. Behaviour of the reference generators: baspacho/testing/TestingUtils.cpp:16-230,
// baspacho/testing/TestingMatGen.cpp:14-201 (RNG consumption order preserved so seeds give the same patterns).
#include "TestingUtils.h"
#include <algorithm>
#include <numeric>
#include "../host/DebugMacros.h"

namespace BaSpaCho::testing_utils {

using std::vector;

vector<int64_t> randomPermutation(size_t size, int64_t seed) {
  std::mt19937 gen(seed);
  vector<int64_t> p(size);
  std::iota(p.begin(), p.end(), 0);
  std::shuffle(p.begin(), p.end(), gen);
  return p;
}

vector<int64_t> randomVec(size_t size, int64_t low, int64_t high, int64_t seed) {
  std::mt19937 gen(seed);
  return randomVec(size, low, high, gen);
}

vector<int64_t> randomVec(size_t size, int64_t low, int64_t high, std::mt19937& gen) {
  std::uniform_int_distribution<int64_t> pick(low, high);
  vector<int64_t> v(size);
  for (auto& x : v) x = pick(gen);
  return v;
}

template <typename T>
vector<T> randomData(size_t size, T low, T high, int64_t seed) {
  std::mt19937 gen(seed);
  return randomData(size, low, high, gen);
}

template <typename T>
vector<T> randomData(size_t size, T low, T high, std::mt19937& gen) {
  std::uniform_real_distribution<> u(low, high);  // always drawn in double, then narrowed
  vector<T> v(size);
  for (auto& x : v) x = (T)u(gen);
  return v;
}

template vector<double> randomData(size_t, double, double, int64_t);
template vector<float> randomData(size_t, float, float, int64_t);
template vector<double> randomData(size_t, double, double, std::mt19937&);
template vector<float> randomData(size_t, float, float, std::mt19937&);

vector<int64_t> randomPartition(int64_t weight, int64_t low, int64_t high, int64_t seed) {
  std::mt19937 gen(seed);
  std::uniform_int_distribution<int64_t> pick(low, high);
  vector<int64_t> parts;
  while (weight > 0) {
    parts.push_back(std::min(weight, pick(gen)));
    weight -= parts.back();
  }
  return parts;
}

ColumnSets randomCols(int64_t size, double fill, int64_t seed) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<> u(0.0, 1.0);
  ColumnSets cols(size);
  for (int64_t i = 0; i < size; i++) {
    cols[i].insert(i);
    for (int64_t j = i + 1; j < size; j++)
      if (u(gen) < fill) cols[i].insert(j);
  }
  return cols;
}

ColumnSets joinColums(const ColumnSets& columns, vector<int64_t> lumpStart) {
  BASPACHO_CHECK_EQ(lumpStart.back(), (int64_t)columns.size());
  ColumnSets out;
  for (size_t a = 0; a + 1 < lumpStart.size(); a++) {
    std::set<int64_t> merged;
    for (int64_t i = lumpStart[a]; i < lumpStart[a + 1]; i++) merged.insert(columns[i].begin(), columns[i].end());
    out.push_back(std::move(merged));
  }
  return out;
}

ColumnSets csrStructToColumns(const SparseStructure& mat) {
  ColumnSets cols(mat.order());
  for (int64_t i = 0; i < mat.order(); i++)
    for (int64_t k = mat.ptrs[i]; k < mat.ptrs[i + 1]; k++) cols[mat.inds[k]].insert(i);
  return cols;
}

SparseStructure columnsToCscStruct(const ColumnSets& columns) {
  SparseStructure s;
  s.ptrs.reserve(columns.size() + 1);
  for (const auto& c : columns) {
    s.ptrs.push_back((int64_t)s.inds.size());
    s.inds.insert(s.inds.end(), c.begin(), c.end());
  }
  s.ptrs.push_back((int64_t)s.inds.size());
  return s;
}

void naiveAddEliminationEntries(ColumnSets& columns, int64_t start, int64_t end) {
  BASPACHO_CHECK_LE(end, (int64_t)columns.size());
  for (int64_t i = start; i < end; i++) {
    const auto& col = columns[i];
    BASPACHO_CHECK(!col.empty() && *col.begin() == i);  // diagonal block expected first
    for (auto a = std::next(col.begin()); a != col.end(); ++a)
      for (auto b = std::next(a); b != col.end(); ++b) columns[*a].insert(*b);
  }
}

ColumnSets makeIndependentElimSet(ColumnSets& columns, int64_t start, int64_t end) {
  ColumnSets out(columns.size());
  for (int64_t i = 0; i < (int64_t)columns.size(); i++) {
    if (i < start || i >= end) {
      out[i] = columns[i];
      continue;
    }
    out[i].insert(i);
    for (int64_t r : columns[i])
      if (r >= end) out[i].insert(r);
  }
  return out;
}

// -------------------------------------------------------------------------------------------------
SparseMatGenerator::SparseMatGenerator(int64_t size, int64_t seed) : gen(seed), columns(size) {
  for (int64_t i = 0; i < size; i++) columns[i].insert(i);
}

void SparseMatGenerator::connectRanges(int64_t begin1, int64_t end1, int64_t begin2, int64_t end2, double fill,
                                       int64_t maxOffset) {
  BASPACHO_CHECK_GE(begin1, 0);
  BASPACHO_CHECK_GE(begin2, 0);
  BASPACHO_CHECK_LE(end1, (int64_t)columns.size());
  BASPACHO_CHECK_LE(end2, (int64_t)columns.size());
  if (begin1 > begin2) {  // canonical order: first range starts first
    connectRanges(begin2, end2, begin1, end1, fill, maxOffset);
    return;
  }
  if (end1 > end2) connectRanges(begin2, end2, end2, end1, fill, maxOffset);  // tail of range 1 past range 2

  std::uniform_real_distribution<> u(0.0, 1.0);
  for (int64_t i = begin1; i < end1; i++) {
    int64_t jFirst = i + std::min(maxOffset, std::max<int64_t>(begin2 - i, 1));
    int64_t jEnd = i + std::min(maxOffset, end2 - i);
    for (int64_t j = jFirst; j < jEnd; j++)
      if (fill >= 1.0 || fill > u(gen)) columns[i].insert(j);
  }
}

void SparseMatGenerator::addSparseConnections(double fill) {
  connectRanges(0, (int64_t)columns.size(), 0, (int64_t)columns.size(), fill);
}

void SparseMatGenerator::addSchurSet(int64_t size, double fill) {
  const int64_t oldSize = (int64_t)columns.size();
  ColumnSets grown(size + oldSize);
  std::uniform_real_distribution<> u(0.0, 1.0);
  for (int64_t i = 0; i < size; i++) {
    grown[i].insert(i);
    for (int64_t j = size; j < size + oldSize; j++)
      if (fill >= 1.0 || fill > u(gen)) grown[i].insert(j);
  }
  for (int64_t i = 0; i < oldSize; i++)
    for (int64_t j : columns[i]) grown[i + size].insert(j + size);
  columns.swap(grown);
}

SparseMatGenerator SparseMatGenerator::genFlat(int64_t size, double fill, int64_t seed) {
  SparseMatGenerator g(size, seed);
  g.connectRanges(0, size, 0, size, fill);
  return g;
}

SparseMatGenerator SparseMatGenerator::genLine(int64_t size, double fill, int64_t /*bandSize (ignored, as in the reference)*/,
                                               int64_t seed) {
  return genFlat(size, fill, seed);
}

SparseMatGenerator SparseMatGenerator::genMeridians(int64_t num, int64_t lineLen, double fill, int64_t bandSize,
                                                    int64_t hairLen, int64_t nPoleHairs, int64_t sPoleHairs, int64_t seed) {
  const int64_t totHairs = nPoleHairs + sPoleHairs;
  const int64_t hairsBase = lineLen * num;
  BASPACHO_CHECK_LE(bandSize, lineLen);
  BASPACHO_CHECK_LE(bandSize, hairLen);
  SparseMatGenerator g(hairsBase + hairLen * totHairs, seed);
  auto band = [&](int64_t a, int64_t b) { g.connectRanges(a, a + bandSize, b, b + bandSize, fill, bandSize); };
  auto meridian = [&](int64_t i) { return lineLen * i; };
  auto hair = [&](int64_t h) { return hairsBase + hairLen * h; };

  for (int64_t i = 0; i < num; i++) g.connectRanges(meridian(i), meridian(i) + lineLen, meridian(i), meridian(i) + lineLen, fill, bandSize);
  for (int64_t h = 0; h < totHairs; h++) g.connectRanges(hair(h), hair(h) + hairLen, hair(h), hair(h) + hairLen, fill, bandSize);
  // meridians meet at the poles (start = north, end = south)
  for (int64_t i = 0; i < num; i++)
    for (int64_t j = 0; j < i; j++) {
      band(meridian(i), meridian(j));
      band(meridian(i) + lineLen - bandSize, meridian(j) + lineLen - bandSize);
    }
  // hairs attach to the meridian ends
  for (int64_t i = 0; i < num; i++) {
    for (int64_t h = 0; h < nPoleHairs; h++) band(meridian(i), hair(h));
    for (int64_t h = 0; h < sPoleHairs; h++) band(meridian(i) + lineLen - bandSize, hair(h + nPoleHairs));
  }
  // hairs of the same pole meet each other
  for (int64_t h = 0; h < nPoleHairs; h++)
    for (int64_t k = 0; k < h; k++) band(hair(k), hair(h));
  for (int64_t h = 0; h < sPoleHairs; h++)
    for (int64_t k = 0; k < h; k++)
      band(hair(h + nPoleHairs), hair(h + nPoleHairs));  // sic: the reference connects the hair with itself here
  return g;
}

SparseMatGenerator SparseMatGenerator::genGrid(int64_t width, int64_t height, double fill, int64_t connMaxDist, int64_t seed) {
  SparseMatGenerator g(width * height, seed);
  std::uniform_real_distribution<> u(0.0, 1.0);
  for (int64_t i = 0; i < width; i++)
    for (int64_t j = 0; j < height; j++) {
      const int64_t me = i * height + j;
      for (int64_t i2 = std::max<int64_t>(i - connMaxDist, 0); i2 < std::min(i + connMaxDist + 1, width); i2++)
        for (int64_t j2 = std::max<int64_t>(j - connMaxDist, 0); j2 < std::min(j + connMaxDist + 1, height); j2++) {
          if (i2 == i && j2 == j) continue;
          if (fill >= 1.0 || fill > u(g.gen)) {
            int64_t other = i2 * height + j2;
            g.columns[std::min(me, other)].insert(std::max(me, other));
          }
        }
    }
  return g;
}

// -------------------------------------------------------------------------------------------------
SparseStructure genBundleAdjustment(int64_t numPts, int64_t numCams, int64_t minObs, double meanExtraObs,
                                    int64_t window, double farProb, int64_t seed) {
  std::mt19937 gen(seed);
  std::poisson_distribution<int64_t> extra(meanExtraObs);
  std::uniform_real_distribution<> u(0.0, 1.0);
  std::uniform_int_distribution<int64_t> anyCam(0, numCams - 1);
  std::uniform_int_distribution<int64_t> nearOff(-window, window);

  // camera rows of the CSR lower triangle: points seen (ascending), then the camera itself
  vector<vector<int64_t>> seenBy(numCams);
  vector<int64_t> mine;
  for (int64_t i = 0; i < numPts; i++) {
    int64_t k = std::min(numCams, minObs + extra(gen));
    int64_t centre = (int64_t)((__int128)i * numCams / numPts);
    mine.clear();
    while ((int64_t)mine.size() < k) {
      int64_t c;
      if (window <= 0 || u(gen) < farProb) {
        c = anyCam(gen);
      } else {
        c = ((centre + nearOff(gen)) % numCams + numCams) % numCams;
      }
      if (std::find(mine.begin(), mine.end(), c) == mine.end()) mine.push_back(c);
    }
    for (int64_t c : mine) seenBy[c].push_back(i);
  }
  SparseStructure ss;
  ss.ptrs.resize(numPts + numCams + 1);
  for (int64_t i = 0; i <= numPts; i++) ss.ptrs[i] = i;
  ss.inds.resize(numPts);
  std::iota(ss.inds.begin(), ss.inds.end(), 0);
  for (int64_t c = 0; c < numCams; c++) {
    ss.inds.insert(ss.inds.end(), seenBy[c].begin(), seenBy[c].end());  // already ascending (points visited in order)
    ss.inds.push_back(numPts + c);
    ss.ptrs[numPts + c + 1] = (int64_t)ss.inds.size();
  }
  return ss;
}

}  // namespace BaSpaCho::testing_utils
