// Seeded synthetic inputs (std::mt19937 + libstdc++ distributions => same numbers as the reference's
// baspacho/testing/TestingUtils.{h,cpp} and TestingMatGen.{h,cpp} when built with the same libstdc++),
// plus the BAL-shaped and sparse-elimination-stress observation graphs of SURVEY.md §8(d).
#pragma once

#include <cstdint>
#include <limits>
#include <random>
#include <set>
#include <vector>
#include "../host/SparseStructure.h"

namespace BaSpaCho::testing_utils {

using ColumnSets = std::vector<std::set<int64_t>>;

std::vector<int64_t> randomPermutation(size_t size, int64_t seed);
std::vector<int64_t> randomVec(size_t size, int64_t low, int64_t high, int64_t seed);
std::vector<int64_t> randomVec(size_t size, int64_t low, int64_t high, std::mt19937& gen);
template <typename T> std::vector<T> randomData(size_t size, T low, T high, int64_t seed);
template <typename T> std::vector<T> randomData(size_t size, T low, T high, std::mt19937& gen);
std::vector<int64_t> randomPartition(int64_t weight, int64_t low, int64_t high, int64_t seed);

ColumnSets randomCols(int64_t size, double fill, int64_t seed);
ColumnSets joinColums(const ColumnSets& columns, std::vector<int64_t> lumpStart);
ColumnSets csrStructToColumns(const SparseStructure& mat);
SparseStructure columnsToCscStruct(const ColumnSets& columns);
void naiveAddEliminationEntries(ColumnSets& columns, int64_t start, int64_t end);
ColumnSets makeIndependentElimSet(ColumnSets& columns, int64_t start, int64_t end);

struct SparseMatGenerator {
  explicit SparseMatGenerator(int64_t size, int64_t seed = 37);

  void connectRanges(int64_t begin1, int64_t end1, int64_t begin2, int64_t end2, double fill,
                     int64_t maxOffset = std::numeric_limits<int64_t>::max());
  void addSparseConnections(double fill);
  void addSchurSet(int64_t size, double fill);

  static SparseMatGenerator genFlat(int64_t size, double fill, int64_t seed = 37);
  static SparseMatGenerator genLine(int64_t size, double fill, int64_t bandSize, int64_t seed = 37);
  static SparseMatGenerator genMeridians(int64_t num, int64_t lineLen, double fill, int64_t bandSize, int64_t hairLen,
                                         int64_t nPoleHairs, int64_t sPoleHairs, int64_t seed = 37);
  static SparseMatGenerator genGrid(int64_t width, int64_t height, double fill, int64_t connMaxDist, int64_t seed = 37);

  std::mt19937 gen;
  ColumnSets columns;
};

// Bundle-adjustment shaped pattern: numPts point blocks first, numCams camera blocks after; point i is
// observed by k_i = min(numCams, minObs + Poisson(meanExtraObs)) distinct cameras, each drawn with
// probability (1-farProb) from a +-window (wrap-around) neighbourhood of camera floor(i*numCams/numPts)
// and with probability farProb uniformly. Returns CSR lower-triangular block pattern (diag included).
SparseStructure genBundleAdjustment(int64_t numPts, int64_t numCams, int64_t minObs, double meanExtraObs,
                                    int64_t window, double farProb, int64_t seed);

}  // namespace BaSpaCho::testing_utils
