// Implementation of the C ABI declared in include/baspacho_b200.h over the C++ host API (Solver.h).
// Included once by capi.cpp (prefix bspb200_, product library) and once by oracle/oracle_capi.cpp
// (prefix oracle_, CPU checker library) - the entry points are identical, only the backends linked
// behind createSolver differ.
#pragma once

#ifndef CAPI
#error "define CAPI(name) before including capi_impl.h"
#endif

#include <sstream>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <unordered_set>
#include <vector>
#include "host/ComputationModel.h"
#include "host/Solver.h"
#include "host/SparseStructure.h"
#include "testing/TestingUtils.h"

namespace capi_detail {

using namespace BaSpaCho;

inline std::string& lastError() {
  static thread_local std::string err;
  return err;
}

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    lastError() = e.what();
  } catch (...) {
    lastError() = "unknown exception";
  }
  return 1;
}

struct SolverBox {
  SolverPtr solver;
  void* stream = nullptr;     // cudaStream_t the device backend runs on (set_stream)
  std::shared_ptr<void> ext;  // library-specific per-solver state (device staging buffers of *_host entry points)
};

struct PatternBox {
  std::vector<int64_t> paramSizes;
  SparseStructure ss;
};

inline const ComputationModel* pickModel(int id) {
  switch (id) {
    case 0: return &ComputationModel::model_OpenBlas_i7_1185g7;
    case 1: return &ComputationModel::model_Cuda117_2080Ti;
    case 2: return &ComputationModel::model_B200;
    case 3: {  // tuning hook: 20 comma-separated coefficients (potrf 4, trsm 6, syge 6, asmbl 4) in BSPB200_MODEL_PARAMS,
               // re-read at every call so tools/model_fit.py can sweep presets inside one process
      static thread_local ComputationModel custom;
      custom = ComputationModel::model_B200;
      if (const char* e = getenv("BSPB200_MODEL_PARAMS")) {
        std::vector<double> v;
        std::stringstream ss(e);
        std::string tok;
        while (std::getline(ss, tok, ',')) v.push_back(atof(tok.c_str()));
        if (v.size() == 20) {
          for (int i = 0; i < 4; i++) custom.potrfParams[i] = v[i];
          for (int i = 0; i < 6; i++) custom.trsmParams[i] = v[4 + i];
          for (int i = 0; i < 6; i++) custom.sygeParams[i] = v[10 + i];
          for (int i = 0; i < 4; i++) custom.asmblParams[i] = v[16 + i];
        }
      }
      return &custom;
    }
    default: return nullptr;
  }
}

inline const std::vector<int64_t>& skelArray(const Solver& s, int which) {
  const auto& k = s.skel();
  switch (which) {
    case 0: return k.spanStart;
    case 1: return k.spanToLump;
    case 2: return k.lumpStart;
    case 3: return k.lumpToSpan;
    case 4: return k.spanOffsetInLump;
    case 5: return k.chainColPtr;
    case 6: return k.chainRowSpan;
    case 7: return k.chainData;
    case 8: return k.chainRowsTillEnd;
    case 9: return k.boardColPtr;
    case 10: return k.boardRowLump;
    case 11: return k.boardChainColOrd;
    case 12: return k.boardRowPtr;
    case 13: return k.boardColLump;
    case 14: return k.boardColOrd;
    case 15: return s.paramToSpan();
    case 16: return s.sparseEliminationRanges();
    default: throw std::runtime_error("bad array id");
  }
}

// Algorithmic work from the skeleton alone:
//   F_factor = sum_lumps s^3/3 + s^2 r + [dense: s r (r+1) | sparse-elim: s * sum_{pairs i<=j} 2 rows_i rows_j]
//   nnz(L)   = sum_lumps s(s+1)/2 + s r ;  F_solve = 4 nnz(L) per RHS
//   elim bytes = 8 * [2 * sum_{elim lumps}(s^2 + s r) + 2 * (distinct target entries touched)]
inline void workEstimate(const Solver& s, double* factorFlops, double* solveFlops, double* nnzL, double* elimBytes,
                         double* elimFlops) {
  const auto& k = s.skel();
  const auto& ranges = s.sparseEliminationRanges();
  int64_t elimEnd = ranges.empty() ? 0 : ranges.back();
  double ff = 0, nz = 0, eb = 0, ef = 0;
  std::unordered_set<int64_t> touched;  // (target chain id * 2^20 + col span offset) keys of target blocks
  for (int64_t l = 0; l < k.numLumps(); l++) {
    double w = (double)k.lumpSize(l), r = (double)(k.lumpTotalRows(l) - k.lumpSize(l));
    nz += w * (w + 1) / 2 + w * r;
    double f = w * w * w / 3 + w * w * r;
    if (l < elimEnd) {
      double pairSum = 0, below = 0;
      int64_t first = k.chainColPtr[l] + (k.lumpToSpan[l + 1] - k.lumpToSpan[l]);
      // sum over ordered pairs i<=j of 2*rows_i*rows_j = (sum rows)^2 + sum rows^2
      double sq = 0;
      for (int64_t c = first; c < k.chainColPtr[l + 1]; c++) {
        double rows = (double)(k.chainRowsTillEnd[c] - k.chainRowsTillEnd[c - 1]);
        below += rows;
        sq += rows * rows;
      }
      pairSum = below * below + sq;
      f += w * pairSum;
      ef += f;
      eb += 8.0 * 2.0 * (w * w + w * r);
      for (int64_t ci = first; ci < k.chainColPtr[l + 1]; ci++) {
        int64_t si = k.chainRowSpan[ci];
        for (int64_t cj = ci; cj < k.chainColPtr[l + 1]; cj++) {
          int64_t sj = k.chainRowSpan[cj];
          int64_t key = sj * (k.numSpans() + 1) + si;
          if (touched.insert(key).second) {
            double ri = (double)(k.spanStart[si + 1] - k.spanStart[si]);
            double rj = (double)(k.spanStart[sj + 1] - k.spanStart[sj]);
            eb += 8.0 * 2.0 * ri * rj;
          }
        }
      }
    } else {
      f += w * r * (r + 1);
    }
    ff += f;
  }
  if (factorFlops) *factorFlops = ff;
  if (solveFlops) *solveFlops = 4.0 * nz;
  if (nnzL) *nnzL = nz;
  if (elimBytes) *elimBytes = eb;
  if (elimFlops) *elimFlops = ef;
}

template <typename T>
void factorRange(const Solver& s, T* data, int64_t startSpan, int64_t endSpan) {
  int64_t nSpans = s.skel().numSpans();
  if (endSpan < 0) endSpan = nSpans;
  if (startSpan == 0) s.factorUpTo(data, endSpan);
  else if (endSpan == nSpans) s.factorFrom(data, startSpan);
  else throw std::runtime_error("factor: range must start at span 0 or end at the last span");
}

template <typename T>
void solveRange(const Solver& s, int mode, const T* data, T* vec, int64_t ld, int nRHS, int64_t startSpan, int64_t endSpan) {
  int64_t nSpans = s.skel().numSpans();
  if (endSpan < 0) endSpan = nSpans;
  bool full = (startSpan == 0 && endSpan == nSpans);
  if (mode == 0) {
    if (!full) throw std::runtime_error("solve(LLt): only the full range is supported (reference Solver.h:61)");
    s.solve(data, vec, ld, nRHS);
  } else if (mode == 1) {
    if (startSpan == 0) s.solveLUpTo(data, endSpan, vec, ld, nRHS);
    else if (endSpan == nSpans) s.solveLFrom(data, startSpan, vec, ld, nRHS);
    else throw std::runtime_error("solveL: range must start at span 0 or end at the last span");
  } else if (mode == 2) {
    if (startSpan == 0) s.solveLtUpTo(data, endSpan, vec, ld, nRHS);
    else if (endSpan == nSpans) s.solveLtFrom(data, startSpan, vec, ld, nRHS);
    else throw std::runtime_error("solveLt: range must start at span 0 or end at the last span");
  } else {
    throw std::runtime_error("bad solve mode");
  }
}

}  // namespace capi_detail

using capi_detail::guarded;
using capi_detail::PatternBox;
using capi_detail::SolverBox;

extern "C" {

const char* CAPI(last_error)(void) { return capi_detail::lastError().c_str(); }

int CAPI(create_solver)(int backend, int num_threads, int find_sparse_elim_ranges, int add_fill_policy,
                        int computation_model, int64_t n_params, const int64_t* param_sizes, const int64_t* ss_ptrs,
                        const int64_t* ss_inds, int64_t n_elim_ranges, const int64_t* elim_ranges, int64_t n_elim_last,
                        const int64_t* elim_last_ids, bspb200_solver** out) {
  return guarded([&] {
    using namespace BaSpaCho;
    Settings st;
    st.backend = (BackendType)backend;
    st.numThreads = num_threads;
    st.findSparseEliminationRanges = find_sparse_elim_ranges != 0;
    st.addFillPolicy = (AddFillPolicy)add_fill_policy;
    st.computationModel = capi_detail::pickModel(computation_model);
    std::vector<int64_t> sizes(param_sizes, param_sizes + n_params);
    SparseStructure ss(std::vector<int64_t>(ss_ptrs, ss_ptrs + n_params + 1),
                       std::vector<int64_t>(ss_inds, ss_inds + ss_ptrs[n_params]));
    std::vector<int64_t> ranges(elim_ranges, elim_ranges + n_elim_ranges);
    std::unordered_set<int64_t> last(elim_last_ids, elim_last_ids + n_elim_last);
    auto box = std::make_unique<SolverBox>();
    box->solver = createSolver(st, sizes, ss, ranges, last);
    *out = reinterpret_cast<bspb200_solver*>(box.release());
  });
}

int CAPI(create_solver_from_skel)(int backend, int num_threads, int64_t n_spans, const int64_t* span_start,
                                  int64_t n_lumps, const int64_t* lump_to_span, const int64_t* col_ptr,
                                  const int64_t* row_ind, int64_t n_elim_ranges, const int64_t* elim_ranges,
                                  const int64_t* permutation, bspb200_solver** out) {
  return guarded([&] {
    using namespace BaSpaCho;
    std::vector<int64_t> spanStart(span_start, span_start + n_spans + 1);
    std::vector<int64_t> lumpToSpan(lump_to_span, lump_to_span + n_lumps + 1);
    std::vector<int64_t> colPtr(col_ptr, col_ptr + n_lumps + 1);
    std::vector<int64_t> rowInd(row_ind, row_ind + col_ptr[n_lumps]);
    CoalescedBlockMatrixSkel skel(spanStart, lumpToSpan, colPtr, rowInd);
    std::vector<int64_t> ranges(elim_ranges, elim_ranges + n_elim_ranges);
    std::vector<int64_t> perm(n_spans);
    if (permutation) perm.assign(permutation, permutation + n_spans); else std::iota(perm.begin(), perm.end(), 0);
    Settings st;
    st.backend = (BackendType)backend;
    st.numThreads = num_threads;
    auto box = std::make_unique<SolverBox>();
    box->solver = SolverPtr(new Solver(std::move(skel), std::move(ranges), std::move(perm),
                                       getBackend(st)));
    *out = reinterpret_cast<bspb200_solver*>(box.release());
  });
}

void CAPI(destroy_solver)(bspb200_solver* s) { delete reinterpret_cast<SolverBox*>(s); }

int64_t CAPI(solver_query)(const bspb200_solver* s, int what) {
  const auto& sv = *reinterpret_cast<const SolverBox*>(s)->solver;
  switch (what) {
    case 0: return sv.order();
    case 1: return sv.dataSize();
    case 2: return sv.skel().numSpans();
    case 3: return sv.skel().numLumps();
    case 4: return sv.canFactorUpToSpan();
    case 5: return sv.elimTempSize();
    case 6: return sv.sparseEliminationRanges().empty() ? 0 : (int64_t)sv.sparseEliminationRanges().size() - 1;
    default: return -1;
  }
}

int64_t CAPI(solver_array)(const bspb200_solver* s, int which, int64_t* out, int64_t cap) {
  int64_t len = -1;
  guarded([&] {
    const auto& v = capi_detail::skelArray(*reinterpret_cast<const SolverBox*>(s)->solver, which);
    len = (int64_t)v.size();
    if (out) std::memcpy(out, v.data(), sizeof(int64_t) * std::min(len, cap));
  });
  return len;
}

int CAPI(densify)(const bspb200_solver* s, int dtype, const void* host_data, void* host_dense, int fill_upper_half,
                  int64_t start_span) {
  return guarded([&] {
    const auto& k = reinterpret_cast<const SolverBox*>(s)->solver->skel();
    if (dtype == 0) k.densify((double*)host_dense, (const double*)host_data, fill_upper_half != 0, start_span);
    else k.densify((float*)host_dense, (const float*)host_data, fill_upper_half != 0, start_span);
  });
}

int CAPI(damp)(const bspb200_solver* s, int dtype, void* host_data, double alpha, double beta) {
  return guarded([&] {
    const auto& k = reinterpret_cast<const SolverBox*>(s)->solver->skel();
    if (dtype == 0) k.damp((double*)host_data, alpha, beta);
    else k.damp((float*)host_data, (float)alpha, (float)beta);
  });
}

int CAPI(block_offset)(const bspb200_solver* s, int64_t row_block, int64_t col_block, int64_t* offset, int64_t* stride,
                       int* flipped) {
  return guarded([&] {
    auto acc = reinterpret_cast<const SolverBox*>(s)->solver->accessor();
    auto [o, st, f] = acc.blockOffset(row_block, col_block);
    *offset = o, *stride = st, *flipped = f ? 1 : 0;
  });
}

int CAPI(work_estimate)(const bspb200_solver* s, double* factor_flops, double* solve_flops_per_rhs, double* nnz_l,
                        double* elim_bytes_f64, double* elim_flops) {
  return guarded([&] {
    capi_detail::workEstimate(*reinterpret_cast<const SolverBox*>(s)->solver, factor_flops, solve_flops_per_rhs, nnz_l,
                              elim_bytes_f64, elim_flops);
  });
}

int CAPI(set_stream)(bspb200_solver* s, void* stream) {
  return guarded([&] {
    auto* box = reinterpret_cast<SolverBox*>(s);
    box->solver->setStream(stream);
    box->stream = stream;
  });
}

int CAPI(set_fused)(bspb200_solver* s, int enabled) {
  return guarded([&] { reinterpret_cast<SolverBox*>(s)->solver->setUseFusedOps(enabled != 0); });
}

int CAPI(factor)(bspb200_solver* s, int dtype, void* data, int64_t start_span, int64_t end_span) {
  return guarded([&] {
    const auto& sv = *reinterpret_cast<SolverBox*>(s)->solver;
    if (dtype == 0) capi_detail::factorRange(sv, (double*)data, start_span, end_span);
    else capi_detail::factorRange(sv, (float*)data, start_span, end_span);
  });
}

int CAPI(factor_batched)(bspb200_solver* s, int dtype, void* const* data_ptrs, int batch, int64_t start_span,
                         int64_t end_span) {
  return guarded([&] {
    const auto& sv = *reinterpret_cast<SolverBox*>(s)->solver;
    if (dtype == 0) {
      std::vector<double*> v((double* const*)data_ptrs, (double* const*)data_ptrs + batch);
      capi_detail::factorRange(sv, &v, start_span, end_span);
    } else {
      std::vector<float*> v((float* const*)data_ptrs, (float* const*)data_ptrs + batch);
      capi_detail::factorRange(sv, &v, start_span, end_span);
    }
  });
}

int CAPI(solve)(bspb200_solver* s, int dtype, int mode, const void* data, void* vec, int64_t ld, int n_rhs,
                int64_t start_span, int64_t end_span) {
  return guarded([&] {
    const auto& sv = *reinterpret_cast<SolverBox*>(s)->solver;
    if (dtype == 0) capi_detail::solveRange(sv, mode, (const double*)data, (double*)vec, ld, n_rhs, start_span, end_span);
    else capi_detail::solveRange(sv, mode, (const float*)data, (float*)vec, ld, n_rhs, start_span, end_span);
  });
}

int CAPI(solve_batched)(bspb200_solver* s, int dtype, int mode, const void* const* data_ptrs, void* const* vec_ptrs,
                        int batch, int64_t ld, int n_rhs, int64_t start_span, int64_t end_span) {
  return guarded([&] {
    const auto& sv = *reinterpret_cast<SolverBox*>(s)->solver;
    if (dtype == 0) {
      std::vector<double*> m((double* const*)data_ptrs, (double* const*)data_ptrs + batch);
      std::vector<double*> v((double* const*)vec_ptrs, (double* const*)vec_ptrs + batch);
      capi_detail::solveRange(sv, mode, (const std::vector<double*>*)&m, &v, ld, n_rhs, start_span, end_span);
    } else {
      std::vector<float*> m((float* const*)data_ptrs, (float* const*)data_ptrs + batch);
      std::vector<float*> v((float* const*)vec_ptrs, (float* const*)vec_ptrs + batch);
      capi_detail::solveRange(sv, mode, (const std::vector<float*>*)&m, &v, ld, n_rhs, start_span, end_span);
    }
  });
}

int CAPI(add_mv_from)(bspb200_solver* s, int dtype, const void* data, int64_t span_index, const void* in_vec,
                      int64_t in_stride, void* out_vec, int64_t out_stride, int n_rhs, double alpha) {
  return guarded([&] {
    const auto& sv = *reinterpret_cast<SolverBox*>(s)->solver;
    if (dtype == 0)
      sv.addMvFrom((const double*)data, span_index, (const double*)in_vec, in_stride, (double*)out_vec, out_stride, n_rhs, alpha);
    else
      sv.addMvFrom((const float*)data, span_index, (const float*)in_vec, in_stride, (float*)out_vec, out_stride, n_rhs, (float)alpha);
  });
}

int CAPI(pseudo_factor_from)(bspb200_solver* s, int dtype, void* data, int64_t span_index) {
  return guarded([&] {
    const auto& sv = *reinterpret_cast<SolverBox*>(s)->solver;
    if (dtype == 0) sv.pseudoFactorFrom((double*)data, span_index);
    else sv.pseudoFactorFrom((float*)data, span_index);
  });
}

int CAPI(do_elimination)(bspb200_solver* s, int dtype, void* data, int range_index) {
  return guarded([&] {
    auto& sv = *reinterpret_cast<SolverBox*>(s)->solver;
    const auto& ranges = sv.sparseEliminationRanges();
    if (range_index < 0 || range_index + 1 >= (int)ranges.size()) throw std::runtime_error("bad elimination range index");
    if (dtype == 0) {
      auto ctx = sv.internalSymbolicContext().createNumericCtx<double>(0, (double*)nullptr);
      ctx->doElimination(sv.internalGetElimCtx(range_index), (double*)data, ranges[range_index], ranges[range_index + 1]);
    } else {
      auto ctx = sv.internalSymbolicContext().createNumericCtx<float>(0, (float*)nullptr);
      ctx->doElimination(sv.internalGetElimCtx(range_index), (float*)data, ranges[range_index], ranges[range_index + 1]);
    }
  });
}

// ---------------------------------------------------------------------------------- synthetic problems
int CAPI(gen_pattern)(int kind, const double* p, int np, int64_t bsize_min, int64_t bsize_max, int64_t seed,
                      bspb200_pattern** out) {
  return guarded([&] {
    using namespace BaSpaCho;
    using namespace BaSpaCho::testing_utils;
    auto need = [&](int n) { if (np < n) throw std::runtime_error("gen_pattern: too few parameters"); };
    auto box = std::make_unique<PatternBox>();
    auto fromGenerator = [&](SparseMatGenerator& g) {  // Bench.cpp:279-288 matGenToSparseProblem
      box->ss = columnsToCscStruct(g.columns).transpose();
      if (bsize_min == bsize_max) box->paramSizes.assign(g.columns.size(), bsize_min);
      else box->paramSizes = randomVec(g.columns.size(), bsize_min, bsize_max, g.gen);
    };
    switch (kind) {
      case 0: { need(2); auto g = SparseMatGenerator::genFlat((int64_t)p[0], p[1], seed); fromGenerator(g); break; }
      case 1: { need(4); auto g = SparseMatGenerator::genGrid((int64_t)p[0], (int64_t)p[1], p[2], (int64_t)p[3], seed); fromGenerator(g); break; }
      case 2: { need(7); auto g = SparseMatGenerator::genMeridians((int64_t)p[0], (int64_t)p[1], p[2], (int64_t)p[3], (int64_t)p[4], (int64_t)p[5], (int64_t)p[6], seed); fromGenerator(g); break; }
      case 3: {
        need(6);
        int64_t nPts = (int64_t)p[0], nCams = (int64_t)p[1];
        box->ss = genBundleAdjustment(nPts, nCams, (int64_t)p[2], p[3], (int64_t)p[4], p[5], seed);
        box->paramSizes.assign(nPts, bsize_min);                       // points first (BaAtLargeBench.cpp:50-57)
        box->paramSizes.insert(box->paramSizes.end(), nCams, bsize_max);  // then cameras
        break;
      }
      case 4: {
        need(2);
        auto cols = randomCols((int64_t)p[0], p[1], seed);
        box->ss = columnsToCscStruct(cols).transpose();
        box->paramSizes = bsize_min == bsize_max ? std::vector<int64_t>(cols.size(), bsize_min)
                                                 : randomVec(cols.size(), bsize_min, bsize_max, seed);
        break;
      }
      case 5: { need(4); auto g = SparseMatGenerator::genFlat((int64_t)p[0], p[1], seed); g.addSchurSet((int64_t)p[2], p[3]); fromGenerator(g); break; }
      default: throw std::runtime_error("gen_pattern: unknown kind");
    }
    *out = reinterpret_cast<bspb200_pattern*>(box.release());
  });
}

int64_t CAPI(pattern_order)(const bspb200_pattern* p) { return reinterpret_cast<const PatternBox*>(p)->ss.order(); }
int64_t CAPI(pattern_nnz)(const bspb200_pattern* p) { return (int64_t)reinterpret_cast<const PatternBox*>(p)->ss.inds.size(); }
int CAPI(pattern_copy)(const bspb200_pattern* p, int64_t* param_sizes, int64_t* ss_ptrs, int64_t* ss_inds) {
  const auto& b = *reinterpret_cast<const PatternBox*>(p);
  std::memcpy(param_sizes, b.paramSizes.data(), b.paramSizes.size() * sizeof(int64_t));
  std::memcpy(ss_ptrs, b.ss.ptrs.data(), b.ss.ptrs.size() * sizeof(int64_t));
  std::memcpy(ss_inds, b.ss.inds.data(), b.ss.inds.size() * sizeof(int64_t));
  return 0;
}
void CAPI(pattern_free)(bspb200_pattern* p) { delete reinterpret_cast<PatternBox*>(p); }

int CAPI(random_data)(int dtype, int64_t size, double low, double high, int64_t seed, void* host_out) {
  return guarded([&] {
    using namespace BaSpaCho::testing_utils;
    if (dtype == 0) {
      auto v = randomData<double>(size, low, high, seed);
      std::memcpy(host_out, v.data(), size * sizeof(double));
    } else {
      auto v = randomData<float>(size, (float)low, (float)high, seed);
      std::memcpy(host_out, v.data(), size * sizeof(float));
    }
  });
}

int CAPI(fill_reducing_permutation)(int64_t n, const int64_t* ptrs, const int64_t* inds, int64_t* perm_out) {
  return guarded([&] {
    BaSpaCho::SparseStructure ss(std::vector<int64_t>(ptrs, ptrs + n + 1), std::vector<int64_t>(inds, inds + ptrs[n]));
    auto perm = ss.fillReducingPermutation();
    std::memcpy(perm_out, perm.data(), n * sizeof(int64_t));
  });
}

}  // extern "C"
