// C ABI of the product library libbaspacho_b200.so (include/baspacho_b200.h). The entry points shared with the
// CPU checker live in capi_impl.h; here: the pieces that only exist with a device (host-buffer convenience call,
// launch counter, version).
#include <cuda_runtime.h>
#include "../../include/baspacho_b200.h"
#include "b200/B200Defs.h"
#include "b200/B200Kernels.h"

#define CAPI(name) bspb200_##name
#include "capi_impl.h"

namespace BaSpaCho {
std::vector<SymElimCtxPtr> b200ElimChunkPlans(SymbolicCtx& sym, const std::vector<int64_t>& bounds);  // B200Ops.cu
}

namespace {

using BaSpaCho::b200::DevBuf;

// device-side staging of the *_host entry points, kept with the solver
struct HostStaging {
  DevBuf<unsigned char> data, vec;
  bool scanned = false;
  std::vector<std::pair<int64_t, int64_t>> wide;  // (data offset, width) of the wide diagonal blocks, ascending
  // pipeline: uploads go on `copy`, the numeric work on the solver's stream, ordered by events
  cudaStream_t copy = nullptr;
  std::vector<cudaEvent_t> events;
  // chunks of the first elimination range (built at the first call)
  bool chunksBuilt = false;
  std::vector<int64_t> chunkBounds;
  std::vector<BaSpaCho::SymElimCtxPtr> chunkPlans;
  int64_t h2dBytes = 0, d2hBytes = 0;  // bytes moved by the last call
  cudaEvent_t event(size_t i) {
    while (events.size() <= i) {
      cudaEvent_t e;
      B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      events.push_back(e);
    }
    return events[i];
  }
  ~HostStaging() {
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    if (copy) cudaStreamDestroy(copy);
  }
};

HostStaging* staging(SolverBox* box) {
  if (!box->ext) box->ext = std::make_shared<HostStaging>();
  auto* stg = static_cast<HostStaging*>(box->ext.get());
  if (!stg->copy) B200_CUDA(cudaStreamCreateWithFlags(&stg->copy, cudaStreamNonBlocking));
  return stg;
}

// uploads [from, to) of the factor data; wide diagonal blocks go up in row bands that stop at the diagonal (their upper
// triangle is a don't-care region of the format, reference CoalescedBlockMatrix.h:38-111)
struct Uploader {
  const BaSpaCho::Solver& sv;
  HostStaging* stg;
  const char* src;
  char* dst;
  size_t es;
  cudaStream_t st;
  void flat(int64_t from, int64_t to) {
    if (to <= from) return;
    B200_CUDA(cudaMemcpyAsync(dst + from * es, src + from * es, (size_t)(to - from) * es, cudaMemcpyHostToDevice, st));
    stg->h2dBytes += (to - from) * (int64_t)es;
  }
  void range(int64_t from, int64_t to) {
    constexpr int64_t kBand = 256;
    int64_t cursor = from;
    for (const auto& ow : stg->wide) {
      const int64_t off = ow.first, w = ow.second;
      if (off + w * w <= from || off >= to) continue;
      flat(cursor, off);
      for (int64_t b0 = 0; b0 < w; b0 += kBand) {
        const int64_t b1 = std::min(w, b0 + kBand);
        B200_CUDA(cudaMemcpy2DAsync(dst + (off + b0 * w) * es, (size_t)w * es, src + (off + b0 * w) * es, (size_t)w * es,
                                    (size_t)b1 * es, (size_t)(b1 - b0), cudaMemcpyHostToDevice, st));
        stg->h2dBytes += b1 * (b1 - b0) * (int64_t)es;
      }
      cursor = off + w * w;
    }
    flat(cursor, to);
  }
};

}  // namespace

template <typename T>
static void factorSolveHostT(SolverBox* box, const T* hostData, T* hostFactorOut, T* hostVec, int64_t ld, int nRhs) {
  using namespace BaSpaCho;
  Solver& sv = *box->solver;
  HostStaging* stg = staging(box);
  cudaStream_t st = (cudaStream_t)box->stream;
  const auto& sk = sv.skel();
  const size_t es = sizeof(T);
  const size_t dataBytes = (size_t)sv.dataSize() * es, vecBytes = (size_t)ld * std::max(0, nRhs) * es;
  if (stg->data.size() < dataBytes) {
    stg->data.resize(dataBytes);
    // the regions that are never uploaded (upper triangles of wide diagonal blocks) must not hold stale device memory:
    // host_factor_out receives the whole buffer
    B200_CUDA(cudaMemsetAsync(stg->data.ptr(), 0, dataBytes, st));
  }
  stg->vec.ensure(std::max<size_t>(vecBytes, 1));
  if (!stg->scanned) {
    constexpr int64_t kMinWidth = 512;
    for (int64_t l = 0; l < sk.numLumps(); l++)
      if (sk.lumpSize(l) >= kMinWidth) stg->wide.emplace_back(sk.lumpDataOffset(l), sk.lumpSize(l));
    stg->scanned = true;
  }
  // chunks of the first elimination range: the elimination of a chunk of columns overlaps the upload of the next ones
  const auto& ranges = sv.sparseEliminationRanges();
  if (!stg->chunksBuilt) {
    stg->chunksBuilt = true;
    const char* e = getenv("BSPB200_HOST_CHUNKS");
    const int want = e ? atoi(e) : 12;
    if (ranges.size() >= 2 && want > 1 && ranges[0] == 0) {
      const int64_t b = ranges[0], en = ranges[1];
      const int64_t bytes = (sk.lumpDataOffset(en) - sk.lumpDataOffset(b)) * (int64_t)es;
      if (bytes >= (int64_t)64 << 20) {
        for (int c = 0; c <= want; c++) stg->chunkBounds.push_back(b + (en - b) * c / want);
        stg->chunkPlans = b200ElimChunkPlans(sv.internalSymbolicContext(), stg->chunkBounds);
      }
    }
  }
  stg->h2dBytes = stg->d2hBytes = 0;
  T* dev = (T*)stg->data.ptr();
  Uploader up{sv, stg, (const char*)hostData, (char*)stg->data.ptr(), es, stg->copy};
  B200_CUDA(cudaEventRecord(stg->event(0), st));  // the copy stream starts behind whatever the solver's stream holds
  B200_CUDA(cudaStreamWaitEvent(stg->copy, stg->event(0), 0));
  if (stg->chunkPlans.empty()) {
    up.range(0, sv.dataSize());
    if (nRhs > 0) B200_CUDA(cudaMemcpyAsync(stg->vec.ptr(), hostVec, vecBytes, cudaMemcpyHostToDevice, stg->copy));
    B200_CUDA(cudaEventRecord(stg->event(1), stg->copy));
    B200_CUDA(cudaStreamWaitEvent(st, stg->event(1), 0));
    sv.factor(dev);
  } else {
    // 1. everything behind the chunked range (the targets of the elimination) and the right-hand sides, 2. the chunks
    const int64_t rangeEnd = sk.lumpDataOffset(stg->chunkBounds.back());
    up.range(rangeEnd, sv.dataSize());
    if (nRhs > 0) B200_CUDA(cudaMemcpyAsync(stg->vec.ptr(), hostVec, vecBytes, cudaMemcpyHostToDevice, stg->copy));
    B200_CUDA(cudaEventRecord(stg->event(1), stg->copy));
    B200_CUDA(cudaStreamWaitEvent(st, stg->event(1), 0));
    auto numCtx = sv.internalSymbolicContext().createNumericCtx<T>(0, (T*)nullptr);
    for (size_t c = 0; c + 1 < stg->chunkBounds.size(); c++) {
      const int64_t b = stg->chunkBounds[c], en = stg->chunkBounds[c + 1];
      up.range(sk.lumpDataOffset(b), sk.lumpDataOffset(en));
      B200_CUDA(cudaEventRecord(stg->event(2 + c), stg->copy));
      B200_CUDA(cudaStreamWaitEvent(st, stg->event(2 + c), 0));
      numCtx->doElimination(*stg->chunkPlans[c], dev, b, en);
    }
    // the rest of the factorization: later elimination ranges (if any) and the dense lumps
    sv.factorFrom(dev, sk.lumpToSpan[stg->chunkBounds.back()]);
  }
  if (nRhs > 0) sv.solve((const T*)dev, (T*)stg->vec.ptr(), ld, nRhs);
  if (hostFactorOut) {
    B200_CUDA(cudaMemcpyAsync(hostFactorOut, dev, dataBytes, cudaMemcpyDeviceToHost, st));
    stg->d2hBytes += (int64_t)dataBytes;
  }
  if (nRhs > 0) {
    B200_CUDA(cudaMemcpyAsync(hostVec, stg->vec.ptr(), vecBytes, cudaMemcpyDeviceToHost, st));
    stg->d2hBytes += (int64_t)vecBytes;
    stg->h2dBytes += (int64_t)vecBytes;
  }
  B200_CUDA(cudaStreamSynchronize(st));
}

template <typename T>
static void factorSolveHostBatchedT(SolverBox* box, const T* const* hostDatas, int batch, T* const* hostVecs, int64_t ld,
                                    int nRhs) {
  using namespace BaSpaCho;
  Solver& sv = *box->solver;
  HostStaging* stg = staging(box);
  cudaStream_t st = (cudaStream_t)box->stream;
  const size_t es = sizeof(T);
  const size_t dataBytes = (size_t)sv.dataSize() * es, vecBytes = (size_t)ld * std::max(0, nRhs) * es;
  stg->data.ensure(dataBytes * batch);
  stg->vec.ensure(std::max<size_t>(vecBytes * batch, 1));
  stg->h2dBytes = stg->d2hBytes = 0;
  // sub-batches: the upload of sub-batch k + 1 (copy stream) overlaps the factorization and solves of sub-batch k
  const char* e = getenv("BSPB200_HOST_SUBBATCH");
  const int sub = std::max(1, e ? atoi(e) : 8);
  B200_CUDA(cudaEventRecord(stg->event(0), st));
  B200_CUDA(cudaStreamWaitEvent(stg->copy, stg->event(0), 0));
  for (int q0 = 0, k = 0; q0 < batch; q0 += sub, k++) {
    const int q1 = std::min(batch, q0 + sub);
    std::vector<T*> datas, vecs;
    for (int q = q0; q < q1; q++) {
      T* d = (T*)(stg->data.ptr() + (size_t)q * dataBytes);
      T* v = (T*)(stg->vec.ptr() + (size_t)q * vecBytes);
      B200_CUDA(cudaMemcpyAsync(d, hostDatas[q], dataBytes, cudaMemcpyHostToDevice, stg->copy));
      if (nRhs > 0) B200_CUDA(cudaMemcpyAsync(v, hostVecs[q], vecBytes, cudaMemcpyHostToDevice, stg->copy));
      stg->h2dBytes += (int64_t)(dataBytes + (nRhs > 0 ? vecBytes : 0));
      datas.push_back(d), vecs.push_back(v);
    }
    B200_CUDA(cudaEventRecord(stg->event(1 + k), stg->copy));
    B200_CUDA(cudaStreamWaitEvent(st, stg->event(1 + k), 0));
    sv.factor(&datas);
    if (nRhs > 0) {
      sv.solve((const std::vector<T*>*)&datas, &vecs, ld, nRhs);
      for (int q = q0; q < q1; q++)
        B200_CUDA(cudaMemcpyAsync(hostVecs[q], vecs[q - q0], vecBytes, cudaMemcpyDeviceToHost, st));
      stg->d2hBytes += (int64_t)vecBytes * (q1 - q0);
    }
  }
  B200_CUDA(cudaStreamSynchronize(st));
}

extern "C" {

const char* bspb200_version(void) { return "baspacho-b200 0.1 (sm_100a)"; }

int64_t bspb200_launch_count(void) { return BaSpaCho::b200::launchCounter().load(); }

int bspb200_profile_enable(int on) {
  return guarded([&] { BaSpaCho::b200::profileEnable(on != 0); });
}

int64_t bspb200_profile_report(char* json_out, int64_t cap) {
  int64_t len = -1;
  guarded([&] {
    std::string js = BaSpaCho::b200::profileReportJson();
    len = (int64_t)js.size();
    if (json_out && cap > 0) {
      int64_t n = std::min<int64_t>(len, cap - 1);
      std::memcpy(json_out, js.data(), n);
      json_out[n] = 0;
    }
  });
  return len;
}

int64_t bspb200_debug_read(int what, void* out, int64_t bytes) {
  int64_t n = -1;
  guarded([&] { n = BaSpaCho::b200::debugRead(what, out, bytes); });
  return n;
}

int bspb200_dev_gemm_nt(int dtype, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                        const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int lower_only, void* stream) {
  return guarded([&] {
    using namespace BaSpaCho::b200;
    if (dtype == 0) {
      Operand<double> a, b, c;
      a.base = (double*)A, b.base = (double*)B, c.base = (double*)C;
      gemmNT<double>((cudaStream_t)stream, 1, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, lower_only != 0);
    } else {
      Operand<float> a, b, c;
      a.base = (float*)A, b.base = (float*)B, c.base = (float*)C;
      gemmNT<float>((cudaStream_t)stream, 1, m, n, k, (float)alpha, a, lda, b, ldb, (float)beta, c, ldc, lower_only != 0);
    }
  });
}

int64_t bspb200_lumpchol_job_list(int block_cols, int block_rows, int segment_len, int lag, int32_t* jobs_out,
                                  int64_t cap_jobs) {
  int64_t n = -1;
  guarded([&] { n = BaSpaCho::b200::lumpCholJobList(block_cols, block_rows, segment_len, lag, jobs_out, cap_jobs); });
  return n;
}

int bspb200_dev_potrf(int dtype, int64_t n, int64_t rows_below, void* A, int64_t ld, void* stream) {
  return guarded([&] {
    using namespace BaSpaCho::b200;
    if (dtype == 0) {
      Operand<double> a;
      a.base = (double*)A;
      potrfTrapezoid<double>((cudaStream_t)stream, 1, n, rows_below, a, ld);
    } else {
      Operand<float> a;
      a.base = (float*)A;
      potrfTrapezoid<float>((cudaStream_t)stream, 1, n, rows_below, a, ld);
    }
  });
}

int bspb200_device_accessor(const bspb200_solver* s, const int64_t** out_ptrs) {
  return guarded([&] {
    const auto acc = reinterpret_cast<const SolverBox*>(s)->solver->deviceAccessor();
    const auto& a = acc.plainAcc;
    const int64_t* p[8] = {a.spanStart, a.spanToLump, a.lumpStart, a.spanOffsetInLump, a.chainColPtr, a.chainRowSpan,
                           a.chainData, acc.permutation};
    for (int i = 0; i < 8; i++) out_ptrs[i] = p[i];
  });
}

int bspb200_factor_solve_host(bspb200_solver* s, int dtype, const void* host_data, void* host_factor_out,
                              void* host_vec, int64_t ld, int n_rhs) {
  return guarded([&] {
    auto* box = reinterpret_cast<SolverBox*>(s);
    if (dtype == 0) factorSolveHostT<double>(box, (const double*)host_data, (double*)host_factor_out, (double*)host_vec, ld, n_rhs);
    else factorSolveHostT<float>(box, (const float*)host_data, (float*)host_factor_out, (float*)host_vec, ld, n_rhs);
  });
}

int bspb200_host_copy_bytes(const bspb200_solver* s, int64_t* h2d_bytes, int64_t* d2h_bytes) {
  return guarded([&] {
    auto* box = reinterpret_cast<const SolverBox*>(s);
    const auto* stg = static_cast<const HostStaging*>(box->ext.get());
    *h2d_bytes = stg ? stg->h2dBytes : 0;
    *d2h_bytes = stg ? stg->d2hBytes : 0;
  });
}

int bspb200_factor_solve_host_batched(bspb200_solver* s, int dtype, const void* const* host_datas, int batch,
                                      void* const* host_vecs, int64_t ld, int n_rhs) {
  return guarded([&] {
    auto* box = reinterpret_cast<SolverBox*>(s);
    if (batch <= 0) return;
    if (dtype == 0) factorSolveHostBatchedT<double>(box, (const double* const*)host_datas, batch, (double* const*)host_vecs, ld, n_rhs);
    else factorSolveHostBatchedT<float>(box, (const float* const*)host_datas, batch, (float* const*)host_vecs, ld, n_rhs);
  });
}

}  // extern "C"
