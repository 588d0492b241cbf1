// C ABI of the product library libbaspacho_b200.so (include/baspacho_b200.h). The entry points shared with the
// CPU checker live in capi_impl.h; here: the pieces that only exist with a device (host-buffer convenience call,
// launch counter, version).
#include <cuda_runtime.h>
#include "../../include/baspacho_b200.h"
#include "b200/B200Defs.h"
#include "b200/B200Kernels.h"

#define CAPI(name) bspb200_##name
#include "capi_impl.h"

namespace {

struct HostStaging {
  BaSpaCho::b200::DevBuf<unsigned char> data, vec;
  bool scanned = false;
  std::vector<std::pair<int64_t, int64_t>> wide;  // (data offset, width) of the wide diagonal blocks, ascending
};

}  // namespace

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

int bspb200_factor_solve_host(bspb200_solver* s, int dtype, const void* host_data, void* host_factor_out,
                              void* host_vec, int64_t ld, int n_rhs) {
  return guarded([&] {
    using namespace BaSpaCho;
    auto* box = reinterpret_cast<SolverBox*>(s);
    const Solver& sv = *box->solver;
    if (!box->ext) box->ext = std::make_shared<HostStaging>();
    auto* stg = static_cast<HostStaging*>(box->ext.get());
    cudaStream_t st = (cudaStream_t)box->stream;
    const size_t es = dtype == 0 ? 8 : 4;
    const size_t dataBytes = (size_t)sv.dataSize() * es, vecBytes = (size_t)ld * std::max(0, n_rhs) * es;
    stg->data.ensure(dataBytes);
    stg->vec.ensure(std::max<size_t>(vecBytes, 1));
    // The upper triangle of a diagonal block is don't-care on input (reference CoalescedBlockMatrix.h: only the
    // lower triangle of a lump's diagonal block is meaningful), so wide diagonal blocks go up in row bands that stop
    // at the diagonal: the 5226-wide camera lump of the BAL-shaped problem is 218 MB as a square, 114 MB as bands.
    {
      const auto& sk = sv.skel();
      const char* src = (const char*)host_data;
      char* dst = (char*)stg->data.ptr();
      constexpr int64_t kMinWidth = 512, kBand = 256;
      int64_t cursor = 0;
      auto flat = [&](int64_t from, int64_t to) {
        if (to > from)
          B200_CUDA(cudaMemcpyAsync(dst + from * es, src + from * es, (size_t)(to - from) * es, cudaMemcpyHostToDevice, st));
      };
      if (!stg->scanned) {
        for (int64_t l = 0; l < sk.numLumps(); l++)
          if (sk.lumpSize(l) >= kMinWidth) stg->wide.emplace_back(sk.lumpDataOffset(l), sk.lumpSize(l));
        stg->scanned = true;
      }
      for (const auto& ow : stg->wide) {
        const int64_t off = ow.first, w = ow.second;
        flat(cursor, off);
        for (int64_t b0 = 0; b0 < w; b0 += kBand) {
          const int64_t b1 = std::min(w, b0 + kBand);
          B200_CUDA(cudaMemcpy2DAsync(dst + (off + b0 * w) * es, (size_t)w * es, src + (off + b0 * w) * es, (size_t)w * es,
                                      (size_t)b1 * es, (size_t)(b1 - b0), cudaMemcpyHostToDevice, st));
        }
        cursor = off + w * w;
      }
      flat(cursor, sv.dataSize());
    }
    if (n_rhs > 0) B200_CUDA(cudaMemcpyAsync(stg->vec.ptr(), host_vec, vecBytes, cudaMemcpyHostToDevice, st));
    if (dtype == 0) {
      sv.factor((double*)stg->data.ptr());
      if (n_rhs > 0) sv.solve((const double*)stg->data.ptr(), (double*)stg->vec.ptr(), ld, n_rhs);
    } else {
      sv.factor((float*)stg->data.ptr());
      if (n_rhs > 0) sv.solve((const float*)stg->data.ptr(), (float*)stg->vec.ptr(), ld, n_rhs);
    }
    if (host_factor_out) B200_CUDA(cudaMemcpyAsync(host_factor_out, stg->data.ptr(), dataBytes, cudaMemcpyDeviceToHost, st));
    if (n_rhs > 0) B200_CUDA(cudaMemcpyAsync(host_vec, stg->vec.ptr(), vecBytes, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
