// Backend that supports analysis only: lets createSolver() build and expose the skeleton on a machine
// without a GPU (CI, the CPU test tier). Every numeric entry point throws - there is deliberately no
// CPU numeric path in the product library.
#include "MatOps.h"

namespace BaSpaCho {
namespace {

struct NoElimCtx : SymElimCtx {};

struct SymbolicOnlyCtx : SymbolicCtx {
  SymElimCtxPtr prepareElimination(int64_t, int64_t) override { return SymElimCtxPtr(new NoElimCtx); }
  NumericCtxBase* createNumericCtxForType(std::type_index, int64_t, int) override {
    throw std::runtime_error("symbolic-only backend: no numeric factorization (use BackendCuda on a B200)");
  }
  SolveCtxBase* createSolveCtxForType(std::type_index, int, int) override {
    throw std::runtime_error("symbolic-only backend: no solve (use BackendCuda on a B200)");
  }
  PermutedCoalescedAccessor deviceAccessor() override {
    throw std::runtime_error("symbolic-only backend: no device accessor");
  }
};

struct SymbolicOnlyOpsImpl : Ops {
  SymbolicCtxPtr createSymbolicCtx(const CoalescedBlockMatrixSkel&, const std::vector<int64_t>&) override {
    return SymbolicCtxPtr(new SymbolicOnlyCtx);
  }
};

}  // namespace

OpsPtr symbolicOnlyOps() { return OpsPtr(new SymbolicOnlyOpsImpl); }

}  // namespace BaSpaCho
