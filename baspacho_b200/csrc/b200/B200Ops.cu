// The B200 backend behind the reference's operator interface (host/MatOps.h == reference MatOps.h):
// b200Ops() -> B200SymbolicCtx -> B200SymElimCtx / B200NumericCtx<T> / B200SolveCtx<T>, for
// T in {double, float, std::vector<double*>, std::vector<float*>} - what reference MatOpsCuda.cu:55-1471
// provides with cuBLAS/cuSOLVER + 11 SIMT kernels, rebuilt on hand-written sm_100a kernels:
//   * every op runs on an explicit stream (reference: stream 0 only, MatOpsCuda.cu:58-62);
//   * workspaces live in the symbolic context and are reused across factor()/solve() calls
//     (reference: cudaMalloc per call, MatOpsCuda.cu:410-414, 1016-1018);
//   * no host<->device traffic inside factor()/solve() except the batch pointer array when it changes
//     (reference: numSpans*8 bytes H2D per lump in prepareAssemble :471-481, pointer arrays before every op);
//   * errors are std::runtime_error, never abort() (reference: CudaDefs.h:27-65).
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include "../host/DebugMacros.h"
#include "../host/MatOps.h"
#include "B200Kernels.h"
#include "B200Plan.h"
#include "B200Sparse.h"
#include "B200Wave.h"

namespace BaSpaCho {
namespace {

using namespace b200;
using std::vector;

thread_local cudaStream_t tlsSyncStream = nullptr;
struct B200SyncOps {
  static void sync() { cudaStreamSynchronize(tlsSyncStream); }
};

struct B200SymElimCtx : SymElimCtx {
  ElimPlan host;  // index vectors are released after upload; scalars stay
  DevBuf<int64_t> dstOff, rowChainOff;
  DevBuf<int32_t> dstStride, dstTaskPtr, rowPtr, rowChainCol, lightList, heavyList;
  DevBuf<int16_t> dstRows, dstCols, rowChainK;
  DevBuf<uint32_t> taskA, taskB;
  DevBuf<uint16_t> taskK;
  DevElimPlan dev;
};

struct B200SymbolicCtx : SymbolicCtx {
  B200SymbolicCtx(const CoalescedBlockMatrixSkel& s, const vector<int64_t>& permutation) : skel(s) {
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess || nDev == 0)
      throw std::runtime_error(std::string("BaSpaCho-B200: no usable CUDA device (") +
                               (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                               "); this library has no CPU fallback");
    dSpanStart.upload(s.spanStart);
    dSpanToLump.upload(s.spanToLump);
    dLumpStart.upload(s.lumpStart);
    dLumpToSpan.upload(s.lumpToSpan);
    dSpanOffsetInLump.upload(s.spanOffsetInLump);
    dChainColPtr.upload(s.chainColPtr);
    dChainRowSpan.upload(s.chainRowSpan);
    dChainData.upload(s.chainData);
    dChainRowsTillEnd.upload(s.chainRowsTillEnd);
    dBoardColPtr.upload(s.boardColPtr);
    dBoardRowLump.upload(s.boardRowLump);
    dBoardChainColOrd.upload(s.boardChainColOrd);
    dPermutation.upload(permutation);
    dsk.spanStart = dSpanStart.ptr(), dsk.spanToLump = dSpanToLump.ptr(), dsk.lumpStart = dLumpStart.ptr();
    dsk.lumpToSpan = dLumpToSpan.ptr(), dsk.spanOffsetInLump = dSpanOffsetInLump.ptr();
    dsk.chainColPtr = dChainColPtr.ptr(), dsk.chainRowSpan = dChainRowSpan.ptr(), dsk.chainData = dChainData.ptr();
    dsk.chainRowsTillEnd = dChainRowsTillEnd.ptr(), dsk.boardColPtr = dBoardColPtr.ptr();
    dsk.boardRowLump = dBoardRowLump.ptr(), dsk.boardChainColOrd = dBoardChainColOrd.ptr();
    dsk.numSpans = s.numSpans(), dsk.numLumps = s.numLumps();
    spanToChainOffset.resize(std::max<int64_t>(1, s.numSpans()));
    if (const char* e = getenv("BSPB200_WAVEFRONT")) useWavefront = atoi(e) != 0;
    if (const char* e = getenv("BSPB200_INVERSE_SOLVE")) useInverseSolve = atoi(e) != 0;
    if (const char* e = getenv("BSPB200_CHAIN_SOLVE")) useChainSolve = atoi(e) != 0;
    if (const char* e = getenv("BSPB200_LANES")) numLanes = std::max(1, std::min(16, atoi(e)));
    // per-op timers insert a device sync after every op: off unless Solver::enableStats() asks for them
    potrfStat.enabled = trsmStat.enabled = sygeStat.enabled = asmblStat.enabled = false;
    solveSparseLStat.enabled = solveSparseLtStat.enabled = pseudoFactorStat.enabled = symmStat.enabled = false;
    solveLStat.enabled = solveLtStat.enabled = solveGemvStat.enabled = solveGemvTStat.enabled = false;
    solveAssVStat.enabled = solveAssVTStat.enabled = false;
  }

  void setStream(void* s) override {
    stream = (cudaStream_t)s;
    tlsSyncStream = stream;
  }

  PermutedCoalescedAccessor deviceAccessor() override {
    PermutedCoalescedAccessor a;
    a.init(dsk.spanStart, dsk.spanToLump, dsk.lumpStart, dsk.spanOffsetInLump, dsk.chainColPtr, dsk.chainRowSpan,
           dsk.chainData, dPermutation.ptr());
    return a;
  }

  SymElimCtxPtr prepareElimination(int64_t lumpsBegin, int64_t lumpsEnd) override {
    auto e = makeElimCtx(lumpsBegin, lumpsEnd);
    elimRegistry.push_back(e.get());  // owned by the Solver, which outlives every numeric/solve context
    return SymElimCtxPtr(e.release());
  }

  // plan of the sparse elimination of [lumpsBegin, lumpsEnd) WITHOUT entering it into the registry of the Solver's
  // ranges (also used for the point chunks of the host-staged pipeline, b200ElimChunkPlans below)
  std::unique_ptr<B200SymElimCtx> makeElimCtx(int64_t lumpsBegin, int64_t lumpsEnd) {
    auto e = std::make_unique<B200SymElimCtx>();
    e->elimStat.enabled = false;
    ElimPlan& p = e->host;
    p = buildElimPlan(skel, lumpsBegin, lumpsEnd);
    e->dstOff.upload(p.dstOff), e->dstStride.upload(p.dstStride), e->dstRows.upload(p.dstRows);
    e->dstCols.upload(p.dstCols), e->dstTaskPtr.upload(p.dstTaskPtr), e->taskA.upload(p.taskA);
    e->taskB.upload(p.taskB), e->taskK.upload(p.taskK), e->rowPtr.upload(p.rowPtr);
    e->rowChainOff.upload(p.rowChainOff), e->rowChainCol.upload(p.rowChainCol), e->rowChainK.upload(p.rowChainK);
    DevElimPlan& d = e->dev;
    d.lumpsBegin = lumpsBegin, d.lumpsEnd = lumpsEnd, d.spanRowBegin = p.spanRowBegin;
    d.uniformLumpSize = p.uniformLumpSize;
    d.factorEntries = p.factorEntries, d.gatherFlops = p.gatherFlops, d.gatherBytes = p.gatherEntries;
    d.numDst = p.numDst(), d.maxDstElems = p.maxDstElems;
    d.uniRows = p.uniRows, d.uniCols = p.uniCols, d.uniK = p.uniK;
    e->lightList.upload(p.lightList), e->heavyList.upload(p.heavyList);
    d.lightList = e->lightList.ptr(), d.heavyList = e->heavyList.ptr();
    d.numLight = (int64_t)p.lightList.size(), d.numHeavy = (int64_t)p.heavyList.size();
    d.lightTasks = 0;
    for (int32_t dl : p.lightList) d.lightTasks += p.dstTaskPtr[dl + 1] - p.dstTaskPtr[dl];
    d.dstOff = e->dstOff.ptr(), d.dstStride = e->dstStride.ptr(), d.dstRows = e->dstRows.ptr();
    d.dstCols = e->dstCols.ptr(), d.dstTaskPtr = e->dstTaskPtr.ptr(), d.taskA = e->taskA.ptr();
    d.taskB = e->taskB.ptr(), d.taskK = e->taskK.ptr();
    d.numRowSpans = (int64_t)p.rowPtr.size() - 1, d.maxRowSpanSize = p.maxRowSpanSize;
    d.rowPtr = e->rowPtr.ptr(), d.rowChainOff = e->rowChainOff.ptr(), d.rowChainCol = e->rowChainCol.ptr();
    d.rowChainK = e->rowChainK.ptr();
    // keep only the scalars on the host
    for (auto* v : {&p.dstOff, &p.rowChainOff}) vector<int64_t>().swap(*v);
    for (auto* v : {&p.dstStride, &p.dstTaskPtr, &p.rowPtr, &p.rowChainCol, &p.lightList, &p.heavyList})
      vector<int32_t>().swap(*v);
    for (auto* v : {&p.dstRows, &p.dstCols, &p.rowChainK}) vector<int16_t>().swap(*v);
    for (auto* v : {&p.taskA, &p.taskB}) vector<uint32_t>().swap(*v);
    vector<uint16_t>().swap(p.taskK);
    return e;
  }

  // sub-plans of a registered elimination range (BSPB200_ELIM_CHUNKS), built once
  std::map<const B200SymElimCtx*, std::vector<std::unique_ptr<B200SymElimCtx>>> elimChunkPlans;
  const std::vector<std::unique_ptr<B200SymElimCtx>>& elimChunks(const B200SymElimCtx* e, int n) {
    auto& v = elimChunkPlans[e];
    if (v.empty()) {
      const int64_t b = e->dev.lumpsBegin, len = e->dev.lumpsEnd - b;
      for (int i = 0; i < n; i++) v.push_back(makeElimCtx(b + len * i / n, b + len * (i + 1) / n));
    }
    return v;
  }

  NumericCtxBase* createNumericCtxForType(std::type_index tIdx, int64_t tempBufSize, int batchSize) override;
  SolveCtxBase* createSolveCtxForType(std::type_index tIdx, int nRHS, int batchSize) override;

  // grow-only scratch shared by the numeric / solve contexts (one host thread per Solver, as in the reference)
  void* scratch(size_t bytes) {
    if (bytes > scratchBytes.size()) {
      B200_CUDA(cudaStreamSynchronize(stream));
      scratchBytes.resize(bytes + bytes / 8);
    }
    return scratchBytes.ptr();
  }

  // wavefront plan of the dense lumps (built at the first full factorization, reused afterwards)
  struct DevWave {
    WavePlan host;  // levels + counts stay on the host; the big arrays are released after upload
    DevBuf<WaveTarget> targets;
    DevBuf<WaveSource> sources;
    DevBuf<int32_t> rowMap;
    DevBuf<WaveTile> tiles;
    DevBuf<WavePanel> panels;
    vector<double> panelFlops;  // per level
  };
  std::unique_ptr<DevWave> wave;
  const DevWave& wavePlan(int64_t firstLump) {
    if (!wave || wave->host.firstLump != firstLump) {
      auto w = std::make_unique<DevWave>();
      w->host = buildWavePlan(skel, firstLump);
      WavePlan& p = w->host;
      for (const WaveLevel& L : p.levels) {
        double f = 0;
        for (int32_t i = L.panelBegin; i < L.panelEnd; i++)
          if (p.panels[i].slab == 0) {
            double n = p.panels[i].n, r = p.panels[i].rows;
            f += n * n * n / 3 + r * n * n;
          }
        w->panelFlops.push_back(f);
      }
      w->targets.upload(p.targets), w->sources.upload(p.sources), w->rowMap.upload(p.rowMap);
      w->tiles.upload(p.tiles), w->panels.upload(p.panels);
      vector<WaveTarget>().swap(p.targets);
      vector<WaveSource>().swap(p.sources);
      vector<int32_t>().swap(p.rowMap);
      vector<WaveTile>().swap(p.tiles);
      vector<WavePanel>().swap(p.panels);
      wave = std::move(w);
    }
    return *wave;
  }

  // ---- lanes: side streams on which independent wide lumps of one tree level are factored concurrently
  struct Lane {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    DevBuf<int64_t> spanToChainOffset;
  };
  std::vector<Lane> lanes;
  cudaEvent_t evLevel = nullptr;
  DevBuf<unsigned char> laneScratch;
  int numLanes = 8;
  void ensureLanes(int batch, size_t tempBytesPerLane) {
    if (lanes.empty()) {
      lanes.resize(numLanes);
      for (Lane& ln : lanes) {
        B200_CUDA(cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking));
        B200_CUDA(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        ln.spanToChainOffset.resize(std::max<int64_t>(1, skel.numSpans()));
      }
      B200_CUDA(cudaEventCreateWithFlags(&evLevel, cudaEventDisableTiming));
    }
    (void)batch;  // the panel kernels' load counters are kept per (device, stream): every lane has its own
    if (laneScratch.size() < tempBytesPerLane * lanes.size()) {
      for (Lane& ln : lanes) B200_CUDA(cudaStreamSynchronize(ln.st));
      laneScratch.resize(tempBytesPerLane * lanes.size());
    }
  }
  // ---- eager (source-driven) schedule of the wide lumps (round 2). The level-synchronous schedule above applies ALL the
  // contributions to a lump when its level comes up: for the lumps at the top of the tree (one per level) that is a long
  // serial run of GEMM + scatter pairs from sources that were finished many levels earlier (GRID 120x120: 15 of 25 ms).
  // Here a contribution is issued as soon as its source level is done: into the targets of the NEXT level on the target's
  // own lane (followed there by the target's factorization), into later targets on a low-priority background lane, so
  // that only the contributions of a lump's last-finished children stay on its critical path. Dependencies are events
  // per wide lump, there is no level barrier. Per target the order of the contributions is fixed (source level, then
  // board order): deterministic.
  struct EagerUpdate {
    int64_t target, boardRow;
  };
  struct EagerPlan {
    int64_t firstLump = -1;
    std::vector<int32_t> level;                          // per lump (dense part)
    std::vector<std::vector<EagerUpdate>> fromLevel;     // updates into WIDE targets by the level of their source
    std::vector<int32_t> laneOf;                         // per lump: lane of a wide lump (position in its level)
    std::vector<int32_t> evIndex;                        // per lump: index of its events, -1 for the small ones
    std::vector<char> hasBg;                             // per lump: receives background contributions
    int64_t numWide = 0;
    int64_t maxTempElems = 1;                            // largest GEMM temp of a contribution into a wide lump
  };
  std::unique_ptr<EagerPlan> eager;
  std::vector<cudaEvent_t> evDone, evBg, evMain;
  std::vector<Lane> bgLanes;
  int numBgLanes = 16;
  const EagerPlan& eagerPlan(int64_t firstLump, const WavePlan& wp) {
    if (!eager || eager->firstLump != firstLump) {
      auto e = std::make_unique<EagerPlan>();
      const int64_t nLumps = skel.numLumps();
      e->firstLump = firstLump;
      e->level.assign(nLumps, 0), e->laneOf.assign(nLumps, 0), e->evIndex.assign(nLumps, -1), e->hasBg.assign(nLumps, 0);
      for (int64_t l = firstLump; l < nLumps; l++) {
        int32_t lv = 0;
        for (int64_t r = skel.boardRowPtr[l], rEnd = skel.boardRowPtr[l + 1] - 1; r < rEnd; r++) {
          const int64_t src = skel.boardColLump[r];
          if (src >= firstLump) lv = std::max(lv, e->level[src] + 1);
        }
        e->level[l] = lv;
      }
      e->fromLevel.resize(wp.levels.size());
      for (size_t lv = 0; lv < wp.levels.size(); lv++)
        for (size_t i = 0; i < wp.levels[lv].bigLumps.size(); i++) {
          const int64_t l = wp.levels[lv].bigLumps[i];
          e->laneOf[l] = (int32_t)(i % std::max(1, numLanes));
          e->evIndex[l] = (int32_t)e->numWide++;
        }
      // targets in level order (then index), per target its sources in board order: grouped by source level below
      for (size_t lv = 0; lv < wp.levels.size(); lv++)
        for (int64_t t : wp.levels[lv].bigLumps)
          for (int64_t r = skel.boardRowPtr[t], rEnd = skel.boardRowPtr[t + 1] - 1; r < rEnd; r++) {
            const int64_t src = skel.boardColLump[r];
            if (src < firstLump) continue;
            e->fromLevel[e->level[src]].push_back(EagerUpdate{t, r});
            if (e->level[src] + 1 < (int32_t)lv) e->hasBg[t] = 1;
            {
              const int64_t ord = skel.boardColOrd[r], cb = skel.chainColPtr[src], bb = skel.boardColPtr[src];
              const int64_t ch0 = skel.boardChainColOrd[bb + ord], ch1 = skel.boardChainColOrd[bb + ord + 1];
              const int64_t chEnd = skel.boardChainColOrd[skel.boardColPtr[src + 1] - 1];
              const int64_t rowBegin = skel.chainRowsTillEnd[cb + ch0 - 1];
              e->maxTempElems = std::max(e->maxTempElems, (skel.chainRowsTillEnd[cb + ch1 - 1] - rowBegin) *
                                                              (skel.chainRowsTillEnd[cb + chEnd - 1] - rowBegin));
            }
          }
      eager = std::move(e);
    }
    return *eager;
  }
  void ensureEager(const EagerPlan& e, size_t numLevels, size_t tempBytesPerLane) {
    auto grow = [](std::vector<cudaEvent_t>& v, size_t n) {
      while (v.size() < n) {
        cudaEvent_t ev;
        B200_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        v.push_back(ev);
      }
    };
    grow(evDone, (size_t)e.numWide), grow(evBg, (size_t)e.numWide), grow(evMain, numLevels);
    if (bgLanes.empty()) {
      if (const char* env = getenv("BSPB200_BG_LANES")) numBgLanes = std::max(1, std::min(16, atoi(env)));
      int lo = 0, hi = 0;
      B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = least urgent
      bgLanes.resize(numBgLanes);
      for (Lane& ln : bgLanes) {
        B200_CUDA(cudaStreamCreateWithPriority(&ln.st, cudaStreamNonBlocking, lo));
        B200_CUDA(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        ln.spanToChainOffset.resize(std::max<int64_t>(1, skel.numSpans()));
      }
    }
    if (bgScratch.size() < tempBytesPerLane * bgLanes.size()) {
      for (Lane& ln : bgLanes) B200_CUDA(cudaStreamSynchronize(ln.st));
      bgScratch.resize(tempBytesPerLane * bgLanes.size());
    }
  }
  // lanes + background lanes each hold one GEMM temp: worth at most a quarter of the free memory
  bool eagerFits(size_t tempBytesPerLane) {
    if (!bgLanes.empty() && bgScratch.size() >= tempBytesPerLane * bgLanes.size()) return true;
    size_t freeB = 0, totalB = 0;
    if (cudaMemGetInfo(&freeB, &totalB) != cudaSuccess) return false;
    const int nb = getenv("BSPB200_BG_LANES") ? std::max(1, std::min(16, atoi(getenv("BSPB200_BG_LANES")))) : numBgLanes;
    return tempBytesPerLane * (size_t)(numLanes + nb) < freeB / 4;
  }
  DevBuf<unsigned char> bgScratch;

  ~B200SymbolicCtx() override {
    for (Lane& ln : lanes) {
      if (ln.st) cudaStreamDestroy(ln.st);
      if (ln.done) cudaEventDestroy(ln.done);
    }
    for (Lane& ln : bgLanes) {
      if (ln.st) cudaStreamDestroy(ln.st);
      if (ln.done) cudaEventDestroy(ln.done);
    }
    for (auto* v : {&evDone, &evBg, &evMain, &evSolve})
      for (cudaEvent_t ev : *v) cudaEventDestroy(ev);
    if (evLevel) cudaEventDestroy(evLevel);
  }

  // ---- dense solves: slots of the 96 x 96 block inverses of every lump wider than one block (built once), the work
  // list of the one-launch inversion, and the flag/ticket state of the flag-chained triangular solve
  static constexpr int64_t kInvBlock = 96;
  struct InvPlan {
    int64_t total = 0;                                       // scratch elements per batch item
    std::map<int64_t, std::pair<int64_t, int64_t>> slot;     // diagonal block offset -> (scratch offset, width)
    DevBuf<InvBlockDesc> list;
    int64_t count = 0, maxBlocks = 0;
  };
  std::unique_ptr<InvPlan> invPlanPtr;
  const InvPlan& invPlan() {
    if (!invPlanPtr) {
      auto p = std::make_unique<InvPlan>();
      const int64_t minBlocks = useChainSolve ? 2 : 6;  // per-step launches: narrow lumps keep the substitution steps
      vector<InvBlockDesc> descs;
      for (int64_t l = 0; l < skel.numLumps(); l++) {
        const int64_t w = skel.lumpSize(l), nblk = (w + kInvBlock - 1) / kInvBlock;
        if (nblk < minBlocks) continue;
        const int64_t off = skel.lumpDataOffset(l);
        p->slot[off] = {p->total, w};
        for (int64_t b = 0; b < nblk; b++)
          descs.push_back(InvBlockDesc{off + b * kInvBlock * w + b * kInvBlock, p->total + b * 2 * kInvBlock * kInvBlock,
                                       (int32_t)w, (int32_t)std::min(kInvBlock, w - b * kInvBlock)});
        p->total += 2 * kInvBlock * kInvBlock * nblk;
        p->maxBlocks = std::max(p->maxBlocks, nblk);
      }
      p->count = (int64_t)descs.size();
      p->list.upload(descs);
      invPlanPtr = std::move(p);
    }
    return *invPlanPtr;
  }
  // vector row of every below-diagonal row of every dense lump (what assembleVec / assembleVecT look up chain by
  // chain): lets the fused solves scatter / gather inside the gemv kernels
  struct BelowRows {
    vector<int64_t> ptr;  // per lump, offset into target (lumps of the elimination ranges: empty)
    DevBuf<int64_t> target;
  };
  std::unique_ptr<BelowRows> belowRowsPtr;
  const BelowRows& belowRows() {
    if (!belowRowsPtr) {
      auto p = std::make_unique<BelowRows>();
      int64_t denseFrom = 0;
      for (const B200SymElimCtx* e : elimRegistry) denseFrom = std::max(denseFrom, e->dev.lumpsEnd);
      vector<int64_t> tgt;
      p->ptr.assign(skel.numLumps() + 1, 0);
      for (int64_t l = 0; l < skel.numLumps(); l++) {
        p->ptr[l] = (int64_t)tgt.size();
        if (l < denseFrom) continue;
        const int64_t cb = skel.chainColPtr[l], ord = skel.boardChainColOrd[skel.boardColPtr[l] + 1];
        const int64_t numChains = skel.boardChainColOrd[skel.boardColPtr[l + 1] - 1];
        for (int64_t ch = cb + ord; ch < cb + numChains; ch++) {
          const int64_t span = skel.chainRowSpan[ch];
          for (int64_t r = skel.spanStart[span]; r < skel.spanStart[span + 1]; r++) tgt.push_back(r);
        }
      }
      p->ptr[skel.numLumps()] = (int64_t)tgt.size();
      if (tgt.empty()) tgt.push_back(0);
      p->target.upload(tgt);
      belowRowsPtr = std::move(p);
    }
    return *belowRowsPtr;
  }
  // ---- fragmented skeletons (every lump a single span of at most 32 scalars): row lists for the gathers and the levels
  // of the block dependency graph for the two substitution directions (built once)
  struct FragPlan {
    bool eligible = false;
    DevBuf<int32_t> rowPtr, rowCol, lvlL, lvlT, tailList;
    DevBuf<int64_t> rowOff;
    vector<int32_t> lvlPtrL, lvlPtrT;    // per level (+1) into lvlL / lvlT
    vector<int32_t> hostLvlL, hostLvlT;  // span lists, ascending inside a level
    FragDev dev;
    double nnz = 0;
  };
  std::unique_ptr<FragPlan> fragPlanPtr;
  const FragPlan& fragPlan() {
    if (!fragPlanPtr) {
      auto p = std::make_unique<FragPlan>();
      const int64_t ns = skel.numSpans();
      bool ok = ns == skel.numLumps() && ns > 0 && ns < (int64_t(1) << 31) && skel.dataSize() < (int64_t(1) << 62);
      for (int64_t s = 0; ok && s < ns; s++) ok = skel.spanStart[s + 1] - skel.spanStart[s] <= 32;
      if (ok) {
        vector<int32_t> cnt(ns + 1, 0), lvF(ns, 0), lvB(ns, 0);
        for (int64_t c = 0; c < ns; c++)
          for (int64_t q = skel.chainColPtr[c] + 1; q < skel.chainColPtr[c + 1]; q++) cnt[skel.chainRowSpan[q] + 1]++;
        for (int64_t s = 0; s < ns; s++) cnt[s + 1] += cnt[s];
        vector<int32_t> rowCol(cnt[ns]), cur(cnt.begin(), cnt.end() - 1);
        vector<int64_t> rowOff(cnt[ns]);
        for (int64_t c = 0; c < ns; c++) {
          const double w = (double)(skel.spanStart[c + 1] - skel.spanStart[c]);
          p->nnz += w * (w + 1) / 2;
          for (int64_t q = skel.chainColPtr[c] + 1; q < skel.chainColPtr[c + 1]; q++) {
            const int64_t r = skel.chainRowSpan[q];
            rowCol[cur[r]] = (int32_t)c, rowOff[cur[r]++] = skel.chainData[q];
            lvF[r] = std::max(lvF[r], lvF[c] + 1);  // columns ascending: lvF[c] is final
            p->nnz += w * (double)(skel.spanStart[r + 1] - skel.spanStart[r]);
          }
        }
        for (int64_t c = ns - 1; c >= 0; c--)
          for (int64_t q = skel.chainColPtr[c] + 1; q < skel.chainColPtr[c + 1]; q++)
            lvB[c] = std::max(lvB[c], lvB[skel.chainRowSpan[q]] + 1);
        auto group = [&](const vector<int32_t>& lv, vector<int32_t>& ptr, vector<int32_t>& list) {
          const int32_t nl = ns ? *std::max_element(lv.begin(), lv.end()) + 1 : 0;
          ptr.assign(nl + 1, 0);
          for (int64_t s = 0; s < ns; s++) ptr[lv[s] + 1]++;
          for (int32_t l = 0; l < nl; l++) ptr[l + 1] += ptr[l];
          list.resize(ns);
          vector<int32_t> at(ptr.begin(), ptr.end() - 1);
          for (int64_t s = 0; s < ns; s++) list[at[lv[s]]++] = (int32_t)s;
        };
        group(lvF, p->lvlPtrL, p->hostLvlL);
        group(lvB, p->lvlPtrT, p->hostLvlT);
        vector<int32_t> all(ns);
        for (int64_t s = 0; s < ns; s++) all[s] = (int32_t)s;
        p->rowPtr.upload(cnt), p->rowCol.upload(rowCol.empty() ? vector<int32_t>(1, 0) : rowCol);
        p->rowOff.upload(rowOff.empty() ? vector<int64_t>(1, 0) : rowOff);
        p->lvlL.upload(p->hostLvlL), p->lvlT.upload(p->hostLvlT), p->tailList.upload(all);
        p->dev.rowPtr = p->rowPtr.ptr(), p->dev.rowCol = p->rowCol.ptr(), p->dev.rowOff = p->rowOff.ptr();
        p->eligible = true;
      }
      fragPlanPtr = std::move(p);
    }
    return *fragPlanPtr;
  }

  // ---- lanes for the dense part of the solves (round 2): the dense lumps are solved level by level of the supernodal
  // tree, the lumps of a level on different lanes (streams) with an event per lump instead of one launch after the
  // other on the solver's stream (GRID 120x120: 99 lumps x 2 directions, ~3.3 us of flag hop per 96 columns each).
  // Backward, a lump only reads the entries of its ancestors and writes its own: no conflict. Forward, the lumps of a
  // level scatter into the rows of common ancestors: every lane scatters into its OWN zero-initialised delta vector,
  // and a lump adds the lanes' deltas of its rows (in lane order: deterministic) before it is solved.
  struct SolvePlan {
    int64_t denseFrom = -1, numDense = 0;
    bool eligible = false;
    vector<int32_t> level, lane;              // per lump
    vector<vector<int64_t>> byLevel;          // lumps of a level, ascending
    vector<vector<int32_t>> srcs, tgts;       // per lump: dense lumps that update it / that own its rows below
  };
  std::unique_ptr<SolvePlan> solvePlanPtr;
  const SolvePlan& solvePlan(int64_t denseFrom) {
    if (!solvePlanPtr || solvePlanPtr->denseFrom != denseFrom) {
      auto p = std::make_unique<SolvePlan>();
      const int64_t nL = skel.numLumps();
      p->denseFrom = denseFrom, p->numDense = nL - denseFrom;
      p->level.assign(nL, 0), p->lane.assign(nL, 0), p->srcs.resize(nL), p->tgts.resize(nL);
      int64_t wide = 0;
      for (int64_t l = denseFrom; l < nL; l++) {
        wide += skel.lumpSize(l) >= 2 * kInvBlock;
        for (int64_t ch = skel.chainColPtr[l]; ch < skel.chainColPtr[l + 1]; ch++) {
          const int64_t t = skel.spanToLump[skel.chainRowSpan[ch]];
          if (t != l && (p->tgts[l].empty() || p->tgts[l].back() != t)) p->tgts[l].push_back((int32_t)t);
        }
        std::sort(p->tgts[l].begin(), p->tgts[l].end());
        p->tgts[l].erase(std::unique(p->tgts[l].begin(), p->tgts[l].end()), p->tgts[l].end());
        for (int32_t t : p->tgts[l]) p->srcs[t].push_back((int32_t)l);
      }
      int32_t maxLevel = 0;
      for (int64_t l = denseFrom; l < nL; l++) {
        int32_t lv = 0;
        for (int32_t sLump : p->srcs[l]) lv = std::max(lv, p->level[sLump] + 1);
        p->level[l] = lv, maxLevel = std::max(maxLevel, lv);
      }
      p->byLevel.resize(maxLevel + 1);
      for (int64_t l = denseFrom; l < nL; l++) {
        p->lane[l] = (int32_t)(p->byLevel[p->level[l]].size() % std::max(1, numLanes));
        p->byLevel[p->level[l]].push_back(l);
      }
      // worth it when launches of chained solves dominate: several wide lumps, not thousands of tiny ones
      static const bool on = !getenv("BSPB200_SOLVE_LANES") || atoi(getenv("BSPB200_SOLVE_LANES")) != 0;
      p->eligible = on && numLanes > 1 && wide >= 4 && p->numDense <= 4096 && (int64_t)p->byLevel.size() < p->numDense;
      solvePlanPtr = std::move(p);
    }
    return *solvePlanPtr;
  }
  std::vector<ChainSync> laneChain;
  std::vector<DevBuf<unsigned>> laneChainBuf;
  int laneChainBatch = 0;
  ChainSync* chainSyncLane(int k, int batch) {
    if (!useChainSolve) return nullptr;
    const InvPlan& ip = invPlan();
    if (ip.maxBlocks == 0) return nullptr;
    if (batch > laneChainBatch || (int)laneChain.size() < numLanes) {
      B200_CUDA(cudaDeviceSynchronize());
      laneChain.assign(numLanes, ChainSync());
      laneChainBuf.clear();
      laneChainBuf.resize(numLanes);
      for (int q = 0; q < numLanes; q++) {
        laneChainBuf[q].resize((size_t)ip.maxBlocks * batch + 1);
        B200_CUDA(cudaMemset(laneChainBuf[q].ptr(), 0, laneChainBuf[q].size() * sizeof(unsigned)));
        laneChain[q].flags = laneChainBuf[q].ptr() + 1, laneChain[q].ticket = laneChainBuf[q].ptr();
        laneChain[q].flagsPerItem = (int)ip.maxBlocks, laneChain[q].ticketBase = 0;
      }
      laneChainBatch = batch;
    }
    return &laneChain[k];
  }
  std::vector<cudaEvent_t> evSolve;
  DevBuf<unsigned char> solveLaneBuf;  // per lane: the delta vector and the staging of the per-step solves
  void* solveLaneScratch(size_t bytes) {
    if (solveLaneBuf.size() < bytes) {
      B200_CUDA(cudaDeviceSynchronize());
      solveLaneBuf.resize(bytes);
      B200_CUDA(cudaMemset(solveLaneBuf.ptr(), 0, bytes));  // the deltas are zero between solves
    }
    return solveLaneBuf.ptr();
  }
  void ensureSolveLanes(int64_t numDense) {
    ensureLanes(1, 0);
    while ((int64_t)evSolve.size() < numDense) {
      cudaEvent_t ev;
      B200_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      evSolve.push_back(ev);
    }
  }

  ChainSync chain;
  DevBuf<unsigned> chainBuf;
  int chainBatch = 0;
  ChainSync* chainSync(int batch) {
    if (!useChainSolve) return nullptr;
    const InvPlan& ip = invPlan();
    if (ip.maxBlocks == 0) return nullptr;
    if (batch > chainBatch) {
      B200_CUDA(cudaStreamSynchronize(stream));
      chainBuf.resize((size_t)ip.maxBlocks * batch + 1);
      B200_CUDA(cudaMemset(chainBuf.ptr(), 0, chainBuf.size() * sizeof(unsigned)));
      chain.flags = chainBuf.ptr() + 1, chain.ticket = chainBuf.ptr();
      chain.flagsPerItem = (int)ip.maxBlocks, chain.ticketBase = 0;  // epoch keeps growing: zeroed flags never match it
      chainBatch = batch;
    }
    return &chain;
  }
  bool useChainSolve = true;

  const CoalescedBlockMatrixSkel& skel;
  cudaStream_t stream = nullptr;
  bool useWavefront = true;
  bool useInverseSolve = true;
  vector<const B200SymElimCtx*> elimRegistry;  // elimination ranges in the order the Solver prepared them
  DevSkel dsk;
  DevBuf<int64_t> dSpanStart, dSpanToLump, dLumpStart, dLumpToSpan, dSpanOffsetInLump, dChainColPtr, dChainRowSpan,
      dChainData, dChainRowsTillEnd, dBoardColPtr, dBoardRowLump, dBoardChainColOrd, dPermutation;
  DevBuf<int64_t> spanToChainOffset;
  DevBuf<unsigned char> scratchBytes;
};

// device array of batch pointers, re-uploaded only when the host vector changes
template <typename T>
struct PtrBatch {
  DevBuf<T*> dev;
  vector<T*> last;
  Mats<T> get(const vector<T*>* v, cudaStream_t st) {
    if (*v != last) {
      dev.ensure(v->size());
      B200_CUDA(cudaMemcpyAsync(dev.ptr(), v->data(), v->size() * sizeof(T*), cudaMemcpyHostToDevice, st));
      last = *v;
    }
    Mats<T> m;
    m.many = dev.ptr(), m.batch = (int)v->size();
    return m;
  }
};

template <typename TT>
struct MatsOf {
  using T = TT;
  Mats<T> get(const T* data, cudaStream_t) {
    Mats<T> m;
    m.one = const_cast<T*>(data);
    return m;
  }
};
template <typename T_>
struct MatsOf<vector<T_*>> {
  using T = T_;
  PtrBatch<T> ptrs;
  Mats<T> get(const vector<T*>* data, cudaStream_t st) { return ptrs.get(data, st); }
};

template <typename TT>
struct B200NumericCtx : NumericCtx<TT> {
  using T = BaseType<TT>;

  B200NumericCtx(B200SymbolicCtx& s, int64_t tempBufSize, int batchSize)
      : sym(s), skel(s.skel), tempSize(tempBufSize), batch(batchSize) {
    tlsSyncStream = sym.stream;
  }

  Work<T> temp() {
    Work<T> w;
    w.base = (T*)sym.scratch((size_t)std::max<int64_t>(1, tempSize) * batch * sizeof(T));
    w.stride = tempSize;
    return w;
  }

  void pseudoFactorSpans(TT* data, int64_t spanBegin, int64_t spanEnd) override {
    auto timer = sym.pseudoFactorStat.template instance<B200SyncOps>();
    Mats<T> m = mats.get(data, sym.stream);
    b200::pseudoFactorSpans<T>(sym.stream, m.batch, sym.dsk, m, spanBegin, spanEnd);
  }

  void doElimination(const SymElimCtx& elimData, TT* data, int64_t lumpsBegin, int64_t lumpsEnd) override {
    const auto* elim = dynamic_cast<const B200SymElimCtx*>(&elimData);
    BASPACHO_CHECK_NOTNULL(elim);
    BASPACHO_CHECK_EQ(elim->dev.lumpsBegin, lumpsBegin);
    BASPACHO_CHECK_EQ(elim->dev.lumpsEnd, lumpsEnd);
    auto timer = elim->elimStat.template instance<B200SyncOps>();
    Mats<T> m = mats.get(data, sym.stream);
    elimFactorLumps<T>(sym.stream, m.batch, sym.dsk, m, lumpsBegin, lumpsEnd, elim->dev.uniformLumpSize,
                       2.0 * elim->dev.factorEntries * sizeof(T));
    elimGather<T>(sym.stream, m.batch, elim->dev, m);
  }

  void potrf(int64_t n, TT* data, int64_t offA) override {
    auto timer = sym.potrfStat.template instance<B200SyncOps>(sizeof(T) + (batch > 1 ? batch * 100 : 0), n);
    sym.potrfBiggestN = std::max(sym.potrfBiggestN, n);
    Mats<T> m = mats.get(data, sym.stream);
    potrfTrapezoid<T>(sym.stream, m.batch, n, 0, opnd(m, offA), n);
  }

  void trsm(int64_t n, int64_t k, TT* data, int64_t offA, int64_t offB) override {
    auto timer = sym.trsmStat.template instance<B200SyncOps>(sizeof(T) + (batch > 1 ? batch * 100 : 0), n, k);
    Mats<T> m = mats.get(data, sym.stream);
    trsmAny<T>(sym.stream, m.batch, n, k, opnd(m, offA), n, opnd(m, offB), n);
  }

  void saveSyrkGemm(int64_t m_, int64_t n, int64_t k, const TT* data, int64_t offset) override {
    auto timer = sym.sygeStat.template instance<B200SyncOps>(sizeof(T) + (batch > 1 ? batch * 100 : 0), m_, n, k);
    BASPACHO_CHECK_LE(m_ * n, tempSize);
    Mats<T> m = mats.get(data, sym.stream);
    // temp(n x m) = B(n x k) * A(m x k)^T, A = first m rows of B
    Operand<T> B = opnd(m, offset);
    gemmNT<T>(sym.stream, m.batch, n, m_, k, T(1), B, k, B, k, T(0), opnd(temp(), 0), m_, false);
    sym.gemmCalls++;
  }

  void prepareAssemble(int64_t targetLump) override {
    int64_t begin = skel.chainColPtr[targetLump];
    b200::prepareAssemble(sym.stream, sym.dsk, sym.spanToChainOffset.ptr(), begin, skel.chainColPtr[targetLump + 1] - begin);
  }

  void assemble(TT* data, int64_t rectRowBegin, int64_t dstStride, int64_t srcColDataOffset, int64_t srcRectWidth,
                int64_t numBlockRows, int64_t numBlockCols) override {
    auto timer = sym.asmblStat.template instance<B200SyncOps>(sizeof(T) + (batch > 1 ? batch * 100 : 0), numBlockRows,
                                                              numBlockCols);
    Mats<T> m = mats.get(data, sym.stream);
    int64_t numRows = skel.chainRowsTillEnd[srcColDataOffset + numBlockRows - 1] - rectRowBegin;
    b200::assemble<T>(sym.stream, m.batch, sym.dsk, sym.spanToChainOffset.ptr(), m, temp(), rectRowBegin, dstStride,
                      srcColDataOffset, srcRectWidth, numBlockRows, numBlockCols, numRows);
  }

  // ---- whole-range factorization in one backend call (MatOps.h addition; same work as Solver::internalFactorRange
  // sequences through the fine-grained ops, reference Solver.cpp:164-219): sparse elimination of the ranges, then per
  // dense lump the contributions of the already factored sources and ONE blocked factorization of the whole lump
  // column (diagonal block + rows below) instead of separate potrf / trsm calls.
  bool hasFusedFactor() override { return true; }

  void fusedFactorRange(TT* data, int64_t startLump, int64_t upToLump) override {
    Mats<T> m = mats.get(data, sym.stream);
    int64_t denseFrom = 0;
    for (const B200SymElimCtx* e : sym.elimRegistry) {
      denseFrom = std::max(denseFrom, e->dev.lumpsEnd);
      if (e->dev.lumpsEnd > upToLump) return;
      if (startLump > e->dev.lumpsBegin) continue;
      // BSPB200_ELIM_CHUNKS=N > 1: the range in N pieces, each factored and gathered before the next, so that a piece's
      // columns are still in L2 when its pair products read them back (sub-plans as in the host-staged pipeline)
      // Pays when the destinations have LONG task lists (stress workload, 1 M points on 200 cameras: gather 1.21 -> 0.85
      // ms with 8 pieces) and costs when they are short, because every piece walks its own list of destinations (BAL,
      // 26 tasks per destination: 0.71 -> 1.40 / 1.75 / 2.56 ms with 4 / 8 / 16 pieces). Default: pieces of ~64 MB of
      // eliminated columns when the average list holds 256 tasks or more.
      static const int envChunks = getenv("BSPB200_ELIM_CHUNKS") ? std::max(1, atoi(getenv("BSPB200_ELIM_CHUNKS"))) : 0;
      int nChunks = envChunks;
      if (nChunks == 0) {
        const double tasks = e->dev.gatherFlops / std::max(1.0, 2.0 * e->dev.uniRows * e->dev.uniCols * e->dev.uniK);
        const bool longLists = e->dev.uniK > 0 && e->dev.numDst > 0 && tasks >= 256.0 * e->dev.numDst;
        nChunks = longLists ? (int)std::min(16.0, e->dev.factorEntries * sizeof(T) / (64.0 * 1024 * 1024)) : 1;
      }
      if (nChunks > 1 && e->dev.lumpsEnd - e->dev.lumpsBegin >= 64 * nChunks) {
        for (const auto& ch : sym.elimChunks(e, nChunks)) {
          elimFactorLumps<T>(sym.stream, m.batch, sym.dsk, m, ch->dev.lumpsBegin, ch->dev.lumpsEnd, ch->dev.uniformLumpSize,
                             2.0 * ch->dev.factorEntries * sizeof(T));
          elimGather<T>(sym.stream, m.batch, ch->dev, m);
        }
        continue;
      }
      elimFactorLumps<T>(sym.stream, m.batch, sym.dsk, m, e->dev.lumpsBegin, e->dev.lumpsEnd, e->dev.uniformLumpSize,
                         2.0 * e->dev.factorEntries * sizeof(T));
      elimGather<T>(sym.stream, m.batch, e->dev, m);
    }
    const int64_t firstSrc = std::max(startLump, denseFrom);
    if (sym.useWavefront && firstSrc == denseFrom && upToLump == skel.numLumps() && firstSrc < skel.numLumps()) {
      // level-of-tree wavefront: per level one gather-GEMM launch + one batched panel launch for the small lumps,
      // the per-lump path for the wide ones
      const auto& wv = sym.wavePlan(firstSrc);
      Operand<T> all = opnd(m, 0);
      {
        static const bool timeline = getenv("BSPB200_PROFILE_TIMELINE") && atoi(getenv("BSPB200_PROFILE_TIMELINE")) != 0;
        static const bool eagerOn = !getenv("BSPB200_EAGER") || atoi(getenv("BSPB200_EAGER")) != 0;
        if (eagerOn && sym.numLanes > 1 && wv.host.numBig >= 2 && (!profileEnabled() || timeline) &&
            sym.eagerFits(eagerTempBytes(sym.eagerPlan(firstSrc, wv.host)))) {
          fusedFactorEager(m, all, wv, firstSrc);
          return;
        }
      }
      for (size_t lv = 0; lv < wv.host.levels.size(); lv++) {
        const WaveLevel& L = wv.host.levels[lv];
        waveUpdate<T>(sym.stream, m.batch, all, wv.tiles.ptr() + L.tileBegin, L.tileEnd - L.tileBegin,
                      wv.targets.ptr(), wv.sources.ptr(), wv.rowMap.ptr());
        static const int order = getenv("BSPB200_WAVE_ORDER") ? atoi(getenv("BSPB200_WAVE_ORDER")) : 0;
        // the wide lumps of a level are independent of each other (and of the level's small lumps): with two or more
        // of them, each runs its whole chain (updates from its sources, blocked factorization of its column) on its
        // own lane - stream, GEMM temp, span-to-chain table, panel counters - between two joins with the main stream
        // (the per-class profiler serialises the lanes to time every launch alone, unless it is asked for the timeline)
        static const bool timeline = getenv("BSPB200_PROFILE_TIMELINE") && atoi(getenv("BSPB200_PROFILE_TIMELINE")) != 0;
        const bool concurrent = sym.numLanes > 1 && L.bigLumps.size() >= 2 && (!profileEnabled() || timeline);
        if (concurrent) {
          sym.ensureLanes(m.batch, laneTempBytes());
          const int used = (int)std::min<size_t>(sym.lanes.size(), L.bigLumps.size());
          B200_CUDA(cudaEventRecord(sym.evLevel, sym.stream));
          for (int k = 0; k < used; k++) B200_CUDA(cudaStreamWaitEvent(sym.lanes[k].st, sym.evLevel, 0));
          for (size_t i = 0; i < L.bigLumps.size(); i++) {
            const int k = (int)(i % used);
            LaneCtx lc = laneCtx(k);
            updateLump(m, L.bigLumps[i], firstSrc, upToLump, &lc);
            factorLumpColumn(m, L.bigLumps[i], &lc);
          }
          potrfTrsmPanelBatch<T>(sym.stream, m.batch, all, wv.panels.ptr() + L.panelBegin, L.panelEnd - L.panelBegin,
                                 L.numSmall, wv.panelFlops[lv]);
          for (int k = 0; k < used; k++) {
            B200_CUDA(cudaEventRecord(sym.lanes[k].done, sym.lanes[k].st));
            B200_CUDA(cudaStreamWaitEvent(sym.stream, sym.lanes[k].done, 0));
          }
          continue;
        }
        if (order == 0)
          for (int64_t l : L.bigLumps) updateLump(m, l, firstSrc, upToLump);
        potrfTrsmPanelBatch<T>(sym.stream, m.batch, all, wv.panels.ptr() + L.panelBegin, L.panelEnd - L.panelBegin,
                               L.numSmall, wv.panelFlops[lv]);
        for (int64_t l : L.bigLumps) {
          if (order != 0) updateLump(m, l, firstSrc, upToLump);
          factorLumpColumn(m, l);
        }
      }
      return;
    }
    for (int64_t l = firstSrc; l < skel.numLumps(); l++) {
      updateLump(m, l, firstSrc, upToLump);
      if (l < upToLump) factorLumpColumn(m, l);
    }
  }

  // the eager schedule (see B200SymbolicCtx::EagerPlan)
  // a lane's temp in the eager schedule: the largest contribution into a WIDE lump (the solver's own temp also covers the
  // elimination and the small lumps, and can be much larger)
  size_t eagerTempBytes(const typename B200SymbolicCtx::EagerPlan& E) const {
    return (size_t)std::min<int64_t>(E.maxTempElems, std::max<int64_t>(1, tempSize)) * batch * sizeof(T);
  }
  void fusedFactorEager(const Mats<T>& m, Operand<T> all, const typename B200SymbolicCtx::DevWave& wv, int64_t firstSrc) {
    const auto& E = sym.eagerPlan(firstSrc, wv.host);
    const size_t nLevels = wv.host.levels.size();
    sym.ensureLanes(m.batch, laneTempBytes());
    sym.ensureEager(E, nLevels, eagerTempBytes(E));
    const int nNear = (int)sym.lanes.size(), nBg = (int)sym.bgLanes.size();
    auto ctxOf = [&](int q) {  // q < nNear: lane, else background lane
      if (q < nNear) return laneCtx(q);
      LaneCtx lc;
      const int b = q - nNear;
      lc.st = sym.bgLanes[b].st;
      lc.temp.base = (T*)(sym.bgScratch.ptr() + (size_t)b * eagerTempBytes(E));
      lc.temp.stride = (int64_t)(eagerTempBytes(E) / sizeof(T) / batch);
      lc.spanToChainOffset = sym.bgLanes[b].spanToChainOffset.ptr();
      return lc;
    };
    // everything queued on the solver's stream so far (the sparse elimination) precedes the lanes
    B200_CUDA(cudaEventRecord(sym.evLevel, sym.stream));
    for (int q = 0; q < nNear + nBg; q++) B200_CUDA(cudaStreamWaitEvent(ctxOf(q).st, sym.evLevel, 0));
    // background lane of a target: by its LEVEL first (a stream is a FIFO: contributions into far targets queued ahead
    // of those into the level that comes up next would hold that level back), then by its position in the level
    auto bgOf = [&](int64_t t) {
      const int perLevel = nBg >= 8 ? 2 : 1, nl = std::max(1, nBg / perLevel);
      return (E.level[t] % nl) * perLevel + E.laneOf[t] % perLevel;
    };
    std::vector<int64_t> prepared(nNear + nBg, -1);   // target whose span-to-chain table the stream holds
    std::vector<int64_t> waited;                      // (stream, event index) pairs already waited for in this level
    std::vector<char> mainWaited(E.numWide, 0);
    for (size_t lv = 0; lv < nLevels; lv++) {
      const WaveLevel& L = wv.host.levels[lv];
      bool mainWork = false;
      if (L.tileEnd > L.tileBegin || L.panelEnd > L.panelBegin) {
        // the small lumps of the level (solver's stream) take contributions from every earlier source
        for (size_t pl = 0; pl < lv; pl++)
          for (int64_t b : wv.host.levels[pl].bigLumps)
            if (!mainWaited[E.evIndex[b]]) {
              B200_CUDA(cudaStreamWaitEvent(sym.stream, sym.evDone[E.evIndex[b]], 0));
              mainWaited[E.evIndex[b]] = 1;
            }
        waveUpdate<T>(sym.stream, m.batch, all, wv.tiles.ptr() + L.tileBegin, L.tileEnd - L.tileBegin, wv.targets.ptr(),
                      wv.sources.ptr(), wv.rowMap.ptr());
        potrfTrsmPanelBatch<T>(sym.stream, m.batch, all, wv.panels.ptr() + L.panelBegin, L.panelEnd - L.panelBegin,
                               L.numSmall, wv.panelFlops[lv]);
        B200_CUDA(cudaEventRecord(sym.evMain[lv], sym.stream));
        mainWork = true;
      }
      // the wide lumps of the level: every contribution is already queued on their lane
      lumpCholSetConcurrency((int)std::min<size_t>(L.bigLumps.size(), (size_t)nNear));
      for (int64_t l : L.bigLumps) {
        LaneCtx lc = laneCtx(E.laneOf[l]);
        factorLumpColumn(m, l, &lc);
        B200_CUDA(cudaEventRecord(sym.evDone[E.evIndex[l]], lc.st));
      }
      lumpCholSetConcurrency(0);
      // contributions of this level's lumps: next-level targets first (their lanes), then the background
      waited.clear();
      for (int pass = 0; pass < 2; pass++)
        for (const auto& u : E.fromLevel[lv]) {
          const int64_t t = u.target, src = skel.boardColLump[u.boardRow];
          const bool near = E.level[t] == (int32_t)lv + 1;
          if (near != (pass == 0)) continue;
          const int q = near ? E.laneOf[t] : nNear + bgOf(t);
          LaneCtx lc = ctxOf(q);
          if (near && E.hasBg[t] && prepared[q] != t) {
            // first near contribution of t: its background contributions (all queued by now) come first
            const int qb = nNear + bgOf(t);
            B200_CUDA(cudaEventRecord(sym.evBg[E.evIndex[t]], ctxOf(qb).st));
            B200_CUDA(cudaStreamWaitEvent(lc.st, sym.evBg[E.evIndex[t]], 0));
          }
          const int64_t evKey = (int64_t)q * (E.numWide + 1) + (E.evIndex[src] >= 0 ? E.evIndex[src] : E.numWide);
          if (std::find(waited.begin(), waited.end(), evKey) == waited.end()) {
            if (E.evIndex[src] >= 0) B200_CUDA(cudaStreamWaitEvent(lc.st, sym.evDone[E.evIndex[src]], 0));
            else if (mainWork) B200_CUDA(cudaStreamWaitEvent(lc.st, sym.evMain[lv], 0));
            waited.push_back(evKey);
          }
          if (prepared[q] != t) {
            const int64_t begin = skel.chainColPtr[t];
            b200::prepareAssemble(lc.st, sym.dsk, lc.spanToChainOffset, begin, skel.chainColPtr[t + 1] - begin);
            prepared[q] = t;
          }
          updateOne(m, t, u.boardRow, lc.st, lc.spanToChainOffset, lc.temp);
        }
    }
    for (int q = 0; q < nNear + nBg; q++) {
      cudaEvent_t ev = q < nNear ? sym.lanes[q].done : sym.bgLanes[q - nNear].done;
      B200_CUDA(cudaEventRecord(ev, ctxOf(q).st));
      B200_CUDA(cudaStreamWaitEvent(sym.stream, ev, 0));
    }
  }

  // where one lump column's chain of kernels runs: the solver's stream and shared workspaces, or a lane
  struct LaneCtx {
    cudaStream_t st;
    Work<T> temp;
    int64_t* spanToChainOffset;
  };
  size_t laneTempBytes() const { return (size_t)std::max<int64_t>(1, tempSize) * batch * sizeof(T); }
  LaneCtx laneCtx(int k) {
    LaneCtx lc;
    lc.st = sym.lanes[k].st;
    lc.temp.base = (T*)(sym.laneScratch.ptr() + (size_t)k * laneTempBytes());
    lc.temp.stride = tempSize;
    lc.spanToChainOffset = sym.lanes[k].spanToChainOffset.ptr();
    return lc;
  }

  // contributions of the already factored sources [firstSrc, min(l, upToLump)) into lump l: GEMM into the temp + scatter
  void updateLump(const Mats<T>& m, int64_t l, int64_t firstSrc, int64_t upToLump, const LaneCtx* lc = nullptr) {
    cudaStream_t st = lc ? lc->st : sym.stream;
    int64_t* s2c = lc ? lc->spanToChainOffset : sym.spanToChainOffset.ptr();
    bool prepared = false;
    for (int64_t r = skel.boardRowPtr[l], rEnd = skel.boardRowPtr[l + 1] - 1; r < rEnd; r++) {
      const int64_t src = skel.boardColLump[r];
      if (src >= upToLump) break;
      if (src < firstSrc) continue;
      if (!prepared) {
        const int64_t begin = skel.chainColPtr[l];
        b200::prepareAssemble(st, sym.dsk, s2c, begin, skel.chainColPtr[l + 1] - begin);
        prepared = true;
      }
      updateOne(m, l, r, st, s2c, lc ? lc->temp : temp());
    }
  }

  // contribution of the source lump of board row r into lump l: GEMM into the temp + scatter
  void updateOne(const Mats<T>& m, int64_t l, int64_t r, cudaStream_t st, int64_t* s2c, const Work<T>& tmp) {
    const int64_t src = skel.boardColLump[r];
    const int64_t ord = skel.boardColOrd[r], cb = skel.chainColPtr[src], bb = skel.boardColPtr[src];
    const int64_t k = skel.lumpSize(src);
    const int64_t ch0 = skel.boardChainColOrd[bb + ord], ch1 = skel.boardChainColOrd[bb + ord + 1];
    const int64_t chEnd = skel.boardChainColOrd[skel.boardColPtr[src + 1] - 1];
    const int64_t rowBegin = skel.chainRowsTillEnd[cb + ch0 - 1];
    const int64_t rowsInBoard = skel.chainRowsTillEnd[cb + ch1 - 1] - rowBegin;
    const int64_t rowsToEnd = skel.chainRowsTillEnd[cb + chEnd - 1] - rowBegin;
    BASPACHO_CHECK_LE(rowsInBoard * rowsToEnd, tempSize);
    Operand<T> B = opnd(m, skel.chainData[cb + ch0]);
    gemmNT<T>(st, m.batch, rowsToEnd, rowsInBoard, k, T(1), B, k, B, k, T(0), opnd(tmp, 0), rowsInBoard, false);
    b200::assemble<T>(st, m.batch, sym.dsk, s2c, m, tmp, rowBegin, skel.lumpSize(l), cb + ch0, rowsInBoard, chEnd - ch0,
                      ch1 - ch0, rowsToEnd);
  }

  void factorLumpColumn(const Mats<T>& m, int64_t l, const LaneCtx* lc = nullptr) {
    const int64_t n = skel.lumpSize(l);
    potrfTrapezoid<T>(lc ? lc->st : sym.stream, m.batch, n, skel.lumpTotalRows(l) - n, opnd(m, skel.lumpDataOffset(l)), n);
  }

  B200SymbolicCtx& sym;
  const CoalescedBlockMatrixSkel& skel;
  int64_t tempSize;
  int batch;
  MatsOf<TT> mats;
};

template <typename TT>
struct B200SolveCtx : SolveCtx<TT> {
  using T = BaseType<TT>;

  B200SolveCtx(B200SymbolicCtx& s, int nRHS_, int batchSize) : sym(s), skel(s.skel), nRHS(nRHS_), batch(batchSize) {
    tlsSyncStream = sym.stream;
  }

  // scratch per batch item: [0] the gemv/assembleVec temp of the reference's solve ctx (order x nRHS), [1] the
  // solved-block staging of the blocked dense triangular solve (order x nRHS), [2] the inverses of the 96 x 96
  // diagonal blocks of the widest lump (2 x 96 x 96 per block)
  int64_t vecStride() const { return std::max<int64_t>(1, skel.order() * nRHS); }
  // block inverses of EVERY wide lump live side by side in the scratch (slot table by diagonal-block offset), so the
  // backward pass of solve() reuses what the forward pass computed
  int64_t invTotal() { return sym.useInverseSolve ? sym.invPlan().total : 0; }
  T* scratchBase() { return (T*)sym.scratch((size_t)(2 * vecStride() + invTotal()) * batch * sizeof(T)); }
  Work<T> temp(int which = 0) {
    Work<T> w;
    w.stride = vecStride();
    w.base = scratchBase() + (size_t)which * w.stride * batch;
    return w;
  }
  Operand<T> invBase() {
    Operand<T> o;
    o.base = scratchBase() + (size_t)2 * vecStride() * batch;
    o.bstride = invTotal();
    return o;
  }
  // scratch operand for the inverses of the lump whose diagonal block starts at offM (null operand: not eligible)
  Operand<T> invScratch(int64_t offM, int64_t n, bool* ready) {
    Operand<T> o;
    *ready = false;
    if (!sym.useInverseSolve) return o;
    const auto& slot = sym.invPlan().slot;
    auto it = slot.find(offM);
    if (it == slot.end() || it->second.second != n) return o;
    o = invBase();
    o.base += it->second.first;
    *ready = invAll || invDone.count(offM) > 0;
    invDone.insert(offM);
    return o;
  }
  std::set<int64_t> invDone;  // inverses computed through this context
  bool invAll = false;        // ... all of them, by the one-launch inversion
  const void* invData = nullptr;
  void invCheckData(const TT* data) {  // a context serves one factor; a different matrix invalidates the inverses
    if (invData != (const void*)data) invDone.clear(), invAll = false;
    invData = data;
  }
  // all block inverses of the factor in ONE launch (instead of one serial 96-step launch per lump inside the chain)
  void invertAll(const TT* data) {
    invCheckData(data);
    if (invAll || !sym.useInverseSolve) return;
    const auto& ip = sym.invPlan();
    if (ip.count > 0) {
      Mats<T> m = mats.get(data, sym.stream);
      invertBlockList<T>(sym.stream, m.batch, ip.list.ptr(), ip.count, opnd(m, 0), invBase());
    }
    invAll = true;
  }

  // ---- whole-range solves (MatOps.h addition): the sequence of Solver::internalSolveLRange / internalSolveLtRange
  // (reference Solver.cpp:268-397) with every block inverse of the factor computed up front in one launch
  bool hasFusedSolve() override { return true; }

  void fusedSolveL(const TT* data, int64_t startLump, int64_t upToLump, TT* C, int64_t ldc) override {
    int64_t denseFrom = 0;
    for (const B200SymElimCtx* e : sym.elimRegistry) {
      denseFrom = std::max(denseFrom, e->dev.lumpsEnd);
      if (e->dev.lumpsEnd > upToLump) return;
      if (startLump > e->dev.lumpsBegin) continue;
      sparseElimSolveL(*e, data, e->dev.lumpsBegin, e->dev.lumpsEnd, C, ldc);
    }
    denseFrom = std::max(denseFrom, startLump);
    if (denseFrom < upToLump) invertAll(data);
    if (lanesUsable(denseFrom, startLump, upToLump)) return lanedSolveL(data, denseFrom, C, ldc);
    for (int64_t l = denseFrom; l < upToLump; l++) {
      const int64_t n = skel.lumpSize(l), start = skel.lumpStart[l], rows = skel.lumpTotalRows(l) - n;
      const int64_t off = skel.lumpDataOffset(l);
      // x_l = L_ll^-1 b_l, then C[rows below] -= L21 x_l (gemv + assembleVec of the reference sequence): inside the
      // chained launch for lumps of two or more blocks, else one gemv kernel that scatters through the row table
      const auto& br = sym.belowRows();
      BASPACHO_CHECK_EQ(br.ptr[l + 1] - br.ptr[l], rows);
      Mats<T> m = mats.get(data, sym.stream), v = vecs.get(C, sym.stream);
      bool ready = false;
      Operand<T> inv = invScratch(off, n, &ready);
      // BSPB200_CHAIN_FUSE_GEMV=0: rows below in a separate gemv launch (GRID solve 11.4 ms vs 10.7 ms fused)
      static const bool fuseBelow = !(getenv("BSPB200_CHAIN_FUSE_GEMV") && atoi(getenv("BSPB200_CHAIN_FUSE_GEMV")) == 0);
      const bool belowDone =
          trsvAny<T>(sym.stream, m.batch, n, opnd(m, off), n, opnd(v, start), ldc, nRHS, false, opnd(temp2(), 0), inv, ready,
                     sym.chainSync(m.batch), fuseBelow ? rows : 0, br.target.ptr() + br.ptr[l], opnd(v, 0));
      if (rows == 0 || belowDone) continue;
      gemvRows<T>(sym.stream, m.batch, rows, n, T(-1), opnd(m, off + n * n), n, opnd(v, start), ldc, opnd(v, 0), 1, ldc, nRHS,
                  true, br.target.ptr() + br.ptr[l]);
    }
  }

  void fusedSolveLt(const TT* data, int64_t startLump, int64_t upToLump, TT* C, int64_t ldc) override {
    int64_t denseFrom = 0;
    for (const B200SymElimCtx* e : sym.elimRegistry) denseFrom = std::max(denseFrom, e->dev.lumpsEnd);
    denseFrom = std::max(denseFrom, startLump);
    if (denseFrom < upToLump) invertAll(data);
    const bool laned = lanesUsable(denseFrom, startLump, upToLump);
    if (laned) lanedSolveLt(data, denseFrom, C, ldc);
    for (int64_t l = upToLump - 1; l >= denseFrom && !laned; l--) {
      const int64_t n = skel.lumpSize(l), start = skel.lumpStart[l], rows = skel.lumpTotalRows(l) - n;
      const int64_t off = skel.lumpDataOffset(l);
      if (rows > 0) {
        // x_l -= L21^T C[rows below]: assembleVecT + gemvT in one product (gather through the row table)
        const auto& br = sym.belowRows();
        BASPACHO_CHECK_EQ(br.ptr[l + 1] - br.ptr[l], rows);
        Mats<T> m = mats.get(data, sym.stream), v = vecs.get(C, sym.stream);
        gemvColsT<T>(sym.stream, m.batch, rows, n, T(-1), opnd(m, off + n * n), n, opnd(v, 0), 1, ldc, opnd(v, start), ldc,
                     nRHS, opnd(temp2(), 0), vecStride(), br.target.ptr() + br.ptr[l]);
      }
      solveLt(data, off, n, C, start, ldc);
    }
    for (int64_t r = (int64_t)sym.elimRegistry.size() - 1; r >= 0; r--) {
      const B200SymElimCtx* e = sym.elimRegistry[r];
      if (e->dev.lumpsEnd > upToLump) continue;
      if (e->dev.lumpsBegin < startLump) return;
      sparseElimSolveLt(*e, data, e->dev.lumpsBegin, e->dev.lumpsEnd, C, ldc);
    }
  }
  Work<T> temp2() { return temp(1); }

  // ---- the dense part of a whole-matrix solve on lanes (see B200SymbolicCtx::SolvePlan)
  bool lanesUsable(int64_t denseFrom, int64_t startLump, int64_t upToLump) {
    static const bool timeline = getenv("BSPB200_PROFILE_TIMELINE") && atoi(getenv("BSPB200_PROFILE_TIMELINE")) != 0;
    if (upToLump != skel.numLumps() || startLump > denseFrom || denseFrom >= upToLump) return false;
    if (profileEnabled() && !timeline) return false;
    if (!sym.useInverseSolve || !sym.useChainSolve) return false;
    return sym.solvePlan(denseFrom).eligible;
  }
  // lane k: [0] its delta vector, [1] its staging for the per-step solves; ldc x nRHS per batch item each
  Work<T> laneWork(int k, int which, int64_t ldc) {
    const int64_t stride = ldc * nRHS;
    const int nLanes = sym.numLanes;
    T* base = (T*)sym.solveLaneScratch((size_t)2 * nLanes * stride * batch * sizeof(T));
    Work<T> w;
    w.stride = stride;
    w.base = base + ((size_t)which * nLanes + k) * stride * batch;
    return w;
  }
  void lanedSolveL(const TT* data, int64_t denseFrom, TT* C, int64_t ldc) {
    const auto& P = sym.solvePlan(denseFrom);
    sym.ensureSolveLanes(P.numDense);
    const auto& br = sym.belowRows();
    Mats<T> m = mats.get(data, sym.stream), v = vecs.get(C, sym.stream);
    const int nLanes = sym.numLanes;
    (void)laneWork(0, 0, ldc);  // allocate (and zero) before anything is queued
    B200_CUDA(cudaEventRecord(sym.evLevel, sym.stream));
    for (int k = 0; k < nLanes; k++) B200_CUDA(cudaStreamWaitEvent(sym.lanes[k].st, sym.evLevel, 0));
    static const bool fuseBelow = !(getenv("BSPB200_CHAIN_FUSE_GEMV") && atoi(getenv("BSPB200_CHAIN_FUSE_GEMV")) == 0);
    for (const auto& lumps : P.byLevel)
      for (int64_t l : lumps) {
        const int k = P.lane[l];
        cudaStream_t st = sym.lanes[k].st;
        const int64_t n = skel.lumpSize(l), start = skel.lumpStart[l], rows = skel.lumpTotalRows(l) - n;
        const int64_t off = skel.lumpDataOffset(l);
        for (int32_t sLump : P.srcs[l])
          if (P.lane[sLump] != k) B200_CUDA(cudaStreamWaitEvent(st, sym.evSolve[sLump - denseFrom], 0));
        if (!P.srcs[l].empty())
          gatherLaneDeltas<T>(st, m.batch, n, nRHS, ldc, opnd(laneWork(0, 0, ldc), start), ldc * nRHS * batch, nLanes,
                              opnd(v, start));
        BASPACHO_CHECK_EQ(br.ptr[l + 1] - br.ptr[l], rows);
        bool ready = false;
        Operand<T> inv = invScratch(off, n, &ready);
        Operand<T> delta = opnd(laneWork(k, 0, ldc), 0);
        const bool belowDone =
            trsvAny<T>(st, m.batch, n, opnd(m, off), n, opnd(v, start), ldc, nRHS, false, opnd(laneWork(k, 1, ldc), 0), inv,
                       ready, sym.chainSyncLane(k, m.batch), fuseBelow ? rows : 0, br.target.ptr() + br.ptr[l], delta);
        if (rows > 0 && !belowDone)
          gemvRows<T>(st, m.batch, rows, n, T(-1), opnd(m, off + n * n), n, opnd(v, start), ldc, delta, 1, ldc, nRHS, true,
                      br.target.ptr() + br.ptr[l]);
        B200_CUDA(cudaEventRecord(sym.evSolve[l - denseFrom], st));
      }
    for (int k = 0; k < nLanes; k++) {
      B200_CUDA(cudaEventRecord(sym.lanes[k].done, sym.lanes[k].st));
      B200_CUDA(cudaStreamWaitEvent(sym.stream, sym.lanes[k].done, 0));
    }
  }
  void lanedSolveLt(const TT* data, int64_t denseFrom, TT* C, int64_t ldc) {
    const auto& P = sym.solvePlan(denseFrom);
    sym.ensureSolveLanes(P.numDense);
    const auto& br = sym.belowRows();
    Mats<T> m = mats.get(data, sym.stream), v = vecs.get(C, sym.stream);
    const int nLanes = sym.numLanes;
    (void)laneWork(0, 0, ldc);
    B200_CUDA(cudaEventRecord(sym.evLevel, sym.stream));
    for (int k = 0; k < nLanes; k++) B200_CUDA(cudaStreamWaitEvent(sym.lanes[k].st, sym.evLevel, 0));
    for (int64_t lv = (int64_t)P.byLevel.size() - 1; lv >= 0; lv--)
      for (int64_t l : P.byLevel[lv]) {
        const int k = P.lane[l];
        cudaStream_t st = sym.lanes[k].st;
        const int64_t n = skel.lumpSize(l), start = skel.lumpStart[l], rows = skel.lumpTotalRows(l) - n;
        const int64_t off = skel.lumpDataOffset(l);
        for (int32_t t : P.tgts[l])
          if (P.lane[t] != k) B200_CUDA(cudaStreamWaitEvent(st, sym.evSolve[t - denseFrom], 0));
        Work<T> stage = laneWork(k, 1, ldc);
        if (rows > 0) {
          BASPACHO_CHECK_EQ(br.ptr[l + 1] - br.ptr[l], rows);
          gemvColsT<T>(st, m.batch, rows, n, T(-1), opnd(m, off + n * n), n, opnd(v, 0), 1, ldc, opnd(v, start), ldc, nRHS,
                       opnd(stage, 0), stage.stride, br.target.ptr() + br.ptr[l]);
        }
        bool ready = false;
        Operand<T> inv = invScratch(off, n, &ready);
        trsvAny<T>(st, m.batch, n, opnd(m, off), n, opnd(v, start), ldc, nRHS, true, opnd(stage, 0), inv, ready,
                   sym.chainSyncLane(k, m.batch));
        B200_CUDA(cudaEventRecord(sym.evSolve[l - denseFrom], st));
      }
    for (int k = 0; k < nLanes; k++) {
      B200_CUDA(cudaEventRecord(sym.lanes[k].done, sym.lanes[k].st));
      B200_CUDA(cudaStreamWaitEvent(sym.stream, sym.lanes[k].done, 0));
    }
  }

  // ---- fragmented whole-range ops (reference MatOps.h:168-183, MatOpsFast.cpp:613-1018): level-scheduled block
  // kernels when every lump is a single small span and nRHS == 1 (SparseKernels.cu)
  bool hasFragmentedOps() override { return nRHS == 1 && sym.fragPlan().eligible; }

  void fragmentedMV(const TT* data, const TT* x, int64_t spanBegin, int64_t spanEnd, TT* y, T alpha) override {
    const auto& fp = sym.fragPlan();
    Mats<T> m = mats.get(data, sym.stream), vx = vecs.get(x, sym.stream), vy = vecs2.get(y, sym.stream);
    fragMV<T>(sym.stream, m.batch, fp.dev, sym.dsk, m, vx, vy, spanBegin, spanEnd, alpha, fp.nnz * sizeof(T));
  }

  void fragmentedSolveL(const TT* data, int64_t spanBegin, int64_t spanEnd, TT* y) override {
    const auto& fp = sym.fragPlan();
    Mats<T> m = mats.get(data, sym.stream), vy = vecs.get(y, sym.stream);
    ProfScope prof(sym.stream, KC_SOLVE_DENSE, 0, fp.nnz * sizeof(T) * m.batch);
    for (size_t l = 0; l + 1 < fp.lvlPtrL.size(); l++) {
      // the spans of a level are stored ascending: those inside [spanBegin, spanEnd) are one contiguous piece
      const int32_t* b = fp.hostLvlL.data() + fp.lvlPtrL[l];
      const int32_t* e = fp.hostLvlL.data() + fp.lvlPtrL[l + 1];
      const int32_t* lo = std::lower_bound(b, e, (int32_t)spanBegin);
      const int32_t* hi = std::lower_bound(lo, e, (int32_t)spanEnd);
      fragSolveLLevel<T>(sym.stream, m.batch, fp.dev, sym.dsk, m, vy, fp.lvlL.ptr() + (lo - fp.hostLvlL.data()),
                         (int)(hi - lo), spanBegin, spanEnd, true);
    }
    // rows behind the range receive the updates of its columns (no diagonal solve)
    fragSolveLLevel<T>(sym.stream, m.batch, fp.dev, sym.dsk, m, vy, fp.tailList.ptr() + spanEnd,
                       (int)(skel.numSpans() - spanEnd), spanBegin, spanEnd, false);
  }

  void fragmentedSolveLt(const TT* data, int64_t spanBegin, int64_t spanEnd, TT* y) override {
    const auto& fp = sym.fragPlan();
    Mats<T> m = mats.get(data, sym.stream), vy = vecs.get(y, sym.stream);
    ProfScope prof(sym.stream, KC_SOLVE_DENSE, 0, fp.nnz * sizeof(T) * m.batch);
    for (size_t l = 0; l + 1 < fp.lvlPtrT.size(); l++) {
      const int32_t* b = fp.hostLvlT.data() + fp.lvlPtrT[l];
      const int32_t* e = fp.hostLvlT.data() + fp.lvlPtrT[l + 1];
      const int32_t* lo = std::lower_bound(b, e, (int32_t)spanBegin);
      const int32_t* hi = std::lower_bound(lo, e, (int32_t)spanEnd);
      fragSolveLtLevel<T>(sym.stream, m.batch, sym.dsk, m, vy, fp.lvlT.ptr() + (lo - fp.hostLvlT.data()), (int)(hi - lo));
    }
  }

  void sparseElimSolveL(const SymElimCtx& elimData, const TT* data, int64_t lumpsBegin, int64_t lumpsEnd, TT* C,
                        int64_t ldc) override {
    auto timer = sym.solveSparseLStat.template instance<B200SyncOps>();
    const auto* elim = dynamic_cast<const B200SymElimCtx*>(&elimData);
    BASPACHO_CHECK_NOTNULL(elim);
    BASPACHO_CHECK_EQ(elim->dev.lumpsBegin, lumpsBegin);
    BASPACHO_CHECK_EQ(elim->dev.lumpsEnd, lumpsEnd);
    Mats<T> m = mats.get(data, sym.stream), v = vecs.get(C, sym.stream);
    elimSolveL<T>(sym.stream, m.batch, sym.dsk, elim->dev, m, v, ldc, nRHS);
  }

  void sparseElimSolveLt(const SymElimCtx& elimData, const TT* data, int64_t lumpsBegin, int64_t lumpsEnd, TT* C,
                         int64_t ldc) override {
    auto timer = sym.solveSparseLtStat.template instance<B200SyncOps>();
    const auto* elim = dynamic_cast<const B200SymElimCtx*>(&elimData);
    BASPACHO_CHECK_NOTNULL(elim);
    BASPACHO_CHECK_EQ(elim->dev.lumpsBegin, lumpsBegin);
    BASPACHO_CHECK_EQ(elim->dev.lumpsEnd, lumpsEnd);
    Mats<T> m = mats.get(data, sym.stream), v = vecs.get(C, sym.stream);
    elimSolveLt<T>(sym.stream, m.batch, sym.dsk, elim->dev, m, v, ldc, nRHS);
  }

  // addMvFrom over a whole sparse-elimination range in two launches (MatOps.h addition; the reference walks the lumps)
  bool hasSparseElimMV() override {
    static const bool on = !getenv("BSPB200_SPARSE_MV") || atoi(getenv("BSPB200_SPARSE_MV")) != 0;
    return on;
  }
  void sparseElimMV(const SymElimCtx& elimData, const TT* data, const TT* in, int64_t inStride, TT* out, int64_t outStride,
                    T alpha) override {
    const auto* elim = dynamic_cast<const B200SymElimCtx*>(&elimData);
    BASPACHO_CHECK_NOTNULL(elim);
    Mats<T> m = mats.get(data, sym.stream), x = vecs.get(in, sym.stream), y = vecs2.get(out, sym.stream);
    elimMV<T>(sym.stream, m.batch, sym.dsk, elim->dev, m, x, inStride, y, outStride, nRHS, alpha);
  }

  void symm(const TT* data, int64_t offM, int64_t n, const TT* C, int64_t offC, int64_t ldc, TT* D, int64_t ldd,
            T alpha) override {
    auto timer = sym.symmStat.template instance<B200SyncOps>();
    Mats<T> m = mats.get(data, sym.stream), x = vecs.get(C, sym.stream), y = vecs2.get(D, sym.stream);
    symmLower<T>(sym.stream, m.batch, n, alpha, opnd(m, offM), opnd(x, offC), ldc, opnd(y, offC), ldd, nRHS);
  }

  void solveL(const TT* data, int64_t offM, int64_t n, TT* C, int64_t offC, int64_t ldc) override {
    auto timer = sym.solveLStat.template instance<B200SyncOps>();
    Mats<T> m = mats.get(data, sym.stream), v = vecs.get(C, sym.stream);
    invCheckData(data);
    bool ready = false;
    Operand<T> inv = invScratch(offM, n, &ready);
    trsvAny<T>(sym.stream, m.batch, n, opnd(m, offM), n, opnd(v, offC), ldc, nRHS, false, opnd(temp2(), 0), inv, ready,
               sym.chainSync(m.batch));
  }

  void solveLt(const TT* data, int64_t offM, int64_t n, TT* C, int64_t offC, int64_t ldc) override {
    auto timer = sym.solveLtStat.template instance<B200SyncOps>();
    Mats<T> m = mats.get(data, sym.stream), v = vecs.get(C, sym.stream);
    invCheckData(data);
    bool ready = false;
    Operand<T> inv = invScratch(offM, n, &ready);
    trsvAny<T>(sym.stream, m.batch, n, opnd(m, offM), n, opnd(v, offC), ldc, nRHS, true, opnd(temp2(), 0), inv, ready,
               sym.chainSync(m.batch));
  }

  void gemv(const TT* data, int64_t offM, int64_t nRows, int64_t nCols, const TT* A, int64_t offA, int64_t lda,
            T alpha) override {
    auto timer = sym.solveGemvStat.template instance<B200SyncOps>();
    Mats<T> m = mats.get(data, sym.stream), v = vecs.get(A, sym.stream);
    gemvRows<T>(sym.stream, m.batch, nRows, nCols, alpha, opnd(m, offM), nCols, opnd(v, offA), lda, opnd(temp(), 0),
                nRHS, 1, nRHS, false);
  }

  void gemvT(const TT* data, int64_t offM, int64_t nRows, int64_t nCols, TT* A, int64_t offA, int64_t lda,
             T alpha) override {
    auto timer = sym.solveGemvTStat.template instance<B200SyncOps>();
    Mats<T> m = mats.get(data, sym.stream), v = vecs.get(A, sym.stream);
    gemvColsT<T>(sym.stream, m.batch, nRows, nCols, alpha, opnd(m, offM), nCols, opnd(temp(), 0), nRHS, 1,
                 opnd(v, offA), lda, nRHS, opnd(temp2(), 0), vecStride());
  }

  int64_t rowsOfChains(int64_t chainColPtr, int64_t numColItems) const {
    return skel.chainRowsTillEnd[chainColPtr + numColItems - 1] - skel.chainRowsTillEnd[chainColPtr - 1];
  }

  void assembleVec(int64_t chainColPtr, int64_t numColItems, TT* C, int64_t ldc) override {
    auto timer = sym.solveAssVStat.template instance<B200SyncOps>();
    if (numColItems <= 0) return;
    Mats<T> v = vecs.get(C, sym.stream);
    b200::assembleVec<T>(sym.stream, v.batch, sym.dsk, temp(), chainColPtr, numColItems,
                         rowsOfChains(chainColPtr, numColItems), v, ldc, nRHS);
  }

  void assembleVecT(const TT* C, int64_t ldc, int64_t chainColPtr, int64_t numColItems) override {
    auto timer = sym.solveAssVTStat.template instance<B200SyncOps>();
    if (numColItems <= 0) return;
    Mats<T> v = vecs.get(C, sym.stream);
    b200::assembleVecT<T>(sym.stream, v.batch, sym.dsk, temp(), chainColPtr, numColItems,
                          rowsOfChains(chainColPtr, numColItems), v, ldc, nRHS);
  }

  B200SymbolicCtx& sym;
  const CoalescedBlockMatrixSkel& skel;
  int nRHS, batch;
  MatsOf<TT> mats, vecs, vecs2;
};

NumericCtxBase* B200SymbolicCtx::createNumericCtxForType(std::type_index tIdx, int64_t tempBufSize, int batchSize) {
  if (tIdx == std::type_index(typeid(double))) {
    BASPACHO_CHECK_EQ(batchSize, 1);
    return new B200NumericCtx<double>(*this, tempBufSize, 1);
  }
  if (tIdx == std::type_index(typeid(float))) {
    BASPACHO_CHECK_EQ(batchSize, 1);
    return new B200NumericCtx<float>(*this, tempBufSize, 1);
  }
  if (tIdx == std::type_index(typeid(vector<double*>)))
    return new B200NumericCtx<vector<double*>>(*this, tempBufSize, batchSize);
  if (tIdx == std::type_index(typeid(vector<float*>)))
    return new B200NumericCtx<vector<float*>>(*this, tempBufSize, batchSize);
  return nullptr;
}

SolveCtxBase* B200SymbolicCtx::createSolveCtxForType(std::type_index tIdx, int nRHS, int batchSize) {
  if (tIdx == std::type_index(typeid(double))) {
    BASPACHO_CHECK_EQ(batchSize, 1);
    return new B200SolveCtx<double>(*this, nRHS, 1);
  }
  if (tIdx == std::type_index(typeid(float))) {
    BASPACHO_CHECK_EQ(batchSize, 1);
    return new B200SolveCtx<float>(*this, nRHS, 1);
  }
  if (tIdx == std::type_index(typeid(vector<double*>))) return new B200SolveCtx<vector<double*>>(*this, nRHS, batchSize);
  if (tIdx == std::type_index(typeid(vector<float*>))) return new B200SolveCtx<vector<float*>>(*this, nRHS, batchSize);
  return nullptr;
}

struct B200Ops : Ops {
  SymbolicCtxPtr createSymbolicCtx(const CoalescedBlockMatrixSkel& skel, const vector<int64_t>& permutation) override {
    return SymbolicCtxPtr(new B200SymbolicCtx(skel, permutation));
  }
};

}  // namespace

OpsPtr b200Ops() { return OpsPtr(new B200Ops); }

// Host-staged pipeline (capi.cpp, bspb200_factor_solve_host): elimination plans of consecutive sub-ranges ("chunks") of
// one elimination range, so that the elimination of a chunk of point columns can start as soon as the chunk has been
// uploaded. The plans are ordinary elimination contexts (run through NumericCtx::doElimination) but are NOT ranges of
// the Solver. Running the chunks in order gives a fixed summation order (deterministic; it differs from the one-plan
// order of factor() in the last bits).
std::vector<SymElimCtxPtr> b200ElimChunkPlans(SymbolicCtx& sym, const std::vector<int64_t>& bounds) {
  auto* bs = dynamic_cast<B200SymbolicCtx*>(&sym);
  BASPACHO_CHECK_NOTNULL(bs);
  std::vector<SymElimCtxPtr> out;
  for (size_t i = 0; i + 1 < bounds.size(); i++) out.emplace_back(bs->makeElimCtx(bounds[i], bounds[i + 1]).release());
  return out;
}

}  // namespace BaSpaCho
