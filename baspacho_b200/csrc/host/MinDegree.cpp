// From-scratch approximate-minimum-degree ordering on a quotient graph (variables + elements),
// after Amestoy/Davis/Duff: element absorption, approximate external degrees via the
// |Le \ Lp| scan, aggressive absorption, mass elimination and hash-based supervariable merging.
//
// Replaces the third-party ordering the reference calls in SparseStructure::fillReducingPermutation
// (Eigen::AMDOrdering, reference baspacho/baspacho/SparseStructure.cpp:313-330; SuiteSparse
// amd_l_order :297-309). The reference's tests pin only the fill quality (SparseStructureTest.cpp:117-152),
// not tie-breaking, so this is a quality-equivalent, not bit-identical, ordering.
#include <algorithm>
#include <cstdint>
#include <vector>
#include "DebugMacros.h"
#include "SparseStructure.h"

namespace BaSpaCho {

namespace {

struct DegreeBuckets {
  std::vector<int64_t> head, next, prev;
  explicit DegreeBuckets(int64_t n) : head(n + 1, -1), next(n, -1), prev(n, -1) {}
  void insert(int64_t v, int64_t d) {
    next[v] = head[d];
    prev[v] = -1;
    if (head[d] >= 0) prev[head[d]] = v;
    head[d] = v;
  }
  void remove(int64_t v, int64_t d) {
    if (prev[v] >= 0) next[prev[v]] = next[v]; else head[d] = next[v];
    if (next[v] >= 0) prev[next[v]] = prev[v];
  }
};

enum NodeState : uint8_t { kVariable, kElement, kDeadElement, kMergedVariable, kEliminated };

}  // namespace

std::vector<int64_t> approximateMinimumDegree(int64_t n, const std::vector<int64_t>& ptrs,
                                              const std::vector<int64_t>& inds) {
  std::vector<std::vector<int64_t>> adjVar(n), adjEl(n), elPattern(n);
  {  // symmetrize, drop diagonal and duplicates
    for (int64_t i = 0; i < n; i++)
      for (int64_t k = ptrs[i]; k < ptrs[i + 1]; k++) {
        int64_t j = inds[k];
        BASPACHO_CHECK_LT(j, n);
        if (j == i) continue;
        adjVar[i].push_back(j);
        adjVar[j].push_back(i);
      }
    for (auto& a : adjVar) {
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end()), a.end());
    }
  }

  std::vector<NodeState> state(n, kVariable);
  std::vector<int64_t> weight(n, 1);          // supervariable size (0 once merged away)
  std::vector<int64_t> degree(n);             // approximate external degree (weighted)
  std::vector<int64_t> elDegree(n, 0);        // weighted |Le| of an element
  std::vector<int64_t> mergedInto(n, -1);     // representative of a merged variable
  std::vector<int64_t> mark(n, -1);           // variable marker (stamp = pivot step)
  std::vector<int64_t> w(n, 0);               // element scan workspace, compared against wflg
  std::vector<int64_t> hashOf(n, 0);
  int64_t wflg = 1;

  DegreeBuckets buckets(n);
  for (int64_t i = 0; i < n; i++) {
    degree[i] = (int64_t)adjVar[i].size();
    buckets.insert(i, degree[i]);
  }

  std::vector<int64_t> order;  // pivots and mass-eliminated variables, in elimination order
  order.reserve(n);
  std::vector<int64_t> Lp, bucketScratch;
  int64_t numEliminated = 0, minDeg = 0, step = 0;

  while (numEliminated < n) {
    while (minDeg <= n && buckets.head[minDeg] < 0) minDeg++;
    BASPACHO_CHECK_LE(minDeg, n);
    int64_t p = buckets.head[minDeg];
    buckets.remove(p, minDeg);
    step++;

    // ---- form the new element p: Lp = (A_p ∪ ⋃ L_e, e in E_p) \ {p}
    Lp.clear();
    mark[p] = step;
    int64_t degLp = 0;
    auto take = [&](int64_t v) {
      if (state[v] != kVariable || mark[v] == step) return;
      mark[v] = step;
      Lp.push_back(v);
      degLp += weight[v];
    };
    for (int64_t v : adjVar[p]) take(v);
    for (int64_t e : adjEl[p]) {
      if (state[e] != kElement) continue;
      for (int64_t v : elPattern[e]) take(v);
      state[e] = kDeadElement;  // absorbed into p
      std::vector<int64_t>().swap(elPattern[e]);
    }
    std::vector<int64_t>().swap(adjVar[p]);
    std::vector<int64_t>().swap(adjEl[p]);
    state[p] = kElement;
    numEliminated += weight[p];
    order.push_back(p);

    for (int64_t v : Lp) buckets.remove(v, degree[v]);

    // ---- w[e] - wflg = |Le \ Lp| for every element touching Lp
    if (wflg + n + 1 < wflg) {  // overflow guard (unreachable for int64, kept for clarity)
      std::fill(w.begin(), w.end(), 0);
      wflg = 1;
    }
    for (int64_t v : Lp)
      for (int64_t e : adjEl[v]) {
        if (state[e] != kElement) continue;
        if (w[e] < wflg) w[e] = elDegree[e] + wflg;
        w[e] -= weight[v];
      }

    // ---- degree update, list pruning, mass elimination
    size_t keep = 0;
    for (size_t idx = 0; idx < Lp.size(); idx++) {
      int64_t v = Lp[idx];
      int64_t deg = 0, hash = 0;
      auto& ev = adjEl[v];
      size_t ne = 0;
      for (int64_t e : ev) {
        if (state[e] != kElement) continue;
        int64_t ext = w[e] - wflg;
        if (ext > 0) {
          deg += ext;
          ev[ne++] = e;
          hash += e;
        } else {  // Le ⊆ Lp: aggressive absorption
          state[e] = kDeadElement;
          std::vector<int64_t>().swap(elPattern[e]);
        }
      }
      ev.resize(ne);
      auto& av = adjVar[v];
      size_t na = 0;
      for (int64_t u : av) {
        if (state[u] != kVariable || mark[u] == step) continue;  // dead, or now covered by element p
        deg += weight[u];
        av[na++] = u;
        hash += u;
      }
      av.resize(na);

      if (deg == 0 && ne == 0) {
        // v is adjacent to nothing outside Lp: eliminate it together with p
        state[v] = kEliminated;
        numEliminated += weight[v];
        degLp -= weight[v];
        order.push_back(v);
        std::vector<int64_t>().swap(av);
        std::vector<int64_t>().swap(ev);
        continue;
      }
      ev.push_back(p);
      hash += p;
      hashOf[v] = hash % n;
      degree[v] = std::min(degree[v], deg);  // |Lp \ v| is added once Lp is final (below)
      Lp[keep++] = v;
    }
    Lp.resize(keep);
    wflg += n + 1;  // invalidates every w[e] set above (elDegree <= n)

    // ---- supervariable detection among Lp (identical adjacency in the quotient graph)
    bucketScratch.assign(Lp.begin(), Lp.end());
    std::sort(bucketScratch.begin(), bucketScratch.end(),
              [&](int64_t a, int64_t b) { return hashOf[a] != hashOf[b] ? hashOf[a] < hashOf[b] : a < b; });
    for (size_t a = 0; a < bucketScratch.size(); a++) {
      int64_t i = bucketScratch[a];
      if (state[i] != kVariable) continue;
      bool sorted_i = false;
      for (size_t b = a + 1; b < bucketScratch.size() && hashOf[bucketScratch[b]] == hashOf[i]; b++) {
        int64_t j = bucketScratch[b];
        if (state[j] != kVariable) continue;
        if (adjVar[i].size() != adjVar[j].size() || adjEl[i].size() != adjEl[j].size()) continue;
        if (!sorted_i) {
          std::sort(adjVar[i].begin(), adjVar[i].end());
          std::sort(adjEl[i].begin(), adjEl[i].end());
          sorted_i = true;
        }
        std::sort(adjVar[j].begin(), adjVar[j].end());
        std::sort(adjEl[j].begin(), adjEl[j].end());
        if (adjVar[i] != adjVar[j] || adjEl[i] != adjEl[j]) continue;
        // j is indistinguishable from i
        weight[i] += weight[j];
        weight[j] = 0;
        state[j] = kMergedVariable;
        mergedInto[j] = i;
        std::vector<int64_t>().swap(adjVar[j]);
        std::vector<int64_t>().swap(adjEl[j]);
      }
    }

    // ---- finalize element p and put the survivors back in the degree lists
    auto& pat = elPattern[p];
    pat.clear();
    for (int64_t v : Lp) {
      if (state[v] != kVariable) continue;
      pat.push_back(v);
      int64_t d = degree[v] + degLp - weight[v];             // + |Lp \ v|
      d = std::min(d, n - numEliminated - weight[v]);
      if (d < 0) d = 0;
      degree[v] = d;
      buckets.insert(v, d);
      if (d < minDeg) minDeg = d;
    }
    elDegree[p] = degLp;
    if (pat.empty()) state[p] = kDeadElement;
  }

  // expand supervariables: merged variables follow their representative
  std::vector<std::vector<int64_t>> members(n);
  for (int64_t j = 0; j < n; j++)
    if (mergedInto[j] >= 0) members[mergedInto[j]].push_back(j);
  std::vector<int64_t> perm;
  perm.reserve(n);
  std::vector<int64_t> stack;
  for (int64_t v : order) {
    stack.push_back(v);
    while (!stack.empty()) {
      int64_t x = stack.back();
      stack.pop_back();
      perm.push_back(x);
      for (auto it = members[x].rbegin(); it != members[x].rend(); ++it) stack.push_back(*it);
    }
  }
  BASPACHO_CHECK_EQ((int64_t)perm.size(), n);
  return perm;
}

}  // namespace BaSpaCho
