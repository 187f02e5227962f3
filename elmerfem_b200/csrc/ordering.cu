// Matrix-structure producer for nodal elements: node graph from the element connectivity, the reference's bandwidth
// optimiser, and the expansion of the (re)numbered node graph to the dofs-per-node CRS structure that
// b200_set_structure takes.  Host-only integer work done once per mesh (no GPU needed, nothing here launches a kernel);
// every output must be bit-identical to what the reference builds, because the row order decides the ILU factors
// and with them the iteration counts.
//
// Reference being replaced:
//   MakeListMatrix, nodal branch        fem/src/ElementUtils.F90:881-891 (rows kept ascending by List_GetMatrixIndex,
//                                       fem/src/ListMatrix.F90:334-386)
//   OptimizeBandwidth                   fem/src/BandwidthOptimize.F90:182-445
//   InitializeMatrix + CRS_SortMatrix   fem/src/ElementUtils.F90:1631-1732, fem/src/CRSMatrix.F90:188-246
// The list matrix of the reference is replaced by flat CRS arrays and array cursors.
#include "common.cuh"
#include "../../include/elmer_b200.h"
#include <algorithm>

using namespace b200;

namespace {

template <class F> int guarded_o(F f) {
  try { f(); return 0; }
  catch (const std::exception &e) { set_last_error(e.what()); fprintf(stderr, "[elmer_b200] %s\n", e.what()); return 1; }
}

// 0-based CRS copy of the node graph with bounds checked once
struct Graph {
  int n = 0;
  std::vector<long long> ptr;
  std::vector<int> adj;
  int degree(int u) const { return (int)(ptr[u + 1] - ptr[u]); }
};

Graph load_graph(int n, const int *rows, const int *cols, int base) {
  B200_REQUIRE(n >= 0 && rows && (cols || n == 0), "node graph: null arrays");
  Graph G; G.n = n; G.ptr.resize((size_t)n + 1);
  for (int i = 0; i <= n; ++i) G.ptr[i] = (long long)rows[i] - base;
  B200_REQUIRE(n == 0 || G.ptr[0] == 0, "node graph: rows[0] must equal index_base");
  const long long nnz = n ? G.ptr[n] : 0;
  G.adj.resize((size_t)nnz);
  for (int i = 0; i < n; ++i) B200_REQUIRE(G.ptr[i + 1] >= G.ptr[i], "node graph: rows not ascending");
  for (long long p = 0; p < nnz; ++p) {
    const int c = cols[p] - base;
    B200_REQUIRE(c >= 0 && c < n, "node graph: column outside 1..k (only nodal graphs are handled)");
    G.adj[(size_t)p] = c;
  }
  return G;
}

// Depth-first labelling from `root` in ascending-neighbour order: level = depth in the search tree
// (Levelize, BandwidthOptimize.F90:375-434).  Returns the deepest level reached.  A root without entries is left
// untouched, as the reference's outer loop never runs for it.
int depth_levels(const Graph &G, int root, std::vector<int> &level, std::vector<char> &seen,
                 std::vector<int> &path, std::vector<long long> &cursor) {
  int deepest = 0;
  if (G.degree(root) == 0) return deepest;
  path.clear(); cursor.clear();
  level[root] = 0; seen[root] = 1;
  path.push_back(root); cursor.push_back(G.ptr[root]);
  while (!path.empty()) {
    const int u = path.back();
    long long c = cursor.back();
    const long long end = G.ptr[u + 1];
    while (c < end && seen[G.adj[(size_t)c]]) ++c;
    if (c == end) { path.pop_back(); cursor.pop_back(); continue; }
    cursor.back() = c + 1;
    const int v = G.adj[(size_t)c];
    const int d = (int)path.size();
    level[v] = d; seen[v] = 1; deepest = std::max(deepest, d);
    path.push_back(v); cursor.push_back(G.ptr[v]);
  }
  return deepest;
}

// first-lowest-degree node among those for which pred holds, starting from candidate `best`
template <class P> int lowest_degree(const Graph &G, int best, P pred) {
  int dmin = G.degree(best);
  for (int i = 0; i < G.n; ++i) if (pred(i) && G.degree(i) < dmin) { best = i; dmin = G.degree(i); }
  return best;
}

// The reference's start-node search (BandwidthOptimize.F90:222-271), 0-based.  When a relabelling from the
// candidate gives a shallower tree, the reference continues from the node whose NUMBER equals the previous
// depth (`StartNode = j`, :268); that is kept, since the ordering has to be the same one.  (A candidate must have a
// strictly lower degree than the start, which already has the lowest degree of all, so in practice the first
// labelling is the only one and the sweep starts at the first node of minimum degree.)
int find_start(const Graph &G, std::vector<int> &level) {
  std::vector<char> seen((size_t)G.n, 0);
  std::vector<int> path; std::vector<long long> cursor;
  int start = lowest_degree(G, 0, [](int) { return true; });
  std::fill(level.begin(), level.end(), 0);
  int deepest = depth_levels(G, start, level, seen, path, cursor);
  for (bool again = true; again;) {
    again = false;
    const int cand = lowest_degree(G, start, [&](int i) { return level[i] == deepest; });
    if (cand == start) break;
    const int before = deepest;
    std::fill(seen.begin(), seen.end(), 0);
    deepest = depth_levels(G, cand, level, seen, path, cursor);
    if (before > deepest) { again = true; start = before - 1; }  // 1-based node number `before`
  }
  return start;
}

// Cuthill-McKee sweep from `start` (neighbours appended in ascending number, not by degree), restarted at the
// lowest-numbered unvisited node when a component is exhausted; returns position -> node
// (BandwidthOptimize.F90:289-307, 349-369).
std::vector<int> sweep_order(const Graph &G, int start) {
  const int n = G.n;
  std::vector<int> order; order.reserve((size_t)n);
  std::vector<char> placed((size_t)n, 0);
  int unplaced_from = 0;
  order.push_back(start); placed[start] = 1;
  for (int head = 0; head < n; ++head) {
    if (head == (int)order.size()) {
      while (placed[unplaced_from]) ++unplaced_from;
      order.push_back(unplaced_from); placed[unplaced_from] = 1;
    }
    const int u = order[(size_t)head];
    for (long long p = G.ptr[u]; p < G.ptr[u + 1]; ++p) {
      const int v = G.adj[(size_t)p];
      if (!placed[v]) { placed[v] = 1; order.push_back(v); }
    }
  }
  return order;
}

int half_bandwidth(const Graph &G, const int *number /* 0-based row -> number, or null */) {
  int hb = 0;
#pragma omp parallel for reduction(max : hb) schedule(static)
  for (int i = 0; i < G.n; ++i) {
    const int a = number ? number[i] : i;
    for (long long p = G.ptr[i]; p < G.ptr[i + 1]; ++p) {
      const int c = G.adj[(size_t)p];
      hb = std::max(hb, std::abs(a - (number ? number[c] : c)));
    }
  }
  return hb;
}

// initial row (0-based) -> current number (1-based) through the permutation pair: the composition
// Reorder(InvInitialReorder(i)) with InvInitialReorder(Perm0(m)) = m, later m overwriting earlier ones
// (ElementUtils.F90:1955-1958).
std::vector<int> compose(int k, int perm_size, const int *perm_initial, const int *perm_now) {
  std::vector<int> map((size_t)k, 0);
  for (int m = 0; m < perm_size; ++m) {
    const int i = perm_initial[m];
    if (i > 0) { B200_REQUIRE(i <= k, "permutation entry exceeds the number of graph rows"); map[(size_t)i - 1] = perm_now[m]; }
  }
  return map;
}

}  // namespace

extern "C" {

int b200_node_graph(const int *n_elems, const int *elem_ptr, const int *elem_nodes, const int *index_base,
                    const int *n_nodes, const int *perm, const int *k, long long *nnz, int *rows, int *cols) {
  return guarded_o([&] {
    B200_REQUIRE(n_elems && elem_ptr && index_base && n_nodes && k && nnz, "b200_node_graph: null argument");
    const int ne = *n_elems, base = *index_base, nn = *n_nodes, K = *k;
    B200_REQUIRE(ne >= 0 && nn >= 0 && K >= 0 && (base == 0 || base == 1), "b200_node_graph: bad sizes");
    auto row_of = [&](int node_id) -> int {  // 0-based row or -1
      const int m = node_id - base;
      B200_REQUIRE(m >= 0 && m < nn, "b200_node_graph: element node outside the mesh");
      const int r = perm ? perm[m] : m + 1;
      B200_REQUIRE(r <= K, "b200_node_graph: permutation entry exceeds k");
      return r > 0 ? r - 1 : -1;
    };
    // pass 1: upper bound of each row's length (every element a node belongs to contributes its active nodes)
    std::vector<long long> off((size_t)K + 1, 0);
    for (int t = 0; t < ne; ++t) {
      int act = 0;
      for (int p = elem_ptr[t]; p < elem_ptr[t + 1]; ++p) act += row_of(elem_nodes[p]) >= 0;
      for (int p = elem_ptr[t]; p < elem_ptr[t + 1]; ++p) { const int r = row_of(elem_nodes[p]); if (r >= 0) off[(size_t)r + 1] += act; }
    }
    for (int i = 0; i < K; ++i) off[(size_t)i + 1] += off[(size_t)i];
    std::vector<int> bucket((size_t)off[(size_t)K]);
    std::vector<long long> fill(off.begin(), off.end() - 1);
    for (int t = 0; t < ne; ++t)
      for (int p = elem_ptr[t]; p < elem_ptr[t + 1]; ++p) {
        const int r = row_of(elem_nodes[p]); if (r < 0) continue;
        for (int q = elem_ptr[t]; q < elem_ptr[t + 1]; ++q) { const int c = row_of(elem_nodes[q]); if (c >= 0) bucket[(size_t)fill[(size_t)r]++] = c; }
      }
    // pass 2: ascending, duplicates removed
    std::vector<int> len((size_t)K, 0);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < K; ++i) {
      int *b = bucket.data() + off[(size_t)i], *e = bucket.data() + fill[(size_t)i];
      std::sort(b, e);
      len[(size_t)i] = (int)(std::unique(b, e) - b);
    }
    long long total = 0;
    for (int i = 0; i < K; ++i) total += len[(size_t)i];
    *nnz = total;
    if (!rows) return;
    B200_REQUIRE(total + base <= 2147483647LL, "b200_node_graph: structure exceeds int32 row pointers");
    rows[0] = base;
    for (int i = 0; i < K; ++i) rows[i + 1] = rows[i] + len[(size_t)i];
    if (!cols) return;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < K; ++i) {
      const int *b = bucket.data() + off[(size_t)i];
      int *o = cols + (rows[i] - base);
      for (int j = 0; j < len[(size_t)i]; ++j) o[j] = b[j] + base;
    }
  });
}

int b200_optimize_bandwidth(const int *k, const int *rows, const int *cols, const int *index_base, const int *perm_size,
                            int *perm, const int *optimize, const int *use_optimized, int *half_bandwidth_out) {
  return guarded_o([&] {
    B200_REQUIRE(k && rows && index_base && perm_size && perm && optimize && use_optimized && half_bandwidth_out,
                 "b200_optimize_bandwidth: null argument");
    const int K = *k, M = *perm_size;
    Graph G = load_graph(K, rows, cols, *index_base);
    const int before = half_bandwidth(G, nullptr) + 1;
    *half_bandwidth_out = before;
    if (!*optimize || K == 0) return;

    std::vector<int> level((size_t)K, 0);
    const int start = find_start(G, level);
    const std::vector<int> order = sweep_order(G, start);
    // reversed positions become the new numbers (1-based)
    std::vector<int> renum((size_t)K);
    for (int i = 0; i < K; ++i) renum[(size_t)order[(size_t)i]] = K - i;

    std::vector<int> old(perm, perm + M), now((size_t)M, 0);
    for (int m = 0; m < M; ++m) if (old[(size_t)m] > 0) {
      B200_REQUIRE(old[(size_t)m] <= K, "b200_optimize_bandwidth: permutation entry exceeds k");
      now[(size_t)m] = renum[(size_t)old[(size_t)m] - 1];
    }
    const std::vector<int> number = compose(K, M, old.data(), now.data());
    const int after = half_bandwidth(G, number.data()) + 1;
    if (before < after && !*use_optimized) return;  // rejected: perm untouched, initial bandwidth reported
    std::copy(now.begin(), now.end(), perm);
    *half_bandwidth_out = after;
  });
}

int b200_initialize_structure(const int *k, const int *rows, const int *cols, const int *index_base, const int *dofs,
                              const int *perm_size, const int *perm_initial, const int *perm, int *out_rows,
                              int *out_cols, int *out_diag) {
  return guarded_o([&] {
    B200_REQUIRE(k && rows && index_base && dofs && out_rows, "b200_initialize_structure: null argument");
    const int K = *k, D = *dofs, base = *index_base;
    B200_REQUIRE(D >= 1, "b200_initialize_structure: dofs must be >= 1");
    B200_REQUIRE((perm_initial == nullptr) == (perm == nullptr), "b200_initialize_structure: need both old and new numbering");
    Graph G = load_graph(K, rows, cols, base);
    std::vector<int> number;  // initial row -> 0-based new node number
    if (perm) {
      B200_REQUIRE(perm_size, "b200_initialize_structure: perm_size missing");
      number = compose(K, *perm_size, perm_initial, perm);
      for (int &v : number) { B200_REQUIRE(v >= 1 && v <= K, "b200_initialize_structure: numbering does not cover every row"); v -= 1; }
    } else {
      number.resize((size_t)K);
      for (int i = 0; i < K; ++i) number[(size_t)i] = i;
    }
    const long long total = (K ? G.ptr[(size_t)K] : 0) * (long long)D * D;
    B200_REQUIRE(total + base <= 2147483647LL && (long long)K * D <= 2147483646LL, "b200_initialize_structure: structure exceeds int32");
    // row pointers: dof row D*number[i]+l has D*degree(i) entries
    std::vector<int> deg_new((size_t)K, 0);
    for (int i = 0; i < K; ++i) deg_new[(size_t)number[(size_t)i]] = G.degree(i);
    out_rows[0] = base;
    for (int j = 0; j < K; ++j)
      for (int l = 0; l < D; ++l) out_rows[(size_t)j * D + l + 1] = out_rows[(size_t)j * D + l] + D * deg_new[(size_t)j];
    if (!out_cols) return;
#pragma omp parallel
    {
      std::vector<int> nb;
#pragma omp for schedule(dynamic, 1024)
      for (int i = 0; i < K; ++i) {
        nb.clear();
        for (long long p = G.ptr[i]; p < G.ptr[i + 1]; ++p) nb.push_back(number[(size_t)G.adj[(size_t)p]]);
        std::sort(nb.begin(), nb.end());
        const int j = number[(size_t)i];
        for (int l = 0; l < D; ++l) {
          const int r = j * D + l;
          int *o = out_cols + (out_rows[r] - base);
          for (size_t a = 0; a < nb.size(); ++a) {
            for (int m = 0; m < D; ++m) o[a * D + m] = nb[a] * D + m + base;
            if (out_diag && nb[a] == j) out_diag[r] = out_rows[r] + (int)a * D + l;
          }
        }
      }
    }
  });
}

}  // extern "C"
