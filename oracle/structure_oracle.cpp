// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU restatement of the reference's matrix-structure producer for nodal elements:
//   MakeListMatrix (plain nodal branch)   fem/src/ElementUtils.F90:881-891 + List_GetMatrixIndex, ListMatrix.F90:334-386
//   ComputeBandwidth / OptimizeBandwidth  fem/src/BandwidthOptimize.F90:71-113, 182-445
//   InitializeMatrix + CRS_SortMatrix     fem/src/ElementUtils.F90:1631-1732, fem/src/CRSMatrix.F90:188-246
// It keeps the reference's data structure (rows of singly linked, ascending entries with Degree and Level) and its
// loop order, including the explicit pointer stack of Levelize and `StartNode = j` at BandwidthOptimize.F90:268.
//
// PARITY UNPINNED: the reference ships no golden permutation or bandwidth for any test case and its Fortran cannot be
// compiled in this image.  What tests/ can check is that the product's independently written ordering code
// (elmerfem_b200/csrc/ordering.cu) returns the same integers as this restatement.
//
// All index arrays are the reference's 1-based Fortran arrays.
#include <cstdlib>
#include <vector>
#include <algorithm>

namespace {

struct Entry { int Index; Entry *Next; };            // ListMatrixEntry_t, Types.F90
struct ListRow { int Degree, Level; Entry *Head; };  // ListMatrix_t

struct ListMatrix {
  std::vector<ListRow> L;  // L[1..n]
  std::vector<Entry *> pool;
  explicit ListMatrix(int n) : L((size_t)n + 1, ListRow{0, 0, nullptr}) {}
  ~ListMatrix() { for (Entry *e : pool) delete e; }
  Entry *new_entry(int ind, Entry *next) { Entry *e = new Entry{ind, next}; pool.push_back(e); return e; }
  // ListMatrix.F90:334-386
  void GetMatrixIndex(int k1, int k2) {
    Entry *CList = L[k1].Head;
    if (!CList) { L[k1].Degree = 1; L[k1].Head = new_entry(k2, nullptr); return; }
    Entry *Prev = nullptr;
    while (CList) { if (CList->Index >= k2) break; Prev = CList; CList = CList->Next; }
    if (CList && CList->Index == k2) return;
    Entry *E = new_entry(k2, CList);
    if (Prev) Prev->Next = E; else L[k1].Head = E;
    L[k1].Degree += 1;
  }
};

void from_crs(ListMatrix &M, int n, const int *rows, const int *cols) {
  for (int i = 1; i <= n; ++i)
    for (int p = rows[i - 1]; p < rows[i]; ++p) M.GetMatrixIndex(i, cols[p - 1]);
}

// BandwidthOptimize.F90:71-113
int ComputeBandwidth(int N, ListMatrix &M, const int *Reorder, const int *InvInitialReorder) {
  int HalfBandWidth = 0;
  for (int i = 1; i <= N; ++i) {
    int j = i;
    if (InvInitialReorder) j = InvInitialReorder[j - 1];
    for (Entry *C = M.L[i].Head; C; C = C->Next) {
      int k = C->Index;
      if (InvInitialReorder) k = InvInitialReorder[k - 1];
      if (Reorder) HalfBandWidth = std::max(HalfBandWidth, std::abs(Reorder[j - 1] - Reorder[k - 1]));
      else HalfBandWidth = std::max(HalfBandWidth, std::abs(j - k));
    }
  }
  return HalfBandWidth;
}

struct Optimizer {
  ListMatrix &M;
  int LocalNodes, MaxLevel = 0, Indx = 0;
  std::vector<char> DoneAlready;
  std::vector<int> PermLocal, DoneIndex;
  Optimizer(ListMatrix &m, int n) : M(m), LocalNodes(n) {}

  // BandwidthOptimize.F90:375-434
  void Levelize(int nin, int Levelin) {
    int n = nin, Level = Levelin;
    std::vector<Entry *> stack;
    Entry *p = M.L[n].Head;
    while (p) {
      stack.push_back(p);
      M.L[n].Level = Level;
      DoneAlready[n] = 1;
      MaxLevel = std::max(MaxLevel, Level);
      p = M.L[n].Head;
      while (true) {
        if (p) {
          n = p->Index;
          if (n <= LocalNodes) {
            if (!DoneAlready[n]) { Level = Level + 1; break; }
          }
        } else if (!stack.empty()) {
          p = stack.back(); stack.pop_back();
          Level = Level - 1;
        } else {
          break;
        }
        p = p->Next;
      }
    }
  }

  // BandwidthOptimize.F90:349-369
  void Renumber(Entry *Current) {
    for (Entry *p = Current; p; p = p->Next) {
      int k = p->Index;
      if (k <= LocalNodes) {
        if (DoneIndex[k] == 0) { PermLocal[Indx] = k; DoneIndex[k] = Indx; Indx = Indx + 1; }
      }
    }
  }
};

int g_new_roots = 0;  // how often the last orc_optimize_bandwidth call took the branch at :266-269 (test coverage only)

}  // namespace

extern "C" {

int orc_optimize_bandwidth_new_roots() { return g_new_roots; }

// ElementUtils.F90:881-891 over elements t = 1..nelem (bulk then boundary, as stored): every pair of the element's
// nodes through Reorder, non-positive skipped.  eptr[nelem+1] 0-based offsets into enodes (1-based node numbers).
// Returns nnz of the k-row list matrix; rows[k+1] / cols (1-based) filled when non-NULL.
long orc_make_list_matrix(int nelem, const int *eptr, const int *enodes, const int *Reorder, int k, int *rows, int *cols) {
  ListMatrix M(k);
  for (int t = 0; t < nelem; ++t) {
    const int *Indexes = enodes + eptr[t];
    const int n = eptr[t + 1] - eptr[t];
    for (int i = 0; i < n; ++i) {
      int k1 = Reorder[Indexes[i] - 1];
      if (k1 <= 0) continue;
      for (int j = 0; j < n; ++j) {
        int k2 = Reorder[Indexes[j] - 1];
        if (k2 <= 0) continue;
        M.GetMatrixIndex(k1, k2);
      }
    }
  }
  long nnz = 0;
  for (int i = 1; i <= k; ++i) nnz += M.L[i].Degree;
  if (rows) {
    rows[0] = 1;
    for (int i = 1; i <= k; ++i) rows[i] = rows[i - 1] + M.L[i].Degree;
    if (cols) for (int i = 1; i <= k; ++i) { int q = rows[i - 1] - 1; for (Entry *e = M.L[i].Head; e; e = e->Next) cols[q++] = e->Index; }
  }
  return nnz;
}

// BandwidthOptimize.F90:182-343.  rows/cols: the list matrix (LocalNodes rows, 1-based).  Perm[permsize] in/out,
// InvInitialReorder[LocalNodes] as CreateMatrix builds it (ElementUtils.F90:1955-1958).  Returns HalfBandWidth.
int orc_optimize_bandwidth(int LocalNodes, const int *rows, const int *cols, int permsize, int *Perm,
                           const int *InvInitialReorder, int Optimize, int UseOptimized) {
  ListMatrix M(LocalNodes);
  from_crs(M, LocalNodes, rows, cols);
  int HalfBandWidth = ComputeBandwidth(LocalNodes, M, nullptr, nullptr) + 1;
  if (!Optimize) return HalfBandWidth;
  const int HalfBandWidthBefore = HalfBandWidth;

  Optimizer O(M, LocalNodes);
  int StartNode = 1;
  int MinDegree = M.L[StartNode].Degree;
  for (int i = 1; i <= LocalNodes; ++i) {
    if (M.L[i].Degree < MinDegree) { StartNode = i; MinDegree = M.L[i].Degree; }
    M.L[i].Level = 0;
  }
  O.DoneAlready.assign((size_t)LocalNodes + 1, 0);
  O.MaxLevel = 0;
  O.Levelize(StartNode, 0);

  bool NewRoot = true;
  g_new_roots = 0;
  while (NewRoot) {
    NewRoot = false;
    MinDegree = M.L[StartNode].Degree;
    int k = StartNode;
    for (int i = 1; i <= LocalNodes; ++i) {
      if (M.L[i].Level == O.MaxLevel) {
        if (M.L[i].Degree < MinDegree) { k = i; MinDegree = M.L[i].Degree; }
      }
    }
    if (k != StartNode) {
      int j = O.MaxLevel;
      O.MaxLevel = 0;
      std::fill(O.DoneAlready.begin(), O.DoneAlready.end(), 0);
      O.Levelize(k, 0);
      if (j > O.MaxLevel) { NewRoot = true; StartNode = j; ++g_new_roots; }  // :266-269, as written
    }
  }

  O.PermLocal.assign((size_t)std::max(permsize, LocalNodes) + 2, 0);
  O.DoneIndex.assign((size_t)LocalNodes + 1, 0);
  O.Indx = 1;
  O.PermLocal[O.Indx] = StartNode;
  O.DoneIndex[StartNode] = O.Indx;
  O.Indx = O.Indx + 1;
  for (int i = 1; i <= LocalNodes; ++i) {
    if (O.PermLocal[i] == 0) {
      for (int j = 1; j <= LocalNodes; ++j) {
        if (O.DoneIndex[j] == 0) { O.PermLocal[O.Indx] = j; O.DoneIndex[j] = O.Indx; O.Indx = O.Indx + 1; break; }
      }
    }
    O.Renumber(M.L[O.PermLocal[i]].Head);
  }

  std::fill(O.DoneIndex.begin(), O.DoneIndex.end(), 0);
  for (int i = 1; i <= LocalNodes; ++i) O.DoneIndex[O.PermLocal[i]] = LocalNodes - i + 1;

  std::vector<int> Saved(Perm, Perm + permsize);  // PermLocal = Perm
  for (int i = 0; i < permsize; ++i) { int k = Saved[i]; Perm[i] = k > 0 ? O.DoneIndex[k] : 0; }

  int HalfBandWidthAfter = ComputeBandwidth(LocalNodes, M, Perm, InvInitialReorder) + 1;
  HalfBandWidth = HalfBandWidthAfter;
  if (HalfBandWidthBefore < HalfBandWidth && !UseOptimized) {
    HalfBandWidth = HalfBandWidthBefore;
    std::copy(Saved.begin(), Saved.end(), Perm);
  }
  return HalfBandWidth;
}

// ElementUtils.F90:1631-1732 followed by CRS_SortMatrix (CRSMatrix.F90:188-246, which also fills Diag).
// Reorder/InvInitialReorder both NULL or both given.  Rows[DOFs*n+1], Cols[DOFs^2*nnz], Diag[DOFs*n], 1-based.
void orc_initialize_matrix(int n, const int *lrows, const int *lcols, int DOFs, const int *Reorder,
                           const int *InvInitialReorder, int *Rows, int *Cols, int *Diag) {
  ListMatrix M(n);
  from_crs(M, n, lrows, lcols);
  Rows[0] = 1;
  for (int i = 1; i <= n; ++i)
    for (int l = 1; l <= DOFs; ++l) {
      int j = Reorder ? Reorder[InvInitialReorder[i - 1] - 1] : i;
      int k1 = DOFs * (j - 1) + l;
      Rows[k1] = DOFs * M.L[i].Degree;  // Rows(k1+1)
    }
  for (int i = 1; i <= DOFs * n; ++i) Rows[i] = Rows[i - 1] + Rows[i];
  for (int i = 1; i <= n; ++i)
    for (int l = 1; l <= DOFs; ++l) {
      int j = Reorder ? Reorder[InvInitialReorder[i - 1] - 1] : i;
      int k1 = DOFs * (j - 1) + l;
      int k2 = Rows[k1 - 1] - 1;
      for (Entry *C = M.L[i].Head; C; C = C->Next) {
        int k = Reorder ? Reorder[InvInitialReorder[C->Index - 1] - 1] : C->Index;
        k = DOFs * (k - 1);
        for (int m = k + 1; m <= k + DOFs; ++m) { k2 = k2 + 1; Cols[k2 - 1] = m; }
      }
    }
  const int N = DOFs * n;
  for (int i = 1; i <= N; ++i) std::sort(Cols + Rows[i - 1] - 1, Cols + Rows[i] - 1);
  for (int i = 1; i <= N; ++i)
    for (int j = Rows[i - 1]; j <= Rows[i] - 1; ++j)
      if (Cols[j - 1] == i) { Diag[i - 1] = j; break; }
}

}  // extern "C"
