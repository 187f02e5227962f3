"""Readers for the on-disk formats either side of the solve path (host tools).

* ElmerGrid meshes: ``mesh.header / mesh.nodes / mesh.elements / mesh.boundary`` and
  ``partitioning.N/part.i.{header,nodes,elements,boundary,shared}`` as read by ElmerAsciiMesh
  (fem/src/MeshUtils.F90:1597-2274; shared-node reader 2178-2230).
* ``Linear System Save`` dumps: ``linsys_a.dat`` (``row col value``, 1-based), ``linsys_b.dat``
  (fem/src/SolverUtils.F90:20313-20446, writer PrintMatrix 14268-14370).
* Ownership and continuous numbering of a partitioned mesh exactly as the MPI path derives them:
  owner of a shared node = first entry of its neighbour list (MeshUtils.F90:2218,
  SParIterSolver.F90:232); dof = DOFs*(node-1)+j (ParallelUtils.F90:180-181); rank r's owned dofs are
  numbered gOffset(r)+1.. in local order (ContinuousNumbering, SParIterSolver.F90:1453-1488).
"""
import os

import numpy as np


class Mesh:
    def __init__(self, nodes_id, xyz, elem_id, elem_body, elem_type, elems, bnd=None):
        self.node_ids = nodes_id          # global node numbers (1-based), file order
        self.xyz = xyz
        self.elem_ids = elem_id
        self.elem_body = elem_body
        self.elem_type = elem_type
        self.elems = elems                # list of int arrays (global node numbers)
        self.bnd = bnd or []              # (id, bc, parent1, parent2, type, nodes)


def _read_nodes(path):
    d = np.loadtxt(path, ndmin=2)
    return d[:, 0].astype(np.int64), d[:, 2:5].astype(np.float64)


def _read_elements(path):
    ids, body, typ, conn = [], [], [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            ids.append(int(t[0])); body.append(int(t[1])); typ.append(int(t[2]))
            conn.append(np.array(t[3:], dtype=np.int64))
    return np.array(ids), np.array(body), np.array(typ), conn


def _read_boundary(path):
    out = []
    if not os.path.exists(path):
        return out
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            out.append((int(t[0]), int(t[1]), int(t[2]), int(t[3]), int(t[4]), np.array(t[5:], dtype=np.int64)))
    return out


def read_mesh(dirname, prefix="mesh"):
    """Serial mesh (prefix 'mesh') or one partition (prefix 'part.i' inside partitioning.N)."""
    nid, xyz = _read_nodes(os.path.join(dirname, prefix + ".nodes"))
    eid, body, typ, conn = _read_elements(os.path.join(dirname, prefix + ".elements"))
    bnd = _read_boundary(os.path.join(dirname, prefix + ".boundary"))
    return Mesh(nid, xyz, eid, body, typ, conn, bnd)


def read_header(dirname, prefix="mesh"):
    with open(os.path.join(dirname, prefix + ".header")) as f:
        t = f.read().split()
    nn, ne, nb, ntypes = int(t[0]), int(t[1]), int(t[2]), int(t[3])
    types = {int(t[4 + 2 * k]): int(t[5 + 2 * k]) for k in range(ntypes)}
    rest = t[4 + 2 * ntypes:]
    nshared = int(rest[0]) if rest else 0
    return dict(nodes=nn, elements=ne, boundary=nb, types=types, shared=nshared)


def read_shared(path):
    """part.i.shared: `globalnode npart p1 p2 ...` (1-based partitions, p1 = owner).  Returns
    {global node: [0-based partitions]} as ParallelInfo%NeighbourList (MeshUtils.F90:2203-2218)."""
    out = {}
    if not os.path.exists(path):
        return out
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            out[int(t[0])] = [int(v) - 1 for v in t[2:2 + int(t[1])]]
    return out


class Partitioning:
    """All partitions of partitioning.N plus the dof ownership and continuous numbering."""

    def __init__(self, dirname, nparts, ndof=1):
        self.nparts, self.ndof = nparts, ndof
        self.parts = [read_mesh(dirname, "part.%d" % (p + 1)) for p in range(nparts)]
        self.shared = [read_shared(os.path.join(dirname, "part.%d.shared" % (p + 1))) for p in range(nparts)]
        # owner of each local node: first neighbour-list entry, else this partition
        self.owner = []
        for p in range(nparts):
            own = np.full(self.parts[p].node_ids.size, p, dtype=np.int32)
            for k, g in enumerate(self.parts[p].node_ids):
                nl = self.shared[p].get(int(g))
                if nl:
                    own[k] = nl[0]
            self.owner.append(own)
        # ContinuousNumbering: owned dofs of rank r get gOffset(r)+0.. in local (file) order
        counts = [int((self.owner[p] == p).sum()) * ndof for p in range(nparts)]
        self.goffset = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.gn = int(self.goffset[-1])
        nn_global = max(int(m.node_ids.max()) for m in self.parts)
        self.cont_of_node = np.full(nn_global + 1, -1, dtype=np.int64)      # global node -> continuous node index
        for p in range(nparts):
            mine = self.parts[p].node_ids[self.owner[p] == p]
            base = self.goffset[p] // ndof
            self.cont_of_node[mine] = base + np.arange(mine.size)
        assert (self.cont_of_node[1:] >= 0).all(), "node without an owner"

    def dof_permutation(self):
        """perm[natural dof (0-based, ndof*(node-1)+j)] = continuous global dof (0-based)."""
        nn = self.cont_of_node.size - 1
        nat = np.arange(nn * self.ndof)
        return self.cont_of_node[1 + nat // self.ndof] * self.ndof + nat % self.ndof

    def owned_rows(self, S):
        """Complete owned rows in continuous numbering of a globally assembled scipy matrix in natural
        dof numbering: what every rank holds after the owners have summed the interface contributions
        (SolverUtils.F90:15461-15579).  Returns per rank (rows, cols, vals) 0-based."""
        import scipy.sparse as sp
        perm = self.dof_permutation()
        C = S.tocoo()
        Sc = sp.csr_matrix((C.data, (perm[C.row], perm[C.col])), shape=S.shape)
        Sc.sort_indices()
        out = []
        for p in range(self.nparts):
            B = Sc[self.goffset[p]:self.goffset[p + 1]]
            out.append((B.indptr.astype(np.int32), B.indices.astype(np.int32), B.data.copy()))
        return out, Sc


def read_linsys(prefix="linsys", dirname="."):
    """`Linear System Save = True` dump -> (scipy CSR, b).  1-based `row col value` triplets."""
    import scipy.sparse as sp
    a = np.loadtxt(os.path.join(dirname, prefix + "_a.dat"), ndmin=2)
    b = np.loadtxt(os.path.join(dirname, prefix + "_b.dat"), ndmin=2)
    bv = b[:, -1] if b.shape[1] > 1 else b[:, 0]
    n = bv.size
    S = sp.csr_matrix((a[:, 2], (a[:, 0].astype(np.int64) - 1, a[:, 1].astype(np.int64) - 1)), shape=(n, n))
    S.sort_indices()
    return S, bv


def write_linsys(S, b, prefix="linsys", dirname="."):
    """Writer in the same format (used to produce fixtures and to hand systems back to Elmer users)."""
    S = S.tocsr(); S.sort_indices()
    with open(os.path.join(dirname, prefix + "_a.dat"), "w") as f:
        for i in range(S.shape[0]):
            for p in range(S.indptr[i], S.indptr[i + 1]):
                f.write("%d %d %.17e\n" % (i + 1, S.indices[p] + 1, S.data[p]))
    with open(os.path.join(dirname, prefix + "_b.dat"), "w") as f:
        for i, v in enumerate(b):
            f.write("%d %.17e\n" % (i + 1, v))
