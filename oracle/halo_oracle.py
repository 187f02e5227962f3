"""TEST INFRASTRUCTURE -- CPU restatement of the reference's only x-halo construction.

Follows elmer_distribute_matrix, fem/src/rocalution.cpp:64-372, for ALL ranks in one process (the MPI
exchanges at 166-248 become list copies), plus the continuous numbering / ownership rules that feed
it (fem/src/SParIterSolver.F90:1453-1488: rank r's owned dofs are numbered gOffset(r)+1.. in local
order; owner = first entry of the neighbour list, SParIterSolver.F90:232).  Plain Python loops: for
small cases only.
"""
import numpy as np


def distribute(parts, index_offset):
    """parts[r] = (rows, cols) of rank r: complete owned rows, 0-based, GLOBAL continuous column ids.
    index_offset[0..np].  Returns per rank the ParallelManager arrays the reference builds:
    boundary_index (local ids, rocalution.cpp:266-274), neighbours, send_offset / recv_offset (205-219),
    received ids in receive order = ghost slots (277-297), interior CSR (local cols) and ghost COO (302-341)."""
    nproc = len(parts)
    boundary = []
    for rank, (rows, cols) in enumerate(parts):
        n = len(rows) - 1
        b = [[] for _ in range(nproc)]
        checked = [set() for _ in range(nproc)]
        lo, hi = index_offset[rank], index_offset[rank + 1]
        for i in range(n):
            for j in range(rows[i], rows[i + 1]):
                c = cols[j]
                if lo <= c < hi:
                    continue
                for r in range(nproc - 1, -1, -1):              # 134-154
                    if r == rank:
                        continue
                    if index_offset[r] <= c < index_offset[r + 1]:
                        if (i + lo) not in checked[r]:
                            b[r].append(i + lo)
                            checked[r].add(i + lo)
                        break
        boundary.append(b)
    out = []
    for rank, (rows, cols) in enumerate(parts):
        n = len(rows) - 1
        lo, hi = index_offset[rank], index_offset[rank + 1]
        neigh = [r for r in range(nproc) if len(boundary[rank][r]) > 0]          # neighbor[r] (147)
        send_offset = [0]
        for r in neigh:
            send_offset.append(send_offset[-1] + len(boundary[rank][r]))
        recv_lists = [boundary[r][rank] for r in neigh]                          # 222-248
        recv_offset = [0]
        for l in recv_lists:
            recv_offset.append(recv_offset[-1] + len(l))
        bnd = [g - lo for r in range(nproc) for g in boundary[rank][r]]          # 266-274
        boundary_index = [g for l in recv_lists for g in l]                      # 277-286
        bmap = {g: i for i, g in enumerate(boundary_index)}                      # 289-297
        irow, icol, grow, gcol = [0], [], [], []
        for i in range(n):
            for j in range(rows[i], rows[i + 1]):
                c = cols[j]
                if c < lo or c >= hi:
                    grow.append(i); gcol.append(bmap[c])                         # 325-330
                else:
                    icol.append(c - lo)                                          # 332-338
            irow.append(len(icol))
        out.append(dict(neigh=np.array(neigh, dtype=np.int32), send_ptr=np.array(send_offset, dtype=np.int32),
                        send_idx=np.array(bnd, dtype=np.int32), recv_ptr=np.array(recv_offset, dtype=np.int32),
                        ghost_gid=np.array(boundary_index, dtype=np.int32), int_rows=np.array(irow, dtype=np.int32),
                        int_cols=np.array(icol, dtype=np.int32), ghost_row=np.array(grow, dtype=np.int32),
                        ghost_col=np.array(gcol, dtype=np.int32)))
    return out


def split_rows(S, offsets):
    """Complete owned rows of a global scipy CSR matrix for contiguous ownership ranges (what the
    ROCSolver bridge assembles at fem/src/SolverUtils.F90:15461-15579)."""
    S = S.tocsr(); S.sort_indices()
    parts = []
    for r in range(len(offsets) - 1):
        B = S[offsets[r]:offsets[r + 1]]
        parts.append((B.indptr.astype(np.int32), B.indices.astype(np.int32), B.data.copy()))
    return parts
