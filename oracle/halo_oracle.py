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


def send_lists_rank(rows, cols, index_offset, rank):
    """The same construction for ONE rank, vectorised (numpy) so that it also runs at benchmark sizes: rank `rank`'s complete owned rows
    (0-based CRS, GLOBAL continuous column ids) -> (neigh, send_ptr, send_idx).  rocalution.cpp:115-154 visits the rows in ascending
    order and appends row i once per foreign owner of one of its columns, so the list towards rank q is the ascending set of rows with a
    column owned by q; 266-274 concatenates the lists in ascending q."""
    rows = np.asarray(rows, dtype=np.int64); cols = np.asarray(cols, dtype=np.int64)
    off = np.asarray(index_offset, dtype=np.int64)
    n = rows.size - 1
    lo, hi = off[rank], off[rank + 1]
    rowid = np.repeat(np.arange(n, dtype=np.int64), np.diff(rows))
    ext = (cols < lo) | (cols >= hi)
    q = np.searchsorted(off, cols[ext], side="right") - 1
    key = np.unique(q * n + rowid[ext])                                       # ascending (q, row)
    qs, idx = key // max(n, 1), key % max(n, 1)
    neigh = np.unique(qs)
    counts = np.array([np.count_nonzero(qs == r) for r in neigh], dtype=np.int64)
    send_ptr = np.concatenate([[0], np.cumsum(counts)])
    return neigh.astype(np.int32), send_ptr.astype(np.int32), idx.astype(np.int32)


def plan_rank(all_send, index_offset, rank):
    """all_send[r] = send_lists_rank(...) of every rank -> the full plan of `rank` in distribute()'s form (222-297): what rank receives
    from neighbour q, in q's send order, are q's boundary rows towards `rank` as global ids; ghost slots follow the receive order."""
    neigh, send_ptr, send_idx = all_send[rank]
    recv, recv_ptr = [], [0]
    for q in neigh:
        nq, pq, iq = all_send[q]
        k = int(np.flatnonzero(nq == rank)[0])
        recv.append(iq[pq[k]:pq[k + 1]].astype(np.int64) + index_offset[q])
        recv_ptr.append(recv_ptr[-1] + recv[-1].size)
    ghost = np.concatenate(recv).astype(np.int32) if recv else np.zeros(0, dtype=np.int32)
    return dict(neigh=neigh, send_ptr=send_ptr, send_idx=send_idx, recv_ptr=np.array(recv_ptr, dtype=np.int32), ghost_gid=ghost)
