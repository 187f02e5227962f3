"""elmerfem_b200 -- B200-native (sm_100a) sparse iterative linear-solve path for Elmer.

The product is the C-ABI shared library ``libelmer_b200.so`` (sources in ``csrc/``, interface in
``include/elmer_b200.h``).  This package is only the ctypes binding the tests and bench.py use to call
that ABI the way Elmer's Fortran would (all scalars by reference, raw 1-based CRS arrays).

There is no CPU path: importing works without a GPU (so that symbol/ABI checks can run), but every
compute entry point fails loudly when no CUDA device is present or the library is missing.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libelmer_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "elmer_b200.h")

METHODS = {"cg": 1, "bicgstab": 2, "bicgstabl": 3, "gcr": 4, "idrs": 5, "gmres": 6, "cgs": 7, "tfqmr": 8, "bicgstab2": 9, "jacobi": 10, "richardson": 11, "sgs": 12}
PRECONDS = {"none": 0, "diagonal": 1, "ilu0": 2, "ilu": 2}
DECLINED = 100

_lib = None


class B200Error(RuntimeError):
    pass


def build(verbose=False):
    """Compile every CUDA source for sm_100a into libelmer_b200.so (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], stdout=out)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error("libelmer_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                        "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vpp = C.POINTER(C.c_void_p)
    ip = C.POINTER(C.c_int)
    dp = C.POINTER(C.c_double)
    L.b200_last_error.restype = C.c_char_p
    sigs = {
        "b200_create": [vpp], "b200_destroy": [vpp], "b200_device_count": [ip], "b200_set_device": [vpp, ip],
        "b200_set_structure": [vpp, ip, ip, ip, ip, ip, ip, ip],
        "b200_set_values": [vpp, dp, dp], "b200_set_values_device": [vpp, C.c_void_p, C.c_void_p],
        "b200_factorize": [vpp], "b200_scale_system": [vpp], "b200_get_values": [vpp, dp],
        "b200_solve": [vpp, dp, dp, ip, dp, ip, ip, dp],
        "b200_solve_device": [vpp, C.c_void_p, C.c_void_p, ip, dp, ip, ip, C.c_void_p],
        "b200_itersolver": [vpp, dp, dp, C.c_char_p, ip, ip],
        "b200_matvec": [vpp, dp, dp], "b200_diag_precondition": [vpp, dp, dp], "b200_lu_precondition": [vpp, dp, dp],
        "b200_dot": [vpp, ip, dp, dp, dp], "b200_nrm2": [vpp, ip, dp, dp],
        "b200_get_ilu_values": [vpp, dp], "b200_set_ilu_order": [vpp, ip], "b200_set_symmetric_ilu": [vpp, ip], "b200_set_ilut": [vpp, ip, dp], "b200_set_bilu_blocks": [vpp, ip], "b200_get_ilu_structure": [vpp, ip, ip, ip, ip], "b200_get_structure": [vpp, ip, ip, ip], "b200_get_levels": [vpp, ip, ip],
        "b200_comm_unique_id": [C.c_char_p], "b200_comm_init": [vpp, ip, ip, C.c_char_p],
        "b200_set_partition": [vpp, ip, ip, ip, ip, ip, ip, ip, ip],
        "b200_get_halo_plan": [vpp, ip, ip, ip, ip, ip, ip],
        "b200_get_stats": [vpp, dp], "b200_time_matvec": [vpp, ip, dp], "b200_time_lu_precondition": [vpp, ip, dp],
        "b200_partition_send_lists": [ip] * 10, "b200_partition_peer_layout": [ip, ip, ip, ip, C.POINTER(C.c_longlong)], "b200_partition_split": [ip] * 14,
        "b200_node_graph": [ip, ip, ip, ip, ip, ip, ip, C.POINTER(C.c_longlong), ip, ip],
        "b200_optimize_bandwidth": [ip] * 9, "b200_initialize_structure": [ip] * 11,
        "b200_itersolver_plan": [C.c_char_p, ip, ip, ip, ip, ip, ip, ip, dp],
        "b200_version": [ip, ip], "b200_vec_len": [vpp, C.POINTER(C.c_longlong)],
    }
    for name, args in sigs.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int
    L.b200_spmv.argtypes = [vpp, ip, ip, ip, dp, dp, dp, ip]
    L.b200_spmv.restype = None
    _lib = L
    return L


def exported_symbols():
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB_PATH]).decode()
    return sorted(line.split()[-1] for line in out.splitlines() if " T " in line)


def header_symbols():
    """Entry points declared in include/elmer_b200.h."""
    import re
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _i(v):
    return C.byref(C.c_int(int(v)))


def _check(rc, what):
    if rc != 0:
        raise B200Error("%s failed: %s" % (what, lib().b200_last_error().decode()))


def fill_ipar_dpar(n, method, tol=1e-8, maxit=1000, minit=0, maxtol=1e20, residual_output=0,
                   bicgstabl_l=2, gcr_restart=None, idrs_s=4, smoothing=False, stopc=1, gmres_restart=10, sgs_omega=None,
                   robust=False, robust_tol=None, robust_limit=None, robust_margin=None, robust_max_bad=None, robust_start=None):
    """HUTI ipar(50)/dpar(10) exactly as IterSolver fills them (fem/src/IterSolve.F90:245-503;
    slots fhutiter/src/huti_fdefs.h:101-155)."""
    ipar = np.zeros(50, dtype=np.int32)
    dpar = np.zeros(10, dtype=np.float64)
    ipar[2] = n
    ipar[3] = {"cg": 4, "bicgstab": 8, "cgs": 7, "tfqmr": 10, "bicgstab2": 8}.get(method, 1)
    ipar[4] = residual_output
    ipar[9] = maxit
    ipar[10] = minit
    ipar[11] = stopc
    ipar[13] = 1
    if method == "gmres":             # IterSolve.F90:346-350
        ipar[14] = gmres_restart
        ipar[3] = 7 + gmres_restart
    if method == "bicgstabl":
        ipar[15] = max(2, bicgstabl_l)
    if method == "gcr":
        ipar[16] = gcr_restart if gcr_restart is not None else min(maxit, 200)
    if method == "idrs":
        ipar[17] = idrs_s
    ipar[27] = 1 if smoothing else 0
    dpar[0] = tol
    dpar[2] = float(np.float32(1.8)) if sgs_omega is None else sgs_omega     # HUTI_SGSPARAM (IterSolve.F90:354-358)
    dpar[1] = maxtol
    if robust:                        # `Linear System Robust` (IterSolve.F90:482-496), defaults as there
        ipar[25] = 1
        dpar[2] = tol ** float(np.float32(2.0) / np.float32(3.0)) if robust_tol is None else robust_tol   # HUTI_TOLERANCE**(2.0/3.0)
        dpar[4] = np.sqrt(tol) if robust_limit is None else robust_limit
        dpar[3] = 1.1 if robust_margin is None else robust_margin
        ipar[26] = maxit // 2 if robust_max_bad is None else robust_max_bad
        ipar[28] = 1 if robust_start is None else robust_start
    return ipar, dpar


class Matrix:
    """One Elmer Matrix_t on the device: owns the opaque handle slot (Matrix_t%SpMV-style)."""

    def __init__(self):
        self._h = C.c_void_p(None)
        _check(lib().b200_create(C.byref(self._h)), "b200_create")
        self.n = 0
        self.nnz = 0

    def close(self):
        if self._h:
            lib().b200_destroy(C.byref(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return C.byref(self._h)

    def set_structure(self, rows, cols, diag, index_base=1, ndeg=1):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        diag = np.ascontiguousarray(diag, dtype=np.int32)
        self.n, self.nnz = rows.size - 1, cols.size
        _check(lib().b200_set_structure(self.handle, _i(self.n), _i(self.nnz), _ip(rows), _ip(cols), _ip(diag),
                                        _i(index_base), _i(ndeg)), "b200_set_structure")

    def set_values(self, vals, prec_vals=None):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        pv = None if prec_vals is None else np.ascontiguousarray(prec_vals, dtype=np.float64)
        _check(lib().b200_set_values(self.handle, _dp(vals), _dp(pv)), "b200_set_values")

    def set_values_device(self, d_vals_ptr, d_prec_ptr=None):
        _check(lib().b200_set_values_device(self.handle, C.c_void_p(d_vals_ptr), C.c_void_p(d_prec_ptr)), "b200_set_values_device")

    def scale_system(self):
        """Linear System Scaling on the device (ScaleLinearSystemDiagonal); b/x of later solves are scaled and back-scaled inside."""
        _check(lib().b200_scale_system(self.handle), "b200_scale_system")

    def values(self):
        out = np.empty(self.nnz)
        _check(lib().b200_get_values(self.handle, _dp(out)), "b200_get_values")
        return out

    def factorize(self):
        _check(lib().b200_factorize(self.handle), "b200_factorize")

    def solve(self, b, x0=None, method="bicgstab", precond="none", P=None, **kw):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        ipar, dpar = fill_ipar_dpar(self.n, method, **kw)
        Pp = None if P is None else np.asfortranarray(P, dtype=np.float64)
        rc = lib().b200_solve(self.handle, _dp(b), _dp(x), _ip(ipar), _dp(dpar), _i(METHODS[method]), _i(PRECONDS[precond]), _dp(Pp))
        _check(rc, "b200_solve")
        return dict(x=x, info=int(ipar[29]), iters=int(ipar[30]), residual=float(dpar[9]), ipar=ipar, dpar=dpar, stats=self.stats())

    def solve_device(self, d_b_ptr, d_x_ptr, method="bicgstab", precond="none", d_P_ptr=None, **kw):
        ipar, dpar = fill_ipar_dpar(self.n, method, **kw)
        rc = lib().b200_solve_device(self.handle, C.c_void_p(d_b_ptr), C.c_void_p(d_x_ptr), _ip(ipar), _dp(dpar),
                                     _i(METHODS[method]), _i(PRECONDS[precond]), C.c_void_p(d_P_ptr))
        _check(rc, "b200_solve_device")
        return dict(info=int(ipar[29]), iters=int(ipar[30]), residual=float(dpar[9]), stats=self.stats())

    def solve_host(self, b_ptr, x_ptr, method="bicgstab", precond="none", P_ptr=None, **kw):
        """b200_solve on raw HOST addresses (e.g. pinned buffers): x in/out in place."""
        ipar, dpar = fill_ipar_dpar(self.n, method, **kw)
        f = lib().b200_solve
        dpv = C.POINTER(C.c_double)
        rc = f(self.handle, C.cast(C.c_void_p(b_ptr), dpv), C.cast(C.c_void_p(x_ptr), dpv), _ip(ipar), _dp(dpar),
               _i(METHODS[method]), _i(PRECONDS[precond]), C.cast(C.c_void_p(P_ptr), dpv))
        _check(rc, "b200_solve")
        return dict(info=int(ipar[29]), iters=int(ipar[30]), residual=float(dpar[9]), stats=self.stats())

    def vec_len(self):
        v = C.c_longlong(0)
        _check(lib().b200_vec_len(self.handle, C.byref(v)), "b200_vec_len")
        return v.value

    def itersolver(self, b, x0, sif, solve_count=0):
        """IterSolver(A,x,b,Solver) through the keyword front-end; returns None when DECLINED."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        sc = C.c_int(solve_count)
        info = (C.c_int * 2)(0, 0)
        rc = lib().b200_itersolver(self.handle, _dp(b), _dp(x), sif.encode(), C.byref(sc), info)
        if rc == DECLINED:
            return None
        _check(rc, "b200_itersolver")
        return dict(x=x, info=int(info[0]), iters=int(info[1]), solve_count=sc.value, stats=self.stats())

    def matvec(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = np.empty(self.n)
        _check(lib().b200_matvec(self.handle, _dp(u), _dp(v)), "b200_matvec")
        return v

    def diag_precondition(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        u = np.empty(self.n)
        _check(lib().b200_diag_precondition(self.handle, _dp(u), _dp(v)), "b200_diag_precondition")
        return u

    def lu_precondition(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        u = np.empty(self.n)
        _check(lib().b200_lu_precondition(self.handle, _dp(u), _dp(v)), "b200_lu_precondition")
        return u

    def dot(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        r = C.c_double(0)
        _check(lib().b200_dot(self.handle, _i(x.size), _dp(x), _dp(y), C.byref(r)), "b200_dot")
        return r.value

    def nrm2(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = C.c_double(0)
        _check(lib().b200_nrm2(self.handle, _i(x.size), _dp(x), C.byref(r)), "b200_nrm2")
        return r.value

    def set_ilu_order(self, order):
        """Fill level n of CRS_IncompleteLU(A, n) ("Linear System Preconditioning = ILUn"); 0 = ILU0."""
        _check(lib().b200_set_ilu_order(self.handle, _i(order)), "b200_set_ilu_order")

    def set_symmetric_ilu(self, flag):
        """A % Cholesky ("Linear System Symmetric ILU"): incomplete Cholesky instead of incomplete LU."""
        _check(lib().b200_set_symmetric_ilu(self.handle, _i(1 if flag else 0)), "b200_set_symmetric_ilu")

    def set_ilut(self, tol, on=True):
        """ILUT with drop tolerance `tol` ("Linear System Preconditioning = ILUT"); on=False returns to ILU(order)."""
        _check(lib().b200_set_ilut(self.handle, _i(1 if on else 0), C.byref(C.c_double(float(tol)))), "b200_set_ilut")

    def set_bilu_blocks(self, blocks):
        """BILU: factorise the block-diagonal part, MOD(i,blocks) == MOD(j,blocks) (CRS_BlockDiagonal); <= 1 switches it off."""
        _check(lib().b200_set_bilu_blocks(self.handle, _i(blocks)), "b200_set_bilu_blocks")

    def ilu_structure(self):
        """ILURows/ILUCols/ILUDiag (caller's index base) of the current ILU order."""
        sizes = np.zeros(2, dtype=np.int32)
        none = C.POINTER(C.c_int)()
        _check(lib().b200_get_ilu_structure(self.handle, _ip(sizes), none, none, none), "b200_get_ilu_structure")
        rows = np.empty(sizes[0] + 1, dtype=np.int32); cols = np.empty(sizes[1], dtype=np.int32); diag = np.empty(sizes[0], dtype=np.int32)
        _check(lib().b200_get_ilu_structure(self.handle, _ip(sizes), _ip(rows), _ip(cols), _ip(diag)), "b200_get_ilu_structure")
        return rows, cols, diag

    def ilu_values(self):
        sizes = np.zeros(2, dtype=np.int32)
        none = C.POINTER(C.c_int)()
        _check(lib().b200_get_ilu_structure(self.handle, _ip(sizes), none, none, none), "b200_get_ilu_structure")
        out = np.empty(int(sizes[1]))
        _check(lib().b200_get_ilu_values(self.handle, _dp(out)), "b200_get_ilu_values")
        return out

    def structure(self):
        rows = np.empty(self.n + 1, dtype=np.int32)
        cols = np.empty(self.nnz, dtype=np.int32)
        diag = np.empty(self.n, dtype=np.int32)
        _check(lib().b200_get_structure(self.handle, _ip(rows), _ip(cols), _ip(diag)), "b200_get_structure")
        return rows, cols, diag

    def levels(self, per_row=False):
        counts = np.zeros(4, dtype=np.int32)
        lev = np.empty(self.n, dtype=np.int32) if per_row else None
        _check(lib().b200_get_levels(self.handle, _ip(counts), None if lev is None else _ip(lev)), "b200_get_levels")
        return dict(forward=int(counts[0]), backward=int(counts[1]), slices_f=int(counts[2]), slices_b=int(counts[3]), level=lev)

    def stats(self):
        s = np.zeros(16)
        _check(lib().b200_get_stats(self.handle, _dp(s)), "b200_get_stats")
        return dict(solve_ms=s[0], matvec=int(s[1]), pcond=int(s[2]), factor_ms=s[3], launches=int(s[4]), h2d=int(s[5]),
                    d2h=int(s[6]), iters=int(s[7]), spmv_ms=s[8], lu_ms=s[9], residual=s[10], sell_entries=int(s[11]),
                    levels_f=int(s[12]), levels_b=int(s[13]), factor_launches=int(s[14]), tri_mode=int(s[15]))

    def time_matvec(self, reps=20):
        ms = C.c_double(0)
        _check(lib().b200_time_matvec(self.handle, _i(reps), C.byref(ms)), "b200_time_matvec")
        return ms.value

    def time_lu(self, reps=10):
        ms = C.c_double(0)
        _check(lib().b200_time_lu_precondition(self.handle, _i(reps), C.byref(ms)), "b200_time_lu_precondition")
        return ms.value

    # ---- multi-GPU -------------------------------------------------------------------------
    def comm_init(self, nranks, rank, unique_id):
        _check(lib().b200_comm_init(self.handle, _i(nranks), _i(rank), unique_id), "b200_comm_init")

    def set_partition(self, gn, rows, cols_global, goffset, index_base=1, ndeg=1):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols_global, dtype=np.int32)
        goffset = np.ascontiguousarray(goffset, dtype=np.int32)
        self.n = rows.size - 1
        self.nnz = cols.size
        _check(lib().b200_set_partition(self.handle, _i(gn), _i(self.n), _i(cols.size), _ip(rows), _ip(cols), _ip(goffset),
                                        _i(index_base), _i(ndeg)), "b200_set_partition")

    def halo_plan(self):
        sizes = np.zeros(3, dtype=np.int32)
        _check(lib().b200_get_halo_plan(self.handle, _ip(sizes), None, None, None, None, None), "b200_get_halo_plan")
        nn, ns, ng = [int(v) for v in sizes]
        neigh = np.zeros(max(nn, 1), dtype=np.int32)
        sp = np.zeros(nn + 1, dtype=np.int32)
        si = np.zeros(max(ns, 1), dtype=np.int32)
        rp = np.zeros(nn + 1, dtype=np.int32)
        gg = np.zeros(max(ng, 1), dtype=np.int32)
        _check(lib().b200_get_halo_plan(self.handle, _ip(sizes), _ip(neigh), _ip(sp), _ip(si), _ip(rp), _ip(gg)), "b200_get_halo_plan")
        return dict(neigh=neigh[:nn], send_ptr=sp, send_idx=si[:ns], recv_ptr=rp, ghost_gid=gg[:ng])


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().b200_comm_unique_id(buf), "b200_comm_unique_id")
    return buf.raw


def spmv_hook(slot, rows, cols, vals, u, reinit=0):
    """Calls b200_spmv exactly as Elmer's matvecsubrext_c does (fem/src/Load.c:806-824)."""
    n = rows.size - 1
    v = np.empty(n)
    lib().b200_spmv(C.byref(slot), _i(n), _ip(rows), _ip(cols), _dp(vals), _dp(u), _dp(v), _i(reinit))
    return v


def partition_send_lists(gn, rows, cols_global, goffset, rank, index_base=1):
    """Host-only: per-destination send lists (global 0-based row ids) of this rank, canonical form of
    rocalution.cpp:121-156.  Returns (counts[nranks], gids concatenated by ascending rank)."""
    rows = np.ascontiguousarray(rows, dtype=np.int32); cols = np.ascontiguousarray(cols_global, dtype=np.int32)
    goffset = np.ascontiguousarray(goffset, dtype=np.int32)
    nr = goffset.size - 1
    cnt = np.zeros(nr, dtype=np.int32)
    args = (_i(gn), _i(rows.size - 1), _ip(rows), _ip(cols), _ip(goffset), _i(index_base), _i(nr), _i(rank))
    _check(lib().b200_partition_send_lists(*args, _ip(cnt), None), "b200_partition_send_lists")
    gid = np.zeros(max(int(cnt.sum()), 1), dtype=np.int32)
    _check(lib().b200_partition_send_lists(*args, _ip(cnt), _ip(gid)), "b200_partition_send_lists")
    return cnt, gid[:int(cnt.sum())]


def partition_peer_layout(nranks, me, r, cnt):
    """Host-only: (index of `me` among r's neighbours, r's neighbour count, r's ghost count, offset of me's segment in r's
    receive area) as the peer-memory halo path derives them from the nranks x nranks send-count matrix."""
    cnt = np.ascontiguousarray(cnt, dtype=np.int32)
    out = (C.c_longlong * 4)()
    _check(lib().b200_partition_peer_layout(_i(nranks), _i(me), _i(r), _ip(cnt), out), "b200_partition_peer_layout")
    return tuple(int(v) for v in out)


def partition_split(rows, cols_global, lo, hi, ghost_gid, index_base=1):
    """Host-only: owned x owned block and ghost block of the complete owned rows (0-based outputs)."""
    rows = np.ascontiguousarray(rows, dtype=np.int32); cols = np.ascontiguousarray(cols_global, dtype=np.int32)
    gg = np.ascontiguousarray(ghost_gid, dtype=np.int32)
    n = rows.size - 1
    sizes = np.zeros(2, dtype=np.int32)
    ggp = _ip(gg) if gg.size else _ip(np.zeros(1, dtype=np.int32))
    head = (_i(n), _ip(rows), _ip(cols), _i(index_base), _i(lo), _i(hi), _i(gg.size), ggp)
    _check(lib().b200_partition_split(*head, _ip(sizes), None, None, None, None, None), "b200_partition_split")
    oo_r = np.zeros(n + 1, dtype=np.int32); oo_c = np.zeros(max(int(sizes[0]), 1), dtype=np.int32); oo_d = np.zeros(max(n, 1), dtype=np.int32)
    g_r = np.zeros(n + 1, dtype=np.int32); g_c = np.zeros(max(int(sizes[1]), 1), dtype=np.int32)
    _check(lib().b200_partition_split(*head, _ip(sizes), _ip(oo_r), _ip(oo_c), _ip(oo_d), _ip(g_r), _ip(g_c)), "b200_partition_split")
    return dict(oo_rows=oo_r, oo_cols=oo_c[:sizes[0]], oo_diag=oo_d[:n], g_rows=g_r, g_cols=g_c[:sizes[1]])


def node_graph(elem_ptr, elem_nodes, n_nodes, perm=None, k=None, index_base=1):
    """Host-only: the list matrix MakeListMatrix builds for plain nodal elements (ElementUtils.F90:881-891) as CRS
    rows/cols in index_base numbering.  perm = Elmer's Perm (1-based rows, 0 = inactive) or None (identity)."""
    ep = np.ascontiguousarray(elem_ptr, dtype=np.int32); en = np.ascontiguousarray(elem_nodes, dtype=np.int32)
    pm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    if k is None:
        k = n_nodes if pm is None else int(pm.max(initial=0))
    nnz = C.c_longlong(0)
    head = (_i(ep.size - 1), _ip(ep), _ip(en), _i(index_base), _i(n_nodes), None if pm is None else _ip(pm), _i(k), C.byref(nnz))
    _check(lib().b200_node_graph(*head, None, None), "b200_node_graph")
    rows = np.zeros(k + 1, dtype=np.int32); cols = np.zeros(max(nnz.value, 1), dtype=np.int32)
    _check(lib().b200_node_graph(*head, _ip(rows), _ip(cols)), "b200_node_graph")
    return rows, cols[:nnz.value]


def optimize_bandwidth(rows, cols, perm, optimize=True, use_optimized=False, index_base=1):
    """Host-only: OptimizeBandwidth (BandwidthOptimize.F90:182-445).  Returns (new Perm, half bandwidth)."""
    rows = np.ascontiguousarray(rows, dtype=np.int32); cols = np.ascontiguousarray(cols, dtype=np.int32)
    pm = np.array(perm, dtype=np.int32)
    hb = C.c_int(0)
    _check(lib().b200_optimize_bandwidth(_i(rows.size - 1), _ip(rows), _ip(cols), _i(index_base), _i(pm.size), _ip(pm),
                                         _i(1 if optimize else 0), _i(1 if use_optimized else 0), C.byref(hb)), "b200_optimize_bandwidth")
    return pm, hb.value


def initialize_structure(rows, cols, dofs, perm_initial=None, perm=None, index_base=1):
    """Host-only: InitializeMatrix + CRS_SortMatrix (ElementUtils.F90:1631-1732): node graph -> (Rows, Cols, Diag)
    with `dofs` interleaved unknowns per node, renumbered from perm_initial to perm when both are given."""
    rows = np.ascontiguousarray(rows, dtype=np.int32); cols = np.ascontiguousarray(cols, dtype=np.int32)
    k = rows.size - 1
    nnz = int(rows[-1] - rows[0])
    R = np.zeros(k * dofs + 1, dtype=np.int32); Cc = np.zeros(max(nnz * dofs * dofs, 1), dtype=np.int32); D = np.zeros(max(k * dofs, 1), dtype=np.int32)
    if perm is None:
        ps, p0, p1 = None, None, None
    else:
        a = np.ascontiguousarray(perm_initial, dtype=np.int32); b = np.ascontiguousarray(perm, dtype=np.int32)
        ps, p0, p1 = _i(a.size), _ip(a), _ip(b)
    _check(lib().b200_initialize_structure(_i(k), _ip(rows), _ip(cols), _i(index_base), _i(dofs), ps, p0, p1, _ip(R), _ip(Cc), _ip(D)),
           "b200_initialize_structure")
    return R, Cc[:nnz * dofs * dofs], D[:k * dofs]


def create_matrix_structure(elem_ptr, elem_nodes, n_nodes, dofs=1, optimize_bw=True, use_optimized=False, perm=None):
    """The nodal path of CreateMatrix (ElementUtils.F90:1745-2170) from the three host-only steps above: initial Perm
    (identity when the equation covers the mesh, 1918-1925), list matrix, OptimizeBandwidth, InitializeMatrix.
    Returns dict(perm, half_bandwidth, rows, cols, diag) with Elmer's 1-based arrays."""
    perm0 = np.arange(1, n_nodes + 1, dtype=np.int32) if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    k = int(perm0.max(initial=0))
    lrows, lcols = node_graph(elem_ptr, elem_nodes, n_nodes, perm0, k)
    perm1, hb = optimize_bandwidth(lrows, lcols, perm0, optimize_bw, use_optimized)
    if optimize_bw:
        R, Cc, D = initialize_structure(lrows, lcols, dofs, perm0, perm1)
    else:
        R, Cc, D = initialize_structure(lrows, lcols, dofs)
    return dict(perm=perm1, half_bandwidth=hb, rows=R, cols=Cc, diag=D, list_rows=lrows, list_cols=lcols)


def itersolver_plan(sif, n, ndeg=1):
    """Host-only: IterSolver's keyword decisions (IterSolve.F90:250-577) as b200_itersolver takes them.  Returns a dict, or
    None with the reason in last_error() when the keyword combination is DECLINED."""
    method = C.c_int(0); pc = C.c_int(0); order = C.c_int(0); blocks = C.c_int(0)
    ipar = np.zeros(50, dtype=np.int32); dpar = np.zeros(10, dtype=np.float64)
    rc = lib().b200_itersolver_plan(sif.encode(), _i(n), _i(ndeg), C.byref(method), C.byref(pc), C.byref(order), C.byref(blocks),
                                    _ip(ipar), _dp(dpar))
    if rc == DECLINED:
        return None
    _check(rc, "b200_itersolver_plan")
    return dict(method=method.value, precond=pc.value, ilu_order=order.value, bilu_blocks=blocks.value, ipar=ipar, dpar=dpar)


def last_error():
    return lib().b200_last_error().decode()
