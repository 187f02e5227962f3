/*
 * elmer_b200.h -- C ABI of the B200-native sparse iterative linear-solve path for Elmer.
 *
 * Drop-in boundary.  Everything is `extern "C"`, every scalar is passed by
 * reference so the entry points are directly callable from Fortran through ISO_C_BINDING
 * (elmerfem_b200/fortran/B200Solve.F90 holds the INTERFACE blocks), and no CUDA, NCCL or torch
 * type appears in a signature.  The matrix handle is an opaque pointer slot owned by the caller
 * (Elmer stores such slots in Matrix_t%SpMV / %AMGX, fem/src/Types.F90:264-267); the library owns
 * all device memory behind it.  Host arrays stay owned by the caller.
 *
 * Reference interfaces each entry point replaces (paths relative to ElmerCSC/elmerfem):
 *
 *   b200_solve            <- IterSolver's method call: IterCall(iterProc, x, b, ipar, dpar, work, mv, pcl, pcr,
 *                            dot, norm, stopc), fem/src/IterSolve.F90:1004-1005, i.e. the HUTI solver entry
 *                            fhutiter/src/huti_interfaces.F90:203-218 with the five callbacks
 *                            CRS_MatrixVectorProd / CRS_DiagPrecondition / CRS_LUPrecondition / ddot / dnrm2 bound
 *                            inside; ipar(50)/dpar(10) are the HUTI arrays verbatim (fhutiter/src/huti_fdefs.h:101-155).
 *                            Same shape as the in-tree GPU bridges ROCSerialSolve / AMGXSolve
 *                            (fem/src/SolverUtils.F90:15314-15333, 15075-15087).
 *   b200_set_structure    <- the CRS container Matrix_t (fem/src/Types.F90:193-283): Rows, Cols, Diag, ndeg.
 *   b200_set_values       <- A%Values / A%PrecValues as read by CRS_IncompleteLU (fem/src/CRSMatrix.F90:3480-3484)
 *                            and CRS_MatrixVectorProd (4772-4776); once per nonlinear iteration.
 *   b200_factorize        <- CRS_IncompleteLU(A,0), fem/src/CRSMatrix.F90:3445-3661, called from IterSolve.F90:741
 *                            under the recompute policy of IterSolve.F90:579-587 (the caller applies the policy).
 *   b200_matvec           <- CRS_MatrixVectorProd(u,v,ipar), fem/src/CRSMatrix.F90:4744-4905.
 *   b200_diag_precondition<- CRS_DiagPrecondition(u,v,ipar), fem/src/CRSMatrix.F90:2279-2326.
 *   b200_lu_precondition  <- CRS_LUPrecondition(u,v,ipar) -> CRS_LUSolve, fem/src/CRSMatrix.F90:4550-4564, 4590-4663.
 *   b200_dot / b200_nrm2  <- ddot / dnrm2 (mathlibs/src/blas) as bound at IterSolve.F90:910-912, and their MPI
 *                            forms SParDotProd / SParNorm (fem/src/SParIterComm.F90:5081-5136) when a partition is set.
 *   b200_spmv             <- the `Matrix Vector Proc` hook: matvecsubrext_c, fem/src/Load.c:806-824, called from
 *                            fem/src/CRSMatrix.F90:1527-1529 and 4778-4781.  Exact signature of that hook.
 *   b200_set_partition /
 *   b200_comm_*           <- SParIterSolver's parallel structure (fem/src/SParIterSolver.F90:123-836, 1409-1594) in the
 *                            "complete owned rows + continuous global numbering" form the ROCSolver/AMGXSolver
 *                            bridges already build (fem/src/SolverUtils.F90:15461-15579, rocalution.cpp:64-372);
 *                            SParMatrixVector's interface exchange (SParIterSolver.F90:2630-2745) becomes a halo
 *                            exchange of x over NCCL.
 *   b200_itersolver       <- IterSolver(A,x,b,Solver) itself (fem/src/IterSolve.F90:159-1047): keyword parsing,
 *                            ipar/dpar filling, x = 1e-8 rule, preconditioner recompute policy, error mapping.
 *
 * Return value of every function: 0 = ok, non-zero = failure (message on stderr and via b200_last_error()).
 * b200_solve additionally reports through ipar(30) = HUTI_INFO exactly as the HUTI routines do; a CUDA/NCCL
 * failure maps to HUTI_HALTED (4).  There is no CPU fallback anywhere in this library.
 */
#ifndef ELMER_B200_H
#define ELMER_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* Linear System Iterative Method (IterSolve.F90:278-315) */
#define B200_METHOD_CG        1   /* huti_dcgsolv        */
#define B200_METHOD_BICGSTAB  2   /* huti_dbicgstabsolv  */
#define B200_METHOD_BICGSTABL 3   /* itermethod_bicgstabl*/
#define B200_METHOD_GCR       4   /* itermethod_gcr      */
#define B200_METHOD_IDRS      5   /* itermethod_idrs     */
#define B200_METHOD_CGS       7   /* huti_dcgssolv (right-oriented preconditioning)   */
#define B200_METHOD_TFQMR     8   /* huti_dtfqmrsolv (preconditioner in the LEFT slot) */
#define B200_METHOD_BICGSTAB2 9   /* huti_dbicgstab_2solv (preconditioner in the LEFT slot) */
#define B200_METHOD_JACOBI    10  /* itermethod_jacobi     (IterativeMethods.F90:297-393), no preconditioner */
#define B200_METHOD_RICHARDSON 11 /* itermethod_richardson (405-521), lumped-matrix scaling, no preconditioner */
#define B200_METHOD_SGS       12  /* itermethod_sgs (IterativeMethods.F90:179-285); Omega in dpar(3) (HUTI_SGSPARAM)   */
#define B200_METHOD_GMRES     6   /* huti_dgmressolv; restart in ipar(15); preconditioner in the LEFT slot as IterSolver does (IterSolve.F90:509-525) */

/* Linear System Preconditioning (IterSolve.F90:529-547) */
#define B200_PRECOND_NONE     0
#define B200_PRECOND_DIAGONAL 1
#define B200_PRECOND_ILU0     2

/* HUTI_INFO codes (fhutiter/src/huti_fdefs.h:18-50) */
#define B200_INFO_CONVERGENCE 1
#define B200_INFO_MAXITER     2
#define B200_INFO_DIVERGENCE  3
#define B200_INFO_HALTED      4

/* ---- lifetime -------------------------------------------------------------------------- */
int b200_create(void **handle);                    /* *handle must be NULL or garbage; set on return   */
int b200_destroy(void **handle);                   /* frees device memory, sets *handle = NULL         */
const char *b200_last_error(void);
int b200_device_count(int *count);                 /* cudaGetDeviceCount; fails loudly without a GPU   */
int b200_set_device(void **handle, const int *device);  /* default: LOCAL_RANK or 0 (amgx.c:161-163)   */

/* ---- matrix ---------------------------------------------------------------------------- */
/* rows[n+1], cols[nnz], diag[n] are Elmer's arrays as they are (index_base = 1) or 0-based copies
 * (index_base = 0, as the ROC/AMGX bridges pass).  Columns must be sorted ascending within each row
 * (CRS_SortMatrix) and every row must hold its diagonal.  The integer arrays are mirrored on the
 * device bit-exactly (b200_get_structure reads them back for the parity tests).  ndeg = Matrix_t%ndeg. */
int b200_set_structure(void **handle, const int *n, const int *nnz, const int *rows, const int *cols,
                       const int *diag, const int *index_base, const int *ndeg);
/* vals[nnz] host; prec_vals may be NULL (A%PrecValues absent). Marks the ILU factor stale. */
int b200_set_values(void **handle, const double *vals, const double *prec_vals);
/* Same, but vals/prec_vals are DEVICE pointers (values assembled or kept on the GPU). */
int b200_set_values_device(void **handle, const double *d_vals, const double *d_prec_vals);
/* `Linear System Scaling` on the device: ScaleLinearSystemDiagonal (fem/src/SolverUtils.F90:12976-13213; real, no
 * Mass/Damp/PrecValues).  Call after b200_set_values with the UNSCALED values: the device copy becomes
 * Values(j) * (D(i) * D(Cols(j))), D(i) = 1/sqrt(|a_ii|) (row abs-sum when a_ii is tiny).  From then on b200_solve /
 * b200_itersolver scale b and x on the way in (b *= D; bnorm = ||b||; b /= bnorm; D *= bnorm; x /= D) and back-scale x on
 * the way out (BackScaleLinearSystemDiagonal, 13515-13643): the caller passes and receives unscaled vectors and never has
 * to scale or un-scale its own matrix.  The next b200_set_values clears the state.  Single-rank handles only. */
int b200_scale_system(void **handle);
int b200_get_values(void **handle, double *vals);   /* device copy of Values (after scaling, if any) */
int b200_factorize(void **handle);                 /* ILU(order) of PrecValues (if given) else Values  */
/* Fill level of the incomplete factorisation, CRS_IncompleteLU(A, ILUn) (fem/src/CRSMatrix.F90:3445-3795; keywords
 * "Linear System Preconditioning = ILU0..ILU9" / "Linear System ILU Order", IterSolve.F90:529-547).  0 (default)
 * factorises on the matrix pattern; n > 0 first adds n rounds of first-order fill (InitializeILU1, 3664-3795).
 * Changing the order drops the current factor. */
int b200_set_ilu_order(void **handle, const int *order);
/* A % Cholesky ("Linear System Symmetric ILU", IterSolve.F90:526): the incomplete factorisation and its solve take the Cholesky branches
 * of CRS_IncompleteLU (fem/src/CRSMatrix.F90:3539-3602: L L^T on the lower part of the ILU(n) pattern, diagonal stored as 1/sqrt) and of
 * CRS_LUSolve (4618-4638: row-oriented forward sweep, column-oriented backward sweep).  b200_get_ilu_values then returns the lower part
 * and the diagonal (the reference leaves the upper part of ILUValues unwritten; here it is 0).  Changing the flag drops the factor. */
int b200_set_symmetric_ilu(void **handle, const int *flag);
/* ILUT ("Linear System Preconditioning = ILUT", "Linear System ILUT Tolerance"; CRS_ILUT, fem/src/CRSMatrix.F90:4144-4340): incomplete LU
 * whose pattern is decided by the values -- entry (i,j) of the eliminated row is kept if |value| >= tol * ||A(i,:)||_2, the diagonal always.
 * flag != 0 selects it (and resets ILU order / BILU), flag == 0 returns to ILU(order).  The factor lives on its own pattern
 * (b200_get_ilu_structure / b200_get_ilu_values after b200_factorize); rows of more than 1024 entries are refused. */
int b200_set_ilut(void **handle, const int *flag, const double *tol);
/* BILU ("Linear System Preconditioning = BILU"): the incomplete factorisation acts on the block-diagonal part of the matrix,
 * entries with MOD(i,blocks) == MOD(j,blocks) (CRS_BlockDiagonal, fem/src/CRSMatrix.F90:2382-2420; IterSolve.F90:745-765), blocks =
 * Solver % Variable % Dofs.  blocks <= 1 switches it off.  Order 0 only through the keyword front-end (see b200_itersolver). */
int b200_set_bilu_blocks(void **handle, const int *blocks);

/* ---- solve ----------------------------------------------------------------------------- */
/* b[n] in, x[n] in/out (initial guess in, solution out), ipar[50] in/out, dpar[10] in.
 * P: n x ipar(18) column-major shadow space for IDR(s) (the shim fills it with RANDOM_NUMBER as
 * IterativeMethods.F90:1640 does) or NULL for a built-in counter-based generator.
 * The ILU0 factor is (re)computed here only if none exists for the current values and
 * precond = ILU0 -- call b200_factorize / b200_set_values to apply IterSolver's recompute policy.
 * `Linear System Robust` (IterSolve.F90:482-496): ipar(26) = 1 with ipar(27) max bad iterations, ipar(29) start
 * iteration, dpar(3) robust tolerance, dpar(4) margin, dpar(5) limit (huti_fdefs.h:132-135, 153-155) is honoured
 * by BiCGStab(l) and IDR(s), the two methods that look at it in the reference. */
int b200_solve(void **handle, const double *b, double *x, int *ipar, double *dpar,
               const int *method, const int *precond, const double *P);
/* Same with b, x, P resident in device memory (used to time the solve without PCIe traffic). */
int b200_solve_device(void **handle, const double *d_b, double *d_x, int *ipar, double *dpar,
                      const int *method, const int *precond, const double *d_P);

/* IterSolver(A,x,b,Solver): `sif` is the text of the Solver section's linear-system keywords, one
 * "Keyword = value" per line (case-insensitive, as in a .sif).  solve_count in/out = A%SolveCount.
 * info_out[0] = HUTI_INFO, info_out[1] = iterations.  Unsupported keywords that would change the
 * algorithm (complex, BILU order > 0, ILUT, left preconditioning for CG/BiCGStab, backward-error stopping criteria ...)
 * make it return B200_DECLINED without touching x so that the caller runs Elmer's own path. */
#define B200_DECLINED 100
int b200_itersolver(void **handle, const double *b, double *x, const char *sif, int *solve_count,
                    int *info_out);
/* Host-only (no GPU needed): what b200_itersolver decides from the keywords alone for a matrix of n rows with ndeg dofs per node --
 * IterSolve.F90:250-577 -- method and preconditioner codes, ILU order (-1: no ILU preconditioner), BILU blocks, and the HUTI
 * ipar[50] / dpar[10] it would pass on.  Returns B200_DECLINED (reason in b200_last_error) exactly when b200_itersolver would. */
int b200_itersolver_plan(const char *sif, const int *n, const int *ndeg, int *method, int *precond, int *ilu_order,
                         int *bilu_blocks, int *ipar, double *dpar);

/* ---- the callbacks, exposed for parity tests and for user code ----------------------------- */
int b200_matvec(void **handle, const double *u, double *v);                /* v = A u            */
int b200_diag_precondition(void **handle, double *u, const double *v);     /* u = D^-1 v         */
int b200_lu_precondition(void **handle, double *u, const double *v);       /* u = (LU)^-1 v      */
int b200_dot(void **handle, const int *n, const double *x, const double *y, double *result);
int b200_nrm2(void **handle, const int *n, const double *x, double *result);
int b200_get_ilu_values(void **handle, double *ilu_vals);                  /* ILUValues          */
/* ILURows/ILUCols/ILUDiag of the current order in the caller's index base.  sizes[0] = n, sizes[1] = entries of
 * the ILU pattern; the arrays may be NULL to query the sizes first. */
int b200_get_ilu_structure(void **handle, int *sizes, int *rows, int *cols, int *diag);
int b200_get_structure(void **handle, int *rows, int *cols, int *diag);    /* device mirror back */
/* level schedule of the triangular solves: counts[0]=forward levels, [1]=backward levels,
 * [2]=forward slices, [3]=backward slices.  level_of_row may be NULL, else int[n] forward levels. */
int b200_get_levels(void **handle, int *counts, int *level_of_row);

/* `Matrix Vector Proc = "libelmer_b200 b200_spmv"` (Load.c:806-824): *spmv is Matrix_t%SpMV,
 * rows/cols are the raw 1-based arrays, u and v host arrays, reinit is always 0 in the reference.
 * The structure is uploaded on the first call; values are re-uploaded whenever their host
 * pointer or a checksum of vals changes (the hook cannot be told), or when *reinit != 0. */
void b200_spmv(void **spmv, int *n, int *rows, int *cols, double *vals, double *u, double *v, int *reinit);

/* ---- multi-GPU (one process per GPU) ----------------------------------------------------- */
/* id: 128 bytes. rank 0 calls b200_comm_unique_id and ships the bytes to the other ranks with
 * whatever the host has (MPI_Bcast in Elmer, torch.distributed in the tests). */
int b200_comm_unique_id(char *id128);
int b200_comm_init(void **handle, const int *nranks, const int *rank, const char *id128);
/* Complete owned rows in continuous global numbering (SParIterSolver.F90:1453-1488): this rank owns
 * global rows goffset[rank] .. goffset[rank+1]-1 (0-based offsets, goffset[nranks] = gn).  rows[n_own+1],
 * cols[nnz] hold GLOBAL column ids in index_base numbering, sorted ascending per row.  Builds the
 * owned x owned block, the ghost block and the halo send/receive lists in the canonical form of
 * rocalution.cpp:121-156, 222-297 (b200_get_halo_plan reads them back, they must be bit-exact). */
int b200_set_partition(void **handle, const int *gn, const int *n_own, const int *nnz, const int *rows,
                       const int *cols, const int *goffset, const int *index_base, const int *ndeg);
/* sizes[0]=nneigh, [1]=nsend, [2]=nghost; then arrays sized by a first call with NULL pointers. */
/* Host-only: where rank `me` writes inside rank r's receive area on the peer-memory halo path, derived from the np x np
 * send-count matrix cnt[s*np + d] every rank holds after b200_set_partition's count exchange.  out[0] = index of `me`
 * among r's neighbours (-1: none), out[1] = r's neighbour count, out[2] = r's ghost count, out[3] = offset of the segment. */
int b200_partition_peer_layout(const int *nranks, const int *me, const int *r, const int *cnt, long long *out);
int b200_get_halo_plan(void **handle, int *sizes, int *neigh, int *send_ptr, int *send_idx,
                       int *recv_ptr, int *ghost_gid);

/* Host-only planning steps of b200_set_partition (no GPU needed; the caller exchanges the lists itself, e.g.
 * with MPI): send_count[nranks] and, if non-NULL, send_gid = the send lists concatenated by ascending
 * destination rank as GLOBAL 0-based row ids (rocalution.cpp:121-156). */
int b200_partition_send_lists(const int *gn, const int *n_own, const int *rows, const int *cols, const int *goffset,
                              const int *index_base, const int *nranks, const int *rank, int *send_count, int *send_gid);
/* Split of the complete owned rows given the ghost ids in receive order (rocalution.cpp:286-338): sizes[0] = nnz of the
 * owned x owned block, sizes[1] = nnz of the ghost block (first call with NULL outputs); then oo_rows[n_own+1],
 * oo_cols, oo_diag (0-based, local columns) and the ghost block g_rows[n_own+1], g_cols (= n_own + ghost slot). */
int b200_partition_split(const int *n_own, const int *rows, const int *cols, const int *index_base, const int *lo, const int *hi,
                         const int *nghost, const int *ghost_gid, int *sizes, int *oo_rows, int *oo_cols, int *oo_diag,
                         int *g_rows, int *g_cols);

/* ---- matrix structure producer (host-only, no GPU needed) ------------------ */
/* Node graph of a nodal discretisation, i.e. the list matrix MakeListMatrix builds for plain nodal elements
 * (fem/src/ElementUtils.F90:881-891; rows ascending and duplicate-free as List_GetMatrixIndex keeps them,
 * fem/src/ListMatrix.F90:334-386).  elem_ptr[n_elems+1] are 0-based offsets into elem_nodes, which holds node numbers in
 * index_base numbering, bulk elements first, then boundary elements, as Mesh % Elements stores them.  perm[n_nodes]
 * is Elmer's Perm (value = 1-based row of the node, <= 0: node not in the equation); NULL = every node, identity
 * (what CreateMatrix uses when the equation covers the mesh, ElementUtils.F90:1918-1925).  k = number of rows.
 * First call with rows = cols = NULL returns *nnz; then rows[k+1], cols[nnz] in index_base numbering. */
int b200_node_graph(const int *n_elems, const int *elem_ptr, const int *elem_nodes, const int *index_base,
                    const int *n_nodes, const int *perm, const int *k, long long *nnz, int *rows, int *cols);
/* OptimizeBandwidth (fem/src/BandwidthOptimize.F90:182-445) on that graph: depth-first level search for the start
 * node (Levelize, 375-434, including the start-node update at 266-269 exactly as written), Cuthill-McKee sweep with
 * neighbours in ascending order (289-307), reversed numbering (312-323); the new numbering is accepted only if the
 * half bandwidth does not grow unless *use_optimized (`Optimize Bandwidth Use Always`) (334-339).  perm[perm_size] is
 * updated in place exactly as the reference updates Perm; *half_bandwidth = the function's result.  With
 * *optimize == 0 only the initial half bandwidth is computed (`Optimize Bandwidth = False`). */
int b200_optimize_bandwidth(const int *k, const int *rows, const int *cols, const int *index_base, const int *perm_size,
                            int *perm, const int *optimize, const int *use_optimized, int *half_bandwidth);
/* InitializeMatrix + CRS_SortMatrix (fem/src/ElementUtils.F90:1631-1732, fem/src/CRSMatrix.F90:188-246): expands the
 * node graph (in the INITIAL numbering perm_initial describes) to the dofs-per-node CRS structure in the numbering
 * of perm (both NULL: no reordering).  out_rows[dofs*k+1], out_cols[dofs*dofs*nnz], out_diag[dofs*k] in index_base
 * numbering, ready for b200_set_structure (ndeg = dofs).  out_cols/out_diag may be NULL to get the row pointers only. */
int b200_initialize_structure(const int *k, const int *rows, const int *cols, const int *index_base, const int *dofs,
                              const int *perm_size, const int *perm_initial, const int *perm, int *out_rows,
                              int *out_cols, int *out_diag);

/* ---- instrumentation --------------------------------------------------------------------- */
/* stats[0] last solve device ms (CUDA events on the solve stream), [1] matvec calls, [2] precond
 * applications, [3] last factorisation device ms, [4] kernels launched by the last solve,
 * [5] last solve H2D bytes, [6] D2H bytes, [7] iterations, [8] last SpMV-only device ms (b200_time_matvec),
 * [9] last LU-solve-only device ms, [10] final residual,
 * [11] SELL entries stored, [12]/[13] forward/backward levels, [14] kernels launched by the last factorisation,
 * [15] triangular-solve kernel in use (0 level sweeps, 1 warp tasks, 2 skewed lanes, 3 wave tiles; negative: not decided yet). */
int b200_get_stats(void **handle, double *stats16);
/* Time `reps` back-to-back v = A u launches on resident vectors with CUDA events; ms_out = mean ms. */
int b200_time_matvec(void **handle, const int *reps, double *ms_out);
int b200_time_lu_precondition(void **handle, const int *reps, double *ms_out);
int b200_version(int *major, int *minor);
/* Length every device vector handed to b200_solve_device must have: n owned rows + ghost entries
 * received from the neighbours (0 on an unpartitioned handle). */
int b200_vec_len(void **handle, long long *len);

#ifdef __cplusplus
}
#endif
#endif /* ELMER_B200_H */
