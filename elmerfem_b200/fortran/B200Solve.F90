!------------------------------------------------------------------------------
!  B200Solve.F90 -- Fortran side of the B200 linear-solve path (ISO_C_BINDING shim).
!
!  Compile against Elmer's module files, exactly like any user solver:
!      elmerf90 B200Solve.F90 -o B200Solve.so -L<dir of libelmer_b200.so> -lelmer_b200
!  and select it in the Solver section of the .sif, leaving every linear-system keyword as it is:
!      Linear System Solver = Iterative
!      Linear System Iterative Method = BiCGStab          ! cg | bicgstab | bicgstabl | gcr | idrs
!      Linear System Preconditioning = ILU0               ! none | diagonal | ilu0
!      Before Linsolve = "B200Solve" "B200BeforeLinsolve"
!
!  `Before Linsolve` is called from SolveSystem (fem/src/SolverUtils.F90:15841-15846) through
!  ExecLinSolveProcs (fem/src/LoadMod.F90:586-614); a non-zero return value skips Elmer's own
!  SolveLinearSystem, a zero return value lets it run.  The shim therefore
!    * returns 0 (DECLINES) for anything the device path does not implement -- Elmer then solves the system
!      itself; nothing inside libelmer_b200 ever falls back to the CPU;
!    * otherwise does what SolveSystem / SolveLinearSystem do around IterSolver (SolverUtils.F90:15848-15865,
!      14748-14751, 14869, 14925-14965): residual-mode change of variables, ScaleLinearSystem, solve,
!      BackScaleLinearSystem, CalculateLoads, BackRotateNTSystem, ComputeChange.
!
!  This file cannot be compiled in the development image (no Fortran compiler); it is checked by
!  tests/test_abi.py for agreement of every BIND(C) name with include/elmer_b200.h.
!------------------------------------------------------------------------------
MODULE B200Interface
  USE ISO_C_BINDING
  IMPLICIT NONE

  INTEGER(C_INT), PARAMETER :: B200_DECLINED = 100

  INTERFACE
    ! int b200_create(void **handle);   handle slot = Matrix_t % SpMV (Types.F90:265), initially 0
    INTEGER(C_INT) FUNCTION b200_create(handle) BIND(C, NAME="b200_create")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
    END FUNCTION
    INTEGER(C_INT) FUNCTION b200_destroy(handle) BIND(C, NAME="b200_destroy")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
    END FUNCTION
    ! Matrix_t: Rows, Cols, Diag as they are (index_base = 1), ndeg = A % ndeg
    INTEGER(C_INT) FUNCTION b200_set_structure(handle, n, nnz, rows, cols, diag, index_base, ndeg) &
        BIND(C, NAME="b200_set_structure")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      INTEGER(C_INT) :: n, nnz, rows(*), cols(*), diag(*), index_base, ndeg
    END FUNCTION
    ! A % Values and (optionally) A % PrecValues; pass C_NULL_PTR when PrecValues is not associated
    INTEGER(C_INT) FUNCTION b200_set_values(handle, vals, prec_vals) BIND(C, NAME="b200_set_values")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      REAL(C_DOUBLE) :: vals(*)
      TYPE(C_PTR), VALUE :: prec_vals
    END FUNCTION
    ! int b200_scale_system(void **handle);   ScaleLinearSystemDiagonal on the device copy (after b200_set_values)
    INTEGER(C_INT) FUNCTION b200_scale_system(handle) BIND(C, NAME="b200_scale_system")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
    END FUNCTION
    INTEGER(C_INT) FUNCTION b200_factorize(handle) BIND(C, NAME="b200_factorize")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
    END FUNCTION
    ! the HUTI call of IterSolver (IterSolve.F90:1004-1005): ipar(50)/dpar(10) verbatim
    INTEGER(C_INT) FUNCTION b200_solve(handle, b, x, ipar, dpar, method, precond, P) BIND(C, NAME="b200_solve")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      REAL(C_DOUBLE) :: b(*), x(*), dpar(*)
      INTEGER(C_INT) :: ipar(*), method, precond
      TYPE(C_PTR), VALUE :: P
    END FUNCTION
    ! IterSolver(A,x,b,Solver) with the keywords passed as text
    INTEGER(C_INT) FUNCTION b200_itersolver(handle, b, x, sif, solve_count, info_out) BIND(C, NAME="b200_itersolver")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      REAL(C_DOUBLE) :: b(*), x(*)
      CHARACTER(KIND=C_CHAR) :: sif(*)
      INTEGER(C_INT) :: solve_count, info_out(2)
    END FUNCTION
    ! host-only: what b200_itersolver would decide for these keywords (rc = B200_DECLINED: Elmer's own path would run)
    INTEGER(C_INT) FUNCTION b200_itersolver_plan(sif, n, ndeg, method, precond, ilu_order, bilu_blocks, ipar, dpar) &
        BIND(C, NAME="b200_itersolver_plan")
      IMPORT; CHARACTER(KIND=C_CHAR) :: sif(*)
      INTEGER(C_INT) :: n, ndeg, method, precond, ilu_order, bilu_blocks, ipar(50)
      REAL(C_DOUBLE) :: dpar(10)
    END FUNCTION
    ! the five HUTI callbacks, host vectors
    INTEGER(C_INT) FUNCTION b200_matvec(handle, u, v) BIND(C, NAME="b200_matvec")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      REAL(C_DOUBLE) :: u(*), v(*)
    END FUNCTION
    INTEGER(C_INT) FUNCTION b200_diag_precondition(handle, u, v) BIND(C, NAME="b200_diag_precondition")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      REAL(C_DOUBLE) :: u(*), v(*)
    END FUNCTION
    INTEGER(C_INT) FUNCTION b200_lu_precondition(handle, u, v) BIND(C, NAME="b200_lu_precondition")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      REAL(C_DOUBLE) :: u(*), v(*)
    END FUNCTION
    INTEGER(C_INT) FUNCTION b200_dot(handle, n, x, y, res) BIND(C, NAME="b200_dot")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      INTEGER(C_INT) :: n
      REAL(C_DOUBLE) :: x(*), y(*), res
    END FUNCTION
    INTEGER(C_INT) FUNCTION b200_nrm2(handle, n, x, res) BIND(C, NAME="b200_nrm2")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      INTEGER(C_INT) :: n
      REAL(C_DOUBLE) :: x(*), res
    END FUNCTION
    ! multi-GPU: one MPI rank per GPU; rank 0 creates the id and MPI_BCASTs the 128 bytes
    INTEGER(C_INT) FUNCTION b200_comm_unique_id(id128) BIND(C, NAME="b200_comm_unique_id")
      IMPORT; CHARACTER(KIND=C_CHAR) :: id128(128)
    END FUNCTION
    INTEGER(C_INT) FUNCTION b200_comm_init(handle, nranks, rank, id128) BIND(C, NAME="b200_comm_init")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      INTEGER(C_INT) :: nranks, rank
      CHARACTER(KIND=C_CHAR) :: id128(128)
    END FUNCTION
    ! complete owned rows in continuous global numbering, as ROCSolver builds them (SolverUtils.F90:15461-15579)
    INTEGER(C_INT) FUNCTION b200_set_partition(handle, gn, n_own, nnz, rows, cols, goffset, index_base, ndeg) &
        BIND(C, NAME="b200_set_partition")
      IMPORT; INTEGER(C_INTPTR_T) :: handle
      INTEGER(C_INT) :: gn, n_own, nnz, rows(*), cols(*), goffset(*), index_base, ndeg
    END FUNCTION
    ! matrix-structure producer (host-only): what CreateMatrix does for nodal elements (ElementUtils.F90:1745-2170);
    ! perm = Solver % Variable % Perm, arrays 1-based (index_base = 1)
    INTEGER(C_INT) FUNCTION b200_node_graph(n_elems, elem_ptr, elem_nodes, index_base, n_nodes, perm, k, nnz, rows, cols) &
        BIND(C, NAME="b200_node_graph")
      IMPORT; INTEGER(C_INT) :: n_elems, elem_ptr(*), elem_nodes(*), index_base, n_nodes, perm(*), k
      INTEGER(C_LONG_LONG) :: nnz
      TYPE(C_PTR), VALUE :: rows, cols          ! C_NULL_PTR on the sizing call
    END FUNCTION
    INTEGER(C_INT) FUNCTION b200_optimize_bandwidth(k, rows, cols, index_base, perm_size, perm, optimize, &
        use_optimized, half_bandwidth) BIND(C, NAME="b200_optimize_bandwidth")
      IMPORT; INTEGER(C_INT) :: k, rows(*), cols(*), index_base, perm_size, perm(*), optimize, use_optimized, half_bandwidth
    END FUNCTION
    INTEGER(C_INT) FUNCTION b200_initialize_structure(k, rows, cols, index_base, dofs, perm_size, perm_initial, perm, &
        out_rows, out_cols, out_diag) BIND(C, NAME="b200_initialize_structure")
      IMPORT; INTEGER(C_INT) :: k, rows(*), cols(*), index_base, dofs, perm_size, perm_initial(*), perm(*)
      INTEGER(C_INT) :: out_rows(*), out_cols(*), out_diag(*)
    END FUNCTION
    FUNCTION b200_last_error() RESULT(msg) BIND(C, NAME="b200_last_error")
      IMPORT; TYPE(C_PTR) :: msg
    END FUNCTION
  END INTERFACE
END MODULE B200Interface


!------------------------------------------------------------------------------
!> `Before Linsolve` procedure.  Returns 1 when the system was solved on the GPU, 0 when declined.
!------------------------------------------------------------------------------
FUNCTION B200BeforeLinsolve( Model, Solver, A, b, x, n, DOFs, Norm ) RESULT(stat)
!------------------------------------------------------------------------------
  USE DefUtils
  USE SolverUtils
  USE B200Interface
  IMPLICIT NONE
  TYPE(Model_t) :: Model
  TYPE(Solver_t) :: Solver
  TYPE(Matrix_t), POINTER :: A
  INTEGER :: n, DOFs
  REAL(KIND=dp), TARGET :: b(n), x(n)
  REAL(KIND=dp) :: Norm
  INTEGER :: stat
!------------------------------------------------------------------------------
  TYPE(ValueList_t), POINTER :: Params
  CHARACTER(LEN=4096) :: sif
  CHARACTER(:), ALLOCATABLE :: str
  LOGICAL :: Found, ScaleSystem, DeviceScaling, L, ResidualMode, BackRotation, CalcLoads
  REAL(KIND=dp), ALLOCATABLE, TARGET :: Res(:)
  REAL(KIND=dp), POINTER :: bb(:), pRes(:)
  TYPE(Variable_t), POINTER :: NodalLoads
  CHARACTER(LEN=32) :: tmp
  INTEGER :: rc, info(2), nnz, ival, base, ndeg
  REAL(KIND=dp) :: rval
  INTEGER(C_INTPTR_T) :: handle
  TYPE(C_PTR) :: prec
  REAL(KIND=dp), POINTER :: pv(:)
  CHARACTER(*), PARAMETER :: Caller = 'B200BeforeLinsolve'

  stat = 0
  Params => Solver % Values

  ! ---- applicability (everything else is left to Elmer's own path)
  IF ( ParEnv % PEs > 1 ) RETURN               ! the MPI variant goes through B200ParallelSolve (INTEGRATION.md)
  IF ( A % FORMAT /= MATRIX_CRS ) RETURN
  IF ( A % COMPLEX ) RETURN
  IF ( ASSOCIATED( A % ConstraintMatrix ) .OR. ASSOCIATED( A % AddMatrix ) ) RETURN
  str = ListGetString( Params, 'Linear System Solver', Found )
  IF ( .NOT. Found ) RETURN
  IF ( str /= 'iterative' ) RETURN
  IF ( ListGetLogical( Params, 'Linear System Skip Scaling', Found ) ) RETURN
  ! branches of SolveSystem / SolveLinearSystem this function does not reproduce (SolverUtils.F90:15805-15892,
  ! 14535-14700, 14740-14790): block and restricted systems, eigen / harmonic analysis, explicit time stepping on a
  ! lumped matrix, Anderson acceleration, change computed in the scaled system, guess normalisation, constraint modes
  IF ( ListGetLogical( Params, 'Linear System Block Mode', Found ) ) RETURN
  IF ( Solver % NOFEigenValues > 0 ) THEN
    IF ( ListGetLogical( Params, 'Eigen Analysis', Found ) ) RETURN
    IF ( ListGetLogical( Params, 'Harmonic Analysis', Found ) ) RETURN
  END IF
  IF ( A % Lumped ) RETURN
  IF ( ListGetLogical( Params, 'Nonlinear System Acceleration', Found ) ) RETURN
  IF ( ListGetLogical( Params, 'Nonlinear System Compute Change in Scaled System', Found ) ) RETURN
  IF ( ListGetLogical( Params, 'Linear System Normalize Guess', Found ) ) RETURN
  IF ( ListGetLogical( Model % Control, 'Constraint Modes Analysis', Found ) ) RETURN
  IF ( ListGetLogical( Params, 'Nonlinear System Constraint Modes', Found ) ) RETURN
  IF ( ListGetLogical( Params, 'Steady State Constraint Modes', Found ) ) RETURN
  IF ( ListGetLogical( Params, 'Run Control Constraint Modes', Found ) ) RETURN

  ! ---- keywords, verbatim, as text (parsed by b200_itersolver exactly as IterSolve.F90:250-583 does)
  sif = ''
  CALL AddStr( 'Linear System Iterative Method' )
  CALL AddStr( 'Linear System Preconditioning' )
  CALL AddInt( 'Linear System Max Iterations' )
  CALL AddInt( 'Linear System Min Iterations' )
  CALL AddInt( 'Linear System Residual Output' )
  CALL AddInt( 'Linear System GCR Restart' )
  CALL AddInt( 'Linear System GMRES Restart' )
  WRITE( tmp, '(I0)' ) Solver % Variable % DOFs
  CALL Append( 'B200 Variable Dofs = ' // TRIM(tmp) )                 ! Blocks of BILU (IterSolve.F90:746)
  CALL AddInt( 'BiCGstabl polynomial degree' )
  CALL AddInt( 'IDRS parameter' )
  CALL AddInt( 'Linear System Precondition Recompute' )
  CALL AddReal( 'Linear System Convergence Tolerance' )
  CALL AddReal( 'Linear System Divergence Limit' )
  CALL AddReal( 'Linear System ILU Order' )
  CALL AddReal( 'Linear System ILUT Tolerance' )
  CALL AddReal( 'SGS Overrelaxation Factor' )
  CALL AddReal( 'Linear System ILU Factor' )
  CALL AddLog( 'Linear System Refactorize' )
  CALL AddLog( 'No Precondition Recompute' )
  CALL AddLog( 'IDRS Smoothing' )
  CALL AddLog( 'Linear System Complex' )
  CALL AddLog( 'Linear System Pseudo Complex' )
  CALL AddLog( 'Linear System Symmetric ILU' )
  CALL AddLog( 'Linear System Left Preconditioning' )
  CALL AddLog( 'Linear System Robust' )
  CALL AddReal( 'Linear System Robust Tolerance' )
  CALL AddReal( 'Linear System Robust Limit' )
  CALL AddReal( 'Linear System Robust Margin' )
  CALL AddInt( 'Linear System Robust Max Iterations' )
  CALL AddInt( 'Linear System Robust Start Iteration' )
  CALL AddLog( 'Linear System Componentwise Backward Error' )
  CALL AddLog( 'Linear System Normwise Backward Error' )
  CALL AddLog( 'Edge Basis' )

  ! ---- what SolveLinearSystem does before IterSolver: default diagonal scaling (SolverUtils.F90:14492-14498, 14748-14751)
  ScaleSystem = ListGetLogical( Params, 'Linear System Scaling', Found )
  IF ( .NOT. Found ) ScaleSystem = .TRUE.
  IF ( ALL( b(1:n) == 0.0_dp ) ) RETURN        ! zero rhs shortcut stays with Elmer (14717-14736)

  ! ---- `Linear System Residual Mode`: SolveSystem calls this procedure BEFORE its change of variables
  ! (SolverUtils.F90:15841-15865), and ComputeChange adds the stored previous solution back (10878-10883), so the
  ! change of variables A dx = b - A x0, dx0 = 0 has to be done here when the solve is taken over.
  ResidualMode = ListGetLogical( Params, 'Linear System Residual Mode', Found )
  bb => b
  IF ( ResidualMode ) THEN
    ALLOCATE( Res(n) )
    pRes => Res
    IF ( ASSOCIATED( Solver % Variable % Perm ) ) CALL RotateNTSystemAll( x, Solver % Variable % Perm, DOFs )
    CALL LinearSystemResidual( A, b, x, pRes )
    bb => Res
    x = 0.0_dp
    IF ( ALL( Res == 0.0_dp ) ) THEN             ! zero rhs shortcut stays with Elmer, as above
      x(1:n) = Solver % Variable % NonlinValues(1:n)     ! stored by SolveSystem at 15836
      RETURN
    END IF
  END IF
  ! 'B200 Device Scaling = True': the device copy is scaled by b200_scale_system (bit-identical values), b and x are
  ! scaled / back-scaled inside b200_itersolver, and the host matrix is never touched (no ScaleLinearSystem /
  ! BackScaleLinearSystem passes over A % Values).  Not with a separate preconditioning matrix.
  DeviceScaling = ScaleSystem .AND. ListGetLogical( Params, 'B200 Device Scaling', Found ) .AND. &
                  .NOT. ASSOCIATED( A % PrecValues )
  IF ( ScaleSystem .AND. .NOT. DeviceScaling ) &
      CALL ScaleLinearSystem( Solver, A, bb, x, RhsScaling=.TRUE., ConstraintScaling=.TRUE. )      ! 14748-14751

  ! ---- device mirror: structure once per matrix, values every call (once per nonlinear iteration)
  handle = A % SpMV
  IF ( handle == 0 ) THEN
    rc = b200_create( handle )
    IF ( rc /= 0 ) CALL Fatal( Caller, 'b200_create failed (no CUDA device?)' )
    nnz = A % Rows(n+1) - 1
    base = 1
    ndeg = A % ndeg
    rc = b200_set_structure( handle, n, nnz, A % Rows, A % Cols, A % Diag, base, ndeg )
    IF ( rc /= 0 ) CALL Fatal( Caller, 'b200_set_structure failed' )
    A % SpMV = handle
  END IF
  prec = C_NULL_PTR
  IF ( ASSOCIATED( A % PrecValues ) ) THEN
    pv => A % PrecValues
    prec = C_LOC( pv(1) )
  END IF
  rc = b200_set_values( handle, A % Values, prec )
  IF ( rc /= 0 ) CALL Fatal( Caller, 'b200_set_values failed' )
  IF ( DeviceScaling ) THEN
    rc = b200_scale_system( handle )
    IF ( rc /= 0 ) CALL Fatal( Caller, 'b200_scale_system failed' )
  END IF

  ! ---- IterSolver on the device
  rc = b200_itersolver( handle, bb, x, TRIM(sif)//C_NULL_CHAR, A % SolveCount, info )

  IF ( rc == B200_DECLINED ) THEN
    ! undo the scaling (and the change of variables) and let Elmer's own path run
    IF ( ScaleSystem .AND. .NOT. DeviceScaling ) CALL BackScaleLinearSystem( Solver, A, bb, x, ConstraintScaling=.TRUE. )
    IF ( ResidualMode ) x(1:n) = Solver % Variable % NonlinValues(1:n)      ! stored by SolveSystem at 15836
    RETURN
  END IF
  IF ( rc /= 0 ) CALL Fatal( Caller, 'b200_itersolver failed' )

  ! ---- error mapping of IterSolve.F90:1016-1038, code by code: convergence sets LinConverged = 1; divergence raises NumericalError
  ! (fatal unless 'Global Abort Not Converged' says otherwise); too many iterations raises it only when 'Linear System Abort Not
  ! Converged' (default True) asks for it; a halted iteration warns and continues; every other code (breakdowns) just continues.
  ! Anything but convergence leaves LinConverged = 0.
  Solver % Variable % LinConverged = 0
  IF ( info(1) == 1 ) THEN                                          ! HUTI_CONVERGENCE
    Solver % Variable % LinConverged = 1
  ELSE IF ( info(1) == 3 ) THEN                                     ! HUTI_DIVERGENCE
    CALL NumericalError( Caller, 'System diverged over maximum tolerance.' )
  ELSE IF ( info(1) == 2 ) THEN                                     ! HUTI_MAXITER
    L = ListGetLogical( Params, 'Linear System Abort Not Converged', Found )
    IF ( .NOT. Found ) L = .TRUE.
    IF ( L ) THEN
      CALL NumericalError( Caller, 'Too many iterations were needed.' )
    ELSE
      CALL Info( Caller, 'Linear iteration did not converge to tolerance', Level=6 )
    END IF
  ELSE IF ( info(1) == 4 ) THEN                                     ! HUTI_HALTED
    CALL Warn( Caller, 'Iteration halted due to problem in algorithm, trying to continue' )
  END IF
  WRITE( Message, '(A,I0,A,I0)' ) 'B200 linear solve: HUTI_INFO=', info(1), ' iterations=', info(2)
  CALL Info( Caller, Message, Level=5 )

  ! ---- what SolveLinearSystem does after IterSolver, in its order (14925-14965): back-scaling, nodal loads,
  ! back-rotation of normal-tangential dofs, ComputeChange (which also applies relaxation and, in residual mode,
  ! adds the previous solution)
  IF ( ScaleSystem .AND. .NOT. DeviceScaling ) CALL BackScaleLinearSystem( Solver, A, bb, x, ConstraintScaling=.TRUE. )
  NodalLoads => VariableGet( Solver % Mesh % Variables, GetVarName( Solver % Variable ) // ' Loads' )
  IF ( ASSOCIATED( NodalLoads ) ) THEN
    CalcLoads = ListGetLogical( Params, 'Calculate Loads', Found )
    IF ( .NOT. Found ) CalcLoads = .TRUE.
    IF ( CalcLoads ) CALL CalculateLoads( Solver, A, x, DOFs, .TRUE., NodalLoads )
  END IF
  BackRotation = ListGetLogical( Params, 'Back Rotate N-T Solution', Found )
  IF ( .NOT. Found ) BackRotation = .TRUE.
  BackRotation = BackRotation .AND. ASSOCIATED( Solver % Variable % Perm )
  IF ( BackRotation ) THEN
    CALL BackRotateNTSystem( x, Solver % Variable % Perm, DOFs )
    IF ( ASSOCIATED( NodalLoads ) ) CALL BackRotateNTSystem( NodalLoads % Values, NodalLoads % Perm, DOFs )
  END IF
  CALL ComputeChange( Solver, .FALSE., n, x, Matrix=A, RHS=bb )
  Norm = Solver % Variable % Norm
  stat = 1

CONTAINS

  SUBROUTINE Append( line )
    CHARACTER(*) :: line
    sif = TRIM(sif) // TRIM(line) // C_NEW_LINE
  END SUBROUTINE Append

  SUBROUTINE AddStr( key )
    CHARACTER(*) :: key
    CHARACTER(:), ALLOCATABLE :: v
    LOGICAL :: f
    v = ListGetString( Params, key, f )
    IF ( f ) CALL Append( key // ' = ' // v )
  END SUBROUTINE AddStr

  SUBROUTINE AddInt( key )
    CHARACTER(*) :: key
    CHARACTER(LEN=32) :: t
    LOGICAL :: f
    INTEGER :: v
    v = ListGetInteger( Params, key, f )
    IF ( f ) THEN
      WRITE( t, '(I0)' ) v
      CALL Append( key // ' = ' // TRIM(t) )
    END IF
  END SUBROUTINE AddInt

  SUBROUTINE AddReal( key )
    CHARACTER(*) :: key
    CHARACTER(LEN=40) :: t
    LOGICAL :: f
    REAL(KIND=dp) :: v
    v = ListGetConstReal( Params, key, f )
    IF ( f ) THEN
      WRITE( t, '(ES25.17E3)' ) v
      CALL Append( key // ' = ' // TRIM(ADJUSTL(t)) )
    END IF
  END SUBROUTINE AddReal

  SUBROUTINE AddLog( key )
    CHARACTER(*) :: key
    LOGICAL :: f, v
    v = ListGetLogical( Params, key, f )
    IF ( f ) THEN
      IF ( v ) THEN
        CALL Append( key // ' = True' )
      ELSE
        CALL Append( key // ' = False' )
      END IF
    END IF
  END SUBROUTINE AddLog
!------------------------------------------------------------------------------
END FUNCTION B200BeforeLinsolve
!------------------------------------------------------------------------------
