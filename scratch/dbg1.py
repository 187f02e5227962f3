import sys, time
sys.path.insert(0, '.')
import numpy as np
from oracle import oracle as O
import elmerfem_b200 as B
A, b = O.heat_cube(24, faces=["x0"]); A = A.copy(); x = np.zeros(A.n); O.scale_system(A, b, x)
M = B.Matrix(); M.set_structure(A.rows, A.cols, A.diag); M.set_values(A.vals)
for method in ["cg", "bicgstab", "bicgstabl", "gcr", "idrs"]:
    ref = O.itersolve(A, b, method=method, precond="none", tol=1e-14, maxit=3, bicgstabl_l=2)
    got = M.solve(b, method=method, precond="none", tol=1e-14, maxit=3, bicgstabl_l=2)
    print(method, ref['info'], got['info'], ref['iters'], got['iters'], np.linalg.norm(got['x']-ref['x'])/np.linalg.norm(ref['x']), ref['residual'], got['residual'])
M.close()
# perf
for ne in [100, 200]:
    t = time.time(); A, b = O.heat_cube(ne, faces=["x0"]); x = np.zeros(A.n); O.scale_system(A, b, x); print('gen', ne, time.time()-t, A.n, A.nnz, flush=True)
    M = B.Matrix(); t = time.time(); M.set_structure(A.rows, A.cols, A.diag); print('set_structure', time.time()-t)
    t = time.time(); M.set_values(A.vals); print('set_values', time.time()-t)
    ms = M.time_matvec(50); bytes_ = 12*A.nnz + 20*A.n + 4
    print('spmv ms', ms, 'GB/s', bytes_/ms/1e6, 'frac', bytes_/ms/1e6/6540.8, flush=True)
    t = time.time(); lv = M.levels(); print('levels', lv['forward'], lv['backward'], lv['slices_f'], time.time()-t)
    t = time.time(); M.factorize(); print('factor wall', time.time()-t, 'dev ms', M.stats()['factor_ms'], flush=True)
    ms = M.time_lu(10); blu = 12*A.nnz + 4*(A.n+1) + 4*A.n + 24*A.n
    print('lu ms', ms, 'GB/s', blu/ms/1e6, flush=True)
    for method, pc in [('cg','diagonal'), ('bicgstab','ilu0'), ('cg','ilu0')]:
        t = time.time(); got = M.solve(b, method=method, precond=pc, tol=1e-8, maxit=5000); w = time.time()-t
        st = got['stats']
        print(method, pc, 'info', got['info'], 'iters', got['iters'], 'solve_ms', st['solve_ms'], 'wall', w, 'it/s', got['iters']/st['solve_ms']*1e3, 'launches', st['launches'], flush=True)
    if ne == 100:
        t = time.time(); ref = O.itersolve(A, b, method='bicgstab', precond='ilu0', tol=1e-8, maxit=5000); print('oracle bicgstab ilu0', ref['iters'], time.time()-t, 'relL2', np.linalg.norm(got['x'] if False else M.solve(b, method='bicgstab', precond='ilu0', tol=1e-8, maxit=5000)['x']-ref['x'])/np.linalg.norm(ref['x']))
    M.close()
