import sys, time, os
sys.path.insert(0, '.')
import numpy as np
from oracle import oracle as O
import elmerfem_b200 as B
ne = int(sys.argv[1])
A, b = O.heat_cube(ne, faces=["x0"]); x = np.zeros(A.n); O.scale_system(A, b, x)
blu = 12*A.nnz + 4*(A.n+1) + 4*A.n + 24*A.n
ilu = O.ilu0(A) if ne <= 100 else None
v = np.random.RandomState(3).standard_normal(A.n)
ref = O.lu_precond(A, ilu, v) if ilu is not None else None
for la, gs, ss, bps in [(1,100,0,0),(1,0,0,0),(1,20,0,0),(2,100,0,0),(2,0,0,0),(2,100,50,0),(3,100,0,0),(3,100,50,0),(2,100,0,1),(4,100,0,0)]:
    os.environ['B200_TRI_LOOKAHEAD']=str(la); os.environ['B200_TRI_BLOCKS_PER_SM']=str(bps)
    os.environ['B200_TRI_GATE_SLEEP']=str(gs); os.environ['B200_TRI_SPIN_SLEEP']=str(ss)
    M = B.Matrix(); M.set_structure(A.rows, A.cols, A.diag); M.set_values(A.vals); M.factorize()
    ms = M.time_lu(5)
    ok = '' if ref is None else ('bitexact=%s' % np.array_equal(M.lu_precondition(v), ref))
    print(ne, 'lookahead', la, 'gate_sleep', gs, 'spin_sleep', ss, 'blocks/SM', bps, 'lu ms', round(ms,3), 'GB/s', round(blu/ms/1e6,1), 'us/level', round(ms*1e3/(2*M.levels()['forward']),3), ok, flush=True)
    M.close()
