'''Device timing of the matrix-free-pattern SpMV and of a constrained CG solve on the assembled n^3 Poisson matrix
(development helper).  usage: python scripts/time_spmv.py [n] [degree]'''
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from nutils_b200 import bspline, points, engine
from bench import make_nodes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nc = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # 3: elasticity matrix (vector-valued space)
ctx = engine.Context.get(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
b1 = [bspline.spline_basis_1d(n, p) for _ in range(3)]
plan = engine.Plan(ctx, b1, points.tensor_gauss(3, 2 * p), make_nodes((n,) * 3), ncomp=nc)
dev = torch.device('cuda', 0)
K = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
f = torch.empty(plan.ndofs, dtype=torch.float64, device=dev)
plan.assemble_rows_device([engine.form_stiffness(3) if nc == 1 else engine.form_elasticity(3, 1., .5 / .3 - 1.)], [engine.form_load(3, nc)], [K], [f])
x = torch.rand(plan.ndofs, dtype=torch.float64, device=dev)
y = torch.empty_like(x)
for _ in range(3):
    plan.spmv_device(K, x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    plan.spmv_device(K, x, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({'kernel': 'k_spmv', 'n': n, 'p': p, 'ncomp': nc, 'ndofs': plan.ndofs, 'nnz': plan.nnz, 'ms': ms, 'algorithmic_GBps': (8 * plan.nnz + 16 * plan.ndofs) / ms * 1e-6}))
if nc != 1:
    sys.exit(0)
# Dirichlet on the first and last dof plane along x, CG to 1e-10
nd = n + p
mask = torch.zeros((nd, nd, nd), dtype=torch.uint8, device=dev)
mask[0] = 1
mask[-1] = 1
u = torch.zeros((nd, nd, nd), dtype=torch.float64, device=dev)
u[-1] = 1.
torch.cuda.synchronize()
t0 = time.perf_counter()
plan.cg_device(K, f, u.clone(), constrained=mask, rtol=1e-2)  # first call: lazy module loading of the CG kernels
torch.cuda.synchronize()
t0 = time.perf_counter()
its, res = plan.cg_device(K, f, u, constrained=mask, rtol=1e-10)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({'solve': 'jacobi-pcg', 'iterations': its, 'resnorm': res, 'seconds': dt, 'ms_per_iteration': dt / max(its, 1) * 1e3, 'u_mean': float(u.mean())}))
