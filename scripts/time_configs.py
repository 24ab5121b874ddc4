'''Device timing of b2_assemble_rows_device for the configurations of BASELINE.json beyond the benchmark one (development helper).'''
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from nutils_b200 import bspline, points, engine
from bench import make_nodes

ctx = engine.Context.get(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
dev = torch.device('cuda', 0)
ctx.set_option('rows_variant', int(os.environ.get('ROWS_VARIANT', 0)))
ctx.set_option('rows_dbg', int(os.environ.get('ROWS_DBG', 0)))
for k in ('rows_gpre_vec', 'rows_sym', 'rows_vecsym', 'rows_vec_offdiag'):
    if k.upper() in os.environ:
        ctx.set_option(k, int(os.environ[k.upper()]))
cases = [dict(n=64, p=1), dict(n=64, p=2), dict(n=48, p=3), dict(n=32, p=4), dict(n=32, p=2, ncomp=3), dict(n=48, p=2, ncomp=3), dict(n=24, p=3, ncomp=3)]
if len(sys.argv) > 1:
    cases = [json.loads(a) for a in sys.argv[1:]]
for c in cases:
    n, p, nc = c['n'], c['p'], c.get('ncomp', 1)
    b1 = [bspline.spline_basis_1d(n, p) for _ in range(3)]
    rules = points.tensor_gauss(3, 2 * p)
    plan = engine.Plan(ctx, b1, rules, make_nodes((n,) * 3), ncomp=nc)
    if nc == 1:
        Ds, Cs = [engine.form_stiffness(3), engine.form_mass(3)], [engine.form_load(3)]
    else:
        C = numpy.zeros((3, 4)); C[2, 0] = -1.
        Ds, Cs = [engine.form_elasticity(3, 1., .5 / .3 - 1.)], [C]
    mats = [torch.empty(plan.nnz, dtype=torch.float64, device=dev) for _ in Ds]
    vecs = [torch.empty(plan.ndofs, dtype=torch.float64, device=dev) for _ in Cs]
    plan.assemble_rows_device(Ds, Cs, mats, vecs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        plan.assemble_rows_device(Ds, Cs, mats, vecs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps(dict(c, ndofs=plan.ndofs, nnz=plan.nnz, ms=ms, dof_per_s=plan.ndofs / ms * 1e3, gbs=8e-6 * plan.nnz * len(Ds) / ms)))
