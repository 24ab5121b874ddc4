'''Device timing of the owner-computes kernel at n^3 (development helper; bench.py is the judged benchmark).'''
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from nutils_b200 import bspline, points, engine
from bench import make_nodes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = engine.Context.get(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
if int(os.environ.get('GRADED', 0)):
    # graded knot vectors: every element has its own coefficient set, the kernel's table-driven S1/S2 variants run everywhere
    rng = numpy.random.RandomState(4)
    b1 = [bspline.spline_basis_1d(n, p, knotvalues=numpy.concatenate([[0.], numpy.cumsum(.3 + rng.rand(n))])) for _ in range(3)]
else:
    b1 = [bspline.spline_basis_1d(n, p) for _ in range(3)]
rules = points.tensor_gauss(3, 2 * p)
plan = engine.Plan(ctx, b1, rules, make_nodes((n,) * 3))
Ds = [engine.form_stiffness(3), engine.form_mass(3)]
Cs = [engine.form_load(3)]
dev = torch.device('cuda', 0)
mats = [torch.empty(plan.nnz, dtype=torch.float64, device=dev) for _ in Ds]
vecs = [torch.empty(plan.ndofs, dtype=torch.float64, device=dev) for _ in Cs]
variant = int(os.environ.get('ROWS_VARIANT', 0))
ctx.set_option('rows_variant', variant)
ctx.set_option('rows_split_forms', int(os.environ.get('ROWS_SPLIT_FORMS', 0)))
ctx.set_option('rows_gpre', int(os.environ.get('ROWS_GPRE', 1)))
if 'ROWS_SYM' in os.environ:
    ctx.set_option('rows_sym', int(os.environ['ROWS_SYM']))
ctx.set_option('rows_dbg', int(os.environ.get('ROWS_DBG', 0)))
for nseg in [int(a) for a in sys.argv[3:]] or [0]:
    ctx.set_option('rows_nseg', nseg)
    for _ in range(2):
        plan.assemble_rows_device(Ds, Cs, mats, vecs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.assemble_rows_device(Ds, Cs, mats, vecs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({'n': n, 'p': p, 'nseg': nseg, 'variant': variant, 'ms': ms, 'dof_per_s': plan.ndofs / ms * 1e3, 'sumM-sumf': float(mats[1].sum() - vecs[0].sum())}))
