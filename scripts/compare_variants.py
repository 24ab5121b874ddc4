'''Development helper: time the owner-computes kernel variants (context option rows_variant; B2_EXPERIMENT builds) and compare
their output with the default variant on the same problem.  usage: compare_variants.py n p v1 v2 ...'''
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from nutils_b200 import bspline, points, engine
from bench import make_nodes

n, p = int(sys.argv[1]), int(sys.argv[2])
variants = [int(v) for v in sys.argv[3:]]
ctx = engine.Context.get(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
b1 = [bspline.spline_basis_1d(n, p) for _ in range(3)]
plan = engine.Plan(ctx, b1, points.tensor_gauss(3, 2 * p), make_nodes((n,) * 3))
Ds = [engine.form_stiffness(3), engine.form_mass(3)]
Cs = [engine.form_load(3)]
dev = torch.device('cuda', 0)


def run(variant, reps=5):
    mats = [torch.zeros(plan.nnz, dtype=torch.float64, device=dev) for _ in Ds]
    vecs = [torch.zeros(plan.ndofs, dtype=torch.float64, device=dev) for _ in Cs]
    ctx.set_option('rows_variant', variant)
    ctx.kernel_time()
    for _ in range(2):
        plan.assemble_rows_device(Ds, Cs, mats, vecs)
    torch.cuda.synchronize()
    ctx.set_option('time_kernels', 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.assemble_rows_device(Ds, Cs, mats, vecs)
    e1.record()
    torch.cuda.synchronize()
    kms, kl = ctx.kernel_time()
    ctx.set_option('time_kernels', 0)
    return e0.elapsed_time(e1) / reps, mats, vecs


ms0, m0, v0 = run(0)
print(json.dumps({'variant': 0, 'ms': ms0}))
for v in variants:
    ms, m, vv = run(v)
    err = [float((a - b).norm() / b.norm()) for a, b in zip(m + vv, m0 + v0)]
    print(json.dumps({'variant': v, 'ms': ms, 'relerr_vs_default': err}))
    del m, vv
