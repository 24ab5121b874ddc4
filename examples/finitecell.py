#!/usr/bin/env python
'''Finite-cell Poisson problem on a ball immersed in a box (BASELINE.json configs[4]), assembled, solved and evaluated on the GPU:

    -div grad u + u = g  in  |x| < R,   du/dn = 0 on |x| = R,   exact solution u = cos(pi |x| / R)

on an n^3 background grid of degree-p splines; cut cells carry an octree quadrature (nutils_b200.fcm -- the reference would
build the same kind of tables with topo.trim, see nutils_b200/adapter.py).  K + M and the load with its pointwise coefficient g
are integrated by the element-set kernel (FP64 tensor cores), the system is solved by Jacobi-PCG on the materialised
pattern of the pruned basis, and the error is measured by evaluating the discrete solution at the quadrature points
(b2_evaluate_elemset_device).  The homogeneous Neumann condition is natural, so no boundary integral is needed; the quadrature
error of the octree (first order in the sub-cell size at the boundary) limits the accuracy.

    python examples/finitecell.py [n] [degree] [depth]
'''

import os
import sys
import time
import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nutils_b200 import bspline, engine, fcm, matrix  # noqa: E402


def main_dirichlet(n=24, degree=2, depth=3, radius=.8, beta=100., rtol=1e-10):
    '''The same solution with a DIRICHLET condition on the immersed sphere, imposed by a penalty term on the trimmed boundary:

        int grad v . grad u dV + beta int_G v u dS = int v g dV + beta int_G v u_D dS,    g = -div grad u,  u_D = cos(pi) = -1

    (du/dn = 0 holds for the manufactured solution, so the penalty form is consistent).  Volume points and surface points live in
    ONE element set; per-point coefficients switch the stiffness off on the surface and the penalty mass off in the volume, and
    per-point scaled normals give every point its measure (|det J| in the volume, the physical weight on the surface).'''
    vol = fcm.octree_ball(n, degree, depth, radius=radius)
    h = 2. / n
    vnormals = numpy.zeros((int(vol[1][-1]), 3))
    vnormals[:, 0] = h      # |det J| |J^-T n| = h^3 / h * h = |det J|: plain volume measure
    elem_ids, qoff, qcoords, qweights, normals, on_surface = fcm.merge_point_sets((vol[0], vol[1], vol[2], vol[3], vnormals), fcm.sphere_surface(n, radius))
    assert numpy.array_equal(elem_ids, vol[0])
    ctx = engine.Context.get(0)
    b1 = [bspline.spline_basis_1d(n, degree) for _ in range(3)]
    v = numpy.linspace(-1, 1, n + 1)
    nodes = numpy.stack(numpy.meshgrid(v, v, v, indexing='ij'))
    plan = engine.ElemSetPlan(ctx, b1, nodes=nodes, elem_ids=elem_ids, qoff=qoff, qcoords=qcoords, qweights=qweights, renumber=vol[4], nbasis_new=vol[5])
    plan.set_normals(normals)
    pts = plan.evaluate()
    r = numpy.linalg.norm(pts['x'], axis=1)
    k = numpy.pi / radius
    uex = numpy.cos(k * r)
    with numpy.errstate(divide='ignore', invalid='ignore'):
        lap = -k * k * numpy.cos(k * r) - numpy.where(r > 0, 2 * k * numpy.sin(k * r) / r, 2 * k * k)
    plan.set_coefficient('matrix', 0, numpy.where(on_surface, 0., 1.))      # stiffness: volume points only
    plan.set_coefficient('matrix', 1, numpy.where(on_surface, beta, 0.))    # penalty mass: surface points only
    plan.set_coefficient('vector', 0, numpy.where(on_surface, beta * -1., -lap))
    A = ctx.device_alloc(8 * plan.nnz)
    f = ctx.device_alloc(8 * plan.ndofs)
    A.zero()
    f.zero()
    plan.assemble_device([engine.form_stiffness(3), engine.form_mass(3)], [engine.form_load(3)], [A, A], [f])
    u = matrix.DeviceMatrix(plan, A).solve(f.to_host(), rtol=rtol, maxiter=20000)
    uh = plan.evaluate([u], x=False, weights=False)['values'][:, 0, 0]
    w = pts['weights']
    vpts = ~on_surface
    err = float(numpy.sqrt((w[vpts] * (uh[vpts] - uex[vpts]) ** 2).sum() / (w[vpts] * uex[vpts] ** 2).sum()))
    berr = float(numpy.sqrt((w[on_surface] * (uh[on_surface] + 1.) ** 2).sum() / w[on_surface].sum()))
    return dict(n=n, degree=degree, depth=depth, ndofs=plan.ndofs, surface_points=int(on_surface.sum()), surface_area=float(w[on_surface].sum()),
                exact_area=4 * numpy.pi * radius ** 2, cg_iterations=plan.last_cg[0], relative_l2_error=err, boundary_rms_error=berr)


def main(n=32, degree=2, depth=3, radius=.8, rtol=1e-10):
    t0 = time.perf_counter()
    elem_ids, qoff, qcoords, qweights, renumber, nbasis = fcm.octree_ball(n, degree, depth, radius=radius)
    t1 = time.perf_counter()
    ctx = engine.Context.get(0)
    b1 = [bspline.spline_basis_1d(n, degree) for _ in range(3)]
    v = numpy.linspace(-1, 1, n + 1)
    nodes = numpy.stack(numpy.meshgrid(v, v, v, indexing='ij'))
    plan = engine.ElemSetPlan(ctx, b1, nodes=nodes, elem_ids=elem_ids, qoff=qoff, qcoords=qcoords, qweights=qweights, renumber=renumber, nbasis_new=nbasis)
    # the physical coordinates of the points (and their weights) come from the device, the coefficient g goes back as a per-point scalar
    pts = plan.evaluate()
    r = numpy.linalg.norm(pts['x'], axis=1)
    k = numpy.pi / radius
    uex = numpy.cos(k * r)
    with numpy.errstate(divide='ignore', invalid='ignore'):
        lap = -k * k * numpy.cos(k * r) - numpy.where(r > 0, 2 * k * numpy.sin(k * r) / r, 2 * k * k)
    plan.set_coefficient('vector', 0, uex - lap)
    A = ctx.device_alloc(8 * plan.nnz)
    f = ctx.device_alloc(8 * plan.ndofs)
    A.zero()
    f.zero()
    plan.assemble_device([engine.form_stiffness(3) + engine.form_mass(3)], [engine.form_load(3)], [A], [f])
    ctx.synchronize()
    t2 = time.perf_counter()
    u = matrix.DeviceMatrix(plan, A).solve(f.to_host(), rtol=rtol)
    t3 = time.perf_counter()
    uh = plan.evaluate([u], x=False, weights=False)['values'][:, 0, 0]
    w = pts['weights']
    err = float(numpy.sqrt((w * (uh - uex) ** 2).sum() / (w * uex ** 2).sum()))
    return dict(n=n, degree=degree, depth=depth, elements=len(elem_ids), points=int(qoff[-1]), ndofs=plan.ndofs, nnz=plan.nnz, cg_iterations=plan.last_cg[0],
                relative_l2_error=err, volume=float(w.sum()), exact_volume=4 / 3 * numpy.pi * radius ** 3,
                seconds=dict(quadrature=t1 - t0, plan_and_assembly=t2 - t1, solve=t3 - t2))


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    d = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    print(main(n, p, d))
    print(main_dirichlet(min(n, 24), p, d))
