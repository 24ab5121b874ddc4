#!/usr/bin/env python
'''Newton's method for a semilinear problem with solution-dependent coefficients, every step on the GPU:

    -div grad u + u^3 = g  on the unit cube,  du/dn = 0,   manufactured solution u = cos(pi x) cos(pi y) cos(pi z).

This is the pattern of SURVEY.md 8f.2 (``System.assemble_jacobian_residual`` of a nonlinear functional, solver.py:357-425:
coefficients that depend on the current iterate are re-evaluated in every Newton step) spelled with this repository's pieces:

    u_q      = Sample.eval of the iterate at the Gauss points                (b2_evaluate_elemset_device)
    jacobian = int grad N_i . grad N_j + 3 u_q^2 N_i N_j                     (two forms, the second with a pointwise coefficient)
    residual = int grad N_i . grad u + (u_q^3 - g) N_i  =  K u + int c_q N_i  (K u by SpMV on the device, the rest a load form)
    du       = -J^-1 r  by Jacobi-PCG on the device

The reference derives jacobian and residual symbolically from one functional; here they are written out.

    python examples/semilinear.py [n] [degree]
'''

import os
import sys
import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nutils_b200 import bspline, points, engine, matrix  # noqa: E402


def main(n=16, degree=2, tol=1e-10, maxiter=12):
    ctx = engine.Context.get(0)
    b1 = [bspline.spline_basis_1d(n, degree) for _ in range(3)]
    v = numpy.linspace(0, 1, n + 1)
    nodes = numpy.stack(numpy.meshgrid(v, v, v, indexing='ij'))
    plan = engine.ElemSetPlan(ctx, b1, nodes=nodes, rules=points.tensor_gauss(3, 2 * degree + 2))   # u^3 raises the degree of the integrand
    pts = plan.evaluate()
    x, w = pts['x'], pts['weights']
    uex = numpy.cos(numpy.pi * x).prod(1)
    g = 3 * numpy.pi ** 2 * uex + uex ** 3
    K, M, L = engine.form_stiffness(3), engine.form_mass(3), engine.form_load(3)
    Kbuf = ctx.device_alloc(8 * plan.nnz)
    Kbuf.zero()
    plan.assemble_device([K], [], [Kbuf], [])                       # the constant part, assembled once
    Kmat = matrix.DeviceMatrix(plan, Kbuf)
    Jbuf = ctx.device_alloc(8 * plan.nnz)
    rbuf = ctx.device_alloc(8 * plan.ndofs)
    u = numpy.zeros(plan.ndofs)
    history = []
    for it in range(maxiter):
        uq = plan.evaluate([u], x=False, weights=False)['values'][:, 0, 0]
        plan.set_coefficient('vector', 0, uq ** 3 - g)
        plan.set_coefficient('matrix', 1, 3 * uq ** 2)
        Jbuf.zero()
        rbuf.zero()
        plan.assemble_device([K, M], [L], [Jbuf, Jbuf], [rbuf])   # both matrix forms accumulate into the same values array
        r = Kmat @ u + rbuf.to_host()
        history.append(float(numpy.linalg.norm(r)))
        if history[-1] <= tol * max(history[0], 1.):
            break
        u = u - matrix.DeviceMatrix(plan, Jbuf).solve(r, rtol=1e-12)
    uq = plan.evaluate([u], x=False, weights=False)['values'][:, 0, 0]
    err = float(numpy.sqrt((w * (uq - uex) ** 2).sum()))
    return dict(ndofs=plan.ndofs, newton_iterations=len(history) - 1, residual_history=history, l2_error=err)


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    print(main(n, p))
