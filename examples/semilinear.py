#!/usr/bin/env python
'''Newton's method for a semilinear problem with solution-dependent coefficients, every step on the GPU:

    -div grad u + u^3 = g  on the unit cube,  du/dn = 0,   manufactured solution u = cos(pi x) cos(pi y) cos(pi z).

This is the pattern of SURVEY.md 8f.2 (``System.assemble_jacobian_residual`` of a nonlinear functional, solver.py:357-425:
coefficients that depend on the current iterate are re-evaluated in every Newton step) spelled with this repository's pieces:

    jacobian = int grad N_i . grad N_j + 3 u_h^2 N_i N_j     (two forms; the coefficient 3 u_h^2 is evaluated INSIDE the kernel from
                                                               the device-resident iterate: b2_elemset_set_coefficient_field)
    residual = int grad N_i . grad u + (u_h^3 - g) N_i  =  K u + int (u_h^3 - g) N_i   (K u by SpMV on the device, the rest two load forms)
    du       = -J^-1 r  by Jacobi-PCG on the device (b2_cg_device on device vectors)

The reference derives jacobian and residual symbolically from one functional; here they are written out.

    python examples/semilinear.py [n] [degree]
'''

import os
import sys
import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nutils_b200 import bspline, points, engine, matrix  # noqa: E402


def main(n=16, degree=2, tol=1e-10, maxiter=12, fused=True):
    '''fused=True: the iterate lives in HBM and its coefficient functions 3 u_h^2 and u_h^3 are evaluated INSIDE the assembly kernel
    (b2_elemset_set_coefficient_field) -- no host round trip inside a Newton step except the residual norm (one scalar).
    fused=False: the round-1 route, evaluation of the iterate at the points (b2_evaluate_elemset_device), coefficients formed on the
    host and attached per point.'''
    import time
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
    history, step_seconds = [], []
    if fused:
        import torch
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        dev = torch.device('cuda', ctx.device)
        u = torch.zeros(plan.ndofs, dtype=torch.float64, device=dev)
        r = torch.empty_like(u)
        nl = torch.empty_like(u)
        du = torch.empty_like(u)
        J = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
        plan.set_coefficient('vector', 0, -g)                        # load form 0: -g N_i (pointwise data, uploaded once)
        plan.set_coefficient_field('vector', 1, u, power=3)          # load form 1: u_h^3 N_i, u_h from the device vector u
        plan.set_coefficient_field('matrix', 1, u, power=2, scale=3.)  # mass form: 3 u_h^2 N_i N_j
        for it in range(maxiter):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            J.zero_()
            nl.zero_()
            plan.assemble_device([K, M], [L, L], [J, J], [nl, nl])    # jacobian = K + 3 u^2 M and the nonlinear part of the residual, one launch
            plan.spmv_device(Kbuf, u, r)
            r += nl
            rnorm = float(torch.linalg.norm(r))                       # the one scalar that crosses PCIe per step
            history.append(rnorm)
            if rnorm <= tol * max(history[0], 1.):
                break
            du.zero_()
            plan.cg_device(J, r, du, rtol=1e-12)
            u -= du
            torch.cuda.synchronize()
            step_seconds.append(time.perf_counter() - t0)
        plan.set_coefficient_field('vector', 1, None)
        plan.set_coefficient_field('matrix', 1, None)
        plan.set_coefficient('vector', 0, None)
        ctx.set_stream(None)
        u = u.cpu().numpy()
    else:
        u = numpy.zeros(plan.ndofs)
        for it in range(maxiter):
            t0 = time.perf_counter()
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
            step_seconds.append(time.perf_counter() - t0)
        plan.set_coefficient('vector', 0, None)
        plan.set_coefficient('matrix', 1, None)
    uq = plan.evaluate([u], x=False, weights=False)['values'][:, 0, 0]
    err = float(numpy.sqrt((w * (uq - uex) ** 2).sum()))
    return dict(ndofs=plan.ndofs, fused=fused, newton_iterations=len(history) - 1, residual_history=history, l2_error=err,
                seconds_per_newton_step=float(numpy.min(step_seconds)) if step_seconds else None)   # the first step carries one-off initialisations


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    print(main(n, p, fused=True))
    print(main(n, p, fused=False))
