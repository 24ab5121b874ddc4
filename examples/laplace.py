#!/usr/bin/env python
'''Laplace's equation on the unit square -- the problem of the reference's examples/laplace.py (BASELINE.json configs[0]),
written against the nutils_b200 API and assembled and solved on the GPU:

    int grad v . grad u dV = int_right v cos(1) cosh(x_1) dS,   u = 0 on the left,  u = cosh(1) sin(x_0) on the top,

with the exact solution u = sin(x_0) cosh(x_1).  The reference spells the forms with its expression namespace
('∇_i(v) ∇_i(u) dV' @ ns) and differentiates the functional symbolically; that machinery is outside the scope of this
repository, so the bilinear and linear forms are written out with basis arrays -- everything from
`domain.integral` on is the accelerated path.

    python examples/laplace.py [nelems] [std|spline] [degree]
'''

import os
import sys
import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nutils_b200 import mesh, function, solver  # noqa: E402


def main(nelems=32, btype='std', degree=1):
    'returns (constraints, solution vector, L2 error) like the reference example'
    cons, u, err2 = solve_laplace(nelems, btype, degree)
    return cons, u, float(numpy.sqrt(err2))


def solve_laplace(nelems, btype, degree):
    domain, geom = mesh.unitsquare(nelems, 'square')
    basis = domain.basis(btype, degree=degree)
    x0, x1 = geom
    J = function.J(geom)
    qd = degree * 2
    g = basis.grad(geom)
    # residual: int grad v . grad u dV - int_right v cos(1) cosh(x_1) dS
    K = domain.integral((g[:, None, :] * g[None, :, :]).sum(-1) * J, degree=qd)
    f = domain.boundary['right'].integral(basis * (numpy.cos(1) * numpy.cosh(x1)) * J, degree=qd)
    # constraints: minimise int_left u^2 dS + int_top (u - cosh(1) sin(x_0))^2 dS over the boundary dofs (solve_constraints, solver.py:562-612)
    cons = solver.solve_constraints([(domain.boundary['left'].integral(function.outer(basis) * J, degree=qd), None),
                                     (domain.boundary['top'].integral(function.outer(basis) * J, degree=qd),
                                      domain.boundary['top'].integral(basis * (numpy.cosh(1) * numpy.sin(x0)) * J, degree=qd))], droptol=1e-15)
    u = solver.LinearSystem(K, [f]).solve(constrain=cons, rtol=1e-13)
    # L2 error against u = sin(x_0) cosh(x_1), integrated with the same rule as the reference does (degree*2):
    # Sample.eval of the discrete solution at the Gauss points, then a weighted sum on the host
    pts = domain.sample('gauss', qd).eval_fields(basis, geom, [u])
    uex = numpy.sin(pts['x'][:, 0]) * numpy.cosh(pts['x'][:, 1])
    err2 = (pts['weights'] * (pts['values'][:, 0, 0] - uex) ** 2).sum()
    return cons, u, err2


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    bt = sys.argv[2] if len(sys.argv) > 2 else 'std'
    p = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    cons, u, err = main(n, bt, p)
    print('{} dofs, {} constrained, L2 error: {:.3e}'.format(len(u), int((~numpy.isnan(cons)).sum()), err))
