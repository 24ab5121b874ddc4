#!/usr/bin/env python
'''Plane strain plate under gravitational pull -- the problem of the reference's examples/elasticity.py (the 2-D origin
of BASELINE.json configs[2]), written against the nutils_b200 API and assembled and solved on the GPU:

    minimise  int (eps_ij sigma_ij - u_i q_i) dV,   sigma = lambda tr(eps) I + 2 mu eps,   q = -e_y,   u = 0 on the top.

The reference differentiates the energy symbolically; here its second derivative (the bilinear form 2 eps(v):sigma(u)) and
first derivative at u = 0 (the load -v.q) are written out.  The traction post-processing of the reference example needs a
second field on the boundary and is not part of this script.

    python examples/elasticity.py [nelems] [std|spline] [degree] [poisson]
'''

import os
import sys
import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nutils_b200 import mesh, function, solver  # noqa: E402


def main(nelems=24, btype='std', degree=2, poisson=.3):
    'returns (constraints [nbasis, 2], displacement coefficients [nbasis, 2]) like cons["u"], args["u"] of the reference example'
    domain, geom = mesh.unitsquare(nelems, 'square')
    u = domain.basis(btype, degree=degree, shape=(2,))       # vector-valued basis, dof = ibasis*2 + component (function.field numbering)
    lmbda, mu = 1., .5 / poisson - 1
    J = function.J(geom)
    qd = degree * 2
    eps = u.symgrad(geom)
    tr = eps.trace(-2, -1)
    # d2/du2 of int eps_ij sigma_ij: 2 (lambda tr(eps_v) tr(eps_u) + 2 mu eps_v : eps_u)
    A = domain.integral(2 * (lmbda * tr[:, None] * tr[None, :] + 2 * mu * (eps[:, None] * eps[None, :]).sum((-2, -1))) * J, degree=qd)
    q = numpy.array([0., -1.])
    b = domain.integral((u * q).sum(-1) * J, degree=qd)        # stationarity: A u = int v.q
    top = domain.boundary['top']
    cons = solver.solve_constraints([(top.integral((u[:, None, :] * u[None, :, :]).sum(-1) * J, degree=qd), None)], droptol=1e-15)
    sol = solver.LinearSystem(A, [b]).solve(constrain=cons, rtol=1e-13)
    return cons.reshape(-1, 2), sol.reshape(-1, 2)


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    bt = sys.argv[2] if len(sys.argv) > 2 else 'std'
    p = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    nu = float(sys.argv[4]) if len(sys.argv) > 4 else .3
    cons, u = main(n, bt, p, nu)
    print('{} dofs, {} constrained, max downward displacement {:.6f}'.format(u.size, int((~numpy.isnan(cons)).sum()), float(-u[:, 1].min())))
