#!/usr/bin/env python
'''3-D Poisson problem assembled and solved without leaving HBM (BASELINE.json configs[1] followed by the solve):

    -div grad u = 1 on a warped n^3 grid,  u = 0 on the left face,  u = 1 on the right face,

degree-p splines.  K and f are assembled by the owner-computes kernel into device memory (Sample.integrate_device),
the Dirichlet data come from boundary projections (element-set kernel), the system is solved by Jacobi-PCG on the
device (matrix.DeviceMatrix.solve).  Nothing but the solution vector crosses PCIe.

    python examples/poisson3d.py [n] [degree]
'''

import os
import sys
import time
import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nutils_b200 import mesh, function, solver  # noqa: E402


def main(n=64, degree=2, warp=.2, rtol=1e-10):
    rng = numpy.random.RandomState(0)
    topo, geom = mesh.rectilinear([numpy.linspace(0, 1, n + 1)] * 3)
    nodes = geom.nodes.copy()
    nodes[:, 1:-1, 1:-1, 1:-1] += warp / n * (rng.rand(3, n - 1, n - 1, n - 1) - .5)   # interior nodes only: the faces stay planar
    geom = topo.nodal_geometry(nodes)
    basis = topo.basis('spline', degree=degree)
    J = function.J(geom)
    g = basis.grad(geom)
    qd = 2 * degree
    t0 = time.perf_counter()
    K = topo.integral((g[:, None, :] * g[None, :, :]).sum(-1) * J, degree=qd)
    f = topo.integral(basis * J, degree=qd)
    mass = function.outer(basis) * J
    cons = solver.solve_constraints([(topo.boundary['left'].integral(mass, degree=qd), None),
                                     (topo.boundary['right'].integral(mass, degree=qd), topo.boundary['right'].integral(basis * J, degree=qd))])
    system = solver.LinearSystem(K, [f])
    Kd, rhs = system.assemble()
    t1 = time.perf_counter()
    u = system.solve(constrain=cons, rtol=rtol)
    t2 = time.perf_counter()
    its, res = Kd.plan.last_cg
    return dict(ndofs=len(basis), nnz=Kd.plan.nnz, constrained=int((~numpy.isnan(cons)).sum()), assemble_s=t1 - t0, solve_s=t2 - t1, iterations=its,
                residual=res, u_min=float(u.min()), u_max=float(u.max()), u_mean=float(u.mean()))


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    print(main(n, p))
