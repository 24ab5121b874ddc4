'''BASELINE.json configs[0]: examples/laplace.py of the reference, assembled AND solved on the B200 path through the
Nutils-style API: volume stiffness (owner-computes kernel / coverage path), Neumann load on the right boundary, Dirichlet
data by L2 projection on the left and top boundaries (boundary mass matrices and loads with coefficient functions:
element-set kernel with the surface measure), constrained solve with the device-resident matrix.  Known answers: the L2
errors asserted by the reference's own unit tests (examples/laplace.py:113-137: 1.63e-3 for nelems=4 std p=1, 8.04e-5 for
nelems=4 spline p=2) and the values the unmodified reference returns here for nelems = 8 and 32 (4.0141e-4, 2.49609e-5).'''

import numpy
import pytest
from nutils_b200 import mesh, function, solver

pytestmark = pytest.mark.gpu


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
    # L2 error against u = sin(x_0) cosh(x_1):  u'Mu - 2 u'b + int uex^2
    M, b = domain.sample('gauss', qd).integrate_sparse([function.outer(basis) * J, basis * (numpy.sin(x0) * numpy.cosh(x1)) * J])
    # the reference integrates the error with the SAME rule (degree*2), so int uex^2 is taken with that rule too
    from nutils_b200 import points
    gx, gw = points.gauss1(qd)
    xs = ((numpy.arange(nelems)[:, None] + gx[None, :]) / nelems).ravel()
    ws = numpy.tile(gw, nelems) / nelems
    uu = (ws * numpy.sin(xs) ** 2).sum() * (ws * numpy.cosh(xs) ** 2).sum()
    err2 = u @ (M @ u) - 2 * u @ b + uu
    return cons, u, err2


@pytest.mark.parametrize('nelems,btype,degree,places,expect', [(4, 'std', 1, 5, 1.63e-3), (4, 'spline', 2, 7, 8.04e-5)])
def test_reference_unit_test_values(nelems, btype, degree, places, expect):
    cons, u, err2 = solve_laplace(nelems, btype, degree)
    assert numpy.isnan(cons).sum() == (nelems + degree - 1 if btype == 'spline' else nelems * degree) ** 2
    assert round(abs(numpy.sqrt(err2) - expect), places) == 0   # unittest.assertAlmostEqual(err, expect, places=places)


@pytest.mark.parametrize('nelems,expect', [(8, 4.0141303937753854e-4), (32, 2.4960897609571972e-05)])
def test_config0_l2_error(nelems, expect):
    cons, u, err2 = solve_laplace(nelems, 'std', 1)
    assert abs(numpy.sqrt(err2) - expect) < 1e-6 * expect
