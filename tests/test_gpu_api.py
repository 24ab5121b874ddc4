'''The Nutils-style API (mesh / topology / function / sample) on the GPU against the golden vectors.'''

import numpy
import pytest

from tests import util
from nutils_b200 import mesh, function, matrix

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(g):
    nelems = tuple(int(n) for n in g['nelems'])
    topo, geom = mesh.rectilinear(nelems)
    geom = topo.nodal_geometry(g['nodes'])
    return topo, geom


@pytest.mark.parametrize('name', ['hex_p2_warp', 'quad_p3', 'line_p2', 'laplace2d_p1_32', 'hex_p2_std'])
def test_scalar_forms(name):
    g = util.load_golden(name)
    topo, geom = _setup(g)
    basis = topo.basis(str(g['btype']), degree=int(g['degree']))
    grad = basis.grad(geom)
    J = function.J(geom)
    qd = int(g['qdegree'])
    K = topo.integral((grad[:, None, :] * grad[None, :, :]).sum(-1) * J, degree=qd)
    M = topo.integral(basis[:, None] * basis[None, :] * J, degree=qd)
    F = topo.integral(basis * J, degree=qd)
    (kv, rp, ci), (mv, mrp, mci), f = function.eval((function.as_csr(K), function.as_csr(M), F))
    assert numpy.array_equal(rp, g['rowptr']) and numpy.array_equal(ci, g['colidx'])
    assert util.relerr(kv, g['K_values']) <= TOL and util.relerr(mv, g['M_values']) <= TOL and util.relerr(f, g['F']) <= TOL
    # Topology.integrate returns dense arrays like the reference (sample.py:160-175)
    if len(basis) <= 400:
        Kd, fd = topo.integrate([(grad[:, None, :] * grad[None, :, :]).sum(-1) * J, basis * J], degree=qd)
        assert Kd.shape == (len(basis),) * 2
        ref = matrix.assemble_csr(g['K_values'], g['rowptr'], g['colidx'], len(basis)).export('dense')
        assert abs(Kd - ref).max() <= TOL * abs(ref).max()
        assert util.relerr(fd, g['F']) <= TOL
    # as_coo
    v, i, j = function.eval(function.as_coo(M))
    assert numpy.array_equal(j, g['colidx']) and (numpy.diff(i) >= 0).all() and util.relerr(v, g['M_values']) <= TOL


@pytest.mark.parametrize('name', ['elast2d_p2_warp', 'elast3d_p1', 'elast3d_p2_warp'])
def test_elasticity(name):
    # the energy of examples/elasticity.py:50-58 written with the array API
    g = util.load_golden(name)
    topo, geom = _setup(g)
    nd = topo.ndims
    u = topo.basis('spline', degree=int(g['degree']), shape=(nd,))
    lm, mu = float(g['lmbda']), float(g['mu'])
    eps = u.symgrad(geom)
    sig = lm * eps.trace(-2, -1)[:, None, None] * numpy.eye(nd) + 2 * mu * eps
    J = function.J(geom)
    q = -numpy.eye(nd)[nd - 1]
    jac = topo.integral(2 * (eps[:, None] * sig[None, :]).sum((-2, -1)) * J, degree=int(g['qdegree']))  # d2/du2 of eps:sig
    res = topo.integral(-(u * q).sum(-1) * J, degree=int(g['qdegree']))
    (kv, rp, ci), f = function.eval((function.as_csr(jac), res))
    assert numpy.array_equal(rp, g['rowptr']) and numpy.array_equal(ci, g['colidx'])
    assert util.relerr(kv, g['K_values']) <= TOL
    assert util.rowsum_relerr(kv, g['K_values'], rp) <= TOL
    assert util.relerr(f, g['F']) <= TOL


def test_integrate_sparse_and_matrix():
    topo, geom = mesh.rectilinear([4, 3])
    basis = topo.basis('spline', degree=2)
    M, f = topo.sample('gauss', 4).integrate_sparse([function.outer(basis) * function.J(geom), basis * function.J(geom)])
    assert isinstance(M, matrix.Matrix) and M.shape == (len(basis),) * 2
    ones = numpy.ones(len(basis))
    numpy.testing.assert_allclose(M @ ones, f, rtol=1e-12)
    assert abs(f.sum() - 12.) < 1e-12  # volume of the 4x3 integer grid (tests/test_sample.py:476-482)
    data, indices, indptr = M.export('csr')
    assert len(data) == len(indices) == indptr[-1]


def test_unsupported_raises():
    topo, geom = mesh.rectilinear([3, 3])
    basis = topo.basis('spline', degree=1)
    with pytest.raises(NotImplementedError):
        topo.integral(basis * basis * function.J(geom), degree=2)        # nonlinear in the basis
    with pytest.raises(NotImplementedError):
        topo.integral(basis, degree=2)                                   # no jacobian
    with pytest.raises(NotImplementedError):
        topo.basis('discont', degree=1)
