'''Device-resident matrix operations (b2_spmv_device, b2_diagonal_device, b2_cg_device) through the C ABI: products and
diagonals against scipy on the exported CSR triplet, constrained CG solves against a direct sparse solve with the
semantics of nutils.matrix.Matrix.solve (matrix/_base.py:100-173).'''

import numpy
import pytest
import scipy.sparse
import scipy.sparse.linalg

from tests import util
from nutils_b200 import engine, mesh, function, _lib

pytestmark = pytest.mark.gpu


def _poisson(n, degree, warp=.2, seed=0, ncomp=1):
    rng = numpy.random.RandomState(seed)
    topo, geom = mesh.rectilinear([numpy.linspace(0, 1, k + 1) for k in n])
    geom = topo.nodal_geometry(geom.nodes + warp / max(n) * (rng.rand(*geom.nodes.shape) - .5))
    basis = topo.basis('spline', degree=degree) if ncomp == 1 else topo.basis('spline', degree=degree, shape=(ncomp,))
    return topo, geom, basis


@pytest.mark.parametrize('plain', [0, 1], ids=['fast', 'plain'])
@pytest.mark.parametrize('n,degree', [((9,), 2), ((7, 6), 3), ((6, 5, 7), 2), ((4, 3, 3), 4), ((5, 4, 6), 1), ((3, 2, 9), 3), ((1, 1, 1), 2), ((40, 3), 1), ((13, 9, 11), 2)])
def test_spmv_and_diagonal(n, degree, plain):
    engine.Context.get(0).set_option('spmv_plain', plain)
    topo, geom, basis = _poisson(n, degree)
    g = basis.grad(geom)
    smp = topo.sample('gauss', 2 * degree)
    K, M = smp.integrate_device([(g[:, None, :] * g[None, :, :]).sum(-1) * function.J(geom), function.outer(basis) * function.J(geom)])
    rng = numpy.random.RandomState(1)
    x = rng.rand(K.shape[1])
    for A in K, M:
        data, indices, indptr = A.export('csr')
        ref = scipy.sparse.csr_matrix((data, indices, indptr), shape=A.shape)
        assert util.relerr(A @ x, ref @ x) <= 1e-13
        assert util.relerr(A.diagonal(), ref.diagonal()) <= 1e-15
    engine.Context.get(0).set_option('spmv_plain', 0)


def test_spmv_vector_valued():
    topo, geom, basis = _poisson((4, 5, 3), 2, ncomp=3)
    lm, mu = 1., .7
    smp = topo.sample('gauss', 4)
    plan = smp.plan(basis, geom)
    vals = plan.ctx.device_alloc(8 * plan.nnz)
    plan.assemble_rows_device([engine.form_elasticity(3, lm, mu)], [], [vals], [])
    rowptr, colidx = plan.csr_pattern()
    ref = scipy.sparse.csr_matrix((vals.to_host(), colidx, rowptr), shape=(plan.ndofs,) * 2)
    x = numpy.random.RandomState(2).rand(plan.ndofs)
    xd, yd = plan.ctx.device_alloc(8 * plan.ndofs), plan.ctx.device_alloc(8 * plan.ndofs)
    xd.from_host(x)
    plan.spmv_device(vals, xd, yd)
    assert util.relerr(yd.to_host(), ref @ x) <= 1e-13
    plan.diagonal_device(vals, yd)
    assert util.relerr(yd.to_host(), ref.diagonal()) <= 1e-15


@pytest.mark.parametrize('n,degree', [((12, 10), 2), ((6, 7, 5), 2), ((5, 5, 4), 3)])
def test_cg_solve_with_constraints(n, degree):
    # Poisson with Dirichlet data on the dofs of the first and last plane along x: K x = f on the free dofs
    topo, geom, basis = _poisson(n, degree)
    g = basis.grad(geom)
    K, f = topo.sample('gauss', 2 * degree).integrate_device([(g[:, None, :] * g[None, :, :]).sum(-1) * function.J(geom), basis * function.J(geom)])
    nd = [b.ndofs for b in basis.bases1d]
    idx = numpy.arange(len(basis)).reshape(nd)
    cons = numpy.full(len(basis), numpy.nan)
    cons[idx[0].ravel()] = 0.
    cons[idx[-1].ravel()] = 1.5
    x = K.solve(f, constrain=cons, rtol=1e-12)
    data, indices, indptr = K.export('csr')
    A = scipy.sparse.csr_matrix((data, indices, indptr), shape=K.shape)
    free = numpy.isnan(cons)
    lhs = numpy.where(free, 0., cons)
    ref = lhs.copy()
    ref[free] += scipy.sparse.linalg.spsolve(A[free][:, free].tocsc(), (f - A @ lhs)[free])
    assert util.relerr(x, ref) <= 1e-9
    assert numpy.array_equal(x[~free], cons[~free])
    # residual criterion of the reference: |A x - b| <= rtol |b_reduced| on the free rows
    assert numpy.linalg.norm((A @ x - f)[free]) <= 2e-12 * numpy.linalg.norm((f - A @ lhs)[free])
    # bool constraints + lhs0
    y = K.solve(f, lhs0=lhs, constrain=~free, rtol=1e-12)
    assert util.relerr(y, x) <= 1e-10


def test_cg_tolerance_not_reached():
    topo, geom, basis = _poisson((8, 8, 8), 2)
    g = basis.grad(geom)
    K, f = topo.sample('gauss', 4).integrate_device([(g[:, None, :] * g[None, :, :] ).sum(-1) * function.J(geom) + function.outer(basis) * function.J(geom), basis * function.J(geom)])
    with pytest.raises(_lib.ToleranceNotReached) as e:
        K.solve(f, rtol=1e-12, maxiter=3)
    assert e.value.best.shape == (len(basis),)


def test_spmv_elemset_pattern():
    # general (materialised) pattern of a trimmed topology
    from tests.test_gpu_elemset import _random_cut, _plan
    prob = _random_cut(3, (5, 4, 4), 2)
    ctx = engine.Context.get(0)
    plan = _plan(ctx, prob)
    vals, _ = plan.assemble_host([engine.form_stiffness(3) + engine.form_mass(3)], [])
    rowptr, colidx = plan.csr_pattern()
    ref = scipy.sparse.csr_matrix((vals[0], colidx, rowptr), shape=(plan.ndofs,) * 2)
    vd = ctx.device_alloc(8 * plan.nnz)
    vd.from_host(vals[0])
    x = numpy.random.RandomState(5).rand(plan.ndofs)
    xd, yd = ctx.device_alloc(8 * plan.ndofs), ctx.device_alloc(8 * plan.ndofs)
    xd.from_host(x)
    plan.spmv_device(vd, xd, yd)
    assert util.relerr(yd.to_host(), ref @ x) <= 1e-13
    plan.diagonal_device(vd, yd)
    assert util.relerr(yd.to_host(), ref.diagonal()) <= 1e-15
    xd.from_host(numpy.zeros(plan.ndofs))
    bd = ctx.device_alloc(8 * plan.ndofs)
    b = ref @ x
    bd.from_host(b)
    plan.cg_device(vd, bd, xd, rtol=1e-8, maxiter=5000)
    assert numpy.linalg.norm(ref @ xd.to_host() - b) <= 1e-7 * numpy.linalg.norm(b)
