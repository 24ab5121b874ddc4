'''Parity of the CUDA path (through the C ABI) with the golden vectors of the reference and with the
C oracle on seeded inputs.  Bar (BASELINE.json north_star): CSR pattern bit-exact; values:
relative Frobenius-norm error and relative row-sum error <= 1e-12; rhs: relative 2-norm error <= 1e-12.'''

import numpy
import pytest

from tests import util
from oracle import c_oracle, fem_oracle
from nutils_b200 import engine, bspline, points

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope='module')
def ctx():
    return engine.Context.get(0)


def _plan(ctx, prob):
    bases = [bspline.Basis1D(prob.degree[d], prob.nelems[d], prob.coeffs[d], prob.setidx[d], prob.start[d], prob.ndofs_d[d]) for d in range(prob.ndims)]
    return engine.Plan(ctx, bases, list(zip(prob.qpts, prob.qwts)), prob.nodes, ncomp=prob.ncomp)


def _forms(g, nd):
    if str(g['kind']) == 'scalar':
        return [engine.form_stiffness(nd), engine.form_mass(nd)], [engine.form_load(nd)], [g['K_values'], g['M_values']]
    C = numpy.zeros((nd, nd + 1))
    C[nd - 1, 0] = 1.
    return [engine.form_elasticity(nd, float(g['lmbda']), float(g['mu']), scale=2.)], [C], [g['K_values']]


@pytest.mark.parametrize('kernel', [0, 1])
@pytest.mark.parametrize('name', util.golden_names())
def test_golden(ctx, name, kernel):
    g = util.load_golden(name)
    prob = util.problem_from_golden(g)
    plan = _plan(ctx, prob)
    ctx.set_option('kernel', kernel)
    try:
        rowptr, colidx = plan.csr_pattern()
        assert rowptr.dtype == numpy.int64 and colidx.dtype == numpy.int64
        assert numpy.array_equal(rowptr, g['rowptr'])
        assert numpy.array_equal(colidx, g['colidx'])
        Ds, Cs, expect = _forms(g, prob.ndims)
        vals, rhs = plan.assemble_host(Ds, Cs)
        for v, ref in zip(vals, expect):
            if len(ref):
                assert util.relerr(v, ref) <= TOL
                assert util.rowsum_relerr(v, ref, rowptr) <= TOL
        assert util.relerr(rhs[0], g['F']) <= TOL
    finally:
        ctx.set_option('kernel', 0)


def _random_problem(seed, nelems, degree, ncomp=1, btype='spline', warp=.25, qdegree=None):
    rng = numpy.random.RandomState(seed)
    nd = len(nelems)
    b1 = util.bases_1d(nelems, degree, btype)
    rules = points.tensor_gauss(nd, 2 * degree if qdegree is None else qdegree)
    verts = [numpy.sort(rng.rand(n + 1)) + numpy.arange(n + 1) for n in nelems]
    X = numpy.stack(numpy.meshgrid(*verts, indexing='ij'))
    X = X + warp * (rng.rand(*X.shape) - .5)
    return fem_oracle.Problem(nelems, [degree] * nd, [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1],
                              [b.ndofs for b in b1], [r[0] for r in rules], [r[1] for r in rules], X, ncomp=ncomp)


CASES = [
    dict(nelems=(40,), degree=1), dict(nelems=(33,), degree=4), dict(nelems=(1,), degree=2),
    dict(nelems=(17, 9), degree=1), dict(nelems=(12, 15), degree=2), dict(nelems=(7, 6), degree=3), dict(nelems=(5, 4), degree=4),
    dict(nelems=(1, 1), degree=2), dict(nelems=(2, 1, 1), degree=1),
    dict(nelems=(9, 8, 7), degree=1), dict(nelems=(10, 9, 11), degree=2), dict(nelems=(5, 4, 6), degree=3), dict(nelems=(3, 2, 3), degree=4),
    dict(nelems=(6, 5, 4), degree=2, btype='std'), dict(nelems=(6, 5, 4), degree=2, qdegree=6),
    dict(nelems=(9, 7), degree=2, ncomp=2), dict(nelems=(5, 4, 5), degree=2, ncomp=3), dict(nelems=(4, 3, 3), degree=3, ncomp=3),
    dict(nelems=(1, 1, 1), degree=2),
]


@pytest.mark.parametrize('kernel', [0, 1])
@pytest.mark.parametrize('case', CASES, ids=lambda c: 'x'.join(map(str, c['nelems'])) + 'p{}c{}'.format(c['degree'], c.get('ncomp', 1)) + c.get('btype', '') + str(c.get('qdegree', '')))
def test_oracle_seeded(ctx, case, kernel):
    prob = _random_problem(seed=hash(str(case)) % 2**31, **case)
    nd, nc = prob.ndims, prob.ncomp
    rng = numpy.random.RandomState(5)
    if nc == 1:
        Ds = [engine.form_stiffness(nd), engine.form_mass(nd), rng.rand(1, nd + 1, 1, nd + 1)]
        Cs = [engine.form_load(nd), rng.rand(1, nd + 1)]
    else:
        Ds = [engine.form_elasticity(nd, 1.3, .7) if nc == nd else engine.form_stiffness(nd, nc), rng.rand(nc, nd + 1, nc, nd + 1)]
        Cs = [rng.rand(nc, nd + 1)]
    mats, vecs = c_oracle.assemble(prob, [('generic', D) for D in Ds], [('generic', C) for C in Cs])
    plan = _plan(ctx, prob)
    ctx.set_option('kernel', kernel)
    try:
        rowptr, colidx = plan.csr_pattern()
        assert numpy.array_equal(rowptr, mats[0][1])
        assert numpy.array_equal(colidx, mats[0][2])
        vals, rhs = plan.assemble_host(Ds, Cs)
    finally:
        ctx.set_option('kernel', 0)
    for v, (ref, _, _) in zip(vals, mats):
        assert util.relerr(v, ref) <= TOL
        assert util.rowsum_relerr(v, ref, rowptr) <= TOL
    for r, ref in zip(rhs, vecs):
        assert util.relerr(r, ref) <= TOL


def test_element_ranges_accumulate(ctx):
    # slabs of elements assembled separately into the same arrays give the full matrix (multi-GPU building block)
    prob = _random_problem(3, (6, 5, 4), 2)
    plan = _plan(ctx, prob)
    D, C = [engine.form_stiffness(3)], [engine.form_load(3)]
    full_v, full_r = plan.assemble_host(D, C)
    n = prob.ntotal
    parts = [plan.assemble_host(D, C, elem_range=r) for r in ((0, n // 3), (n // 3, n // 3), (n // 3, n))]
    assert util.relerr(sum(p[0][0] for p in parts), full_v[0]) <= 1e-14
    assert util.relerr(sum(p[1][0] for p in parts), full_r[0]) <= 1e-14


def test_properties_large(ctx):
    # size-independent properties at a size the oracle would not finish quickly: 48^3, p=2
    n = 48
    b1 = util.bases_1d((n,) * 3, 2, 'spline')
    rules = points.tensor_gauss(3, 4)
    rng = numpy.random.RandomState(0)
    X = numpy.stack(numpy.meshgrid(*[numpy.linspace(0, 1, n + 1)] * 3, indexing='ij'))
    X = X + .2 / n * (rng.rand(*X.shape) - .5)
    X[:, 0] = numpy.stack(numpy.meshgrid(*[numpy.linspace(0, 1, n + 1)] * 2, indexing='ij'))[[0, 0, 1]] * [[[0.]], [[1.]], [[1.]]]  # flat x=0 face
    X[0, 0] = 0.
    plan = engine.Plan(ctx, b1, rules, X)
    (K, M), (f,) = plan.assemble_host([engine.form_stiffness(3), engine.form_mass(3)], [engine.form_load(3)])
    rowptr, colidx = plan.csr_pattern()
    assert plan.nnz == (5 * (n + 2) - 6) ** 3
    rows = numpy.repeat(numpy.arange(plan.ndofs), numpy.diff(rowptr))
    # partition of unity: K 1 = 0, sum M = volume = sum f
    Ksum = numpy.bincount(rows, weights=K, minlength=plan.ndofs)
    assert abs(Ksum).max() <= 1e-12 * abs(K).max() * 125
    assert abs(M.sum() - f.sum()) <= 1e-12 * f.sum()
    Msum = numpy.bincount(rows, weights=M, minlength=plan.ndofs)
    assert util.relerr(Msum, f) <= 1e-12
    # symmetry up to rounding: K[i,j] == K[j,i]
    import scipy.sparse
    A = scipy.sparse.csr_matrix((K, colidx, rowptr), shape=(plan.ndofs,) * 2)
    assert abs(A - A.T).max() <= 1e-12 * abs(K).max()


def test_errors(ctx):
    from nutils_b200._lib import B200Error
    prob = _random_problem(1, (3, 3), 2)
    plan = _plan(ctx, prob)
    with pytest.raises(B200Error):
        plan.assemble_host([engine.form_mass(2)], [], elem_range=(5, 100))
    with pytest.raises(ValueError):
        engine.Plan(ctx, plan.bases, plan.rules, prob.nodes[:, :-1])
