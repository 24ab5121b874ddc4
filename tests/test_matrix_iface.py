'''nutils_b200.matrix.Matrix against the interface of nutils.matrix.Matrix (matrix/_base.py:33-302): arithmetic, transpose,
submatrix, export, constrained solves -- checked against dense numpy; and as a backend object of the UNMODIFIED reference
(``nutils.matrix.backend(obj)``, matrix/__init__.py:20-27) under solver.System.'''

import numpy
import pytest

from tests import util_ref
from nutils_b200 import matrix


def _random(n, m, seed, density=.3):
    rng = numpy.random.RandomState(seed)
    A = rng.rand(n, m) * (rng.rand(n, m) < density)
    rows, cols = numpy.nonzero(A)
    rowptr = numpy.concatenate([[0], numpy.cumsum(numpy.bincount(rows, minlength=n))])
    return matrix.assemble_csr(A[rows, cols], rowptr, cols, m), A


def test_arithmetic_export_transpose():
    A, a = _random(7, 5, 1)
    B, b = _random(7, 5, 2)
    assert numpy.allclose((A + B).export('dense'), a + b) and numpy.allclose((A - B).export('dense'), a - b)
    assert numpy.allclose((A * 2.5).export('dense'), 2.5 * a) and numpy.allclose((3 * A).export('dense'), 3 * a)
    assert numpy.allclose((A / 4).export('dense'), a / 4) and numpy.allclose((-A).export('dense'), -a)
    assert numpy.allclose(A.T.export('dense'), a.T) and A.T.shape == (5, 7)
    data, (row, col) = A.export('coo')
    assert numpy.array_equal(a[row, col], data)
    data, indices, indptr = (A + B).export('csr')
    assert indptr.dtype == numpy.int64 and indices.dtype == numpy.int64
    for r in range(7):
        assert (numpy.diff(indices[indptr[r]:indptr[r + 1]]) > 0).all()
    x = numpy.arange(5.)
    assert numpy.allclose(A @ x, a @ x) and numpy.allclose(A @ numpy.stack([x, 2 * x], 1), a @ numpy.stack([x, 2 * x], 1))
    with pytest.raises(matrix.MatrixError):
        A @ numpy.arange(4.)
    with pytest.raises(matrix.MatrixError):
        A + A.T
    assert A.size == 35 and numpy.array_equal(A.rowsupp(), abs(a).sum(1) > 0)


def test_submatrix():
    A, a = _random(8, 8, 3)
    rows = numpy.array([True, False] * 4)
    cols = numpy.array([0, 3, 4])
    S = A.submatrix(rows, cols)
    assert numpy.allclose(S.export('dense'), a[rows][:, cols])
    assert A.submatrix(rows, cols) is S                      # cached (matrix/_base.py:277-282)
    assert A.submatrix(numpy.ones(8, bool), numpy.ones(8, bool)) is A


@pytest.mark.parametrize('solver', ['direct', 'arnoldi', 'gmres'])
def test_solve_constraints(solver):
    rng = numpy.random.RandomState(4)
    n = 12
    a = rng.rand(n, n) * (rng.rand(n, n) < .4)
    a = a + a.T + n * numpy.eye(n)
    rows, cols = numpy.nonzero(a)
    A = matrix.assemble_csr(a[rows, cols], numpy.concatenate([[0], numpy.cumsum(numpy.bincount(rows, minlength=n))]), cols, n)
    b = rng.rand(n)
    kw = dict(atol=1e-10) if solver == 'gmres' else {}
    assert numpy.allclose(A.solve(b, solver=solver, **kw), numpy.linalg.solve(a, b), atol=1e-8)
    cons = numpy.full(n, numpy.nan)
    cons[[1, 5]] = [.5, -2.]
    x = A.solve(b, constrain=cons, solver=solver, **kw)
    free = numpy.isnan(cons)
    assert numpy.array_equal(x[~free], cons[~free])
    assert numpy.allclose((a @ x - b)[free], 0, atol=1e-8)
    mask = ~free
    lhs0 = rng.rand(n)
    y = A.solve(b, constrain=mask, lhs0=lhs0, solver=solver, symmetric=True, **kw)
    assert numpy.array_equal(y[mask], lhs0[mask]) and numpy.allclose((a @ y - b)[free], 0, atol=1e-8)
    assert numpy.allclose(A.diagonal(), numpy.diag(a))


@pytest.mark.skipif(not util_ref.have_reference(), reason='baseline/_ref not installed')
def test_backend_object_of_the_reference():
    'solver.System of the unmodified reference with this module as matrix backend: examples/laplace.py reproduces its numbers'
    nutils = util_ref.reference()
    import laplace
    from nutils import matrix as refmatrix
    cons0, u0, err0 = laplace.main(nelems=5)
    with refmatrix.backend(matrix):
        cons1, u1, err1 = laplace.main(nelems=5)
    assert numpy.nanmax(abs(cons0 - cons1)) <= 1e-12 and abs(u0 - u1).max() <= 1e-10 and abs(err0 - err1) <= 1e-10


@pytest.mark.gpu
def test_gpu_mirror_matmul_and_cg():
    'the device mirror of a host Matrix: b2_pattern_create_csr + b2_spmv_device / b2_cg_device'
    from nutils_b200 import engine
    rng = numpy.random.RandomState(7)
    n = 300
    a = rng.rand(n, n) * (rng.rand(n, n) < .05)
    a = a + a.T + n * .2 * numpy.eye(n)
    rows, cols = numpy.nonzero(a)
    A = matrix.assemble_csr(a[rows, cols], numpy.concatenate([[0], numpy.cumsum(numpy.bincount(rows, minlength=n))]), cols, n)
    n0 = engine.Context.get(0).launch_count
    x = rng.rand(n)
    assert numpy.allclose(A @ x, a @ x, rtol=1e-13)
    assert engine.Context.get(0).launch_count > n0
    b = rng.rand(n)
    cons = numpy.full(n, numpy.nan)
    cons[::7] = 1.
    y = A.solve(b, constrain=cons, solver='cg', rtol=1e-12)
    free = numpy.isnan(cons)
    assert numpy.array_equal(y[~free], cons[~free]) and abs((a @ y - b)[free]).max() <= 1e-9
    R, r = _random(40, 25, 9)
    v = rng.rand(25)
    assert numpy.allclose(R @ v, r @ v, rtol=1e-13)   # rectangular
    assert numpy.allclose(A.todevice().diagonal(), numpy.diag(a))
