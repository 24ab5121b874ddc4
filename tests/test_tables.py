'''Host-side table producers (bspline.py, points.py): properties, and equality with the reference's
own tables when the reference tree is available (build container only).'''

import os
import sys
import numpy
import pytest

from nutils_b200 import bspline, points

REF = '/root/reference/src'
SHIMS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'shims')


@pytest.mark.parametrize('p', [0, 1, 2, 3, 4])
@pytest.mark.parametrize('cont', [-1, 0])
def test_partition_of_unity(p, cont):
    if p == 0 and cont == 0:
        pytest.skip('degree 0 has no C0 variant')
    b = bspline.spline_basis_1d(6, p, continuity=cont)
    x = numpy.linspace(0, 1, 7)
    tab = b.tabulate(x)
    for e in range(b.nelems):
        s = b.setidx[e]
        numpy.testing.assert_allclose(tab[s, 0].sum(0), 1., atol=1e-14)
        numpy.testing.assert_allclose(tab[s, 1].sum(0), 0., atol=1e-13)
    assert b.start[-1] + p + 1 == b.ndofs


def test_continuity_across_elements():
    b = bspline.spline_basis_1d(5, 3)
    t = b.tabulate(numpy.array([0., 1.]))
    for e in range(b.nelems - 1):
        # function with global dof i: value at right end of element e == value at left end of e+1
        for i in range(b.start[e + 1], b.start[e] + 4):
            vl = t[b.setidx[e], 0, i - b.start[e], 1]
            vr = t[b.setidx[e + 1], 0, i - b.start[e + 1], 0]
            assert abs(vl - vr) < 1e-14


@pytest.mark.parametrize('degree', range(0, 10))
def test_gauss_exactness(degree):
    # reference tests/test_quadrature.py:8-49: exactness on the line up to the stated degree
    x, w = points.gauss1(degree)
    assert len(x) == degree // 2 + 1
    for k in range(degree + 1):
        assert abs((w * x**k).sum() - 1 / (k + 1)) < 1e-14


def test_tensor_points_order():
    rules = points.tensor_gauss(3, (2, 4, 0))
    pts, wts = points.tensor_points(rules)
    assert pts.shape == (2 * 3 * 1, 3)
    assert abs(wts.sum() - 1) < 1e-14
    assert pts[1, 1] > pts[0, 1] and pts[0, 0] == pts[1, 0]  # last dimension fastest


def test_invalid_arguments():
    with pytest.raises(ValueError):
        bspline.spline_basis_1d(0, 2)
    with pytest.raises(ValueError):
        bspline.local_polynomials([0., 1., 1.])
    with pytest.raises(ValueError):
        points.gauss1(-1)


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present (GPU box)')
@pytest.mark.parametrize('p', [1, 2, 3, 4])
@pytest.mark.parametrize('btype', ['spline', 'std'])
def test_tables_equal_reference(p, btype):
    for path in (REF, SHIMS):
        if path not in sys.path:
            sys.path.insert(0, path)
    from nutils import mesh, points as refpoints
    topo, geom = mesh.rectilinear([7, 3])
    ref = topo.basis(btype, degree=p)
    for d, n in enumerate((7, 3)):
        mine = bspline.spline_basis_1d(n, p, continuity={'spline': -1, 'std': 0}[btype])
        for e in range(n):
            assert numpy.array_equal(mine.coeffs[mine.setidx[e]], numpy.asarray(ref._coeffs[d][e]))
        assert numpy.array_equal(mine.start, numpy.asarray(ref._start_dofs[d]))
        assert mine.ndofs == ref._dofs_shape[d]
    x, w = points.gauss1(2 * p)
    rx, rw = refpoints.gauss1(2 * p)
    numpy.testing.assert_allclose(x, numpy.asarray(rx)[:, 0], atol=1e-15)
    numpy.testing.assert_allclose(w, numpy.asarray(rw), atol=1e-15)
