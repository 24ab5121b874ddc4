'''nutils_b200.adapter: tables read out of the reference's own objects (build container only -- the reference tree
does not exist on the GPU box) describe the same problem the reference assembles: fed to the oracle they reproduce
the reference's CSR matrix of a TRIMMED topology with a pruned basis.'''

import os
import sys
import numpy
import pytest

from tests import util
from oracle import fem_oracle
from nutils_b200 import adapter, points

REF = '/root/reference/src'
SHIMS = os.path.join(util.ROOT, 'oracle', 'shims')
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present (GPU box)')


def _reference():
    for path in (REF, SHIMS):
        if path not in sys.path:
            sys.path.insert(0, path)
    os.environ.setdefault('NUTILS_MATRIX', 'scipy')
    import nutils
    return nutils


def _problem(t, qdegree):
    b1 = t['bases']
    nd = len(b1)
    rules = points.tensor_gauss(nd, qdegree)
    kw = {k: t[k] for k in ('elem_ids', 'qoff', 'qcoords', 'qweights', 'renumber', 'nbasis_new') if k in t}
    return fem_oracle.Problem(t['nelems'], [b.degree for b in b1], [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1],
                              [b.ndofs for b in b1], [r[0] for r in rules], [r[1] for r in rules], t['nodes'], **kw)


@pytest.mark.parametrize('trimmed', [False, True])
def test_tables_reproduce_reference(trimmed):
    nutils = _reference()
    from nutils import mesh, function
    verts = [numpy.linspace(-1, 1, 6), numpy.linspace(-1, 1, 5) ** 3]
    topo, geom = mesh.rectilinear(verts)
    if trimmed:
        topo = topo.trim(.8 - numpy.linalg.norm(geom), maxrefine=2)
    basis = topo.basis('spline', degree=2)
    g = basis.grad(geom)
    K = topo.integral((g[:, None, :] * g[None, :, :]).sum(-1) * function.J(geom), degree=4)
    F = topo.integral(basis * function.J(geom), degree=4)
    (kv, rp, ci), f = function.eval((function.as_csr(K), F))
    t = adapter.tables_from_reference(topo, basis, 4, vertices=verts)
    assert ('elem_ids' in t) == trimmed
    prob = _problem(t, 4)
    mats, vecs = fem_oracle.assemble(prob, [('stiffness',)], [('load',)])
    v, rowptr, colidx = mats[0]
    assert numpy.array_equal(rowptr, rp) and numpy.array_equal(colidx, ci)
    assert util.relerr(v, kv) <= 1e-12
    assert util.relerr(vecs[0], f) <= 1e-12


@pytest.mark.parametrize('degree', [2, 3])
def test_nurbs_tables_reproduce_reference(degree):
    # the NURBS construction of examples/platewithhole.py:66-86: tables read out by the adapter, assembled by the oracle,
    # against the reference's own mass matrix and load of the rational basis on the NURBS geometry
    _reference()
    from nutils import mesh, function
    topo, geom0 = mesh.rectilinear([1, 2])
    bsplinebasis = topo.basis('spline', degree=2)
    controlweights = numpy.ones(12)
    controlweights[1:3] = .5 + .25 * numpy.sqrt(2)
    weightfunc = bsplinebasis @ controlweights
    nurbsbasis = bsplinebasis * controlweights / weightfunc
    A, B, C = (0, 0, 0), ((2**.5 - 1) * .5, .3 * 1.5 / 2, 1), (.5, 1.5 / 2, 1)
    geom = nurbsbasis @ numpy.array([[A, B, C, C], [C, C, B, A]]).T.reshape(-1, 2)
    topo = topo.refine(1)
    bases, scale, gbases, gctrl, gweights = adapter.nurbs_tables_from_reference(topo, weightfunc, geom, degree)
    abasis = topo.basis('spline', degree=degree)
    nbasis = abasis * scale / weightfunc
    J = function.J(geom)
    (mv, rp, ci), f = function.eval((function.as_csr(topo.integral(nbasis[:, None] * nbasis[None, :] * J, degree=8)), topo.integral(nbasis * J, degree=8)))
    rules = points.tensor_gauss(2, 8)
    prob = fem_oracle.Problem(tuple(topo.shape), [b.degree for b in bases], [b.coeffs for b in bases], [b.setidx for b in bases], [b.start for b in bases],
                              [b.ndofs for b in bases], [r[0] for r in rules], [r[1] for r in rules], numpy.zeros((2, 0, 0)), scale=scale, rational=2,
                              geom_spline=dict(degree=[b.degree for b in gbases], coeffs=[b.coeffs for b in gbases], setidx=[b.setidx for b in gbases],
                                               start=[b.start for b in gbases], ndofs_d=[b.ndofs for b in gbases], ctrl=gctrl, weights=gweights))
    mats, vecs = fem_oracle.assemble(prob, [('mass',)], [('load',)])
    assert numpy.array_equal(mats[0][1], rp) and numpy.array_equal(mats[0][2], ci)
    assert util.relerr(mats[0][0], mv) <= 1e-11 and util.relerr(vecs[0], f) <= 1e-11


def test_immersed_boundary_tables_reproduce_reference():
    # the trimmed boundary of a finite-cell topology: facets of the cut-cell mosaics, not element faces
    _reference()
    from nutils import mesh, function
    verts = [numpy.linspace(-1, 1, 5), numpy.linspace(-1, 1, 4)]
    topo0, geom = mesh.rectilinear(verts)
    topo = topo0.trim(.7 - numpy.linalg.norm(geom), maxrefine=2, name='trimmed')
    bt = topo.boundary['trimmed']
    basis = topo.basis('spline', degree=2)
    J = function.J(geom)
    (bv, rp, ci), f, length = function.eval((function.as_csr(bt.integral(basis[:, None] * basis[None, :] * J, degree=4)), bt.integral(basis * J, degree=4), bt.integral(J, degree=4)))
    assert abs(length - 2 * numpy.pi * .7) < 2e-2
    elem_ids, qoff, qcoords, qweights, normals = adapter.immersed_boundary_tables_from_reference(bt, topo0, 4)
    t = adapter.tables_from_reference(topo, basis, 4, vertices=verts)
    b1 = t['bases']
    rules = points.tensor_gauss(2, 4)
    prob = fem_oracle.Problem(t['nelems'], [b.degree for b in b1], [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1], [b.ndofs for b in b1],
                              [r[0] for r in rules], [r[1] for r in rules], t['nodes'], elem_ids=elem_ids, qoff=qoff, qcoords=qcoords, qweights=qweights,
                              normals=normals, renumber=t['renumber'], nbasis_new=t['nbasis_new'])
    mats, vecs = fem_oracle.assemble(prob, [('mass',)], [('load',)])
    assert numpy.array_equal(mats[0][1], rp) and numpy.array_equal(mats[0][2], ci)
    assert util.relerr(mats[0][0], bv) <= 1e-12 and util.relerr(vecs[0], f) <= 1e-12
    assert abs(vecs[0].sum() - length) <= 1e-12 * length


def test_rejects_unstructured_basis():
    _reference()
    from nutils import mesh
    topo, geom = mesh.unitsquare(2, 'triangle')
    with pytest.raises(NotImplementedError):
        adapter.tables_from_reference(topo, topo.basis('std', degree=1), 2, vertices=[[0, 1], [0, 1]])
