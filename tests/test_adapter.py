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


def test_rejects_unstructured_basis():
    _reference()
    from nutils import mesh
    topo, geom = mesh.unitsquare(2, 'triangle')
    with pytest.raises(NotImplementedError):
        adapter.tables_from_reference(topo, topo.basis('std', degree=1), 2, vertices=[[0, 1], [0, 1]])
