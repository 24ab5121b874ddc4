'''General dof maps: simplex and mixed meshes.  Goldens were written by the unmodified reference
(oracle/make_golden.py general_case: mesh.unitsquare(..., 'triangle' | 'mixed'), basis('std'), function.eval(as_csr(...))) together
with the element tables that nutils_b200.adapter.general_tables_from_reference read out of the reference's objects.
CPU: the oracle restatement on those tables reproduces the reference.  GPU: engine.GeneralPlan -- b2_pattern_general (device
radix sort / unique of (row, col) keys) and b2_assemble_general_host -- against the same files: pattern bit-exact, values 1e-12.'''

import numpy
import pytest

from tests import util
from oracle import fem_oracle
from nutils_b200 import engine

NAMES = ['general_tri_p1', 'general_tri_p2', 'general_mixed_p2', 'general_tri_elast_p2']


def _tables(g):
    types = [dict(weights=g['weights_%d' % k], phi=g['phi_%d' % k], dphi=g['dphi_%d' % k], gdphi=g['gdphi_%d' % k]) for k in range(int(g['ntypes']))]
    dofs = [g['dofs'][a:b] for a, b in zip(g['dofoff'][:-1], g['dofoff'][1:])]
    verts = [g['vertcoords'][a:b] for a, b in zip(g['vertoff'][:-1], g['vertoff'][1:])]
    return types, dofs, verts


def _forms(g):
    nc = int(g['ncomp'])
    if nc == 1:
        return [engine.form_stiffness(2), engine.form_mass(2)], [engine.form_load(2)], [g['K_values'], g['M_values']]
    C = numpy.zeros((2, 3))
    C[:, 0] = -g['load']
    return [engine.form_elasticity(2, float(g['lmbda']), float(g['mu']))], [C], [g['K_values']]


@pytest.mark.parametrize('name', NAMES)
def test_oracle_general(name):
    g = util.load_golden(name)
    types, dofs, verts = _tables(g)
    Ds, Cs, expect = _forms(g)
    mats, vecs = fem_oracle.assemble_general(2, types, g['etype'], dofs, verts, int(g['nbasis']), [('generic', D) for D in Ds], [('generic', C) for C in Cs], ncomp=int(g['ncomp']))
    assert numpy.array_equal(mats[0][1], g['rowptr']) and numpy.array_equal(mats[0][2], g['colidx'])
    for (v, _, _), ref in zip(mats, expect):
        assert util.relerr(v, ref) <= 1e-13
    assert util.relerr(vecs[0], g['F']) <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_gpu_general(name):
    g = util.load_golden(name)
    types, dofs, verts = _tables(g)
    Ds, Cs, expect = _forms(g)
    ctx = engine.Context.get(0)
    n0 = ctx.launch_count
    plan = engine.GeneralPlan(ctx, 2, types, g['etype'], dofs, verts, int(g['nbasis']), ncomp=int(g['ncomp']))
    rowptr, colidx = plan.csr_pattern()
    assert rowptr.dtype == numpy.int64 and colidx.dtype == numpy.int64
    assert numpy.array_equal(rowptr, g['rowptr']) and numpy.array_equal(colidx, g['colidx'])
    vals, rhs = plan.assemble_host(Ds, Cs)
    assert ctx.launch_count > n0
    for v, ref in zip(vals, expect):
        assert util.relerr(v, ref) <= 1e-12 and util.rowsum_relerr(v, ref, rowptr) <= 1e-12
    assert util.relerr(rhs[0], g['F']) <= 1e-12
    # the assembled matrix stays usable on the device: y = K x against the host product
    from nutils_b200 import matrix
    K = matrix.assemble_csr(vals[0], rowptr, colidx, plan.ndofs)
    x = numpy.linspace(0, 1, plan.ndofs)
    assert numpy.allclose(K.todevice() @ x, K._sp() @ x, rtol=1e-12, atol=1e-12 * abs(vals[0]).max())


@pytest.mark.gpu
def test_gpu_general_pattern_random():
    'b2_pattern_general on random ragged dof lists against numpy unique (the reference algorithm, evaluable.py:5646-5682)'
    rng = numpy.random.RandomState(11)
    nb, ne = 500, 2000
    dofs = [numpy.sort(rng.choice(nb, size=rng.randint(1, 9), replace=False)) for _ in range(ne)]
    ctx = engine.Context.get(0)
    types = [dict(weights=numpy.ones(1), phi=numpy.ones((1, k)), dphi=numpy.zeros((1, k, 2)), gdphi=numpy.zeros((1, 3, 2))) for k in range(1, 9)]
    etype = numpy.array([len(d) - 1 for d in dofs])
    verts = [numpy.array([[0., 0], [1, 0], [0, 1]])] * ne
    plan = engine.GeneralPlan(ctx, 2, types, etype, dofs, verts, nb)
    rowptr, colidx = plan.csr_pattern()
    keys = numpy.unique(numpy.concatenate([(d[:, None] * nb + d[None, :]).ravel() for d in dofs]))
    assert numpy.array_equal(colidx, keys % nb)
    assert numpy.array_equal(rowptr, numpy.searchsorted(keys // nb, numpy.arange(nb + 1)))
