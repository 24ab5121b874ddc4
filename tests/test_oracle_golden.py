'''Pins both oracle restatements (numpy and C) to the golden vectors written by the unmodified
reference (oracle/make_golden.py): CSR pattern bit-exact, values to 1e-13 relative.'''

import numpy
import pytest

from tests import util
from oracle import fem_oracle, c_oracle

TOL = 1e-13


def _forms(g, prob):
    if str(g['kind']) == 'scalar':
        return [('stiffness',), ('mass',)], [('load',)]
    nd = prob.ndims
    c = numpy.zeros((nd, nd + 1))
    c[nd - 1, 0] = 1.  # -u_i q_i with q = -e_last (examples/elasticity.py:57)
    return [('elasticity', float(g['lmbda']), float(g['mu']))], [('generic', c)]


def _check(g, mats, vecs):
    expect = [g['K_values'], g['M_values']] if str(g['kind']) == 'scalar' else [g['K_values']]
    for (v, rp, ci), ref in zip(mats, expect):
        assert rp.dtype == numpy.int64 and ci.dtype == numpy.int64
        assert numpy.array_equal(rp, g['rowptr'])
        assert numpy.array_equal(ci, g['colidx'])
        if len(ref):  # mass1d_known stores no stiffness
            assert util.relerr(v, ref) <= TOL
            assert util.rowsum_relerr(v, ref, rp) <= TOL
    assert util.relerr(vecs[0], g['F']) <= TOL


@pytest.mark.parametrize('name', util.golden_names())
def test_numpy_oracle(name):
    g = util.load_golden(name)
    prob = util.problem_from_golden(g)
    if prob.ntotal > 600:
        pytest.skip('python element loop: small cases only')
    mf, vf = _forms(g, prob)
    _check(g, *fem_oracle.assemble(prob, mf, vf))


@pytest.mark.parametrize('name', util.golden_names())
def test_c_oracle(name):
    g = util.load_golden(name)
    prob = util.problem_from_golden(g)
    mf, vf = _forms(g, prob)
    _check(g, *c_oracle.assemble(prob, mf, vf))


@pytest.mark.parametrize('name', util.elemset_golden_names())
def test_numpy_oracle_elemset(name):
    # trimmed topologies (ragged points, pruned numbering), NURBS: the numpy restatement against the reference's output
    g = util.load_golden(name)
    prob = util.elemset_problem_from_golden(g)
    Ds, Cs, expect, F = util.elemset_forms(g, prob.ndims)
    mc, vc = util.elemset_coefs(g, len(Ds), len(Cs))
    mats, vecs = fem_oracle.assemble(prob, [('generic', D) for D in Ds], [('generic', C) for C in Cs], matrix_coefs=mc, vector_coefs=vc)
    for (v, rp, ci), ref in zip(mats, expect):
        assert rp.dtype == numpy.int64 and ci.dtype == numpy.int64
        assert numpy.array_equal(rp, g['rowptr'])
        assert numpy.array_equal(ci, g['colidx'])
        assert util.relerr(v, ref) <= 1e-12
        assert util.rowsum_relerr(v, ref, rp) <= 1e-12
    assert util.relerr(vecs[0], F) <= 1e-12
    if 'volume' in g:
        # sum of the mass matrix = volume of the trimmed domain (partition of unity survives the pruning)
        assert abs(mats[1][0].sum() - float(g['volume'])) <= 1e-13 * abs(float(g['volume']))


@pytest.mark.parametrize('name', util.eval_golden_names())
def test_numpy_oracle_eval(name):
    # Sample.eval of the reference: coordinates, w |det J|, a discrete field and its physical gradient at the Gauss points
    g = util.load_golden(name)
    prob = util.elemset_problem_from_golden(g)
    x, w, v, gr = fem_oracle.evaluate(prob, [g['coefs']])
    assert util.relerr(x, g['x']) <= 1e-13 and util.relerr(w, g['wdet']) <= 1e-13
    assert util.relerr(v[:, 0].reshape(g['values'].shape), g['values']) <= 1e-13
    assert util.relerr(gr[:, 0].reshape(g['grads'].shape), g['grads']) <= 1e-13


def test_known_answer_mass_1d():
    # tests/test_function.py:1574-1585 of the reference: exact COO of the 1-D p=1 mass matrix
    g = util.load_golden('mass1d_known')
    prob = util.problem_from_golden(g)
    (mats, vecs) = c_oracle.assemble(prob, [('mass',)], [])
    v, rp, ci = mats[0]
    assert rp.tolist() == [0, 2, 5, 7]
    assert ci.tolist() == [0, 1, 0, 1, 2, 1, 2]
    numpy.testing.assert_allclose(v, [1 / 3, 1 / 6, 1 / 6, 2 / 3, 1 / 6, 1 / 6, 1 / 3], rtol=1e-15)


def test_generic_form_equals_named_forms():
    # the coefficient-tensor form (what the CUDA path consumes) against the reference-shaped einsums
    from nutils_b200 import engine
    g = util.load_golden('elast3d_p2_warp')
    prob = util.problem_from_golden(g)
    lm, mu = float(g['lmbda']), float(g['mu'])
    D = engine.form_elasticity(3, lm, mu, scale=2.)
    a = c_oracle.assemble(prob, [('elasticity', lm, mu)], [])[0][0][0]
    b = c_oracle.assemble(prob, [('generic', D)], [])[0][0][0]
    assert util.relerr(b, a) <= 1e-14
    g = util.load_golden('hex_p2_warp')
    prob = util.problem_from_golden(g)
    mats = c_oracle.assemble(prob, [('generic', engine.form_stiffness(3)), ('generic', engine.form_mass(3))], [('generic', engine.form_load(3))])
    assert util.relerr(mats[0][0][0], g['K_values']) <= TOL
    assert util.relerr(mats[0][1][0], g['M_values']) <= TOL
    assert util.relerr(mats[1][0], g['F']) <= TOL


def test_sort_is_stable_and_unique():
    rng = numpy.random.RandomState(0)
    rows = rng.randint(0, 50, size=5000).astype(numpy.int64)
    cols = rng.randint(0, 40, size=5000).astype(numpy.int64)
    vals = rng.rand(5000)
    (d_c,), rp_c, ci_c = c_oracle.coo_to_csr([vals], rows, cols, 50, 40)
    d_n, rp_n, ci_n = fem_oracle.coo_to_csr(vals, rows, cols, 50, 40)
    assert numpy.array_equal(rp_c, rp_n) and numpy.array_equal(ci_c, ci_n)
    # same summation order (array order) => bitwise equal sums
    assert numpy.array_equal(d_c, d_n)


def test_empty_coo():
    (d,), rp, ci = c_oracle.coo_to_csr([numpy.zeros(0)], numpy.zeros(0, dtype=numpy.int64), numpy.zeros(0, dtype=numpy.int64), 3, 3)
    assert len(d) == 0 and len(ci) == 0 and rp.tolist() == [0, 0, 0, 0]
