'''The example scripts of this repository on the B200 path, against known answers of the reference.

examples/laplace.py -- BASELINE.json configs[0], the reference's examples/laplace.py: volume stiffness, Neumann load on the
right boundary, Dirichlet data by L2 projection on the left and top boundaries (boundary mass matrices and loads with
coefficient functions: element-set kernel with the surface measure), constrained solve with the device-resident matrix,
error by Sample.eval on the device.  Known answers: the L2 errors asserted by the reference's own unit tests
(examples/laplace.py:113-137: 1.63e-3 for nelems=4 std p=1, 8.04e-5 for nelems=4 spline p=2) and the values the unmodified
reference returns here for nelems = 8 and 32 (4.0141e-4, 2.49609e-5).
examples/elasticity.py -- the reference's plane-strain example: constraints and displacements of the unmodified reference.
examples/poisson3d.py, examples/finitecell.py -- properties (maximum principle, convergence of the octree quadrature).'''

import numpy
import pytest

pytestmark = pytest.mark.gpu


from examples.laplace import solve_laplace  # noqa: E402


@pytest.mark.parametrize('nelems,btype,degree,places,expect', [(4, 'std', 1, 5, 1.63e-3), (4, 'spline', 2, 7, 8.04e-5)])
def test_reference_unit_test_values(nelems, btype, degree, places, expect):
    cons, u, err2 = solve_laplace(nelems, btype, degree)
    assert numpy.isnan(cons).sum() == (nelems + degree - 1 if btype == 'spline' else nelems * degree) ** 2
    assert round(abs(numpy.sqrt(err2) - expect), places) == 0   # unittest.assertAlmostEqual(err, expect, places=places)


@pytest.mark.parametrize('nelems,expect', [(8, 4.0141303937753854e-4), (32, 2.4960897609571972e-05)])
def test_config0_l2_error(nelems, expect):
    cons, u, err2 = solve_laplace(nelems, 'std', 1)
    assert abs(numpy.sqrt(err2) - expect) < 1e-6 * expect


def test_poisson3d_example():
    # examples/poisson3d.py: assembly, boundary projections and the constrained solve on the device; maximum principle:
    # -div grad u = 1 with u = 0 / 1 on two faces and natural conditions elsewhere keeps u >= 0 and above the harmonic part
    from examples import poisson3d
    r = poisson3d.main(n=12, degree=2)
    assert r['ndofs'] == 14 ** 3 and r['constrained'] == 2 * 14 ** 2
    assert r['residual'] <= 1e-9 and r['iterations'] > 0
    assert r['u_min'] >= -1e-9 and 1. <= r['u_max'] < 1.2
    # 1-D analogue u = x + x (1 - x) / 2: mean 2/3 - 1/12 ... = 0.5833; the warped 3-D solution stays close to it
    assert abs(r['u_mean'] - (0.5 + 1. / 12)) < 2e-2


@pytest.mark.parametrize('name', ['example_elasticity_p1', 'example_elasticity_p2', 'example_elasticity_spline'])
def test_elasticity_example(name):
    # examples/elasticity.py of the reference (plate clamped at the top under gravity; its unit tests use these sizes): the
    # constraint vector and the displacement solution the unmodified reference returns, reproduced on the GPU path
    from tests import util
    from examples import elasticity
    g = util.load_golden(name)
    cons, u = elasticity.main(int(g['nelems']), str(g['btype']), int(g['degree']), float(g['poisson']))
    assert numpy.array_equal(numpy.isnan(cons), numpy.isnan(g['cons']))
    assert numpy.allclose(cons[~numpy.isnan(cons)], g['cons'][~numpy.isnan(g['cons'])], atol=1e-13)
    assert abs(u - g['u']).max() <= 1e-9 * abs(g['u']).max()


def test_finitecell_example():
    # examples/finitecell.py (BASELINE configs[4] at small size): assembled on the tensor-core kernel with a pointwise load
    # coefficient, solved by CG on the general pattern, evaluated on the device; the error is the octree quadrature error
    # and halves (at least) with every extra level
    from examples import finitecell
    coarse = finitecell.main(n=10, degree=2, depth=1)
    fine = finitecell.main(n=10, degree=2, depth=3)
    assert abs(fine['volume'] - fine['exact_volume']) < abs(coarse['volume'] - coarse['exact_volume'])
    assert fine['relative_l2_error'] < .5 * coarse['relative_l2_error'] and fine['relative_l2_error'] < 5e-2
    assert fine['cg_iterations'] > 0


def test_semilinear_example():
    # examples/semilinear.py: Newton with solution-dependent pointwise coefficients, every step on the device; quadratic
    # convergence of the residual and an L2 error that drops with the mesh (degree 2: third order)
    from examples import semilinear
    a = semilinear.main(n=6, degree=2)
    b = semilinear.main(n=12, degree=2)
    for r in a, b:
        h = r['residual_history']
        assert r['newton_iterations'] <= 8 and h[-1] <= 1e-10 * max(h[0], 1.)
        assert h[-1] < h[-2] ** 1.5 or h[-1] < 1e-12    # superlinear at the end
    assert b['l2_error'] < a['l2_error'] / 5
    # the in-kernel coefficients (b2_elemset_set_coefficient_field, the default) and the evaluate-then-attach route agree
    c = semilinear.main(n=6, degree=2, fused=False)
    assert a['fused'] and not c['fused'] and c['newton_iterations'] == a['newton_iterations']
    assert numpy.allclose(a['residual_history'][:-1], c['residual_history'][:-1], rtol=1e-8) and abs(a['l2_error'] - c['l2_error']) <= 1e-9


def test_field_coefficient_against_pointwise():
    'b2_elemset_set_coefficient_field: scale * u_h^power inside the kernel == the same coefficient evaluated at the points and attached per point'
    import torch
    from nutils_b200 import engine, bspline, points
    ctx = engine.Context.get(0)
    rng = numpy.random.RandomState(2)
    b1 = [bspline.spline_basis_1d(n, 2) for n in (5, 4, 3)]
    verts = [numpy.linspace(0, 1, n + 1) for n in (5, 4, 3)]
    nodes = numpy.stack(numpy.meshgrid(*verts, indexing='ij')) + .02 * (rng.rand(3, 6, 5, 4) - .5)
    plan = engine.ElemSetPlan(ctx, b1, nodes=nodes, rules=points.tensor_gauss(3, 6))
    u = rng.rand(plan.ndofs)
    uq = plan.evaluate([u], x=False, weights=False)['values'][:, 0, 0]
    M, L = engine.form_mass(3), engine.form_load(3)
    plan.set_coefficient('matrix', 0, 3 * uq ** 2)
    plan.set_coefficient('vector', 0, .5 * uq ** 3)
    (ref_m,), (ref_v,) = plan.assemble_host([M], [L])
    plan.set_coefficient('matrix', 0, None)
    plan.set_coefficient('vector', 0, None)
    ud = ctx.device_alloc(8 * plan.ndofs)
    ud.from_host(u)
    plan.set_coefficient_field('matrix', 0, ud, power=2, scale=3.)
    plan.set_coefficient_field('vector', 0, ud, power=3, scale=.5)
    (m,), (v,) = plan.assemble_host([M], [L])
    plan.set_coefficient_field('matrix', 0, None)
    plan.set_coefficient_field('vector', 0, None)
    assert numpy.linalg.norm(m - ref_m) <= 1e-13 * numpy.linalg.norm(ref_m) and numpy.linalg.norm(v - ref_v) <= 1e-13 * numpy.linalg.norm(ref_v)
    (m0,), _ = plan.assemble_host([M], [])
    assert abs(m0 - ref_m).max() > 1e-3 * abs(ref_m).max()   # and it is gone again


def test_finitecell_dirichlet_example():
    # Dirichlet data on the IMMERSED sphere by a penalty term on the trimmed boundary: volume and surface points in one element
    # set, per-point normals for the measures, per-point coefficients to switch the forms
    from examples import finitecell
    r = finitecell.main_dirichlet(n=10, degree=2, depth=3)
    assert abs(r['surface_area'] - r['exact_area']) <= 1e-10 * r['exact_area']
    assert r['relative_l2_error'] < 6e-2 and r['boundary_rms_error'] < 6e-2 and r['cg_iterations'] > 0
