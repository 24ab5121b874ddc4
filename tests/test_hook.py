'''nutils_b200.hook: the reference's own Sample.integral routed through ``__nutils_dispatch__`` (_util.py:813-832).

CPU tests run the UNMODIFIED reference from baseline/_ref with the hook installed and the device backend replaced by the
oracle: they pin the host-side logic -- dispatch, the evaluable node (derivatives, as_csr, System), recognition by probing,
fall-through.  The -m gpu tests run the same with the real backend (libb200fem.so) and count kernel launches.'''

import numpy
import pytest

from tests import util, util_ref

pytestmark = pytest.mark.skipif(not util_ref.have_reference(), reason='baseline/_ref not installed (scripts/install_reference.py)')


@pytest.fixture
def hooked():
    util_ref.reference()
    from nutils_b200 import hook
    saved = hook.BACKEND
    yield hook
    hook.uninstall()
    hook.set_backend(saved)
    for k in ('offered', 'accelerated', 'declined'):
        hook.STATS[k] = 0
    hook.STATS['reasons'].clear()


def _poisson(nutils, shape=(4, 3, 3), degree=2, warp=.15):
    from nutils import mesh, function
    rng = numpy.random.RandomState(3)
    verts = [numpy.linspace(0, 1, n + 1) ** (1. + .2 * k) for k, n in enumerate(shape)]
    topo, geom0 = mesh.rectilinear(verts)
    X = numpy.stack(numpy.meshgrid(*verts, indexing='ij'))
    X = X + warp / max(shape) * (rng.rand(*X.shape) - .5)
    lin = topo.basis('spline', degree=1)
    geom = (lin * X.reshape(len(shape), -1)).sum(-1)   # the construction of mesh.rectilinear (mesh.py:55-57) with warped nodes
    basis = topo.basis('spline', degree=degree)
    g = basis.grad(geom)
    J = function.J(geom)
    qd = 2 * degree
    K = topo.integral((g[:, None, :] * g[None, :, :]).sum(-1) * J, degree=qd)
    M = topo.integral(basis[:, None] * basis[None, :] * J, degree=qd)
    F = topo.integral(basis * J, degree=qd)
    return function.eval((function.as_csr(K), function.as_csr(M), F))


def _compare(ref, got):
    for a, b in zip(ref, got):
        if isinstance(a, tuple):
            assert numpy.array_equal(a[1], b[1]) and numpy.array_equal(a[2], b[2]), 'CSR pattern differs from the stock evaluation'
            assert b[1].dtype == numpy.int64 and b[2].dtype == numpy.int64
            assert util.relerr(b[0], a[0]) <= 1e-12
        else:
            assert util.relerr(b, a) <= 1e-12


@pytest.mark.parametrize('shape,degree', [((4, 3, 3), 2), ((5, 4), 3), ((6,), 2), ((3, 3, 2), 1)])
def test_as_csr_hooked_equals_stock(hooked, shape, degree):
    nutils = util_ref.reference()
    ref = _poisson(nutils, shape, degree)
    hooked.install()
    be = util_ref.OracleBackend()
    hooked.set_backend(be)
    got = _poisson(nutils, shape, degree)
    assert hooked.STATS['accelerated'] == 3 and hooked.STATS['declined'] == 0, hooked.STATS
    assert be.calls == 3
    _compare(ref, got)


def _elasticity_system(nutils):
    'the 3-D extension of examples/elasticity.py (BASELINE.json configs[2]) through solver.System'
    from nutils import mesh, function
    from nutils.solver import System
    from nutils.expression_v2 import Namespace
    topo, geom = mesh.rectilinear([numpy.linspace(0, 1, 4), numpy.linspace(0, 1.5, 4) ** 1.2, numpy.linspace(0, 1, 3)])
    ns = Namespace()
    ns.δ = function.eye(3)
    ns.x = geom
    ns.define_for('x', gradient='∇', normal='n', jacobians=('dV', 'dS'))
    ns.u = topo.field('u', btype='spline', degree=2, shape=[3])
    ns.λ = 1.
    ns.μ = .5 / .3 - 1
    ns.ε_ij = '.5 (∇_i(u_j) + ∇_j(u_i))'
    ns.σ_ij = 'λ ε_kk δ_ij + 2 μ ε_ij'
    ns.q_i = '-δ_i2'
    energy = topo.integral('(.5 ε_ij σ_ij - u_i q_i) dV' @ ns, degree=4)
    system = System(energy, trial='u')
    jac, res = system.assemble_jacobian_residual({'u': numpy.zeros(system.trial_shapes[0])})
    return jac.export('csr'), res


def test_system_elasticity(hooked):
    nutils = util_ref.reference()
    (rv, rci, rrp), rres = _elasticity_system(nutils)
    hooked.install()
    be = util_ref.OracleBackend()
    hooked.set_backend(be)
    (gv, gci, grp), gres = _elasticity_system(nutils)
    assert hooked.STATS['accelerated'] >= 2 and be.calls >= 2, hooked.STATS
    assert numpy.array_equal(rci, gci) and numpy.array_equal(rrp, grp)
    assert util.relerr(gv, rv) <= 1e-12 and util.relerr(gres, rres) <= 1e-12


@pytest.mark.parametrize('shape,side', [((4, 3, 3), 'right'), ((5, 4), 'top'), ((5, 4), 'left'), ((3, 3, 2), 'front'), ((6,), 'right')])
def test_boundary_side_hooked_equals_stock(hooked, shape, side):
    'boundary mass matrix and load of one side of the topology (topology.py:2049-2057): the faces of the adjacent elements, surface measure'
    nutils = util_ref.reference()
    from nutils import mesh, function

    def run():
        rng = numpy.random.RandomState(5)
        verts = [numpy.linspace(0, 1, n + 1) ** (1. + .3 * k) for k, n in enumerate(shape)]
        topo, geom0 = mesh.rectilinear(verts)
        X = numpy.stack(numpy.meshgrid(*verts, indexing='ij'))
        X = X + .1 / max(shape) * (rng.rand(*X.shape) - .5)
        geom = (topo.basis('spline', degree=1) * X.reshape(len(shape), -1)).sum(-1)
        basis = topo.basis('spline', degree=2)
        J = function.J(geom)
        btopo = topo.boundary[side]
        M = btopo.integral(basis[:, None] * basis[None, :] * J, degree=4)
        F = btopo.integral(2.5 * basis * J, degree=4)
        return function.eval((function.as_csr(M), F))
    ref = run()
    hooked.install()
    be = util_ref.OracleBackend()
    hooked.set_backend(be)
    got = run()
    assert hooked.STATS['accelerated'] == 2 and be.calls == 2, hooked.STATS
    _compare(ref, got)


def _trimmed(nutils, n=4, degree=2, maxrefine=1):
    'finite-cell Poisson forms on a ball trimmed out of a box (topology.py:1598-1660): PrunedBasis, ragged cut-cell points'
    from nutils import mesh, function
    topo0, geom = mesh.rectilinear([numpy.linspace(-1, 1, n + 1)] * 3)
    topo = topo0.trim(.77 - (geom ** 2).sum(), maxrefine=maxrefine)
    basis = topo.basis('spline', degree=degree)
    g = basis.grad(geom)
    J = function.J(geom)
    K = topo.integral((g[:, None, :] * g[None, :, :]).sum(-1) * J, degree=2 * degree)
    M = topo.integral(basis[:, None] * basis[None, :] * J, degree=2 * degree)
    F = topo.integral(basis * J, degree=2 * degree)
    return function.eval((function.as_csr(K), function.as_csr(M), F))


def test_trimmed_topology_hooked_equals_stock(hooked):
    'a trimmed topology arrives as an element set: the reference\'s own cut-cell points and pruned numbering, element-set kernels'
    nutils = util_ref.reference()
    ref = _trimmed(nutils)
    hooked.install()
    be = util_ref.OracleBackend()
    hooked.set_backend(be)
    got = _trimmed(nutils)
    assert hooked.STATS['accelerated'] >= 1 and be.calls >= 1, hooked.STATS
    _compare(ref, got)


@pytest.mark.gpu
def test_gpu_trimmed_topology_hooked_equals_stock(hooked):
    nutils = util_ref.reference()
    from nutils_b200 import engine
    ref = _trimmed(nutils)
    hooked.install()
    ctx = engine.Context.get(0)
    n0 = ctx.launch_count
    got = _trimmed(nutils)
    assert hooked.STATS['accelerated'] >= 1, hooked.STATS
    assert ctx.launch_count > n0, 'the integrals did not run on the GPU'
    _compare(ref, got)


def test_declines_position_dependent_coefficient(hooked):
    'an integrand outside the closed form falls through to the stock evaluable: same numbers, no backend call'
    nutils = util_ref.reference()
    from nutils import mesh, function

    def run():
        topo, geom = mesh.rectilinear([numpy.linspace(0, 1, 5), numpy.linspace(0, 1, 4)])
        basis = topo.basis('spline', degree=2)
        M = topo.integral(numpy.cosh(geom[1]) * basis[:, None] * basis[None, :] * function.J(geom), degree=6)
        return function.eval(function.as_csr(M))
    ref = run()
    hooked.install()
    be = util_ref.OracleBackend()
    hooked.set_backend(be)
    got = run()
    assert be.calls == 0 and hooked.STATS['declined'] == 1 and hooked.STATS['accelerated'] == 0, hooked.STATS
    assert all(numpy.array_equal(a, b) for a, b in zip(ref, got))


def test_reference_laplace_example(hooked):
    'examples/laplace.py of the reference, unmodified: the domain integral and the constant-coefficient boundary matrices are assembled by the backend, the rest declines'
    util_ref.reference()
    import laplace
    cons0, u0, err0 = laplace.main(nelems=6)
    hooked.install()
    be = util_ref.OracleBackend()
    hooked.set_backend(be)
    cons1, u1, err1 = laplace.main(nelems=6)
    assert be.calls >= 3 and hooked.STATS['accelerated'] >= 3, hooked.STATS
    assert numpy.array_equal(numpy.isnan(cons0), numpy.isnan(cons1))
    assert numpy.nanmax(abs(cons0 - cons1)) <= 1e-13 and abs(u0 - u1).max() <= 1e-12 and abs(err0 - err1) <= 1e-12


# ---- the real backend ----------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_gpu_reference_laplace_example(hooked):
    "/root/reference/examples/laplace.py main(nelems=32), unmodified, with the hook active: BASELINE.json configs[0]"
    util_ref.reference()
    import laplace
    from nutils_b200 import engine
    hooked.install()
    ctx = engine.Context.get(0)
    n0 = ctx.launch_count
    cons, u, err = laplace.main(nelems=32)
    assert ctx.launch_count > n0, 'no kernel was launched: the integrals did not run on the GPU'
    assert hooked.STATS['accelerated'] >= 3, hooked.STATS   # the domain integral and the two constant-coefficient boundary matrices
    assert abs(err - 2.496e-05) <= 5e-9
    assert len(u) == 33 * 33 and int((~numpy.isnan(cons)).sum()) == 65


@pytest.mark.gpu
def test_gpu_as_csr_hooked_equals_stock(hooked):
    'function.eval(as_csr(topo.integral(...))) hooked vs stock at 8^3 p=2: bit-exact pattern, values to 1e-12'
    nutils = util_ref.reference()
    from nutils_b200 import engine
    ref = _poisson(nutils, (8, 8, 8), 2)
    hooked.install()
    ctx = engine.Context.get(0)
    n0 = ctx.launch_count
    got = _poisson(nutils, (8, 8, 8), 2)
    assert ctx.launch_count >= n0 + 3 and hooked.STATS['accelerated'] == 3, hooked.STATS
    _compare(ref, got)


@pytest.mark.gpu
def test_gpu_system_elasticity(hooked):
    nutils = util_ref.reference()
    from nutils_b200 import engine
    (rv, rci, rrp), rres = _elasticity_system(nutils)
    hooked.install()
    n0 = engine.Context.get(0).launch_count
    (gv, gci, grp), gres = _elasticity_system(nutils)
    assert engine.Context.get(0).launch_count > n0 and hooked.STATS['accelerated'] >= 2
    assert numpy.array_equal(rci, gci) and numpy.array_equal(rrp, grp)
    assert util.relerr(gv, rv) <= 1e-12 and util.relerr(gres, rres) <= 1e-12
