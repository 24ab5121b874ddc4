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
    # run on torch's current stream: several tests fill torch tensors right before handing them to the C ABI
    import torch
    c = engine.Context.get(0)
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    yield c
    c.set_stream(None)


def _plan(ctx, prob):
    bases = [bspline.Basis1D(prob.degree[d], prob.nelems[d], prob.coeffs[d], prob.setidx[d], prob.start[d], prob.ndofs_d[d]) for d in range(prob.ndims)]
    return engine.Plan(ctx, bases, list(zip(prob.qpts, prob.qwts)), prob.nodes, ncomp=prob.ncomp)


def _forms(g, nd):
    if str(g['kind']) == 'scalar':
        return [engine.form_stiffness(nd), engine.form_mass(nd)], [engine.form_load(nd)], [g['K_values'], g['M_values']]
    C = numpy.zeros((nd, nd + 1))
    C[nd - 1, 0] = 1.
    return [engine.form_elasticity(nd, float(g['lmbda']), float(g['mu']), scale=2.)], [C], [g['K_values']]


# (kernel, path): owner-computes rows kernel / its coverage fallback / element-scatter specialised / element-scatter generic
MODES = [(0, 0), (1, 0), (0, 1), (1, 1)]
MODE_IDS = ['rows', 'rows-generic', 'scatter', 'scatter-generic']


@pytest.mark.parametrize('mode', MODES, ids=MODE_IDS)
@pytest.mark.parametrize('name', util.golden_names())
def test_golden(ctx, name, mode):
    g = util.load_golden(name)
    prob = util.problem_from_golden(g)
    plan = _plan(ctx, prob)
    ctx.set_option('kernel', mode[0])
    ctx.set_option('path', mode[1])
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
        ctx.set_option('path', 0)


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
    dict(nelems=(1, 1, 1), degree=2), dict(nelems=(3, 13, 18), degree=2), dict(nelems=(7, 11, 5), degree=1), dict(nelems=(1, 2, 9), degree=2), dict(nelems=(7, 9, 8), degree=3), dict(nelems=(2, 1, 3), degree=3),
    dict(nelems=(6, 5, 7), degree=4), dict(nelems=(1, 1, 1), degree=4), dict(nelems=(2, 7, 1), degree=4), dict(nelems=(3, 2, 3), degree=4, ncomp=3),
]


@pytest.mark.parametrize('mode', MODES, ids=MODE_IDS)
@pytest.mark.parametrize('case', CASES, ids=lambda c: 'x'.join(map(str, c['nelems'])) + 'p{}c{}'.format(c['degree'], c.get('ncomp', 1)) + c.get('btype', '') + str(c.get('qdegree', '')))
def test_oracle_seeded(ctx, case, mode):
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
    ctx.set_option('kernel', mode[0])
    ctx.set_option('path', mode[1])
    try:
        rowptr, colidx = plan.csr_pattern()
        assert numpy.array_equal(rowptr, mats[0][1])
        assert numpy.array_equal(colidx, mats[0][2])
        vals, rhs = plan.assemble_host(Ds, Cs)
    finally:
        ctx.set_option('kernel', 0)
        ctx.set_option('path', 0)
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


@pytest.mark.parametrize('nelems,degree,btype,ncomp', [((9, 7), 3, 'std', 1), ((5, 4, 3), 3, 'spline', 1), ((6, 5), 2, 'spline', 2)])
def test_element_ranges_accumulate_coverage(ctx, nelems, degree, btype, ncomp):
    # the same contract on the coverage route (element-set kernel on the element set "everything": the element range becomes a
    # range of the selection), against the oracle
    prob = _random_problem(5, nelems, degree, ncomp=ncomp, btype=btype)
    plan = _plan(ctx, prob)
    nd = prob.ndims
    D, C = [engine.form_stiffness(nd, ncomp) + .5 * engine.form_mass(nd, ncomp)], [engine.form_load(nd, ncomp)]
    mats, vecs = c_oracle.assemble(prob, [('generic', D[0])], [('generic', C[0])])
    n = prob.ntotal
    parts = [plan.assemble_host(D, C, elem_range=r) for r in ((0, n // 3), (n // 3, n // 3), (n // 3, n))]
    assert util.relerr(sum(p[0][0] for p in parts), mats[0][0]) <= TOL
    assert util.relerr(sum(p[1][0] for p in parts), vecs[0]) <= TOL


ROWS_CASES = [
    dict(nelems=(10, 9, 11), degree=2), dict(nelems=(9, 8, 7), degree=1), dict(nelems=(3, 13, 18), degree=2), dict(nelems=(1, 1, 1), degree=2),
    dict(nelems=(2, 1, 1), degree=1), dict(nelems=(6, 5, 4), degree=2, btype='std'), dict(nelems=(5, 4, 6), degree=3), dict(nelems=(12, 15), degree=2),
    dict(nelems=(5, 4, 5), degree=2, ncomp=3), dict(nelems=(8, 7, 10), degree=3), dict(nelems=(1, 1, 1), degree=3), dict(nelems=(3, 2, 2), degree=4),
    dict(nelems=(6, 7, 5), degree=2, ncomp=3, vector_forms=True), dict(nelems=(9, 4, 9), degree=1, ncomp=3, vector_forms=True),
    dict(nelems=(1, 1, 1), degree=2, ncomp=3, vector_forms=True), dict(nelems=(3, 3, 2), degree=3, ncomp=3, vector_forms=True),
    dict(nelems=(5, 6, 4), degree=4), dict(nelems=(1, 1, 1), degree=4), dict(nelems=(9, 3, 5), degree=4),
    dict(nelems=(4, 3, 3), degree=4, ncomp=3, vector_forms=True), dict(nelems=(7, 4, 5), degree=3, ncomp=3, vector_forms=True),
]


@pytest.mark.parametrize('nseg', [0, 1, 3])
@pytest.mark.parametrize('case', ROWS_CASES, ids=lambda c: 'x'.join(map(str, c['nelems'])) + 'p{}c{}'.format(c['degree'], c.get('ncomp', 1)) + c.get('btype', '') + ('v' if c.get('vector_forms') else ''))
def test_rows_plane_ranges(ctx, case, nseg):
    case = dict(case)
    '''b2_assemble_rows_device: physical coefficients (anisotropic conductivity, density, load) against the oracle; disjoint
    plane ranges written into one poisoned array give the full result (the multi-GPU decomposition: no exchange, no zero-fill),
    and rows outside a range are not touched.  Covers the specialised owner-computes kernel (3-D scalar p=1,2 splines) and the
    coverage path (everything else).'''
    import torch
    prob = _random_problem(seed=hash(str(case)) % 2**31, **{k: v for k, v in case.items() if k != 'vector_forms'})
    nd, nc = prob.ndims, prob.ncomp
    rng = numpy.random.RandomState(11)
    A = rng.rand(nd, nd)
    Ds = [engine.form_stiffness(nd, nc, conductivity=A @ A.T + numpy.eye(nd)), 2.5 * engine.form_mass(nd, nc)]
    Cs = [-.75 * engine.form_load(nd, nc)]
    if nc == nd == 3 and case.get('vector_forms'):
        # elasticity, an arbitrary (non-symmetric) gradient-gradient coupling of the components, a full component mass matrix
        Dgg = numpy.zeros((nc, nd + 1, nc, nd + 1))
        Dgg[:, 1:, :, 1:] = rng.rand(nc, nd, nc, nd) - .5
        Dm = numpy.zeros((nc, nd + 1, nc, nd + 1))
        Dm[:, 0, :, 0] = rng.rand(nc, nc)
        Ds = [engine.form_elasticity(nd, 1.3, .7), Dgg, Dm]
        Cs = [engine.form_load(nd, nc, f=[.3, -1.1, 2.])]
    mats, vecs = c_oracle.assemble(prob, [('generic', D) for D in Ds], [('generic', C) for C in Cs])
    plan = _plan(ctx, prob)
    rowptr, _ = plan.csr_pattern()
    ndof0 = prob.ndofs_d[0]
    dev = torch.device('cuda', 0)
    poison = -7.25
    vals = [torch.full((plan.nnz,), poison, dtype=torch.float64, device=dev) for _ in Ds]
    rhs = [torch.full((plan.ndofs,), poison, dtype=torch.float64, device=dev) for _ in Cs]
    cuts = sorted(set([0, ndof0 // 3, ndof0 // 3 + 1, ndof0]))
    ctx.set_option('rows_nseg', nseg)
    try:
        per_plane = plan.ndofs // ndof0
        for i, (p0, p1) in enumerate(zip(cuts[:-1], cuts[1:])):
            plan.assemble_rows_device(Ds, Cs, vals, rhs, plane_range=(p0, p1))
            ctx.synchronize()
            if i == 0 and p1 < ndof0:
                # nothing outside the planes [p0, p1) was written
                assert bool((vals[0][int(rowptr[p1 * per_plane]):] == poison).all())
                assert bool((rhs[0][p1 * per_plane:] == poison).all())
    finally:
        ctx.set_option('rows_nseg', 0)
    for v, (ref, _, _) in zip(vals, mats):
        v = v.cpu().numpy()
        assert util.relerr(v, ref) <= TOL
        assert util.rowsum_relerr(v, ref, rowptr) <= TOL
    assert util.relerr(rhs[0].cpu().numpy(), vecs[0]) <= TOL


VEC_CASES = [dict(nelems=(6, 7, 5), degree=2), dict(nelems=(9, 4, 9), degree=1), dict(nelems=(1, 1, 1), degree=2), dict(nelems=(2, 1, 3), degree=1),
             dict(nelems=(5, 4, 4), degree=3), dict(nelems=(4, 3, 3), degree=4), dict(nelems=(1, 1, 1), degree=4), dict(nelems=(11, 5, 6), degree=2)]


@pytest.mark.parametrize('form', ['elasticity', 'symmetric', 'general'])
@pytest.mark.parametrize('case', VEC_CASES, ids=lambda c: 'x'.join(map(str, c['nelems'])) + 'p{}'.format(c['degree']))
def test_rows_vector_blocks(ctx, case, form):
    '''Vector-valued (3-component) stiffness-like forms through the owner-computes kernel ITSELF (context option kernel = 2: the
    specialised kernel or an error; one matrix form per call, as the kernel takes at most two).  Symmetric forms (D[c][x][e][y] =
    D[e][y][c][x]: elasticity, a random symmetrised coupling) are integrated by blocks -- diagonal blocks on the symmetric scalar
    pipeline, blocks above the diagonal with transposed stores, blocks below never; a non-symmetric coupling takes all nine
    blocks.  Against the oracle, over disjoint plane ranges written into one poisoned array, and symmetric route against the
    nine-block route (rows_vecsym = 0).'''
    import torch
    prob = _random_problem(seed=hash(str(case) + form) % 2**31, ncomp=3, **case)
    rng = numpy.random.RandomState(5)
    Dgg = numpy.zeros((3, 4, 3, 4))
    Dgg[:, 1:, :, 1:] = rng.rand(3, 3, 3, 3) - .5
    D = {'elasticity': engine.form_elasticity(3, 1.3, .7), 'symmetric': Dgg + Dgg.transpose(2, 3, 0, 1) + 2 * engine.form_elasticity(3, .1, 1.), 'general': Dgg}[form]
    Cs = [engine.form_load(3, 3, f=[.3, -1.1, 2.])]
    mats, vecs = c_oracle.assemble(prob, [('generic', D)], [('generic', C) for C in Cs])
    plan = _plan(ctx, prob)
    rowptr, _ = plan.csr_pattern()
    ndof0 = prob.ndofs_d[0]
    dev = torch.device('cuda', 0)
    poison = -7.25
    cuts = sorted(set([0, ndof0 // 3, ndof0 // 3 + 1, ndof0]))
    per_plane = plan.ndofs // ndof0
    results = []
    ctx.set_option('kernel', 2)
    try:
        for vecsym in (1, 0):
            ctx.set_option('rows_vecsym', vecsym)
            vals = [torch.full((plan.nnz,), poison, dtype=torch.float64, device=dev)]
            rhs = [torch.full((plan.ndofs,), poison, dtype=torch.float64, device=dev)]
            for i, (p0, p1) in enumerate(zip(cuts[:-1], cuts[1:])):
                plan.assemble_rows_device([D], Cs, vals, rhs, plane_range=(p0, p1))
                ctx.synchronize()
                if i == 0 and p1 < ndof0:
                    assert bool((vals[0][int(rowptr[p1 * per_plane]):] == poison).all())
                    assert bool((rhs[0][p1 * per_plane:] == poison).all())
            v = vals[0].cpu().numpy()
            assert util.relerr(v, mats[0][0]) <= TOL
            assert util.rowsum_relerr(v, mats[0][0], rowptr) <= TOL
            assert util.relerr(rhs[0].cpu().numpy(), vecs[0]) <= TOL
            results.append(v)
    finally:
        ctx.set_option('kernel', 0)
        ctx.set_option('rows_vecsym', 1)
    assert util.relerr(results[0], results[1]) <= TOL


def test_rows_nonuniform_knots_and_long_march(ctx):
    '''Graded knot vectors give every element its own coefficient set (the kernel's table-driven variants instead of the
    constant-bank fast path), and 700 element layers exceed one marching segment (512 layers), forcing a split.'''
    import torch
    rng = numpy.random.RandomState(4)
    dev = torch.device('cuda', 0)
    for nelems, degree in (((9, 8, 7), 2), ((700, 2, 3), 2), ((6, 9, 5), 1), ((5, 6, 4), 3)):
        knots = [numpy.concatenate([[0.], numpy.cumsum(.3 + rng.rand(n))]) for n in nelems]
        b1 = [bspline.spline_basis_1d(n, degree, knotvalues=k) for n, k in zip(nelems, knots)]
        assert degree == 1 or len(b1[1].coeffs) > 3 or nelems[1] <= 3  # (hat functions are the same polynomials on every element)
        rules = points.tensor_gauss(3, 2 * degree)
        X = numpy.stack(numpy.meshgrid(*knots, indexing='ij'))
        X = X + .1 * (rng.rand(*X.shape) - .5)
        prob = fem_oracle.Problem(nelems, [degree] * 3, [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1],
                                  [b.ndofs for b in b1], [r[0] for r in rules], [r[1] for r in rules], X)
        Ds, Cs = [engine.form_stiffness(3), engine.form_mass(3)], [engine.form_load(3)]
        mats, vecs = c_oracle.assemble(prob, [('generic', D) for D in Ds], [('generic', C) for C in Cs])
        plan = _plan(ctx, prob)
        rowptr, colidx = plan.csr_pattern()
        assert numpy.array_equal(rowptr, mats[0][1]) and numpy.array_equal(colidx, mats[0][2])
        vals = [torch.full((plan.nnz,), 9., dtype=torch.float64, device=dev) for _ in Ds]
        rhs = [torch.full((plan.ndofs,), 9., dtype=torch.float64, device=dev)]
        ctx.set_option('kernel', 2)  # the specialised kernel or an error
        try:
            plan.assemble_rows_device(Ds, Cs, vals, rhs)
        finally:
            ctx.set_option('kernel', 0)
        for v, (ref, _, _) in zip(vals, mats):
            assert util.relerr(v.cpu().numpy(), ref) <= TOL
        assert util.relerr(rhs[0].cpu().numpy(), vecs[0]) <= TOL


def test_rows_single_forms(ctx):
    # K only, M only, f only through the owner-computes kernel
    import torch
    prob = _random_problem(21, (7, 6, 9), 2)
    plan = _plan(ctx, prob)
    Ds = [engine.form_stiffness(3), engine.form_mass(3)]
    Cs = [engine.form_load(3)]
    mats, vecs = c_oracle.assemble(prob, [('generic', D) for D in Ds], [('generic', C) for C in Cs])
    dev = torch.device('cuda', 0)
    ctx.set_option('kernel', 2)  # specialised kernel or error
    try:
        for m in range(2):
            v = torch.full((plan.nnz,), 3., dtype=torch.float64, device=dev)
            plan.assemble_rows_device([Ds[m]], [], [v], [])
            assert util.relerr(v.cpu().numpy(), mats[m][0]) <= TOL
        r = torch.full((plan.ndofs,), 3., dtype=torch.float64, device=dev)
        plan.assemble_rows_device([], Cs, [], [r])
        assert util.relerr(r.cpu().numpy(), vecs[0]) <= TOL
    finally:
        ctx.set_option('kernel', 0)


def test_properties_large(ctx):
    # size-independent properties at 48^3, p=2 (the value-by-value oracle comparison at bench scale is test_bench_workload_against_oracle)
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


def test_full_size_properties(ctx):
    '''BASELINE.json configs[1] at its full size (128^3, p=2, K+M+f; the oracle would need hours and ~100 GB): size-independent
    properties checked on the device.  (1) the two independent CUDA paths -- owner-computes rows kernel and element-scatter
    kernel -- agree to 1e-12; (2) partition of unity: K 1 = 0, M 1 = f, sum f = volume; (3) scaling the nodes by 2 scales K by
    exactly 2 and M, f by exactly 8 (power-of-two scaling commutes with every rounding), bitwise.'''
    import torch
    from bench import make_nodes
    n, p = 128, 2
    b1 = util.bases_1d((n,) * 3, p, 'spline')
    rules = points.tensor_gauss(3, 2 * p)
    X = make_nodes((n,) * 3)
    plan = engine.Plan(ctx, b1, rules, X)
    assert plan.ndofs == 130 ** 3 and plan.nnz == 644 ** 3
    dev = torch.device('cuda', 0)
    Ds, Cs = [engine.form_stiffness(3), engine.form_mass(3)], [engine.form_load(3)]
    K, M = (torch.empty(plan.nnz, dtype=torch.float64, device=dev) for _ in range(2))
    f = torch.empty(plan.ndofs, dtype=torch.float64, device=dev)
    ctx.set_option('kernel', 2)
    try:
        plan.assemble_rows_device(Ds, Cs, [K, M], [f])
    finally:
        ctx.set_option('kernel', 0)
    ctx.synchronize()
    # (2) row sums
    rowptr = torch.empty(plan.ndofs + 1, dtype=torch.int64, device=dev)
    colidx = torch.empty(plan.nnz, dtype=torch.int64, device=dev)
    plan.csr_pattern_device(rowptr, colidx)
    ctx.synchronize()
    del colidx
    lengths = rowptr[1:] - rowptr[:-1]
    Ksum = torch.segment_reduce(K, 'sum', lengths=lengths)
    Msum = torch.segment_reduce(M, 'sum', lengths=lengths)
    assert float(Ksum.abs().max()) <= 1e-12 * 125 * float(K.abs().max())
    assert float((Msum - f).norm() / f.norm()) <= 1e-12
    volume = float(f.sum())
    assert abs(volume - 1.) < 1e-3 and abs(float(M.sum()) - volume) <= 1e-12 * volume  # warped unit cube
    # (1) element-scatter path
    K2, M2 = torch.zeros_like(K), torch.zeros_like(M)
    f2 = torch.zeros_like(f)
    torch.cuda.synchronize()  # the zero-fill runs on torch's stream, the assembly on the context's own
    plan.assemble_device(Ds, Cs, [K2, M2], [f2])
    ctx.synchronize()
    for a, b in ((K, K2), (M, M2), (f, f2)):
        assert float((a - b).norm() / b.norm()) <= TOL
    del K2, M2, f2, Ksum, Msum
    # (3) exact scaling
    plan2 = engine.Plan(ctx, b1, rules, 2. * X)
    K3, M3 = torch.empty_like(K), torch.empty_like(M)
    f3 = torch.empty_like(f)
    plan2.assemble_rows_device(Ds, Cs, [K3, M3], [f3])
    ctx.synchronize()
    assert bool((K3 == 2. * K).all()) and bool((M3 == 8. * M).all()) and bool((f3 == 8. * f).all())


def test_errors(ctx):
    from nutils_b200._lib import B200Error
    prob = _random_problem(1, (3, 3), 2)
    plan = _plan(ctx, prob)
    with pytest.raises(B200Error):
        plan.assemble_host([engine.form_mass(2)], [], elem_range=(5, 100))
    with pytest.raises(ValueError):
        engine.Plan(ctx, plan.bases, plan.rules, prob.nodes[:, :-1])


@pytest.mark.parametrize('workload,degree,n', [('poisson', 1, 40), ('poisson', 2, 40), ('poisson', 3, 24), ('poisson', 4, 16), ('elasticity', 2, 24), ('elasticity', 1, 24)])
def test_bench_workload_against_oracle(ctx, workload, degree, n):
    '''The TIMED configuration of bench.py (same geometry generator, forms, entry point and kernels: k_geom3d + k_rows3d through
    b2_assemble_host) at the size of bench.py's CPU sample, value by value against the C oracle: pattern bit-exact, relative
    Frobenius / row-sum / rhs errors <= 1e-12.  bench.py repeats this comparison inside every run (`parity`).'''
    import bench
    rate, dt, ndofs, cores, mats, vecs, (prob, b1, rules) = bench.port_assemble(workload, n, degree)
    Ds, Cs = bench.structured_forms(workload)
    plan = engine.Plan(ctx, b1, rules, prob.nodes, ncomp=prob.ncomp)
    n0 = ctx.launch_count
    vals, rhs = plan.assemble_host(Ds, Cs)
    assert ctx.launch_count > n0
    rec = bench.parity_record(vals, [m[0] for m in mats], plan.csr_pattern(), (mats[0][1], mats[0][2]), rhs, vecs, 'test')
    assert rec['ok'], rec
    # and the same numbers without the precomputed geometry (in-kernel evaluation): the two routes are interchangeable
    ctx.set_option('rows_gpre', 0)
    try:
        vals0, rhs0 = plan.assemble_host(Ds, Cs)
    finally:
        ctx.set_option('rows_gpre', 1)
    for a, b in zip(vals + rhs, vals0 + rhs0):
        assert util.relerr(a, b) <= 1e-13
