'''Parity of the element-set CUDA path (b2_elemset_* / b2_pattern_create_elemset / b2_assemble_elemset_*, through the
C ABI) with the golden vectors of the reference: trimmed topologies with ragged cut-cell quadrature and pruned dof
numbering (finite cell method, BASELINE config 5; examples/platewithhole.py FCM mode) and rational splines on a
NURBS geometry (examples/platewithhole.py NURBS mode, degree 2 and BASELINE config 4's degree 4).
Bar: CSR pattern bit-exact; values: relative Frobenius and row-sum error <= 1e-12; rhs <= 1e-12.'''

import numpy
import pytest

from tests import util
from oracle import fem_oracle
from nutils_b200 import engine, bspline, points

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope='module')
def ctx():
    import torch
    c = engine.Context.get(0)
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    yield c
    c.set_stream(None)


def _bases(prob):
    return [bspline.Basis1D(prob.degree[d], prob.nelems[d], prob.coeffs[d], prob.setidx[d], prob.start[d], prob.ndofs_d[d]) for d in range(prob.ndims)]


def _plan(ctx, prob):
    gs = None
    if prob.geom_spline is not None:
        g = prob.geom_spline
        gb = [bspline.Basis1D(g['degree'][d], prob.nelems[d], g['coeffs'][d], g['setidx'][d], g['start'][d], g['ndofs_d'][d]) for d in range(prob.ndims)]
        gs = gb, g['ctrl'], g['weights']
    plan = engine.ElemSetPlan(ctx, _bases(prob), nodes=None if gs else prob.nodes, ncomp=prob.ncomp, rules=list(zip(prob.qpts, prob.qwts)),
                              elem_ids=prob.elem_ids, qoff=prob.qoff, qcoords=prob.qcoords, qweights=prob.qweights,
                              renumber=prob.renumber, nbasis_new=prob.nbasis_new, scale=prob.scale, rational=prob.rational, geom_spline=gs)
    if prob.face_dim is not None:
        plan.set_faces(prob.face_dim)
    if prob.normals is not None:
        plan.set_normals(prob.normals)
    return plan


@pytest.fixture(params=[1, 0], ids=['mma', 'fma'])
def mma(request, ctx):
    'block entries on the FP64 tensor cores (default) or by the scalar FMA loop'
    ctx.set_option('elemset_mma', request.param)
    yield request.param
    ctx.set_option('elemset_mma', 1)


@pytest.mark.parametrize('name', util.elemset_golden_names())
def test_golden(ctx, name, mma):
    g = util.load_golden(name)
    prob = util.elemset_problem_from_golden(g)
    plan = _plan(ctx, prob)
    rowptr, colidx = plan.csr_pattern()
    assert rowptr.dtype == numpy.int64 and colidx.dtype == numpy.int64
    assert numpy.array_equal(rowptr, g['rowptr'])
    assert numpy.array_equal(colidx, g['colidx'])
    assert plan.ndofs == int(g['ndofs']) and plan.nnz == len(g['colidx'])
    Ds, Cs, expect, F = util.elemset_forms(g, prob.ndims)
    mc, vc = util.elemset_coefs(g, len(Ds), len(Cs))
    for k in range(len(Ds)):
        plan.set_coefficient('matrix', k, None if mc is None else mc[k])
    for k in range(len(Cs)):
        plan.set_coefficient('vector', k, None if vc is None else vc[k])
    vals, rhs = plan.assemble_host(Ds, Cs)
    for v, ref in zip(vals, expect):
        assert util.relerr(v, ref) <= TOL
        assert util.rowsum_relerr(v, ref, rowptr) <= TOL
    assert util.relerr(rhs[0], F) <= TOL
    if 'area' in g and mc is not None:
        # without the coefficient the boundary mass matrix sums to the area of the face (partition of unity on the boundary)
        plan.set_coefficient('matrix', 0, None)
        assert abs(plan.assemble_host(Ds, [])[0][0].sum() - float(g['area'])) <= 1e-12 * abs(float(g['area']))
    if 'volume' in g:
        assert abs(vals[1].sum() - float(g['volume'])) <= 1e-12 * abs(float(g['volume']))
    for k in range(plan.ndofs + 1):
        if k % 37 == 0 or k == plan.ndofs:
            assert plan.row_offset(k) == rowptr[k]


@pytest.mark.parametrize('name', ['hex_p2_warp', 'quad_p3', 'elast3d_p1', 'line_p3_std', 'hex_p2_std', 'quad_p4_warp', 'hex_p3_warp', 'elast3d_p2_warp', 'elast2d_p2_warp'])
def test_full_selection_equals_structured(ctx, name, mma):
    # an element set that selects everything (tensor rule, identity numbering) reproduces the structured goldens:
    # the general pattern construction against the analytic one, the element-set kernel against the reference
    g = util.load_golden(name)
    prob = util.problem_from_golden(g)
    plan = engine.ElemSetPlan(ctx, _bases(prob), nodes=prob.nodes, ncomp=prob.ncomp, rules=list(zip(prob.qpts, prob.qwts)))
    rowptr, colidx = plan.csr_pattern()
    assert numpy.array_equal(rowptr, g['rowptr']) and numpy.array_equal(colidx, g['colidx'])
    nd = prob.ndims
    if str(g['kind']) == 'scalar':
        Ds, Cs, expect = [engine.form_stiffness(nd), engine.form_mass(nd)], [engine.form_load(nd)], [g['K_values'], g['M_values']]
    else:
        C = numpy.zeros((nd, nd + 1))
        C[nd - 1, 0] = 1.
        Ds, Cs, expect = [engine.form_elasticity(nd, float(g['lmbda']), float(g['mu']), scale=2.)], [C], [g['K_values']]
    vals, rhs = plan.assemble_host(Ds, Cs)
    for v, ref in zip(vals, expect):
        assert util.relerr(v, ref) <= TOL
    assert util.relerr(rhs[0], g['F']) <= TOL


def _random_cut(seed, nelems, degree, keep=.6, maxpts=40):
    'random element subset with random ragged point sets (positive weights), pruned numbering derived like PrunedBasis does'
    rng = numpy.random.RandomState(seed)
    nd = len(nelems)
    b1 = util.bases_1d(nelems, degree, 'spline')
    ntot = int(numpy.prod(nelems))
    elem_ids = numpy.sort(rng.choice(ntot, size=max(1, int(keep * ntot)), replace=False))
    npts = rng.randint(0, maxpts, size=len(elem_ids))     # empty point sets are legal (fully trimmed away but kept)
    qoff = numpy.concatenate([[0], numpy.cumsum(npts)])
    qcoords = rng.rand(qoff[-1], nd)
    qweights = rng.rand(qoff[-1]) / maxpts
    rules = points.tensor_gauss(nd, 2 * degree)
    verts = [numpy.sort(rng.rand(n + 1)) + numpy.arange(n + 1) for n in nelems]
    X = numpy.stack(numpy.meshgrid(*verts, indexing='ij'))
    X = X + .25 * (rng.rand(*X.shape) - .5)
    prob = fem_oracle.Problem(nelems, [degree] * nd, [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1], [b.ndofs for b in b1],
                              [r[0] for r in rules], [r[1] for r in rules], X)
    # PrunedBasis (function.py:3121-3122): sorted unique parent dofs of the kept elements, inverse map with missing = len
    prob.elem_ids = None
    dofs = numpy.unique(numpy.concatenate([fem_oracle.element_data(prob, int(e))[0] for e in elem_ids]))
    renumber = numpy.full(int(numpy.prod(prob.ndofs_d)), len(dofs), dtype=numpy.int64)
    renumber[dofs] = numpy.arange(len(dofs))
    return fem_oracle.Problem(nelems, [degree] * nd, [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1], [b.ndofs for b in b1],
                              [r[0] for r in rules], [r[1] for r in rules], X, elem_ids=elem_ids, qoff=qoff, qcoords=qcoords, qweights=qweights,
                              renumber=renumber, nbasis_new=len(dofs))


@pytest.mark.parametrize('seed,nelems,degree', [(0, (5, 4, 3), 2), (1, (7, 6), 3), (2, (9,), 2), (3, (4, 3, 4), 1), (4, (5, 5), 4), (5, (3, 3, 3), 3)])
def test_random_cut_against_oracle(ctx, seed, nelems, degree, mma):
    prob = _random_cut(seed, nelems, degree)
    nd = prob.ndims
    rng = numpy.random.RandomState(100 + seed)
    D = rng.rand(1, nd + 1, 1, nd + 1)        # a full coefficient tensor: values, gradients and the mixed terms
    C = rng.rand(1, nd + 1)
    mats, vecs = fem_oracle.assemble(prob, [('generic', D), ('mass',)], [('generic', C)])
    plan = _plan(ctx, prob)
    rowptr, colidx = plan.csr_pattern()
    assert numpy.array_equal(rowptr, mats[0][1]) and numpy.array_equal(colidx, mats[0][2])
    vals, rhs = plan.assemble_host([D, engine.form_mass(nd)], [C])
    for v, (ref, _, _) in zip(vals, mats):
        assert util.relerr(v, ref) <= TOL
        assert util.rowsum_relerr(v, ref, rowptr) <= TOL
    assert util.relerr(rhs[0], vecs[0]) <= TOL


def test_device_accumulate_in_two_halves(ctx):
    # b2_assemble_elemset_device accumulates: two halves of the selection add up to the whole (what multi-GPU sharding uses)
    import torch
    prob = _random_cut(7, (5, 4, 4), 2)
    plan = _plan(ctx, prob)
    nd = prob.ndims
    Ds, Cs = [engine.form_stiffness(nd)], [engine.form_load(nd)]
    whole, rhs_whole = plan.assemble_host(Ds, Cs)
    dev = torch.device('cuda', 0)
    v = torch.zeros(plan.nnz, dtype=torch.float64, device=dev)
    r = torch.zeros(plan.ndofs, dtype=torch.float64, device=dev)
    half = plan.nsel // 2
    plan.assemble_device(Ds, Cs, [v], [r], sel_range=(0, half))
    plan.assemble_device(Ds, Cs, [v], [r], sel_range=(half, plan.nsel))
    torch.cuda.synchronize()
    assert util.relerr(v.cpu().numpy(), whole[0]) <= TOL
    assert util.relerr(r.cpu().numpy(), rhs_whole[0]) <= TOL


def test_invalid_arguments(ctx):
    from nutils_b200 import _lib
    b1 = util.bases_1d((3, 3), 2, 'spline')
    nodes = numpy.stack(numpy.meshgrid(numpy.arange(4.), numpy.arange(4.), indexing='ij'))
    rules = points.tensor_gauss(2, 4)
    with pytest.raises(_lib.B200Error):   # elem_ids not increasing
        engine.ElemSetPlan(ctx, b1, nodes=nodes, rules=rules, elem_ids=[3, 1])
    with pytest.raises(_lib.B200Error):   # renumber not monotone
        ren = numpy.arange(25)[::-1].copy()
        engine.ElemSetPlan(ctx, b1, nodes=nodes, rules=rules, renumber=ren, nbasis_new=25)
    with pytest.raises(_lib.B200Error):   # rational mode 2 without a rational geometry
        plan = engine.ElemSetPlan(ctx, b1, nodes=nodes, rules=rules, scale=numpy.ones(25), rational=2)
        plan.assemble_host([engine.form_mass(2)], [])


@pytest.mark.parametrize('n,degree,depth', [(20, 2, 2), (12, 3, 1), (10, 4, 1)])
def test_finite_cell_ball_properties(ctx, n, degree, depth):
    # size-independent properties on a synthetic finite-cell workload (octree quadrature of nutils_b200.fcm):
    # partition of unity survives trimming and pruning: sum(M) = sum(f) = quadrature volume, K has zero row sums and is symmetric
    from nutils_b200 import fcm
    elem_ids, qoff, qc, qw, ren, nbn = fcm.octree_ball(n, degree, depth)
    b1 = util.bases_1d((n,) * 3, degree, 'spline')
    v = numpy.linspace(-1, 1, n + 1)
    nodes = numpy.stack(numpy.meshgrid(v, v, v, indexing='ij'))
    plan = engine.ElemSetPlan(ctx, b1, nodes=nodes, elem_ids=elem_ids, qoff=qoff, qcoords=qc, qweights=qw, renumber=ren, nbasis_new=nbn)
    (K, M), (f,) = plan.assemble_host([engine.form_stiffness(3), engine.form_mass(3)], [engine.form_load(3)])
    vol = qw.sum() * (2. / n) ** 3
    assert abs(M.sum() - vol) <= 1e-12 * vol and abs(f.sum() - vol) <= 1e-12 * vol
    assert abs(vol - 4 / 3 * numpy.pi * .8 ** 3) < .02 * vol
    rowptr, colidx = plan.csr_pattern()
    import scipy.sparse
    A = scipy.sparse.csr_matrix((K, colidx, rowptr), shape=(plan.ndofs,) * 2)
    assert abs(A @ numpy.ones(plan.ndofs)).max() <= 1e-11 * abs(K).max()
    assert abs(A - A.T).max() <= 1e-12 * abs(K).max()
    # every kept function has support on a kept element: no empty rows
    assert (numpy.diff(rowptr) > 0).all()


@pytest.mark.parametrize('name', util.eval_golden_names())
def test_evaluate_golden(ctx, name):
    # b2_evaluate_elemset_device against Sample.eval of the reference
    g = util.load_golden(name)
    prob = util.elemset_problem_from_golden(g)
    plan = _plan(ctx, prob)
    out = plan.evaluate([g['coefs']], grads=True)
    assert util.relerr(out['x'], g['x']) <= TOL and util.relerr(out['weights'], g['wdet']) <= TOL
    assert util.relerr(out['values'][:, 0].reshape(g['values'].shape), g['values']) <= TOL
    assert util.relerr(out['grads'][:, 0].reshape(g['grads'].shape), g['grads']) <= TOL


@pytest.mark.parametrize('name', ['fcm3d_ball_p2', 'nurbs_plate_p4', 'bnd3d_right_p2_warp', 'fcm2d_plate_p2', 'bnd1d_right_p2'])
def test_evaluate_against_oracle(ctx, name):
    # ragged cut-cell points, rational functions on a NURBS geometry, boundary faces (weights = surface measure), two fields
    g = util.load_golden(name)
    prob = util.elemset_problem_from_golden(g)
    plan = _plan(ctx, prob)
    rng = numpy.random.RandomState(3)
    fields = [rng.rand(plan.ndofs) - .5, numpy.ones(plan.ndofs)]
    x, w, v, gr = fem_oracle.evaluate(prob, fields)
    out = plan.evaluate(fields, grads=True)
    assert util.relerr(out['x'], x) <= TOL and util.relerr(out['weights'], w) <= TOL
    assert util.relerr(out['values'], v) <= TOL
    assert abs(out['grads'] - gr).max() <= 1e-10 * max(abs(gr).max(), 1.)
    if not prob.rational or True:
        # partition of unity: the all-ones field evaluates to 1 with zero gradient (NURBS included)
        assert abs(out['values'][:, 1] - 1.).max() <= 1e-12
        assert abs(out['grads'][:, 1]).max() <= 1e-9
    # weights only (no fields): the measure of the sample
    assert abs(plan.evaluate(x=False)['weights'].sum() - w.sum()) <= 1e-12 * abs(w.sum())
