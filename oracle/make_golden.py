#!/usr/bin/env python
'''Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE.  Usage (in the build container, where /root/reference exists):

    python oracle/make_golden.py [--only NAME]

The reference (evalf/nutils @ 37d1cc5, /root/reference/src) is imported with the
dependency stand-ins of oracle/shims on the path (treelog/ags/stringly/appdirs are
no-ops; nutils_poly is restated in oracle/shims/nutils_poly.py).  Every case goes
through the reference's own public API:

    mesh.rectilinear -> topo.basis / topo.field -> topo.integral(...)
    -> function.eval(function.as_csr(...))          (function.py:2432-2452)
    or solver.System(...).assemble_jacobian_residual  (solver.py:357-386)

and the resulting (values, rowptr, colidx) / rhs are stored next to the plain
inputs (vertex arrays, degree, quadrature degree, form parameters) that the
oracle and the CUDA path consume.  /root/reference does not exist on the GPU box;
the committed .npz files are what travels.
'''

import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('B200FEM_REFERENCE', '/root/reference/src')
sys.path[:0] = [os.path.join(HERE, 'shims'), REF]
os.environ.setdefault('NUTILS_MATRIX', 'scipy')

import numpy  # noqa: E402
from nutils import mesh, function, evaluable  # noqa: E402
from nutils.solver import System  # noqa: E402
from nutils.expression_v2 import Namespace  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def _verts(n, kind, seed):
    'per-dimension vertex arrays'
    rng = numpy.random.RandomState(seed)
    out = []
    for k, ni in enumerate(n):
        if kind == 'uniform':
            v = numpy.linspace(0, 1, ni + 1)
        elif kind == 'graded':
            v = numpy.linspace(0, 1, ni + 1) ** (1.3 + .2 * k) * (1 + .5 * k)
        else:
            raise ValueError(kind)
        out.append(v)
    return out


def _nodes(verts, warp, seed):
    'nodal coordinates (ndims, *shape) of a (possibly warped) multilinear geometry'
    X = numpy.stack(numpy.meshgrid(*verts, indexing='ij'))
    if warp:
        rng = numpy.random.RandomState(seed)
        h = min(numpy.diff(v).min() for v in verts)
        X = X + warp * h * (rng.rand(*X.shape) - .5)
    return X


def _geom(topo, X):
    # same construction as mesh.rectilinear for non-integer vertices (mesh.py:55-57)
    funcsp = topo.basis('spline', degree=1)
    return (funcsp * X.reshape(X.shape[0], -1)).sum(-1)


def scalar_case(name, n, degree, btype='spline', qdegree=None, vkind='graded', warp=0., seed=0):
    ndims = len(n)
    verts = _verts(n, vkind, seed)
    X = _nodes(verts, warp, seed)
    topo, geom0 = mesh.rectilinear(verts)
    geom = _geom(topo, X) if warp else geom0
    basis = topo.basis(btype, degree=degree)
    qd = 2 * degree if qdegree is None else qdegree
    J = function.J(geom)
    g = basis.grad(geom)
    K = topo.integral((g[:, None, :] * g[None, :, :]).sum(-1) * J, degree=qd)
    M = topo.integral(basis[:, None] * basis[None, :] * J, degree=qd)
    F = topo.integral(basis * J, degree=qd)
    (kv, krp, kci), (mv, mrp, mci), f = function.eval((function.as_csr(K), function.as_csr(M), F))
    assert (krp == mrp).all() and (kci == mci).all()
    return dict(kind='scalar', name=name, ndims=ndims, nelems=numpy.array(n), degree=degree, btype=btype, qdegree=qd,
                nodes=X, ndofs=len(basis), K_values=kv, M_values=mv, rowptr=krp, colidx=kci, F=f)


def elasticity_case(name, n, degree, lmbda=1., poisson=.3, vkind='graded', warp=0., seed=0):
    # the 2-D examples/elasticity.py:43-58 energy, in len(n) dimensions
    ndims = len(n)
    verts = _verts(n, vkind, seed)
    X = _nodes(verts, warp, seed)
    topo, geom0 = mesh.rectilinear(verts)
    geom = _geom(topo, X) if warp else geom0
    mu = .5 / poisson - 1
    ns = Namespace()
    ns.δ = function.eye(ndims)
    ns.x = geom
    ns.define_for('x', gradient='∇', jacobians=('dV',))
    ns.u = topo.field('u', btype='spline', degree=degree, shape=[ndims])
    ns.λ = lmbda
    ns.μ = mu
    ns.ε_ij = '.5 (∇_i(u_j) + ∇_j(u_i))'
    ns.σ_ij = 'λ ε_kk δ_ij + 2 μ ε_ij'
    ns.E = 'ε_ij σ_ij'
    ns.q_i = '-δ_i{}'.format(ndims - 1)
    energy = topo.integral('(E - u_i q_i) dV' @ ns, degree=2 * degree)
    system = System(energy, trial='u')
    nbasis = len(topo.basis('spline', degree=degree))
    jac, res = system.assemble_jacobian_residual(arguments={'u': numpy.zeros((nbasis, ndims))})[:2]
    data, indices, indptr = jac.export('csr')
    return dict(kind='elasticity', name=name, ndims=ndims, nelems=numpy.array(n), degree=degree, btype='spline', qdegree=2 * degree,
                nodes=X, ndofs=nbasis * ndims, lmbda=lmbda, mu=mu,
                K_values=numpy.asarray(data), rowptr=numpy.asarray(indptr, dtype=numpy.int64), colidx=numpy.asarray(indices, dtype=numpy.int64),
                F=numpy.asarray(res).ravel())


# ---- element sets: trimmed topologies (finite cell method) and NURBS -----------------------------------

def _ragged_points(topo, qdegree):
    'per-element points of topo.sample("gauss", qdegree) in element-local coordinates: qoff, qcoords, qweights'
    smp = topo.sample('gauss', qdegree)
    coords, weights, qoff = [], [], [0]
    for i in range(len(topo)):
        p = smp.points.get(i)
        coords.append(numpy.asarray(p.coords))
        weights.append(numpy.asarray(p.weights))
        qoff.append(qoff[-1] + len(weights[-1]))
    return numpy.array(qoff, dtype=numpy.int64), numpy.concatenate(coords), numpy.concatenate(weights)


def _elasticity_system(topo, geom, basis, ndims, lmbda, mu, qdegree, load):
    'jacobian and residual (at u=0) of  int (grad_j(v_i) sigma_ij - v_i q_i) dV  (examples/platewithhole.py:118-151)'
    ns = Namespace()
    ns.δ = function.eye(ndims)
    ns.x = geom
    ns.define_for('x', gradient='∇', jacobians=('dV',))
    ns.λ = lmbda
    ns.μ = mu
    ns.u = function.field('u', basis, shape=[ndims])
    ns.v = function.field('v', basis, shape=[ndims])
    ns.ε_ij = '(∇_j(u_i) + ∇_i(u_j)) / 2'
    ns.σ_ij = 'λ ε_kk δ_ij + 2 μ ε_ij'
    ns.q = numpy.asarray(load, dtype=float)
    res = topo.integral('(∇_j(v_i) σ_ij - v_i q_i) dV' @ ns, degree=qdegree)
    system = System(res, trial='u', test='v')
    jac, r = system.assemble_jacobian_residual(arguments={'u': numpy.zeros((len(basis), ndims))})[:2]
    data, indices, indptr = jac.export('csr')
    return numpy.asarray(data), numpy.asarray(indptr, dtype=numpy.int64), numpy.asarray(indices, dtype=numpy.int64), numpy.asarray(r).ravel()


def fcm_scalar_case(name, n, degree, radius, maxrefine, ball=True):
    'config 5 at toy size: Poisson on a ball (or a box with a spherical hole) cut out of a structured grid, spline basis'
    ndims = len(n)
    verts = [numpy.linspace(-1, 1, ni + 1) for ni in n]
    topo0, geom = mesh.rectilinear(verts)
    levelset = radius - numpy.linalg.norm(geom) if ball else numpy.linalg.norm(geom) - radius
    topo = topo0.trim(levelset, maxrefine=maxrefine)
    basis = topo.basis('spline', degree=degree)
    qd = 2 * degree
    J = function.J(geom)
    g = basis.grad(geom)
    K = topo.integral((g[:, None, :] * g[None, :, :]).sum(-1) * J, degree=qd)
    M = topo.integral(basis[:, None] * basis[None, :] * J, degree=qd)
    F = topo.integral(basis * J, degree=qd)
    (kv, krp, kci), (mv, mrp, mci), f = function.eval((function.as_csr(K), function.as_csr(M), F))
    assert (krp == mrp).all() and (kci == mci).all()
    qoff, qcoords, qweights = _ragged_points(topo, qd)
    return dict(kind='elemset_scalar', name=name, ndims=ndims, nelems=numpy.array(n), degree=degree, btype='spline', qdegree=qd,
                nodes=numpy.stack(numpy.meshgrid(*verts, indexing='ij')), ndofs=len(basis), ncomp=1,
                elem_ids=numpy.asarray(basis._transmap, dtype=numpy.int64), renumber=numpy.asarray(basis._renumber, dtype=numpy.int64), nbasis_new=len(basis),
                qoff=qoff, qcoords=qcoords, qweights=qweights, volume=function.eval(topo.integral(J, degree=qd)),
                K_values=kv, M_values=mv, rowptr=krp, colidx=kci, F=f)


def fcm_plate_case(name, nelems=4, degree=2, maxrefine=2, radius=.5, poisson=.3):
    'examples/platewithhole.py FCM mode (its own unit test uses nelems=4, spline): 2-D plane-strain elasticity on a trimmed square'
    topo0, geom = mesh.unitsquare(nelems, 'square')
    topo = topo0.trim(numpy.linalg.norm(geom) - radius, maxrefine=maxrefine, name='hole')
    basis = topo.basis('spline', degree=degree)
    lmbda, mu = 2 * poisson, 1 - poisson
    qd = 2 * degree
    kv, rp, ci, r = _elasticity_system(topo, geom, basis, 2, lmbda, mu, qd, load=[.3, -1.])
    qoff, qcoords, qweights = _ragged_points(topo, qd)
    verts = [numpy.linspace(0, 1, nelems + 1)] * 2
    return dict(kind='elemset_elasticity', name=name, ndims=2, nelems=numpy.array([nelems] * 2), degree=degree, btype='spline', qdegree=qd,
                nodes=numpy.stack(numpy.meshgrid(*verts, indexing='ij')), ndofs=2 * len(basis), ncomp=2, lmbda=lmbda, mu=mu, load=numpy.array([.3, -1.]),
                elem_ids=numpy.asarray(basis._transmap, dtype=numpy.int64), renumber=numpy.asarray(basis._renumber, dtype=numpy.int64), nbasis_new=len(basis),
                qoff=qoff, qcoords=qcoords, qweights=qweights,
                K_values=kv, rowptr=rp, colidx=ci, F=r)


def nurbs_plate_case(name, nrefine=2, degree=2, radius=.5, poisson=.3, qdegree=10):
    '''examples/platewithhole.py NURBS mode (:66-86): a 1x2 quadratic NURBS patch mapped onto the quarter plate with hole,
    refined nrefine times; analysis basis = B-splines of `degree` on the refined topology times projected control weights,
    divided by the patch's weight function.  degree=2 is the example itself, degree=4 is BASELINE config 4.'''
    topo, geom0 = mesh.rectilinear([1, 2])
    bsplinebasis = topo.basis('spline', degree=2)
    controlweights = numpy.ones(12)
    controlweights[1:3] = .5 + .25 * numpy.sqrt(2)
    weightfunc = bsplinebasis @ controlweights
    nurbsbasis = bsplinebasis * controlweights / weightfunc
    A = 0, 0, 0
    B = (2**.5 - 1) * radius, .3 * (radius + 1) / 2, 1
    C = radius, (radius + 1) / 2, 1
    controlpoints = numpy.array([[A, B, C, C], [C, C, B, A]]).T.reshape(-1, 2)
    geom = nurbsbasis @ controlpoints
    topo = topo.refine(nrefine)

    def project(target, pbasis):
        sqr = topo.integral((function.field('w', pbasis) - target)**2, degree=9)
        return System(sqr, trial='w').solve()['w']
    # the patch's weight function and weighted coordinates in the quadratic B-splines of the refined topology (nested
    # spaces: the projection is exact) -- this is the representation handed to b2_geom_create_spline
    fine2 = topo.basis('spline', degree=2)
    gweights = project(weightfunc, fine2)
    gctrl = numpy.stack([project(weightfunc * geom[i], fine2) for i in range(2)]) / gweights
    # analysis basis, as in the example (:79-83)
    abasis = topo.basis('spline', degree=degree)
    cw = project(weightfunc, abasis)
    nbasis = abasis * cw / weightfunc
    lmbda, mu = 2 * poisson, 1 - poisson
    kv, rp, ci, r = _elasticity_system(topo, geom, nbasis, 2, lmbda, mu, qdegree, load=[.3, -1.])
    return dict(kind='elemset_elasticity', name=name, ndims=2, nelems=numpy.array(topo.shape), degree=degree, btype='spline', qdegree=qdegree,
                ndofs=2 * len(abasis), ncomp=2, lmbda=lmbda, mu=mu, load=numpy.array([.3, -1.]),
                scale=cw, rational=2, gdegree=2, gctrl=gctrl, gweights=gweights,
                K_values=kv, rowptr=rp, colidx=ci, F=r)


# ---- boundary integrals and pointwise coefficients --------------------------------------------------------

BNAMES = {1: ['left', 'right'], 2: ['left', 'right', 'bottom', 'top'], 3: ['left', 'right', 'bottom', 'top', 'front', 'back']}


def _face_tables(shape, bname, qdegree):
    '''volume elements adjacent to a boundary of a structured grid and the face Gauss points in their local coordinates:
    elem_ids (increasing), face_dim, qoff, qcoords, qweights'''
    from nutils import points as refpoints
    ndims = len(shape)
    k = BNAMES[ndims].index(bname)
    dim, side = k // 2, k % 2
    grids = [numpy.arange(n) for n in shape]
    grids[dim] = numpy.array([shape[dim] - 1 if side else 0])
    idx = numpy.stack(numpy.meshgrid(*grids, indexing='ij'), -1).reshape(-1, ndims)
    elem_ids = numpy.sort(numpy.ravel_multi_index(idx.T, shape))
    rx, rw = refpoints.gauss1(qdegree)
    x1, w1 = numpy.asarray(rx)[:, 0], numpy.asarray(rw)
    tang = [d for d in range(ndims) if d != dim]
    if tang:
        mesh_ = numpy.stack(numpy.meshgrid(*[x1] * len(tang), indexing='ij'), -1).reshape(-1, len(tang))
        wts = numpy.ones(len(mesh_))
        for j in range(len(tang)):
            wts = wts * w1[numpy.stack(numpy.meshgrid(*[numpy.arange(len(x1))] * len(tang), indexing='ij'), -1).reshape(-1, len(tang))[:, j]]
    else:
        mesh_, wts = numpy.zeros((1, 0)), numpy.ones(1)
    xi = numpy.empty((len(mesh_), ndims))
    xi[:, dim] = float(side)
    for j, d in enumerate(tang):
        xi[:, d] = mesh_[:, j]
    nq = len(xi)
    return (elem_ids, numpy.full(len(elem_ids), dim, dtype=numpy.int8), numpy.arange(len(elem_ids) + 1, dtype=numpy.int64) * nq,
            numpy.tile(xi, (len(elem_ids), 1)), numpy.tile(wts, len(elem_ids)))


def _physical_points(nodes, shape, elem_ids, qoff, qcoords):
    'x at the points: multilinear map of the nodal geometry'
    ndims = len(shape)
    out = numpy.empty_like(qcoords)
    for k, e in enumerate(elem_ids):
        idx = numpy.unravel_index(e, shape)
        xi = qcoords[qoff[k]:qoff[k + 1]]
        X = nodes[(slice(None),) + tuple(slice(i, i + 2) for i in idx)]  # (ndims, 2, 2, ...)
        acc = X
        for d in range(ndims):   # contract one reference direction at a time
            x = xi[:, d]
            if d == 0:
                acc = acc[:, 0][None] * (1 - x).reshape((-1,) + (1,) * (acc.ndim - 1)) + acc[:, 1][None] * x.reshape((-1,) + (1,) * (acc.ndim - 1))
            else:
                acc = acc[:, :, 0] * (1 - x).reshape((-1,) + (1,) * (acc.ndim - 2)) + acc[:, :, 1] * x.reshape((-1,) + (1,) * (acc.ndim - 2))
        out[qoff[k]:qoff[k + 1]] = acc
    return out


def _coef_g(x):
    'the coefficient function of the boundary / variable-coefficient goldens, numpy and nutils spelling'
    ndims = x.shape[-1]
    return 1. + .5 * numpy.prod(x, axis=-1) + numpy.cosh(x[..., 0])


def boundary_case(name, n, degree, bname, btype='spline', warp=0., seed=0):
    '''boundary mass matrix and boundary load vector with a coefficient function, the two ingredients of Neumann terms and of the
    boundary projections of solve_constraints (examples/laplace.py:60-86): int_G g N_i N_j dS, int_G g N_i dS'''
    ndims = len(n)
    verts = _verts(n, 'graded', seed)
    X = _nodes(verts, warp, seed)
    topo, geom0 = mesh.rectilinear(verts)
    geom = _geom(topo, X) if warp else geom0
    basis = topo.basis(btype, degree=degree)
    qd = 2 * degree
    btopo = topo.boundary[bname]
    g = 1. + .5 * numpy.prod(geom, axis=0) + numpy.cosh(geom[0])
    J = function.J(geom)
    B = btopo.integral(basis[:, None] * basis[None, :] * g * J, degree=qd)
    F = btopo.integral(basis * g * J, degree=qd)
    area = function.eval(btopo.integral(J, degree=qd))
    (bv, rp, ci), f = function.eval((function.as_csr(B), F))
    elem_ids, face_dim, qoff, qcoords, qweights = _face_tables(tuple(n), bname, qd)
    xq = _physical_points(X, tuple(n), elem_ids, qoff, qcoords)
    return dict(kind='elemset_boundary', name=name, ndims=ndims, nelems=numpy.array(n), degree=degree, btype=btype, qdegree=qd, nodes=X, ndofs=len(basis), ncomp=1,
                elem_ids=elem_ids, face_dim=face_dim, qoff=qoff, qcoords=qcoords, qweights=qweights, coef=_coef_g(xq), area=area,
                M_values=bv, rowptr=rp, colidx=ci, F=f)


def varcoef_case(name, n, degree, warp=.25, seed=0):
    'volume integrals with a coefficient function: int g grad N_i . grad N_j dV, int g N_i dV (SURVEY 8d: variable coefficient)'
    ndims = len(n)
    verts = _verts(n, 'graded', seed)
    X = _nodes(verts, warp, seed)
    topo, geom0 = mesh.rectilinear(verts)
    geom = _geom(topo, X) if warp else geom0
    basis = topo.basis('spline', degree=degree)
    qd = 2 * degree
    g = 1. + .5 * numpy.prod(geom, axis=0) + numpy.cosh(geom[0])
    J = function.J(geom)
    gr = basis.grad(geom)
    K = topo.integral((gr[:, None, :] * gr[None, :, :]).sum(-1) * g * J, degree=qd)
    F = topo.integral(basis * g * J, degree=qd)
    (kv, rp, ci), f = function.eval((function.as_csr(K), F))
    from nutils import points as refpoints
    rx, rw = refpoints.gauss1(qd)
    x1 = numpy.asarray(rx)[:, 0]
    xi = numpy.stack(numpy.meshgrid(*[x1] * ndims, indexing='ij'), -1).reshape(-1, ndims)
    nel = int(numpy.prod(n))
    qoff = numpy.arange(nel + 1, dtype=numpy.int64) * len(xi)
    xq = _physical_points(X, tuple(n), numpy.arange(nel), qoff, numpy.tile(xi, (nel, 1)))
    return dict(kind='elemset_varcoef', name=name, ndims=ndims, nelems=numpy.array(n), degree=degree, btype='spline', qdegree=qd, nodes=X, ndofs=len(basis), ncomp=1,
                coef=_coef_g(xq), K_values=kv, rowptr=rp, colidx=ci, F=f)


def immersed_boundary_case(name, n, degree, radius, maxrefine, ball=True):
    '''integrals over the TRIMMED boundary of a finite-cell topology (Nitsche / penalty terms, immersed Neumann data): boundary mass
    matrix and load with a coefficient function, and the surface area, on the sphere / circle cut out of a structured grid'''
    sys.path.insert(0, os.path.dirname(HERE))
    from nutils_b200 import adapter
    ndims = len(n)
    verts = [numpy.linspace(-1, 1, ni + 1) for ni in n]
    topo0, geom = mesh.rectilinear(verts)
    levelset = radius - numpy.linalg.norm(geom) if ball else numpy.linalg.norm(geom) - radius
    topo = topo0.trim(levelset, maxrefine=maxrefine, name='trimmed')
    bt = topo.boundary['trimmed']
    basis = topo.basis('spline', degree=degree)
    qd = 2 * degree
    g = 1. + .5 * numpy.prod(geom, axis=0) + numpy.cosh(geom[0])
    J = function.J(geom)
    B = bt.integral(basis[:, None] * basis[None, :] * g * J, degree=qd)
    F = bt.integral(basis * g * J, degree=qd)
    area = function.eval(bt.integral(J, degree=qd))
    (bv, rp, ci), f = function.eval((function.as_csr(B), F))
    elem_ids, qoff, qcoords, qweights, normals = adapter.immersed_boundary_tables_from_reference(bt, topo0, qd)
    X = numpy.stack(numpy.meshgrid(*verts, indexing='ij'))
    xq = _physical_points(X, tuple(n), elem_ids, qoff, qcoords)
    return dict(kind='elemset_boundary', name=name, ndims=ndims, nelems=numpy.array(n), degree=degree, btype='spline', qdegree=qd, nodes=X, ndofs=len(basis), ncomp=1,
                elem_ids=elem_ids, qoff=qoff, qcoords=qcoords, qweights=qweights, normals=normals, coef=_coef_g(xq), area=area,
                renumber=numpy.asarray(basis._renumber, dtype=numpy.int64), nbasis_new=len(basis),
                M_values=bv, rowptr=rp, colidx=ci, F=f)


def eval_case(name, n, degree, ncomp=1, warp=.25, seed=0):
    'Sample.eval of the reference (sample.py:192-215): coordinates, w |det J|, a discrete field and its gradient at the Gauss points'
    ndims = len(n)
    verts = _verts(n, 'graded', seed)
    X = _nodes(verts, warp, seed)
    topo, geom0 = mesh.rectilinear(verts)
    geom = _geom(topo, X) if warp else geom0
    basis = topo.basis('spline', degree=degree)
    rng = numpy.random.RandomState(seed + 100)
    coefs = rng.rand(len(basis), ncomp) - .5
    u = (basis[:, None] * coefs).sum(0)                        # [ncomp]
    qd = 2 * degree
    smp = topo.sample('gauss', qd)
    x, w, val, grad = smp.eval([geom, function.J(geom), u, function.grad(u, geom)])
    # the reference multiplies J by the point weights only inside integrals: do it here
    weights = numpy.concatenate([numpy.asarray(smp.points.get(i).weights) for i in range(len(topo))])
    return dict(kind='elemset_eval', name=name, ndims=ndims, nelems=numpy.array(n), degree=degree, btype='spline', qdegree=qd, nodes=X, ndofs=len(basis) * ncomp, ncomp=ncomp,
                coefs=coefs.ravel(), x=x, wdet=w * weights, values=val, grads=grad)


def example_elasticity_case(name, nelems, degree, btype='std', poisson=.25):
    'examples/elasticity.py:43-61 (plane strain plate clamped at the top under gravity): constraints and displacement solution'
    domain, geom = mesh.unitsquare(nelems, 'square')
    ns = Namespace()
    ns.δ = function.eye(domain.ndims)
    ns.x = geom
    ns.define_for('x', gradient='∇', normal='n', jacobians=('dV', 'dS'))
    ns.u = domain.field('u', btype=btype, degree=degree, shape=[2])
    ns.λ = 1
    ns.μ = .5 / poisson - 1
    ns.ε_ij = '.5 (∇_i(u_j) + ∇_j(u_i))'
    ns.σ_ij = 'λ ε_kk δ_ij + 2 μ ε_ij'
    ns.E = 'ε_ij σ_ij'
    ns.q_i = '-δ_i1'
    sqr = domain.boundary['top'].integral('u_k u_k dS' @ ns, degree=degree * 2)
    cons = System(sqr, trial='u').solve_constraints(droptol=1e-15)
    energy = domain.integral('(E - u_i q_i) dV' @ ns, degree=degree * 2)
    args = System(energy, trial='u').solve(constrain=cons)
    return dict(kind='example_elasticity', name=name, nelems=nelems, degree=degree, btype=btype, poisson=poisson,
                cons=numpy.asarray(cons['u']), u=numpy.asarray(args['u']), ndofs=numpy.asarray(args['u']).size)


def known_answer_mass_1d():
    # tests/test_function.py:1574-1585 (known-answer COO of a 1-D p=1 mass matrix)
    topo, geom = mesh.line([0, 1, 2], bnames=['a', 'b'], space='X')
    basis = topo.basis('std', degree=1)
    M = topo.integral(basis[:, None] * basis[None, :] * function.J(geom), degree=2)
    v, rp, ci = function.eval(function.as_csr(M))
    F = function.eval(topo.integral(basis * function.J(geom), degree=2))
    X = numpy.array([[0., 1., 2.]])
    return dict(kind='scalar', name='mass1d_known', ndims=1, nelems=numpy.array([2]), degree=1, btype='std', qdegree=2,
                nodes=X, ndofs=3, K_values=numpy.zeros(0), M_values=v, rowptr=rp, colidx=ci, F=F)



# ---- general dof maps: simplex and mixed meshes ------------------------------------------------------

def general_case(name, nelems, etype, degree, ncomp=1):
    '''mesh.unitsquare(nelems, 'triangle' | 'mixed') (mesh.py:686-783) with basis('std') (topology.py:2493): K, M, f through
    function.eval(as_csr(...)); next to them the element tables read out of the reference's objects by
    nutils_b200.adapter.general_tables_from_reference -- the inputs of engine.GeneralPlan / fem_oracle.assemble_general'''
    sys.path.insert(0, os.path.dirname(HERE))
    from nutils_b200 import adapter
    topo, geom = mesh.unitsquare(nelems, etype)
    basis = topo.basis('std', degree=degree)
    J = function.J(geom)
    qd = 2 * degree
    t = adapter.general_tables_from_reference(topo, basis, geom, qd)
    d = dict(kind='general', name=name, ndims=2, degree=degree, qdegree=qd, ncomp=ncomp, nbasis=len(basis), ndofs=len(basis) * ncomp, ntypes=len(t['types']),
             etype=t['etype'], dofoff=numpy.concatenate([[0], numpy.cumsum([len(x) for x in t['dofs']])]), dofs=numpy.concatenate(t['dofs']),
             vertoff=numpy.concatenate([[0], numpy.cumsum([len(x) for x in t['verts']])]), vertcoords=numpy.concatenate(t['verts']))
    for k, ty in enumerate(t['types']):
        for key in ('weights', 'phi', 'dphi', 'gdphi'):
            d['{}_{}'.format(key, k)] = ty[key]
    if ncomp == 1:
        g = basis.grad(geom)
        K = topo.integral((g[:, None, :] * g[None, :, :]).sum(-1) * J, degree=qd)
        M = topo.integral(basis[:, None] * basis[None, :] * J, degree=qd)
        F = topo.integral(basis * J, degree=qd)
        (kv, rp, ci), (mv, mrp, mci), f = function.eval((function.as_csr(K), function.as_csr(M), F))
        assert numpy.array_equal(rp, mrp) and numpy.array_equal(ci, mci)
        d.update(K_values=kv, M_values=mv, rowptr=rp, colidx=ci, F=f)
    else:
        lmbda, mu = .6, .7
        kv, rp, ci, r = _elasticity_system(topo, geom, basis, ncomp, lmbda, mu, qd, load=[.3, -1.])
        d.update(K_values=kv, rowptr=rp, colidx=ci, F=r, lmbda=lmbda, mu=mu, load=numpy.array([.3, -1.]))
    return d


CASES = {
    'mass1d_known': known_answer_mass_1d,
    'line_p2': lambda: scalar_case('line_p2', (7,), 2),
    'line_p3_std': lambda: scalar_case('line_p3_std', (5,), 3, btype='std'),
    'laplace2d_p1_32': lambda: scalar_case('laplace2d_p1_32', (32, 32), 1, btype='std', vkind='uniform'),  # config 1 mesh/basis (examples/laplace.py)
    'quad_p2_warp': lambda: scalar_case('quad_p2_warp', (5, 4), 2, warp=.3, seed=1),
    'quad_p3': lambda: scalar_case('quad_p3', (4, 3), 3),
    'quad_p4_warp': lambda: scalar_case('quad_p4_warp', (3, 3), 4, warp=.2, seed=4),
    'hex_p1_warp': lambda: scalar_case('hex_p1_warp', (4, 3, 5), 1, warp=.3, seed=2),
    'hex_p2': lambda: scalar_case('hex_p2', (4, 3, 2), 2),
    'hex_p2_warp': lambda: scalar_case('hex_p2_warp', (5, 4, 6), 2, warp=.3, seed=3),   # config 2 at toy size, general geometry
    'hex_p2_uniform8': lambda: scalar_case('hex_p2_uniform8', (8, 8, 8), 2, vkind='uniform'),  # config 2 mesh at n=8
    'hex_p3_warp': lambda: scalar_case('hex_p3_warp', (3, 2, 3), 3, warp=.2, seed=5),
    'hex_p2_std': lambda: scalar_case('hex_p2_std', (3, 2, 2), 2, btype='std'),
    'hex_p2_underint': lambda: scalar_case('hex_p2_underint', (3, 3, 2), 2, qdegree=2),
    'elast2d_p2_warp': lambda: elasticity_case('elast2d_p2_warp', (4, 5), 2, warp=.3, seed=6),
    'elast3d_p1': lambda: elasticity_case('elast3d_p1', (3, 2, 3), 1),
    'fcm3d_ball_p2': lambda: fcm_scalar_case('fcm3d_ball_p2', (4, 4, 4), 2, radius=.8, maxrefine=1),                 # config 5 at toy size
    'fcm3d_hole_p1': lambda: fcm_scalar_case('fcm3d_hole_p1', (3, 4, 3), 1, radius=.7, maxrefine=2, ball=False),
    'fcm2d_disc_p3': lambda: fcm_scalar_case('fcm2d_disc_p3', (6, 5), 3, radius=.75, maxrefine=2),
    'fcm2d_plate_p2': lambda: fcm_plate_case('fcm2d_plate_p2'),                                                      # examples/platewithhole.py FCM, its test size
    'nurbs_plate_p2': lambda: nurbs_plate_case('nurbs_plate_p2', nrefine=2, degree=2),                               # examples/platewithhole.py NURBS
    'nurbs_plate_p4': lambda: nurbs_plate_case('nurbs_plate_p4', nrefine=2, degree=4),                               # config 4 at toy size
    'bnd2d_top_p1_std': lambda: boundary_case('bnd2d_top_p1_std', (8, 6), 1, 'top', btype='std'),                   # examples/laplace.py Neumann side
    'bnd2d_left_p2_warp': lambda: boundary_case('bnd2d_left_p2_warp', (5, 7), 2, 'left', warp=.3, seed=8),
    'bnd3d_right_p2_warp': lambda: boundary_case('bnd3d_right_p2_warp', (4, 3, 5), 2, 'right', warp=.3, seed=9),
    'bnd3d_front_p3': lambda: boundary_case('bnd3d_front_p3', (3, 4, 2), 3, 'front', warp=.2, seed=10),
    'bnd1d_right_p2': lambda: boundary_case('bnd1d_right_p2', (6,), 2, 'right'),
    'fcm3d_sphere_surface_p2': lambda: immersed_boundary_case('fcm3d_sphere_surface_p2', (4, 4, 4), 2, radius=.8, maxrefine=1),   # the immersed boundary of config 5
    'fcm2d_circle_boundary_p2': lambda: immersed_boundary_case('fcm2d_circle_boundary_p2', (6, 5), 2, radius=.75, maxrefine=2, ball=False),
    'varcoef3d_p2': lambda: varcoef_case('varcoef3d_p2', (4, 3, 3), 2, seed=11),
    'varcoef2d_p3': lambda: varcoef_case('varcoef2d_p3', (5, 4), 3, seed=12),
    'eval3d_p2': lambda: eval_case('eval3d_p2', (3, 4, 2), 2, seed=13),
    'eval2d_p3_vec': lambda: eval_case('eval2d_p3_vec', (4, 3), 3, ncomp=2, seed=14),
    'example_elasticity_p1': lambda: example_elasticity_case('example_elasticity_p1', 4, 1),    # the reference's own test sizes (examples/elasticity.py:96-135)
    'example_elasticity_p2': lambda: example_elasticity_case('example_elasticity_p2', 4, 2),
    'example_elasticity_spline': lambda: example_elasticity_case('example_elasticity_spline', 6, 2, btype='spline', poisson=.3),
    'elast3d_p2_warp': lambda: elasticity_case('elast3d_p2_warp', (3, 3, 2), 2, warp=.3, seed=7),  # config 3 at toy size
    'general_tri_p1': lambda: general_case('general_tri_p1', 4, 'triangle', 1),     # SimplexTopology.basis_std (topology.py:2493)
    'general_tri_p2': lambda: general_case('general_tri_p2', 4, 'triangle', 2),
    'general_mixed_p2': lambda: general_case('general_mixed_p2', 5, 'mixed', 2),   # mesh.py:737-753
    'general_tri_elast_p2': lambda: general_case('general_tri_elast_p2', 3, 'triangle', 2, ncomp=2),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    for name, f in CASES.items():
        if args.only and name != args.only:
            continue
        d = f()
        path = os.path.join(OUT, name + '.npz')
        numpy.savez_compressed(path, **d)
        print('{:20s} ndofs={:6d} nnz={:8d} {:8.1f} kB'.format(name, int(d['ndofs']), len(d['colidx']) if 'colidx' in d else 0, os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
