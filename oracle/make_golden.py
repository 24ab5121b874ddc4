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
    'elast3d_p2_warp': lambda: elasticity_case('elast3d_p2_warp', (3, 3, 2), 2, warp=.3, seed=7),  # config 3 at toy size
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
        print('{:20s} ndofs={:6d} nnz={:8d} {:8.1f} kB'.format(name, int(d['ndofs']), len(d['colidx']), os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
