'''Host-side logic of the Nutils-style API that needs no device: the closed-form integrand algebra (function.Array),
coefficient functions (function.PointFunction), boundary topologies and the tables a boundary sample produces.'''

import numpy
import pytest

from nutils_b200 import mesh, function, engine


def _setup(n=(4, 3), btype='spline', degree=2):
    topo, geom = mesh.rectilinear([numpy.linspace(0, 1, k + 1) for k in n])
    return topo, geom, topo.basis(btype, degree=degree)


def test_form_tensors():
    topo, geom, basis = _setup()
    J = function.J(geom)
    g = basis.grad(geom)
    K = topo.integral((g[:, None, :] * g[None, :, :]).sum(-1) * J, degree=4)
    M = topo.integral(function.outer(basis) * J, degree=4)
    F = topo.integral(basis * J, degree=4)
    assert K.kind == M.kind == 'matrix' and F.kind == 'vector'
    assert numpy.array_equal(K.tensor, engine.form_stiffness(2)) and numpy.array_equal(M.tensor, engine.form_mass(2))
    assert numpy.array_equal(F.tensor, engine.form_load(2))
    assert numpy.array_equal((2 * K - M).tensor, 2 * engine.form_stiffness(2) - engine.form_mass(2))


def test_elasticity_tensor_from_symgrad():
    topo, geom, _ = _setup((3, 3, 2))
    u = topo.basis('spline', degree=2, shape=(3,))
    lm, mu = 1.3, .7
    eps = u.symgrad(geom)                                   # [dof, i, j]
    tr = eps.trace(-2, -1)
    energy = (lm * tr[:, None] * tr[None, :] + 2 * mu * (eps[:, None] * eps[None, :]).sum((-2, -1))) * function.J(geom)
    D = topo.integral(energy, degree=4).tensor
    assert numpy.allclose(D, engine.form_elasticity(3, lm, mu), atol=1e-15)


def test_point_functions():
    topo, geom, basis = _setup()
    x0, x1 = geom
    f = numpy.cos(1) * numpy.cosh(x1) + x0 ** 2 / (1 + x0 * x1) - 3
    x = numpy.random.RandomState(0).rand(7, 5, 2)
    assert numpy.allclose(f(x), numpy.cos(1) * numpy.cosh(x[..., 1]) + x[..., 0] ** 2 / (1 + x[..., 0] * x[..., 1]) - 3)
    a = basis * f * function.J(geom)
    assert a.coef is not None and a.shape == (len(basis),)
    b = f * (function.outer(basis) * function.J(geom))
    assert b.coef is not None and b.shape == (len(basis),) * 2
    c = (basis * x0) * x1                                    # coefficients multiply
    assert numpy.allclose(c.coef(x), x[..., 0] * x[..., 1])
    with pytest.raises(NotImplementedError):                 # different coefficient functions cannot be summed in one integral
        basis * x0 + basis * x1
    assert (basis * 2. + basis).coef is None


def test_boundary_tables():
    topo, geom, basis = _setup((4, 3, 2))
    assert topo.boundary['left'].faces == ((0, 0),) and topo.boundary['back,top'].faces == ((2, 1), (1, 1))
    with pytest.raises(KeyError):
        topo.boundary['inside']
    smp = topo.boundary['top'].sample('gauss', 4)
    elem_ids, xi, w = smp._face_tables(1, 1)
    assert elem_ids.tolist() == sorted(numpy.ravel_multi_index((i, 2, k), (4, 3, 2)) for i in range(4) for k in range(2))
    assert (xi[:, 1] == 1.).all() and len(w) == 9 and abs(w.sum() - 1.) < 1e-15
    x = smp._physical_points(geom.nodes, elem_ids, xi)
    assert x.shape == (8, 9, 3) and numpy.allclose(x[..., 1], 1.)
    # corner points of the face rule map into the element's corner cell
    assert (x[0, :, 0] >= 0).all() and (x[0, :, 0] <= .25).all()


def test_outside_the_closed_form():
    topo, geom, basis = _setup()
    with pytest.raises(NotImplementedError):
        function.outer(basis) * basis[:, None, None]       # trilinear in the basis: not a (bi)linear form
    with pytest.raises(NotImplementedError):
        topo.integral(basis, degree=4)                       # no jacobian: not an integral over the geometry
    with pytest.raises(NotImplementedError):
        topo.basis('discont', degree=1)


def test_fcm_generators():
    # own finite-cell table generators (octree volume quadrature, sphere surface quadrature, merged point sets)
    from nutils_b200 import fcm
    n, p = 8, 2
    elem_ids, qoff, qc, qw, ren, nbn = fcm.octree_ball(n, p, 2)
    assert (numpy.diff(elem_ids) > 0).all() and len(qoff) == len(elem_ids) + 1 and qoff[-1] == len(qw) == len(qc)
    assert (qc >= 0).all() and (qc <= 1).all() and (qw > 0).all()
    vol = qw.sum() * (2. / n) ** 3
    assert abs(vol - 4 / 3 * numpy.pi * .8 ** 3) < .03 * vol
    kept = ren[(ren >= 0) & (ren < nbn)]
    assert len(kept) == nbn and (numpy.diff(kept) > 0).all()              # monotone pruned numbering, like PrunedBasis
    se, sq, sx, sw, sn = fcm.sphere_surface(n)
    assert abs(sw.sum() - 4 * numpy.pi * .64) < 1e-12 and set(se) <= set(elem_ids)
    assert (sx >= 0).all() and (sx <= 1).all()
    vn = numpy.zeros((len(qw), 3))
    vn[:, 0] = 2. / n
    me, mq, mx, mw, mn, on_surface = fcm.merge_point_sets((elem_ids, qoff, qc, qw, vn), (se, sq, sx, sw, sn))
    assert numpy.array_equal(me, elem_ids) and mq[-1] == len(qw) + len(sw) and on_surface.sum() == len(sw)
    assert abs(mw[on_surface].sum() - sw.sum()) < 1e-12 and abs(mw[~on_surface].sum() - qw.sum()) < 1e-12
    # per element the volume points come first
    for k in numpy.searchsorted(elem_ids, se):
        seg = on_surface[mq[k]:mq[k + 1]]
        assert seg[-1] and (numpy.diff(seg.astype(int)) >= 0).all()
