'''Octree quadrature for level-set domains on a structured grid: a SIMPLE stand-in for the cut-cell machinery of the reference.

The reference trims a topology with ``Topology.trim`` (src/nutils/topology.py:1598-1660) and builds boundary-fitted mosaics of
simplices in the cut cells (element.py) -- host Python that stays the reference's; its tables enter this repository through
``nutils_b200.adapter``.  For workloads that must be generated WITHOUT the reference (the GPU box has no copy: timing scripts,
``examples/finitecell.py``) this module provides the classical finite-cell alternative: cut cells are subdivided `depth` times,
sub-cells whose centre lies inside the domain carry a scaled Gauss rule (Parvizian, Duester, Rank 2007).  It produces the SAME
kind of tables (kept elements, ragged points in element-local coordinates, pruned numbering) -- not the same numbers as the
reference's mosaic; parity of the kernels on the reference's own cut-cell points is pinned by the goldens in tests/golden.
'''

import numpy
from . import points


def octree_ball(n, degree, L, radius=.8):
    'elem_ids, qoff, qcoords, qweights, renumber, nbasis_new for the ball |x| < radius in [-1,1]^3'
    v = numpy.linspace(-1, 1, n + 1)
    gx, gw = points.gauss1(2 * degree)
    g3 = numpy.stack(numpy.meshgrid(gx, gx, gx, indexing='ij'), -1).reshape(-1, 3)
    w3 = numpy.einsum('i,j,k->ijk', gw, gw, gw).ravel()
    X = numpy.stack(numpy.meshgrid(v, v, v, indexing='ij'), -1)
    inside = numpy.linalg.norm(X, axis=-1) < radius
    c = [inside[i:n + i, j:n + j, k:n + k] for i in (0, 1) for j in (0, 1) for k in (0, 1)]
    nin = numpy.sum(c, axis=0)
    full = nin == 8
    cut = (nin > 0) & ~full
    # a cell without inside vertices may still touch the ball: ignored (thin slivers), as a midpoint rule would
    m = 2 ** L
    sub = (numpy.stack(numpy.meshgrid(*[numpy.arange(m)] * 3, indexing='ij'), -1).reshape(-1, 1, 3) + g3[None]) / m   # [m^3, nq, 3]
    subc = (numpy.stack(numpy.meshgrid(*[numpy.arange(m)] * 3, indexing='ij'), -1).reshape(-1, 3) + .5) / m          # sub-cell centres
    elem_ids, coords, weights, qoff = [], [], [], [0]
    h = 2. / n
    for e in numpy.flatnonzero((full | cut).ravel()):
        i, j, k = numpy.unravel_index(e, (n, n, n))
        if full[i, j, k]:
            coords.append(g3)
            weights.append(w3)
        else:
            org = numpy.array([v[i], v[j], v[k]])
            keep = numpy.linalg.norm(org + subc * h, axis=-1) < radius
            coords.append(sub[keep].reshape(-1, 3))
            weights.append(numpy.tile(w3 / m ** 3, keep.sum()))
        elem_ids.append(e)
        qoff.append(qoff[-1] + len(weights[-1]))
    elem_ids = numpy.array(elem_ids, dtype=numpy.int64)
    nd = n + degree
    used = numpy.zeros((nd, nd, nd), dtype=bool)
    ii, jj, kk = numpy.unravel_index(elem_ids, (n, n, n))
    for a in range(degree + 1):
        for b in range(degree + 1):
            for c_ in range(degree + 1):
                used[ii + a, jj + b, kk + c_] = True
    dofs = numpy.flatnonzero(used.ravel())
    renumber = numpy.full(nd ** 3, len(dofs), dtype=numpy.int64)
    renumber[dofs] = numpy.arange(len(dofs))
    return elem_ids, numpy.array(qoff, dtype=numpy.int64), numpy.concatenate(coords), numpy.concatenate(weights), renumber, len(dofs)


def sphere_surface(n, radius=.8, ntheta=None):
    '''Quadrature of the sphere |x| = radius immersed in the n^3 grid on [-1, 1]^3, as tables for an immersed-boundary element set:

        elem_ids, qoff, qcoords, qweights, normals

    Product rule on the sphere (Gauss in cos(theta), trapezoid in phi: exact for polynomials of the matching degree), every point
    assigned to the grid cell that contains it.  The weights are PHYSICAL surface weights; the scaled reference normal that makes
    |det J| |J^-T n| = 1 on the uniform grid (J = h I with h the cell size) is n = e_x / h^2, so b2_elemset_set_normals reproduces them.
    The union of the cells is a subset of the cut cells of octree_ball with the same radius.'''
    ntheta = ntheta or 4 * n
    nphi = 2 * ntheta
    ct, wt = numpy.polynomial.legendre.leggauss(ntheta)
    phi = (numpy.arange(nphi) + .5) * (2 * numpy.pi / nphi)
    st = numpy.sqrt(1 - ct ** 2)
    x = radius * numpy.stack([numpy.outer(st, numpy.cos(phi)), numpy.outer(st, numpy.sin(phi)), numpy.outer(ct, numpy.ones(nphi))], -1).reshape(-1, 3)
    w = radius ** 2 * numpy.outer(wt, numpy.full(nphi, 2 * numpy.pi / nphi)).ravel()
    h = 2. / n
    cell = numpy.clip(numpy.floor((x + 1) / h).astype(int), 0, n - 1)
    xi = (x + 1) / h - cell
    ids = numpy.ravel_multi_index(cell.T, (n, n, n))
    order = numpy.argsort(ids, kind='stable')
    ids, xi, w = ids[order], xi[order], w[order]
    elem_ids, counts = numpy.unique(ids, return_counts=True)
    qoff = numpy.concatenate([[0], numpy.cumsum(counts)]).astype(numpy.int64)
    normals = numpy.zeros_like(xi)
    normals[:, 0] = 1. / h ** 2
    return elem_ids.astype(numpy.int64), qoff, xi, w, normals


def merge_point_sets(a, b):
    '''Union of two ragged point sets on the same grid, each given as (elem_ids, qoff, qcoords, qweights, normals): per element the
    points of `a` followed by the points of `b`.  Also returns the boolean array "point comes from b", so that per-point
    coefficients can switch forms on and off -- e.g. volume points (a) carry the stiffness and the load, surface points (b) a
    penalty term, in ONE assembly into ONE pattern.'''
    ea, qa, xa, wa, na = a
    eb, qb, xb, wb, nb = b
    elem_ids = numpy.union1d(ea, eb)
    ia = {int(e): k for k, e in enumerate(ea)}
    ib = {int(e): k for k, e in enumerate(eb)}
    xs, ws, ns, fb, qoff = [], [], [], [], [0]
    for e in elem_ids:
        cnt = 0
        for src, idx, (q, x, w, n) in ((0, ia, (qa, xa, wa, na)), (1, ib, (qb, xb, wb, nb))):
            k = idx.get(int(e))
            if k is None:
                continue
            sl = slice(q[k], q[k + 1])
            xs.append(x[sl]); ws.append(w[sl]); ns.append(n[sl])
            fb.append(numpy.full(q[k + 1] - q[k], bool(src)))
            cnt += q[k + 1] - q[k]
        qoff.append(qoff[-1] + cnt)
    return elem_ids, numpy.array(qoff, dtype=numpy.int64), numpy.concatenate(xs), numpy.concatenate(ws), numpy.concatenate(ns), numpy.concatenate(fb)
