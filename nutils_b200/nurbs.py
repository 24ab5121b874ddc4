'''NURBS tables of BASELINE.json configs[3] (host side, one-off): the plate with a circular hole of the reference's
``examples/platewithhole.py`` in NURBS mode (:51-86) at any refinement level and analysis degree.

The reference builds a 1 x 2 quadratic rational patch, refines the topology `nrefine` times and L2-projects the patch's
weight function onto the refined spline space to obtain the control weights of the analysis basis
``N_i c_i / W`` (platewithhole.py:79-83).  Here the same tables come from 1-D operations only (the patch is a tensor
product, so are the refined spaces):

* geometry: knot insertion of the homogeneous control points (w X, w) into the refined quadratic space -- exact, the
  spaces are nested -- via 1-D collocation matrices;  -> ``b2_geom_create_spline(ctrl, weights)``
* analysis weights: the L2 projection factorises into 1-D mass matrices, ``c = (M0^-1 G0) w (M1^-1 G1)^T``.
  -> ``b2_elemset_create(scale=c, rational=2)``

tests/test_tables.py pins the result to the arrays the unmodified reference produced for tests/golden/nurbs_plate_p*.npz.
'''

import numpy
from . import bspline, points


def eval_basis_1d(b, t):
    'dense matrix N[k, i] = N_i(t_k) of a bspline.Basis1D at parameters t in [0, nelems] (element e = [e, e+1])'
    t = numpy.asarray(t, dtype=float)
    e = numpy.clip(numpy.floor(t).astype(int), 0, b.nelems - 1)
    xi = t - e
    out = numpy.zeros((len(t), b.ndofs))
    for k in range(len(t)):
        c = b.coeffs[b.setidx[e[k]]]
        for a in range(b.degree + 1):
            out[k, b.start[e[k]] + a] = numpy.polyval(c[a], xi[k])
    return out


def greville(b):
    'one collocation parameter per function: the centroid of its support, nudged into the interior (Schoenberg-Whitney)'
    lo = numpy.full(b.ndofs, numpy.inf)
    hi = numpy.full(b.ndofs, -numpy.inf)
    for e in range(b.nelems):
        for a in range(b.degree + 1):
            i = b.start[e] + a
            lo[i] = min(lo[i], e)
            hi[i] = max(hi[i], e + 1)
    # weighted towards the function's maximum: first/last functions interpolate the end points
    idx = numpy.arange(b.ndofs)
    return lo + (hi - lo) * (idx / max(b.ndofs - 1, 1))


def refinement_matrix(coarse, fine, ratio):
    'R[i, j]: coefficients of the coarse function j in the fine space (fine element size = coarse / ratio; nested spaces)'
    t = greville(fine)
    A_f = eval_basis_1d(fine, t)
    A_c = eval_basis_1d(coarse, t / ratio)
    return numpy.linalg.solve(A_f, A_c)


def cross_mass(b_row, b_col, qdegree=12):
    'G[i, j] = int N^row_i N^col_j dt over the common element grid'
    assert b_row.nelems == b_col.nelems
    x, w = points.gauss1(qdegree)
    G = numpy.zeros((b_row.ndofs, b_col.ndofs))
    for e in range(b_row.nelems):
        cr = b_row.coeffs[b_row.setidx[e]]
        cc = b_col.coeffs[b_col.setidx[e]]
        vr = numpy.stack([numpy.polyval(cr[a], x) for a in range(b_row.degree + 1)])
        vc = numpy.stack([numpy.polyval(cc[a], x) for a in range(b_col.degree + 1)])
        G[b_row.start[e]:b_row.start[e] + b_row.degree + 1, b_col.start[e]:b_col.start[e] + b_col.degree + 1] += numpy.einsum('q,aq,bq->ab', w, vr, vc)
    return G


def plate_with_hole(nrefine, degree, radius=.5):
    '''Tables of the quarter plate with hole (platewithhole.py:66-86), refined `nrefine` times, analysis degree `degree`:
    dict(shape, bases, scale, gbases, gctrl, gweights) -- the arguments of engine.ElemSetPlan(..., scale=scale, rational=2,
    geom_spline=(gbases, gctrl, gweights)).'''
    cshape = (1, 2)
    cb = [bspline.spline_basis_1d(n, 2) for n in cshape]
    w = numpy.ones(12)
    w[1:3] = .5 + .25 * numpy.sqrt(2)
    A = 0, 0, 0
    B = (2**.5 - 1) * radius, .3 * (radius + 1) / 2, 1
    C = radius, (radius + 1) / 2, 1
    ctrl = numpy.array([[A, B, C, C], [C, C, B, A]]).T.reshape(-1, 2)
    ratio = 2 ** nrefine
    shape = tuple(n * ratio for n in cshape)
    gb = [bspline.spline_basis_1d(n, 2) for n in shape]
    R = [refinement_matrix(c, f, ratio) for c, f in zip(cb, gb)]
    W = w.reshape(3, 4)
    gweights = R[0] @ W @ R[1].T
    gctrl = numpy.stack([(R[0] @ (W * ctrl[:, i].reshape(3, 4)) @ R[1].T) / gweights for i in range(2)])
    ab = [bspline.spline_basis_1d(n, degree) for n in shape]
    P = []
    for a, g in zip(ab, gb):
        M = cross_mass(a, a)
        P.append(numpy.linalg.solve(M, cross_mass(a, g)))
    scale = P[0] @ gweights @ P[1].T
    return dict(shape=shape, bases=ab, scale=scale.ravel(), gbases=gb, gctrl=gctrl.reshape(2, -1), gweights=gweights.ravel())
