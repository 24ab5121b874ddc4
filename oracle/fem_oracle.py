'''CPU restatement (numpy) of the reference's element-integration hot path.

TEST INFRASTRUCTURE ONLY -- may be imported from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference
legs, as the checker or the reported CPU baseline; never from nutils_b200/.

Reference: evalf/nutils @ 37d1cc5 (version 10a8), /root/reference/src/nutils.
The reference generates a Python element loop (evaluable.compile,
evaluable.py:6532-6838; a listing is in SURVEY.md appendix A).  This module
restates that loop step by step for structured tensor-product spline spaces on
a multilinear nodal geometry, and the sparse post-processing exactly:

  element index -> (ix,iy,iz)            transformseq.py:563-579 (C-order ravel)
  coefficient rows per dimension         function.py:3080-3100 (StructuredBasis.f_dofs_coeffs)
  N, dN/dxi at the quadrature points     evaluable.py:4328-4374 (Polyval), :4584-4634 (PolyGrad)
                                         -- evaluated as products of the 1-D factors, which is the
                                         same polynomial the reference expands with nutils_poly.MulPlan
  dofs of the element                    function.py:3087-3093 (C-order ravel of per-dim ranges, modulo ndofs)
  geometry x = sum_v phi_v X_v, J        mesh.py:55-57 (degree-1 spline times nodal coordinates),
                                         function.py:1266-1295 (_Jacobian.lower)
  J^-1, det J                            evaluable.py:1403 (Inverse) -> numeric.py:221-242, evaluable.py:1463
  physical gradient dN J^-1              function.py:1207-1231 (_Gradient.lower)
  integrand contractions + weights       sample.py:951-956 (_Integral.lower), generated einsums
  COO append of block, rows, cols        evaluable.py:5383-5501 (LoopConcatenate), :5322-5343, :3474-3487
  vector dofs ibasis*ncomp+comp          function.py:2623-2626 (field)
  flat=row*ncols+col, stable argsort,
  unique, accumulate, rowptr             evaluable.py:588-616, :5646-5652, :3570-3577 -> numeric.py:434-460,
                                         evaluable.py:5655-5676 -> numeric.py:687-711
  RHS scatter-add                        evaluable.py:3582-3620 (numpy.add.at)

Element sets (trimmed / subset topologies, SubsetTopology topology.py:1598-1660): the loop runs over the kept
elements only, with per-element point sets (pointsseq.py:324-332), the pruned numbering of
function.PrunedBasis (function.py:3118-3133: dofs = renumber[parent dofs]), rational functions
c_i B_i / W (examples/platewithhole.py:66-83) and a (rational) spline geometry whose Jacobian is the
derivative of the spline map (function.py:1207-1231, 1266-1295).

Parity pin: tests/test_oracle_golden.py checks this module against the
golden vectors in tests/golden/*.npz, which were produced by the unmodified
reference itself (oracle/make_golden.py) -- CSR pattern bit-exact, values to
1e-13 relative.  The same tests pin the C restatement (oracle/fem_oracle.c).
'''

import itertools
import numpy


class Problem:
    '''Plain-array description of one structured assembly problem.

    ndims      : int
    nelems     : int[ndims]
    degree     : int[ndims]
    coeffs     : list of float64[nsets_d, p_d+1, p_d+1]   (highest power first, local xi in [0,1])
    setidx     : list of int[nelems_d]
    start      : list of int[nelems_d]
    ndofs_d    : int[ndims]
    ncomp      : int (vector-valued fields: dof = ibasis*ncomp + comp)
    qpts, qwts : list of float64[nq_d]  (tensor quadrature, C order)
    nodes      : float64[ndims, nelems_0+1, ...]  nodal coordinates of the multilinear geometry
    '''

    def __init__(self, nelems, degree, coeffs, setidx, start, ndofs_d, qpts, qwts, nodes, ncomp=1, elem_ids=None, qoff=None, qcoords=None,
                 qweights=None, renumber=None, nbasis_new=None, scale=None, rational=0, geom_spline=None, face_dim=None, normals=None):
        self.ndims = len(nelems)
        self.nelems = tuple(int(n) for n in nelems)
        self.degree = tuple(int(p) for p in degree)
        self.coeffs = [numpy.asarray(c, dtype=float) for c in coeffs]
        self.setidx = [numpy.asarray(s, dtype=numpy.int64) for s in setidx]
        self.start = [numpy.asarray(s, dtype=numpy.int64) for s in start]
        self.ndofs_d = tuple(int(n) for n in ndofs_d)
        self.qpts = [numpy.asarray(x, dtype=float) for x in qpts]
        self.qwts = [numpy.asarray(w, dtype=float) for w in qwts]
        self.nodes = numpy.asarray(nodes, dtype=float)
        self.ncomp = int(ncomp)
        # element set (all optional): kept elements, ragged points in element-local coordinates, pruned numbering,
        # numerator weights c_i and the denominator of rational functions (0 none, 1 own weight function, 2 the geometry's)
        self.elem_ids = None if elem_ids is None else numpy.asarray(elem_ids, dtype=numpy.int64)
        self.qoff = None if qoff is None else numpy.asarray(qoff, dtype=numpy.int64)
        self.qcoords = None if qcoords is None else numpy.asarray(qcoords, dtype=float)
        self.qweights = None if qweights is None else numpy.asarray(qweights, dtype=float)
        self.renumber = None if renumber is None else numpy.asarray(renumber, dtype=numpy.int64)
        self.nbasis_new = None if renumber is None else int(nbasis_new)
        self.scale = None if scale is None else numpy.asarray(scale, dtype=float)
        self.rational = int(rational)
        # spline geometry: dict(degree, coeffs, setidx, start, ndofs_d, ctrl[ndims, nbasis_g], weights[nbasis_g] or None)
        self.geom_spline = geom_spline
        # boundary integrals: face_dim[isel] = -1 (volume) or the reference direction normal to the face the element's points lie
        # on; the measure is then the surface measure |det J| |J^-T e_dim| (function.J on a boundary sample, function.py:2291-2316)
        self.face_dim = None if face_dim is None else numpy.asarray(face_dim, dtype=numpy.int64)
        # immersed boundaries (topo.boundary['trimmed']): per point the reference normal of its facet scaled by the facet's reference
        # measure; the surface measure is |det J| |J^-T n| (Nanson's formula applied to the facet transform, transform.py)
        self.normals = None if normals is None else numpy.asarray(normals, dtype=float)
        assert geom_spline is not None or self.nodes.shape == (self.ndims,) + tuple(n + 1 for n in self.nelems)

    @property
    def nbasis(self):
        return int(numpy.prod(self.ndofs_d)) if self.renumber is None else self.nbasis_new

    @property
    def nsel(self):
        return self.ntotal if self.elem_ids is None else len(self.elem_ids)

    @property
    def ndofs(self):
        return self.nbasis * self.ncomp

    @property
    def ntotal(self):
        return int(numpy.prod(self.nelems))


def _polyval_rows(c, x):
    'rows of c are polynomials (highest power first); returns [nrows, len(x)] values and derivatives'
    p = c.shape[1] - 1
    val = numpy.stack([numpy.polyval(row, x) for row in c])
    if p:
        der = numpy.stack([numpy.polyval(row[:-1] * numpy.arange(p, 0, -1.), x) for row in c])
    else:
        der = numpy.zeros_like(val)
    return val, der


def _tensor(factors):
    'C-order tensor product of per-dim [n_d, q_d] tables -> [prod q, prod n]'
    out = numpy.ones((1, 1))
    for f in factors:
        out = (out[:, None, :, None] * f.T[None, :, None, :]).reshape(out.shape[0] * f.shape[1], out.shape[1] * f.shape[0])
    return out


def _tensor_points(factors, npts):
    'product of per-dim [n_d, npts] tables evaluated at the SAME list of points -> [npts, prod n]'
    out = numpy.ones((npts, 1))
    for f in factors:
        out = (out[:, :, None] * f.T[:, None, :]).reshape(npts, out.shape[1] * f.shape[0])
    return out


def _basis_at(coeffs, setidx, start, degree, ndofs_d, idx, xi):
    'N[nq, n_e], dN/dxi[nq, n_e, nd] and parent dofs of a tensor-product space on element idx at local points xi[nq, nd]'
    nd = len(idx)
    vals, ders, ranges = [], [], []
    for d in range(nd):
        v, g = _polyval_rows(coeffs[d][setidx[d][idx[d]]], xi[:, d])
        vals.append(v)
        ders.append(g)
        ranges.append((start[d][idx[d]] + numpy.arange(degree[d] + 1)) % ndofs_d[d])
    N = _tensor_points(vals, len(xi))
    dN = numpy.stack([_tensor_points([ders[k] if k == d else vals[k] for k in range(nd)], len(xi)) for d in range(nd)], axis=-1)
    dofs = ranges[0]
    for d in range(1, nd):
        dofs = (dofs[:, None] * ndofs_d[d] + ranges[d][None, :]).ravel()
    return N, dN, dofs


def element_data(prob, isel, return_x=False):
    '''Everything the generated loop computes for one (selected) element before the integrand:
    dofs[n_e] (-1: dropped by the pruned numbering), N[nq,n_e], grad[nq,n_e,ndims] (physical), wdet[nq]
    (and the physical coordinates x[nq,ndims] with return_x).'''
    nd = prob.ndims
    ielem = isel if prob.elem_ids is None else int(prob.elem_ids[isel])
    idx = numpy.unravel_index(ielem, prob.nelems)
    if prob.qoff is None:
        # uniform tensor rule, C-order outer product (pointsseq.py:420-428)
        grids = numpy.meshgrid(*prob.qpts, indexing='ij')
        xi = numpy.stack([g.ravel() for g in grids], axis=-1)
        w = _tensor([wq[None, :] for wq in prob.qwts])[:, 0]
    else:
        xi = prob.qcoords[prob.qoff[isel]:prob.qoff[isel + 1]]
        w = prob.qweights[prob.qoff[isel]:prob.qoff[isel + 1]]
    N, dN, dofs = _basis_at(prob.coeffs, prob.setidx, prob.start, prob.degree, prob.ndofs_d, idx, xi)
    Wg = dWg = None
    if prob.geom_spline is None:
        # multilinear geometry: shape functions of the degree-1 spline on this element are (1-xi, xi) per dim
        lin_v = [numpy.stack([1 - xi[:, d], xi[:, d]]) for d in range(nd)]
        lin_d = [numpy.stack([-numpy.ones(len(xi)), numpy.ones(len(xi))]) for d in range(nd)]
        X = prob.nodes[(slice(None),) + tuple(slice(i, i + 2) for i in idx)].reshape(nd, -1)  # [ndims, 2^nd] C-order vertices
        J = numpy.stack([_tensor_points([lin_d[k] if k == d else lin_v[k] for k in range(nd)], len(xi)) @ X.T for d in range(nd)], axis=-1)  # [nq, i, k] = dx_i/dxi_k
        xphys = _tensor_points(lin_v, len(xi)) @ X.T
    else:
        g = prob.geom_spline
        Ng, dNg, gdofs = _basis_at(g['coeffs'], g['setidx'], g['start'], g['degree'], g['ndofs_d'], idx, xi)
        ctrl = numpy.asarray(g['ctrl'], dtype=float)[:, gdofs]  # [ndims, n_g]
        if g.get('weights') is None:
            J = numpy.einsum('qak,ia->qik', dNg, ctrl)
            xphys = Ng @ ctrl.T
        else:
            wts = numpy.asarray(g['weights'], dtype=float)[gdofs]
            Wg = Ng @ wts
            dWg = numpy.einsum('qak,a->qk', dNg, wts)
            Xh = Ng @ (ctrl * wts).T                            # [nq, ndims]
            dXh = numpy.einsum('qak,ia->qik', dNg, ctrl * wts)
            x = Xh / Wg[:, None]
            J = (dXh - x[:, :, None] * dWg[:, None, :]) / Wg[:, None, None]
            xphys = x
    if prob.scale is not None:
        c = prob.scale[dofs]
        N = N * c
        dN = dN * c[None, :, None]
    if prob.rational:
        if prob.rational == 1:
            W, dW = N.sum(1), dN.sum(1)                         # sum_j c_j B_j and its xi-gradient
        else:
            W, dW = Wg, dWg
        N = N / W[:, None]
        dN = (dN - N[:, :, None] * dW[:, None, :]) / W[:, None, None]
    Jinv = numpy.linalg.inv(J)
    det = numpy.linalg.det(J)
    grad = numpy.einsum('qak,qki->qai', dN, Jinv)
    meas = abs(det)
    if prob.normals is not None:
        nr = prob.normals[prob.qoff[isel]:prob.qoff[isel + 1]]
        meas = meas * numpy.linalg.norm(numpy.einsum('qki,qk->qi', Jinv, nr), axis=-1)
    elif prob.face_dim is not None and prob.face_dim[isel] >= 0:
        meas = meas * numpy.linalg.norm(Jinv[:, prob.face_dim[isel], :], axis=-1)
    if prob.renumber is not None:
        dofs = prob.renumber[dofs]
        dofs = numpy.where((dofs >= 0) & (dofs < prob.nbasis_new), dofs, -1)
    if return_x:
        return dofs, N, grad, w * meas, xphys
    return dofs, N, grad, w * meas


def _vector_dofs(dofs, ncomp):
    return (dofs[:, None] * ncomp + numpy.arange(ncomp)[None, :]).ravel()


def element_matrix(form, N, grad, wdet, ncomp=1):
    '''Dense element block for a bilinear form.

    form = ('mass',) | ('stiffness',) | ('elasticity', lmbda, mu) | ('generic', D)
    with D[ncomp, ndims+1, ncomp, ndims+1] acting on (value, gradient) of test and trial.'''
    kind = form[0]
    nd = grad.shape[-1]
    if kind == 'mass':
        assert ncomp == 1
        return numpy.einsum('qa,qb,q->ab', N, N, wdet)
    if kind == 'stiffness':
        assert ncomp == 1
        return numpy.einsum('qai,qbi,q->ab', grad, grad, wdet)
    if kind == 'elasticity':
        # second derivative of the energy  eps_ij sigma_ij,  sigma = lmbda tr(eps) I + 2 mu eps  (examples/elasticity.py:50-58)
        assert ncomp == nd
        lmbda, mu = form[1], form[2]
        n_e = N.shape[1]
        eye = numpy.eye(nd)
        # du_i/dx_j for basis (a, c): grad[q,a,j] delta_ic
        G = numpy.einsum('qaj,ic->qacij', grad, eye)
        eps = .5 * (G + G.swapaxes(-1, -2))
        sig = lmbda * numpy.einsum('qackk,ij->qacij', eps, eye) + 2 * mu * eps
        blk = 2 * numpy.einsum('qacij,qbdij,q->acbd', eps, sig, wdet)
        return blk.reshape(n_e * ncomp, n_e * ncomp)
    if kind == 'generic':
        D = numpy.asarray(form[1], dtype=float)
        B = numpy.concatenate([N[:, :, None], grad], axis=-1)  # [q, a, alpha]
        n_e = N.shape[1]
        blk = numpy.einsum('qax,cxdy,qby,q->acbd', B, D, B, wdet)
        return blk.reshape(n_e * ncomp, n_e * ncomp)
    raise ValueError(kind)


def element_vector(form, N, grad, wdet, ncomp=1):
    '''Element load vector.  form = ('load',) for int N_a, or ('generic', c) with c[ncomp, ndims+1].'''
    kind = form[0]
    if kind == 'load':
        assert ncomp == 1
        return numpy.einsum('qa,q->a', N, wdet)
    if kind == 'generic':
        c = numpy.asarray(form[1], dtype=float)
        B = numpy.concatenate([N[:, :, None], grad], axis=-1)
        return numpy.einsum('qax,cx,q->ac', B, c, wdet).ravel()
    raise ValueError(kind)


def evaluate(prob, fields=()):
    '''Sample.eval (sample.py:192-215, 959-975) in point order: x[npoints, ndims], weights w |det J| [npoints], and for every
    coefficient vector in `fields` the discrete field values [npoints, nfields, ncomp] and physical gradients
    [npoints, nfields, ncomp, ndims].'''
    nc = prob.ncomp
    xs, ws, vals, grads = [], [], [], []
    for isel in range(prob.nsel):
        dofs, N, grad, wdet, x = element_data(prob, isel, return_x=True)
        xs.append(x)
        ws.append(wdet)
        vdofs = (dofs[:, None] * nc + numpy.arange(nc)[None, :])            # [n_e, ncomp]
        C = numpy.stack([numpy.asarray(f, dtype=float)[vdofs] for f in fields]) if len(fields) else numpy.zeros((0,) + vdofs.shape)
        vals.append(numpy.einsum('qa,fac->qfc', N, C))
        grads.append(numpy.einsum('qak,fac->qfck', grad, C))
    return numpy.concatenate(xs), numpy.concatenate(ws), numpy.concatenate(vals), numpy.concatenate(grads)


def coo_to_csr(values, rows, cols, nrows, ncols):
    '''evaluable.py:588-616 + :5646-5682: sorted unique CSR with structural zeros kept.'''
    flat = rows.astype(numpy.int64) * ncols + cols.astype(numpy.int64)
    order = numpy.argsort(flat, kind='stable')                      # ArgSort, evaluable.py:5577-5581
    sflat = flat[order]
    mask = numpy.ones(len(sflat), dtype=bool)                       # UniqueMask, :5604-5612
    mask[1:] = sflat[1:] != sflat[:-1]
    inverse = numpy.empty(len(flat), dtype=numpy.int64)             # UniqueInverse, :5632-5640
    inverse[order] = numpy.cumsum(mask) - 1
    nnz = int(mask.sum())
    data = numpy.bincount(inverse, weights=values, minlength=nnz)   # numeric.accumulate, numeric.py:448-455
    ukeys = sflat[mask]
    urows, ucols = numpy.divmod(ukeys, ncols)
    rowptr = numpy.searchsorted(urows, numpy.arange(nrows + 1)).astype(numpy.int64)  # numeric.compress_indices, numeric.py:687-711
    return data, rowptr, ucols.astype(numpy.int64)


def assemble(prob, matrix_forms=(), vector_forms=(), matrix_coefs=None, vector_coefs=None):
    '''Run the element loop and the sparse post-processing.

    matrix_coefs / vector_coefs: per form None or one scalar per quadrature point (in point order) multiplying the integrand.
    Returns ([(values, rowptr, colidx), ...], [rhs, ...]).'''
    nc = prob.ncomp
    n_e = int(numpy.prod([p + 1 for p in prob.degree])) * nc
    nel = prob.nsel
    vals = [numpy.empty((nel, n_e, n_e)) for _ in matrix_forms]
    rows = numpy.empty((nel, n_e, n_e), dtype=numpy.int64)
    cols = numpy.empty((nel, n_e, n_e), dtype=numpy.int64)
    rhs = [numpy.zeros(prob.ndofs) for _ in vector_forms]
    qpos = 0
    for ielem in range(nel):
        dofs, N, grad, wdet = element_data(prob, ielem)
        pts = slice(qpos, qpos + len(wdet))
        qpos += len(wdet)
        vdofs = _vector_dofs(dofs, nc)
        rows[ielem] = vdofs[:, None]
        cols[ielem] = vdofs[None, :]
        for k, form in enumerate(matrix_forms):
            c = 1. if matrix_coefs is None or matrix_coefs[k] is None else numpy.asarray(matrix_coefs[k])[pts]
            vals[k][ielem] = element_matrix(form, N, grad, wdet * c, nc)
        for k, form in enumerate(vector_forms):
            c = 1. if vector_coefs is None or vector_coefs[k] is None else numpy.asarray(vector_coefs[k])[pts]
            numpy.add.at(rhs[k], vdofs, element_vector(form, N, grad, wdet * c, nc))
    mats = [coo_to_csr(v.ravel(), rows.ravel(), cols.ravel(), prob.ndofs, prob.ndofs) for v in vals]
    return mats, rhs


def assemble_general(ndims, types, etype, dofs, verts, nbasis, matrix_forms=(), vector_forms=(), ncomp=1):
    '''General dof maps (simplex / mixed meshes; TEST INFRASTRUCTURE): the reference's element loop for bases given per element
    (function.PlainBasis, function.py:2881-2913: Inflate(Polyval(coeffs, points), dofs)), a (multi)linear geometry
    (x = sum_v X_v phi_v, J = sum_v X_v (x) dphi_v; function.py:1207-1295), einsum + weighted sum (sample.py:951-956), and the
    sparse post-processing of coo_to_csr (evaluable.py:588-616, 5646-5682).  Forms: ('generic', D[nc, na, nc, na]) /
    ('generic', C[nc, na]).  Returns ([(values, rowptr, colidx)...], [rhs...]).'''
    na = ndims + 1
    rows, cols, vals = [], [], [[] for _ in matrix_forms]
    rhs = [numpy.zeros(nbasis * ncomp) for _ in vector_forms]
    for e in range(len(etype)):
        t = types[etype[e]]
        X = numpy.asarray(verts[e], dtype=float)                                   # [nvert, nd]
        J = numpy.einsum('vi,qvk->qik', X, t['gdphi'])                               # dx_i / dxi_k
        Jinv = numpy.linalg.inv(J)
        wdet = t['weights'] * abs(numpy.linalg.det(J))
        g = numpy.concatenate([t['phi'][:, :, None], numpy.einsum('qak,qkj->qaj', t['dphi'], Jinv)], axis=2)   # [nq, nfun, na]
        d = numpy.asarray(dofs[e], dtype=numpy.int64)
        full = (d[:, None] * ncomp + numpy.arange(ncomp)[None, :]).ravel()
        rows.append(numpy.repeat(full, len(full)))
        cols.append(numpy.tile(full, len(full)))
        for m, (kind, D) in enumerate(matrix_forms):
            blk = numpy.einsum('q,qax,cxey,qby->acbe', wdet, g, numpy.asarray(D, dtype=float).reshape(ncomp, na, ncomp, na), g)
            vals[m].append(blk.reshape(len(full), len(full)).ravel())
        for v, (kind, C) in enumerate(vector_forms):
            numpy.add.at(rhs[v], full, numpy.einsum('q,qax,cx->ac', wdet, g, numpy.asarray(C, dtype=float).reshape(ncomp, na)).ravel())
    rows = numpy.concatenate(rows) if rows else numpy.zeros(0, dtype=numpy.int64)
    cols = numpy.concatenate(cols) if cols else numpy.zeros(0, dtype=numpy.int64)
    n = nbasis * ncomp
    mats = [coo_to_csr(numpy.concatenate(v) if v else numpy.zeros(0), rows, cols, n, n) for v in vals]
    if not matrix_forms:
        mats = []
    return mats, rhs
