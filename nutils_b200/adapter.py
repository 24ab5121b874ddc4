'''Table extraction from evalf/nutils objects: the reference-side half of the drop-in.

Everything the CUDA path consumes is a plain array that the reference already holds inside its own
objects; this module reads those arrays out (duck-typed: the nutils modules are reached through the objects passed in) so that a
topology built, refined or TRIMMED by the reference's host code (topology.trim and the cut-cell
mosaics of element.py stay where they are: host Python, outside the measured path) can be integrated
through the C ABI:

    StructuredBasis._coeffs / _start_dofs / _dofs_shape   function.py:3051-3058   -> bspline.Basis1D per dimension
    PrunedBasis._transmap / _renumber                     function.py:3118-3123   -> elem_ids, renumber, nbasis_new
    Sample.points.get(ielem).coords / .weights            pointsseq.py:324-332    -> qoff, qcoords, qweights
    vertex arrays of mesh.rectilinear                     mesh.py:34-57           -> nodes

`plan_from_reference` turns them into an :class:`engine.ElemSetPlan` (trimmed / pruned) or an
:class:`engine.Plan` (plain structured).  INTEGRATION.md shows where the reference would call it.
'''

import numpy
from . import bspline, engine


def bases1d_from_structured_basis(basis):
    '''per-dimension bspline.Basis1D tables of a nutils ``function.StructuredBasis`` (function.py:3040-3100).

    ``_coeffs[d][e]`` is the (p+1) x (p+1) coefficient array of element e along d, highest power first
    (topology.py:2327-2361); equal arrays are shared into coefficient sets.'''
    out = []
    for coeffs_d, start_d, stop_d, ndofs_d in zip(basis._coeffs, basis._start_dofs, basis._stop_dofs, basis._dofs_shape):
        sets, setidx, index = [], [], {}
        for c in coeffs_d:
            a = numpy.asarray(c, dtype=float)
            key = a.shape, a.tobytes()
            k = index.get(key)
            if k is None:
                k = index[key] = len(sets)
                sets.append(a)
            setidx.append(k)
        start = numpy.asarray(start_d, dtype=numpy.int64)
        stop = numpy.asarray(stop_d, dtype=numpy.int64)
        p = sets[0].shape[0] - 1
        if any(s.shape != (p + 1, p + 1) for s in sets) or (stop - start != p + 1).any():
            raise NotImplementedError('elements with a varying number of functions along one dimension')
        if (start + p >= ndofs_d).any():
            raise NotImplementedError('periodic bases are outside the accelerated path')
        out.append(bspline.Basis1D(p, len(setidx), numpy.stack(sets), numpy.array(setidx, dtype=numpy.int32), start, int(ndofs_d)))
    return out


def ragged_points(sample, nelems=None):
    '(qoff int64[n+1], qcoords float64[npoints, ndims], qweights float64[npoints]) of a nutils Sample, element-local coordinates'
    n = sample.nelems if nelems is None else nelems
    coords, weights, qoff = [], [], [0]
    for i in range(n):
        p = sample.points.get(i)
        coords.append(numpy.asarray(p.coords, dtype=float))
        weights.append(numpy.asarray(p.weights, dtype=float))
        qoff.append(qoff[-1] + len(weights[-1]))
    ndims = coords[0].shape[1] if coords else 0
    return (numpy.array(qoff, dtype=numpy.int64), numpy.concatenate(coords) if coords else numpy.zeros((0, ndims)),
            numpy.concatenate(weights) if weights else numpy.zeros(0))


def tables_from_reference(topo, basis, degree, vertices=None, nodes=None, ischeme='gauss'):
    '''Arrays describing ``topo.integral(<form in basis> * J(geom), degree=degree)`` for a structured topology or a
    trimmed / subset topology of one, with a StructuredBasis or a PrunedBasis of one.

    vertices : per-dimension vertex arrays (what was passed to mesh.rectilinear), or
    nodes    : explicit nodal coordinates float64[ndims, n0+1, ...] of a multilinear geometry.'''
    parent = getattr(basis, '_parent', basis)
    if not hasattr(parent, '_start_dofs'):
        raise NotImplementedError('only (pruned) structured bases are on the accelerated path')
    bases = bases1d_from_structured_basis(parent)
    shape = tuple(parent._transforms_shape)
    if nodes is None:
        if vertices is None:
            raise ValueError('either vertices or nodes are needed')
        nodes = numpy.stack(numpy.meshgrid(*[numpy.asarray(v, dtype=float) for v in vertices], indexing='ij'))
    t = dict(bases=bases, nelems=shape, nodes=numpy.asarray(nodes, dtype=float))
    if parent is not basis:
        t['elem_ids'] = numpy.asarray(basis._transmap, dtype=numpy.int64)
        t['renumber'] = numpy.asarray(basis._renumber, dtype=numpy.int64)
        t['nbasis_new'] = len(basis)
        if len(topo) != len(t['elem_ids']):
            raise ValueError('basis and topology have different elements')
        t['qoff'], t['qcoords'], t['qweights'] = ragged_points(topo.sample(ischeme, degree), len(topo))
    return t


def immersed_boundary_tables_from_reference(btopo, basetopo, degree, ischeme='gauss'):
    '''Points of a boundary topology whose facets are NOT element faces -- the trimmed boundary ``topo.boundary['name']`` of a
    finite-cell topology -- grouped per volume element of the structured base topology:

        elem_ids, qoff, qcoords, qweights, normals

    Every boundary element's transform chain is split into the base element and the tail that maps facet coordinates into the
    element (``transforms.index_with_tail``, transformseq.py); the points are pushed through the tail (``transform.apply``) and
    the scaled reference normal is the generalised cross product of the columns of the tail's linear map, so that
    |det J| |J^-T n| is the surface measure the reference obtains from function.J on the boundary sample (function.py:2291-2316).'''
    import importlib
    transform = importlib.import_module(type(btopo).__module__.partition('.')[0] + '.transform')
    sample = btopo.sample(ischeme, degree)
    nd = basetopo.ndims
    per_elem = {}
    for i in range(len(btopo)):
        ielem, tail = basetopo.transforms.index_with_tail(btopo.transforms[i])
        pts = sample.points.get(i)
        xi = numpy.asarray(transform.apply(tail, numpy.asarray(pts.coords)), dtype=float)
        L = numpy.eye(nd)
        for item in tail:
            L = L @ numpy.asarray(item.linear, dtype=float)
        if L.shape != (nd, nd - 1):
            raise NotImplementedError('boundary elements must be facets (codimension one)')
        if nd == 3:
            n = numpy.cross(L[:, 0], L[:, 1])
        elif nd == 2:
            n = numpy.array([L[1, 0], -L[0, 0]])
        else:
            n = numpy.ones(1)
        entry = per_elem.setdefault(int(ielem), ([], [], []))
        entry[0].append(xi)
        entry[1].append(numpy.asarray(pts.weights, dtype=float))
        entry[2].append(numpy.tile(n, (len(xi), 1)))
    elem_ids = numpy.array(sorted(per_elem), dtype=numpy.int64)
    qoff = numpy.concatenate([[0], numpy.cumsum([sum(len(w) for w in per_elem[e][1]) for e in elem_ids])]).astype(numpy.int64)
    cat = lambda k: numpy.concatenate([a for e in elem_ids for a in per_elem[e][k]]) if len(elem_ids) else numpy.zeros((0, nd) if k != 1 else 0)
    return elem_ids, qoff, cat(0), cat(1), cat(2)


def nurbs_tables_from_reference(topo, weightfunc, geom, degree, geometry_degree=2):
    '''Tables of a NURBS discretisation built the way examples/platewithhole.py:66-86 builds it: a coarse rational patch
    (`weightfunc`, `geom` = nurbsbasis @ controlpoints) on a topology that has since been refined.

    Returns (bases, scale, gbases, gctrl, gweights): the analysis space is ``topo.basis('spline', degree)`` scaled by the
    projected control weights `scale` and divided by the patch's weight function (``rational=2``); the geometry is the patch
    itself, represented exactly in the B-splines of degree `geometry_degree` of the refined topology (nested spaces) by its
    weights `gweights` and control points `gctrl` -- the arguments of b2_geom_create_spline.  The projections are the
    reference's own (``System(...).solve()`` of a least-squares functional, platewithhole.py:79-82); the nutils modules are
    taken from the objects passed in.'''
    import importlib
    nutils = importlib.import_module(type(topo).__module__.partition('.')[0])
    function = importlib.import_module(nutils.__name__ + '.function')
    System = importlib.import_module(nutils.__name__ + '.solver').System

    def project(target, pbasis):
        sqr = topo.integral((function.field('w', pbasis) - target)**2, degree=9)
        return numpy.asarray(System(sqr, trial='w').solve()['w'])
    gbasis = topo.basis('spline', degree=geometry_degree)
    gweights = project(weightfunc, gbasis)
    gctrl = numpy.stack([project(weightfunc * geom[i], gbasis) for i in range(topo.ndims)]) / gweights
    abasis = topo.basis('spline', degree=degree)
    scale = project(weightfunc, abasis)
    return bases1d_from_structured_basis(abasis), scale, bases1d_from_structured_basis(gbasis), gctrl, gweights


def plan_from_reference(ctx, topo, basis, degree, vertices=None, nodes=None, ncomp=1):
    'engine.Plan (structured) or engine.ElemSetPlan (trimmed / pruned) for reference objects; see tables_from_reference'
    from . import points
    t = tables_from_reference(topo, basis, degree, vertices=vertices, nodes=nodes)
    if 'elem_ids' in t:
        return engine.ElemSetPlan(ctx, t['bases'], nodes=t['nodes'], ncomp=ncomp, elem_ids=t['elem_ids'], qoff=t['qoff'], qcoords=t['qcoords'],
                                  qweights=t['qweights'], renumber=t['renumber'], nbasis_new=t['nbasis_new'])
    return engine.Plan(ctx, t['bases'], points.tensor_gauss(len(t['bases']), degree), t['nodes'], ncomp=ncomp)


# ---- general dof maps: simplex and mixed meshes, bases given per element ---------------------------------------

def _poly_powers(nvars, degree):
    '''exponents of the coefficients of a polynomial in `nvars` variables of total degree `degree` in the reference's order:
    descending in (k_{n-1}, ..., k_0) -- last variable most significant, highest power first (evaluable.py:4328-4350)'''
    import itertools
    pw = [k for k in itertools.product(range(degree + 1), repeat=nvars) if sum(k) <= degree]
    pw.sort(key=lambda k: k[::-1], reverse=True)
    return numpy.array(pw, dtype=int).reshape(len(pw), nvars)


def poly_tabulate(coeffs, points):
    '(values[nq, nfun], gradients[nq, nfun, nvars]) of polynomials given as coefficient rows coeffs[nfun, ncoeffs] (reference order)'
    import math
    coeffs = numpy.asarray(coeffs, dtype=float)
    points = numpy.asarray(points, dtype=float)
    nvars = points.shape[1]
    degree = 0
    while math.comb(nvars + degree, nvars) < coeffs.shape[1]:
        degree += 1
    if math.comb(nvars + degree, nvars) != coeffs.shape[1]:
        raise ValueError('invalid number of polynomial coefficients')
    pw = _poly_powers(nvars, degree)
    mono = numpy.prod(points[:, None, :] ** pw[None], axis=2)                      # [nq, ncoeffs]
    vals = mono @ coeffs.T
    grads = numpy.empty((len(points), len(coeffs), nvars))
    for k in range(nvars):
        dpw = pw.copy()
        dpw[:, k] = numpy.maximum(dpw[:, k] - 1, 0)
        dmono = pw[:, k][None] * numpy.prod(points[:, None, :] ** dpw[None], axis=2)
        grads[:, :, k] = dmono @ coeffs.T
    return vals, grads


def _vertex_shape_gradients(vertices, points):
    'reference gradients [nq, nvert, nd] of the (multi)linear nodal shape functions of a reference element with the given vertices'
    import itertools
    vertices = numpy.asarray(vertices, dtype=float)
    nv, nd = vertices.shape
    if nv == nd + 1:
        pw = numpy.concatenate([numpy.zeros((1, nd), dtype=int), numpy.eye(nd, dtype=int)])
    elif nv == 2 ** nd:
        pw = numpy.array(list(itertools.product((0, 1), repeat=nd)), dtype=int)
    else:
        raise NotImplementedError('reference element with {} vertices in {} dimensions'.format(nv, nd))
    V = numpy.prod(vertices[:, None, :] ** pw[None], axis=2)   # V[v, m] = monomial m at vertex v; shape function v = sum_m inv(V)[m, v] monomial m
    coef = numpy.linalg.inv(V)
    out = numpy.empty((len(points), nv, nd))
    for k in range(nd):
        dpw = pw.copy()
        dpw[:, k] = numpy.maximum(dpw[:, k] - 1, 0)
        dmono = pw[:, k][None] * numpy.prod(numpy.asarray(points)[:, None, :] ** dpw[None], axis=2)
        out[:, :, k] = dmono @ coef
    return out


def general_tables_from_reference(topo, basis, geom, degree, ischeme='gauss'):
    '''Tables of ``topo.integral(<form in basis> * J(geom), degree=degree)`` for ANY topology and basis of the reference -- simplex
    (topology.py:2493), mixed (mesh.py:737-753) -- with a piecewise (multi)linear geometry: the arguments of
    :class:`engine.GeneralPlan`.  Per element the reference's own objects are read out: ``basis.get_dofs`` /
    ``get_coefficients`` (function.py:2795-2830), ``sample.points.get(i)`` (pointsseq.py:324-332), ``topo.references[i].vertices``;
    elements with identical reference data share one type table.'''
    import importlib
    pkg = type(topo).__module__.partition('.')[0]
    smp = importlib.import_module(pkg + '.sample')
    pseq = importlib.import_module(pkg + '.pointsseq')
    pts = importlib.import_module(pkg + '.points')
    types_mod = importlib.import_module(pkg + '.types')
    sample = topo.sample(ischeme, degree)
    nd = topo.ndims
    types, index, etype, dofs = [], {}, [], []
    refverts = []
    for i in range(len(topo)):
        p = sample.points.get(i)
        coords, weights = numpy.asarray(p.coords, dtype=float), numpy.asarray(p.weights, dtype=float)
        coeffs = numpy.asarray(basis.get_coefficients(i), dtype=float)
        verts = numpy.asarray(topo.references[i].vertices, dtype=float)
        key = coords.tobytes(), weights.tobytes(), coeffs.shape, coeffs.tobytes(), verts.tobytes()
        t = index.get(key)
        if t is None:
            phi, dphi = poly_tabulate(coeffs, coords)
            t = index[key] = len(types)
            types.append(dict(weights=weights, phi=phi, dphi=dphi, gdphi=_vertex_shape_gradients(verts, coords), vertices=verts))
        etype.append(t)
        dofs.append(numpy.asarray(basis.get_dofs(i), dtype=numpy.int64))
        refverts.append(pts.CoordsPoints(types_mod.arraydata(verts)))
    # physical vertex coordinates: the geometry evaluated by the reference at the reference vertices of every element
    vsample = smp.Sample.new(sample.spaces[0], sample.transforms, pseq.PointsSequence.from_iter(refverts, nd))
    x = numpy.asarray(vsample.eval(geom))
    verts, pos = [], 0
    for i in range(len(topo)):
        nv = len(types[etype[i]]['vertices'])
        verts.append(x[pos:pos + nv])
        pos += nv
    return dict(ndims=nd, types=types, etype=numpy.array(etype, dtype=numpy.int32), dofs=dofs, verts=verts, nbasis=len(basis))
