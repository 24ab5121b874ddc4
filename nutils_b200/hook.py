'''The reference-side half of the drop-in: evalf/nutils' own ``Sample.integral`` routed to the CUDA path.

The reference offers exactly one interception point on the integration path, the ``__nutils_dispatch__`` protocol
(src/nutils/_util.py:813-832): a positional argument whose TYPE defines ``__nutils_dispatch__`` is offered the call first
and returns ``NotImplemented`` to decline.  ``Sample.integral`` is a dispatched method (sample.py:177) and ``self`` is
positional argument 0, so

    import nutils_b200.hook; nutils_b200.hook.install()

gives ``nutils.sample.Sample`` that attribute (nothing else of the reference is touched) and every
``topo.integral(...)`` / ``sample.integral(...)`` / ``topo.integrate(...)`` of an UNMODIFIED script lands in
:func:`_dispatch`.  What comes back is a ``function.Array`` like the reference's own ``_Integral`` (sample.py:944-956)
whose lowering is an evaluable node (:class:`B200Sum`) with the semantics of the stock ``LoopSum`` over the elements:

* it supports everything ``solver.System`` does with an integral before evaluating it -- ``evaluable.derivative`` to the
  test and trial arguments (solver.py:239-251), ``simplified``, ``replace_arguments``, ``as_csr`` (evaluable.py:5679);
* when it is about to be EVALUATED (``_optimized_for_numpy`` / ``_assparse``, i.e. inside ``evaluable.compile``,
  evaluable.py:6532) the integrand is recognised by PROBING: the stock per-element evaluable is evaluated for a handful
  of elements and fitted with the closed form the kernels assemble,

      A[(i,c),(j,e)] = sum_q w_q |det J| sum_xy D[c][x][e][y] d_x N_i d_y N_j ,   b[(i,c)] = sum_q w_q |det J| sum_x C[c][x] d_x N_i

  (d_0 = value, d_k = d/dx_k) with CONSTANT D, C; the fit is verified on further elements;
* recognised: values come from ``b2_assemble_host`` (engine.Plan), the CSR pattern from ``b2_pattern_export_host``;
* anything else -- integrands with solution-dependent or position-dependent coefficients, samples that are not plain
  tensor Gauss samples of a structured topology, bases that are not StructuredBasis -- DECLINES: the node turns itself
  back into the stock ``LoopSum`` and the reference evaluates it as if the hook were not there (the reference's own
  convention for "not mine", _util.py:826-829).

The recognition is numerical, so it does not depend on how the script spells the integrand (namespace expressions,
``function.outer(basis.grad(geom))``, fields and ``System``): mass, Laplace / anisotropic diffusion, elasticity, loads.
'''

import importlib
import itertools
import typing

import numpy

from . import adapter, engine

_NT = {}            # the reference's modules, filled by install()
_REGISTRY = {}      # token -> _Info (python-side state of an integral; evaluable nodes only carry the token)
_PLANS = {}         # plan id -> recognised closed form (B200Values nodes only carry the id)
_COUNTER = itertools.count(1)
STATS = {'offered': 0, 'accelerated': 0, 'declined': 0, 'reasons': []}
PROBE_TOL = 1e-10   # relative misfit of the closed form on the probe elements above which the integrand is declined


class Declined(Exception):
    'the integral is not one the CUDA path assembles; the stock evaluable takes over'


# ---- backend: the only place that touches the device (tests on a CPU-only box substitute the oracle here) --------------

class GpuBackend:
    'engine.Plan per (basis, rules, geometry); values through b2_assemble_host, pattern through b2_pattern_export_host'

    def __init__(self, device=0):
        self.device = device
        self._plans = {}

    def plan(self, spec):
        key = spec['key']
        plan = self._plans.get(key)
        if plan is None:
            ctx = engine.Context.get(self.device)
            if 'face' in spec:
                ids, xi, w = spec['elem_ids'], spec['xi'], spec['weights']
                plan = engine.ElemSetPlan(ctx, spec['bases'], nodes=spec['nodes'], ncomp=spec['ncomp'], elem_ids=ids,
                                          qoff=numpy.arange(len(ids) + 1, dtype=numpy.int64) * len(w), qcoords=numpy.tile(xi, (len(ids), 1)), qweights=numpy.tile(w, len(ids)))
                plan.set_faces(numpy.full(len(ids), spec['face']['dim'], dtype=numpy.int8))
            elif 'elemset' in spec:
                es = spec['elemset']
                plan = engine.ElemSetPlan(ctx, spec['bases'], nodes=spec['nodes'], ncomp=spec['ncomp'], elem_ids=es['elem_ids'], qoff=es['qoff'], qcoords=es['qcoords'],
                                          qweights=es['qweights'], renumber=spec['renumber'], nbasis_new=spec['nbasis_new'])
            else:
                plan = engine.Plan(ctx, spec['bases'], spec['rules'], spec['nodes'], ncomp=spec['ncomp'])
            self._plans[key] = plan
        return plan

    def pattern(self, spec):
        return self.plan(spec).csr_pattern()

    def assemble(self, spec, Ds, Cs):
        return self.plan(spec).assemble_host(Ds, Cs)


BACKEND = GpuBackend()


def set_backend(backend):
    global BACKEND
    BACKEND = backend


# ---- host-side jets of the local basis functions (recognition only; the assembly never runs on the host) ---------------

def element_jets(bases, nodes, eidx, xi, nidx=None):
    '''Values and PHYSICAL gradients of the local functions of element `eidx` (index per dimension) at local points
    xi[nq, nd] for a multilinear nodal geometry: returns (phi[nq, n_e, 1+nd], |det J|[nq], J^-1[nq, nd, nd]); local functions in C
    order.  `nidx` is the position of the element in `nodes` when that array covers only part of the grid (default: eidx).'''
    nidx = eidx if nidx is None else nidx
    nd = len(bases)
    nq = len(xi)
    N = numpy.ones((nq, 1))
    dN = [numpy.ones((nq, 1)) for _ in range(nd)]
    for d, b in enumerate(bases):
        c = b.coeffs[b.setidx[eidx[d]]]
        p = b.degree
        v = numpy.stack([numpy.polyval(c[a], xi[:, d]) for a in range(p + 1)], 1)
        g = numpy.stack([numpy.polyval(c[a, :-1] * numpy.arange(p, 0, -1.), xi[:, d]) if p else numpy.zeros(nq) for a in range(p + 1)], 1)
        N = (N[:, :, None] * v[:, None, :]).reshape(nq, -1)
        dN = [(dN[k][:, :, None] * (g if k == d else v)[:, None, :]).reshape(nq, -1) for k in range(nd)]
    J = numpy.zeros((nq, nd, nd))
    for corner in itertools.product((0, 1), repeat=nd):
        X = nodes[(slice(None),) + tuple(e + c for e, c in zip(nidx, corner))]
        for k in range(nd):
            dphi = numpy.ones(nq)
            for d in range(nd):
                dphi = dphi * ((1. if corner[d] else -1.) if d == k else (xi[:, d] if corner[d] else 1. - xi[:, d]))
            J[:, :, k] += X[None, :] * dphi[:, None]
    Jinv = numpy.linalg.inv(J)
    det = abs(numpy.linalg.det(J))
    phi = numpy.empty((nq, N.shape[1], 1 + nd))
    phi[:, :, 0] = N
    phi[:, :, 1:] = numpy.einsum('kqa,qkj->qaj', numpy.stack(dN), Jinv)
    return phi, det, Jinv


def element_dofs(bases, eidx):
    'global (scalar) dofs of the local functions of element `eidx`, C order (function.py:3080-3093)'
    dofs = numpy.zeros(1, dtype=numpy.int64)
    for b, e in zip(bases, eidx):
        dofs = (dofs[:, None] * b.ndofs + (b.start[e] + numpy.arange(b.degree + 1))[None, :]).ravel()
    return dofs


# ---- python-side state of one intercepted integral ---------------------------------------------------------------------

class _Info:
    def __init__(self, sample, integrand, shape, coords, weights, rules, face=None):
        self.sample = sample
        self.integrand = integrand
        self.elemset = None         # dict(elem_ids, qoff, qcoords, qweights): a subset of the elements and / or per-element point sets
        if face is not None and 'elemset' in face:
            self.elemset, face = face['elemset'], None
        self.face = face            # None: volume sample; dict(dim, side, index): a side of the topology, `shape` holds the tangential dimensions
        if face is not None:
            # points in the coordinates of the adjacent VOLUME element: the fixed coordinate is 0 or 1 (the order of the tangential
            # points does not matter for a sum)
            coords = numpy.insert(coords, face['dim'], float(face['side']), axis=1)
        self.shape = shape          # elements per dimension
        self.coords = coords        # points of one element [nq, nd]
        self.weights = weights
        self.rules = rules          # per-dimension (points, weights)
        self._tree = None
        self._nodes = None
        self._digest = None

    def tree(self):
        'bases and geometries that occur in the integrand (function.Array tree walk)'
        if self._tree is None:
            F = _NT['function']
            bases, geoms, seen = [], [], set()
            stack = [self.integrand]
            while stack:
                a = stack.pop()
                if id(a) in seen:
                    continue
                seen.add(id(a))
                if isinstance(a, F.Basis):
                    if not any(a is b for b in bases):
                        bases.append(a)
                    continue
                name = type(a).__name__
                if name in ('_Jacobian', '_Gradient', '_SurfaceGradient', '_Normal', '_ExteriorNormal') and hasattr(a, '_geom'):
                    geoms.append(a._geom)
                for k, v in vars(a).items():
                    if k.startswith('_Array__'):
                        continue
                    if isinstance(v, F.Array):
                        stack.append(v)
                    elif isinstance(v, (tuple, list)):
                        stack.extend(w for w in v if isinstance(w, F.Array))
                    elif isinstance(v, dict):
                        stack.extend(w for w in v.values() if isinstance(w, F.Array))
            self._tree = bases, geoms
        return self._tree

    def volume_shape(self, nfixed=1):
        'elements per dimension of the volume grid the element indices refer to (for a side: `nfixed` elements along the fixed dimension)'
        if self.face is None:
            return self.shape
        return self.shape[:self.face['dim']] + (nfixed,) + self.shape[self.face['dim']:]

    def corner_sample(self):
        'a sample whose points are the 2^nd corners of every (adjacent volume) element'
        trans = self.sample.transforms
        if self.face is not None:
            # the layer of volume elements behind the side: the fixed axis becomes a dimension axis of length one
            tseq = importlib.import_module(_NT['nutils'].__name__ + '.transformseq')
            t0 = trans[0]
            axes = tuple(tseq.DimAxis(a.i, a.j, a.mod, False) if not a.isdim else a for a in t0._axes)
            slab = tseq.StructuredTransforms(t0._root, axes, t0._nrefine)
            trans = (slab, slab)
        nd = len(self.volume_shape())
        corners = numpy.array(list(itertools.product((0., 1.), repeat=nd)))
        pts = _NT['pointsseq'].PointsSequence.uniform(_NT['points'].CoordsPoints(_NT['types'].arraydata(corners)), self.sample.nelems)
        return _NT['sample'].Sample.new(self.sample.spaces[0], trans, pts)

    def nodes(self):
        '''nodal coordinates float64[nd, n0+1, ...] of the geometry of the integrand, evaluated by the reference at the
        element corners (one-off, host); declines if the integrand has no or more than one geometry, or if the geometry is
        discontinuous across elements.  Whether it is MULTILINEAR inside the elements is settled by the probing.'''
        if self._nodes is None:
            bases, geoms = self.tree()
            if not geoms:
                raise Declined('no geometry in the integrand')
            vshape = self.volume_shape()
            nd = len(vshape)
            cs = self.corner_sample()
            vals = []
            for g in geoms:
                if g.shape[-1] != nd:
                    raise Declined('geometry dimension differs from the topology dimension')
                if g.arguments:
                    raise Declined('geometry depends on arguments')
                x = numpy.asarray(cs.eval(g))
                x = x.reshape(x.shape[0], -1, nd)[:, 0, :]   # broadcast copies of the geometry carry leading axes
                vals.append(x)
            scale = numpy.nanmax(abs(vals[0])) or 1.
            if any(abs(v - vals[0]).max() > 1e-13 * scale for v in vals[1:]):
                raise Declined('more than one geometry in the integrand')
            if self.elemset is not None:
                # only the kept elements have corners: scatter them into the full node grid (the rest is never read)
                ids = self.elemset['elem_ids']
                full = numpy.full(vshape + (2,) * nd + (nd,), numpy.nan)
                full.reshape((-1,) + (2,) * nd + (nd,))[ids] = vals[0].reshape((len(ids),) + (2,) * nd + (nd,))
                vals[0] = full
            x = vals[0].reshape(vshape + (2,) * nd + (nd,))
            nodes = numpy.full((nd,) + tuple(n + 1 for n in vshape), numpy.nan)
            for corner in itertools.product((0, 1), repeat=nd):
                sl = tuple(slice(c, n + c) for c, n in zip(corner, vshape))
                v = numpy.moveaxis(x[(Ellipsis,) + corner + (slice(None),)], -1, 0)
                old = nodes[(slice(None),) + sl]
                if (abs(numpy.where(numpy.isnan(old) | numpy.isnan(v), 0., old - v)) > 1e-12 * scale).any():
                    raise Declined('geometry is discontinuous across elements')
                nodes[(slice(None),) + sl] = numpy.where(numpy.isnan(v), old, v)
            if self.elemset is not None:
                nodes = numpy.nan_to_num(nodes)
            self._nodes = nodes
        return self._nodes

    def probes(self, rng):
        'elements (positions in the sample) on which the integrand is probed: corners, centre and pseudo-random ones'
        nd = len(self.volume_shape())
        if self.elemset is not None:
            n = self.sample.nelems
            picks = [0, n - 1, n // 2, n // 3, (2 * n) // 3] + [int(rng.randint(n)) for _ in range(12)]
            out = []
            for k in picks:
                if k not in out:
                    out.append(k)
            return out[:14]
        if not self.shape:
            return [0]
        return [int(numpy.ravel_multi_index(e, self.shape)) for e in _probe_elements(self.shape, rng, 10 + 2 * nd)]

    def element(self, k):
        '(index of the volume element per dimension, its position in the node array, local points, weights) of sample element k'
        if self.elemset is not None:
            es = self.elemset
            e = tuple(int(i) for i in numpy.unravel_index(int(es['elem_ids'][k]), self.shape))
            sl = slice(int(es['qoff'][k]), int(es['qoff'][k + 1]))
            return e, e, es['qcoords'][sl], es['qweights'][sl]
        t = tuple(int(i) for i in numpy.unravel_index(k, self.shape)) if self.shape else ()
        if self.face is None:
            return t, t, self.coords, self.weights
        f = self.face
        return t[:f['dim']] + (f['index'],) + t[f['dim']:], t[:f['dim']] + (0,) + t[f['dim']:], self.coords, self.weights

    def nodes_digest(self):
        if self._digest is None:
            import hashlib
            self._digest = hashlib.sha1(numpy.ascontiguousarray(self.nodes()).tobytes()).hexdigest()
        return self._digest


def _sample_info(sample):
    'shape / points of a sample the CUDA path covers, else Declined'
    if type(sample).__name__ != '_DefaultIndex' or len(sample.spaces) != 1:
        raise Declined('sample type {}'.format(type(sample).__name__))
    trans = sample.transforms[0]
    elem_ids = None
    if type(trans).__name__ == 'MaskedTransforms' and type(trans._parent).__name__ == 'StructuredTransforms':
        # a subset of the elements of a structured topology: trimmed / subset topologies (topology.py:2615-2740)
        elem_ids = numpy.asarray(trans._indices, dtype=numpy.int64)
        trans = trans._parent
    if type(trans).__name__ != 'StructuredTransforms':
        raise Declined('not a sample of a structured topology')
    axes = trans._axes
    fixed = [k for k, a in enumerate(axes) if not getattr(a, 'isdim', False)]
    if any(getattr(a, 'isperiodic', False) for a in axes) or len(fixed) > 1 or (fixed and type(axes[fixed[0]]).__name__ != 'IntAxis'):
        raise Declined('not a volume or boundary sample of a non-periodic structured topology')
    face = None
    if fixed:
        # a side of the topology (topology.py:2049-2057): the elements are the faces of the adjacent volume elements
        a = axes[fixed[0]]
        if len(a) != 1:
            raise Declined('interfaces are not boundary sides')
        face = dict(dim=fixed[0], side=int(bool(a.side)), index=int(a.map(0)))
    shape = tuple(len(a) for k, a in enumerate(axes) if k not in fixed)
    if sample.nelems == 0:
        raise Declined('empty sample')
    if type(sample.points).__name__ != '_Uniform' or elem_ids is not None:
        # cut cells carry their own point sets (pointsseq.py:324-332, points.py:257-337): ragged tables of an element set
        if face is not None:
            raise Declined('ragged points on a boundary side')
        nd = len(shape)
        if not 1 <= nd <= 3:
            raise Declined('dimension')
        qc, qw, qoff = [], [], [0]
        try:
            for i in range(sample.nelems):
                p = sample.points.get(i)
                qc.append(numpy.asarray(p.coords, dtype=float).reshape(-1, nd))
                qw.append(numpy.asarray(p.weights, dtype=float))
                qoff.append(qoff[-1] + len(qw[-1]))
        except Exception:
            raise Declined('points without weights')
        es = dict(elem_ids=numpy.arange(sample.nelems, dtype=numpy.int64) if elem_ids is None else elem_ids, qoff=numpy.array(qoff, dtype=numpy.int64),
                  qcoords=numpy.concatenate(qc), qweights=numpy.concatenate(qw))
        return shape, None, None, None, dict(elemset=es)
    pts = sample.points.get(0)
    try:
        coords = numpy.asarray(pts.coords, dtype=float)
        weights = numpy.asarray(pts.weights, dtype=float)
    except Exception:
        raise Declined('points without weights')
    nd = len(shape)
    if coords.ndim != 2 or coords.shape[1] != nd or not (1 <= nd <= 3 or (face and nd == 0)) or len(axes) > 3:
        raise Declined('dimension')
    # tensor rule?  per-dimension points ascending, C-order outer product (points.py:144-164)
    rules = []
    nq = [len(numpy.unique(numpy.round(coords[:, d], 14))) for d in range(nd)]
    if int(numpy.prod(nq)) != len(weights):
        raise Declined('points are not a tensor rule')
    W = weights.reshape(nq)
    X = coords.reshape(tuple(nq) + (nd,))
    for d in range(nd):
        idx = [0] * nd
        idx[d] = slice(None)
        x = X[tuple(idx) + (d,)]
        w = W.sum(axis=tuple(k for k in range(nd) if k != d))
        rules.append((numpy.array(x), numpy.array(w)))
    from . import points as _points
    if nd == 0:
        return shape, coords, weights, rules, face
    tc, tw = _points.tensor_points(rules)
    if abs(tc - coords).max() > 1e-14 or abs(tw - weights).max() > 1e-14 * abs(weights).max() or any(len(r[0]) > 5 for r in rules):
        raise Declined('points are not a tensor rule')
    return shape, coords, weights, rules, face


# ---- dispatch ----------------------------------------------------------------------------------------------------------

def _dispatch(cls, func, args, kwargs):
    'the __nutils_dispatch__ classmethod given to nutils.sample.Sample (_util.py:813-832)'
    S = _NT['sample'].Sample
    if func is not S.integral or len(args) != 2 or kwargs:
        return NotImplemented
    sample, integrand = args
    STATS['offered'] += 1
    try:
        shape, coords, weights, rules, face = _sample_info(sample)
    except Declined as e:
        _note_declined(str(e))
        return NotImplemented
    integrand = _NT['function'].Array.cast(integrand)
    if integrand.dtype == complex:
        _note_declined('complex integrand')
        return NotImplemented
    token = next(_COUNTER)
    _REGISTRY[token] = _Info(sample, integrand, shape, coords, weights, rules, face)
    return _NT['Integral'](integrand, sample, token)


def _note_declined(reason):
    STATS['declined'] += 1
    if len(STATS['reasons']) < 100:
        STATS['reasons'].append(reason)


def install(nutils=None):
    '''Activate the hook on the reference package `nutils` (default: ``import nutils``).  Idempotent.'''
    nt = nutils or importlib.import_module('nutils')
    if _NT.get('nutils') is nt:
        return
    for name in ('sample', 'function', 'evaluable', 'points', 'pointsseq', 'types', '_util', 'util'):
        try:
            _NT[name] = importlib.import_module(nt.__name__ + '.' + name)
        except ImportError:
            if name != 'util':
                raise
    _NT['nutils'] = nt
    _define_classes()
    _NT['sample'].Sample.__nutils_dispatch__ = classmethod(_dispatch)


def uninstall():
    if 'sample' in _NT and '__nutils_dispatch__' in vars(_NT['sample'].Sample):
        del _NT['sample'].Sample.__nutils_dispatch__
    _NT.clear()


# ---- recognition by probing --------------------------------------------------------------------------------------------

def _probe_elements(shape, rng, n):
    'corner, centre and pseudo-random elements (index tuples), no duplicates'
    out = []
    for c in itertools.product(*[(0, m - 1) for m in shape]):
        if c not in out:
            out.append(c)
    c = tuple(m // 2 for m in shape)
    if c not in out:
        out.append(c)
    for _ in range(4 * n):
        if len(out) >= n:
            break
        c = tuple(int(rng.randint(m)) for m in shape)
        if c not in out:
            out.append(c)
    return out


def _recognise(node):
    'closed form of a B200Sum node -> dict(spec, kind, D or C, ...) or Declined'
    ev = _NT['evaluable']
    info = _REGISTRY[node.token]
    if any(isinstance(a, ev.Argument) for a in node.arguments):
        raise Declined('integrand depends on arguments (nonlinear or not yet linearised)')
    groups = node.groups
    if len(groups) not in (1, 2) or any(g not in (1, 2) for g in groups):
        raise Declined('{} array axes groups'.format(len(groups)))
    dims = [int(n.__index__()) for n in node.shape]
    nd = len(info.volume_shape())
    na = nd + 1
    # axes of every group: basis axis [, component axis]
    pos = 0
    gaxes = []
    for g in groups:
        gaxes.append(tuple(range(pos, pos + g)))
        pos += g
    ncomp = [dims[ax[1]] if len(ax) == 2 else 1 for ax in gaxes]
    if len(groups) == 2 and ncomp[0] != ncomp[1]:
        raise Declined('test and trial spaces with different numbers of components')
    nc = ncomp[0]
    if nc > 3:
        raise Declined('more than three components')
    fbases, _ = info.tree()
    def on_topology(b):
        if info.elemset is not None:   # the basis of a trimmed / subset topology is a PrunedBasis of a structured one (function.py:3103-3133)
            parent = getattr(b, '_parent', None)
            if parent is None:
                return hasattr(b, '_start_dofs') and tuple(b._transforms_shape) == info.shape and len(info.elemset['elem_ids']) == int(numpy.prod(info.shape))
            return hasattr(parent, '_start_dofs') and tuple(parent._transforms_shape) == info.shape and numpy.array_equal(numpy.asarray(b._transmap), info.elemset['elem_ids'])
        bshape = tuple(b._transforms_shape)
        if info.face is None:
            return bshape == info.shape
        f = info.face
        return len(bshape) == nd and bshape == info.volume_shape(bshape[f['dim']]) and f['index'] == (bshape[f['dim']] - 1 if f['side'] else 0)
    cands = [b for b in fbases if (hasattr(b, '_start_dofs') or hasattr(getattr(b, '_parent', None), '_start_dofs')) and on_topology(b) and all(len(b) == dims[ax[0]] for ax in gaxes)]
    if not cands:
        raise Declined('no structured basis of the topology matches the array axes')
    nodes = info.nodes()
    # stock sparse chunks of the per-element integral with the loop index turned into an argument
    probe_arg = ev.InRange(ev.Argument('_b200_ielem', (), int), node.length)
    index = node.index
    chunks = tuple(tuple(_replace_index(a, index, probe_arg) for a in chunk) for chunk in node.func.simplified._assparse)
    if not chunks:
        raise Declined('integrand is identically zero')
    f = ev.compile(chunks)
    rng = numpy.random.RandomState(len(dims) * 7919 + info.sample.nelems)
    elems = info.probes(rng)
    data = []
    for ielem in elems:
        data.append([tuple(numpy.asarray(a).ravel() for a in chunk) for chunk in f({'_b200_ielem': ielem})])
    last = None
    for basis in cands:
        try:
            return _fit(node, info, basis, nodes, elems, data, gaxes, nc)
        except Declined as e:
            last = e
    raise last


def _fit(node, info, basis, nodes, elems, data, gaxes, nc):
    nd = len(info.volume_shape())
    na = nd + 1
    parent = getattr(basis, '_parent', basis)
    bases = adapter.bases1d_from_structured_basis(parent)
    renumber = numpy.asarray(basis._renumber, dtype=numpy.int64) if parent is not basis else None
    k = len(gaxes)
    face = info.face
    A_rows, rhs = [], []
    blocks = []
    for ielem, chunks in zip(elems, data):
        e, nidx, xi, wq = info.element(ielem)
        ldofs = element_dofs(bases, e)
        if renumber is not None:
            ldofs = renumber[ldofs]
        lookup = {int(d): a for a, d in enumerate(ldofs)}
        ne = len(ldofs)
        T = numpy.zeros((ne, nc) * k)
        for chunk in chunks:
            *idx, vals = chunk
            if not len(vals):
                continue
            # entries outside the element's own functions are tolerated when they are exact zeros (dense chunks)
            keep = numpy.ones(len(vals), dtype=bool)
            for ax in gaxes:
                keep &= numpy.isin(idx[ax[0]], ldofs)
            if (vals[~keep] != 0).any():
                raise Declined('dofs of the integrand are not those of the structured basis')
            idx = [i[keep] for i in idx]
            vals = vals[keep]
            where = []
            for ax in gaxes:
                where.append(numpy.array([lookup[int(i)] for i in idx[ax[0]]], dtype=int))
                where.append(idx[ax[1]] if len(ax) == 2 else numpy.zeros(len(vals), dtype=int))
            numpy.add.at(T, tuple(where), vals)
        phi, det, Jinv = element_jets(bases, nodes, e, xi, nidx)
        wdet = wq * det
        if face is not None:   # surface measure |det J| |J^-T e_k| (function.py:2291-2316 on a boundary sample)
            wdet = wdet * numpy.linalg.norm(Jinv[:, face['dim'], :], axis=1)
        if k == 2:
            E = numpy.einsum('q,qax,qby->xyab', wdet, phi, phi).reshape(na * na, ne * ne)
        else:
            E = numpy.einsum('q,qax->xa', wdet, phi)
        blocks.append((T, E))
    nfit = max(len(blocks) - 4, (len(blocks) + 1) // 2)   # the last elements are the hold-out set
    scale = max(abs(T).max() for T, E in blocks)
    if scale == 0.:
        raise Declined('integrand vanishes on the probe elements')
    if k == 2:
        D = numpy.zeros((nc, na, nc, na))
        for c in range(nc):
            for e_ in range(nc):
                A = numpy.concatenate([E.T for T, E in blocks[:nfit]])
                b = numpy.concatenate([T[:, c, :, e_].ravel() for T, E in blocks[:nfit]])
                sol = numpy.linalg.lstsq(A, b, rcond=None)[0]
                D[c, :, e_, :] = sol.reshape(na, na)
        for T, E in blocks:
            model = numpy.einsum('cxey,xyab->acbe', D, E.reshape(na, na, T.shape[0], T.shape[0]))
            if abs(model - T).max() > PROBE_TOL * scale:
                raise Declined('integrand is not a constant-coefficient bilinear form in (N, grad N)')
        D[abs(D) < 1e-13 * abs(D).max()] = 0.
        Dt = D.transpose(2, 3, 0, 1)
        if abs(D - Dt).max() <= 1e-13 * abs(D).max():
            D = .5 * (D + Dt)
        coef = D
    else:
        C = numpy.zeros((nc, na))
        for c in range(nc):
            A = numpy.concatenate([E.T for T, E in blocks[:nfit]])
            b = numpy.concatenate([T[:, c] for T, E in blocks[:nfit]])
            C[c] = numpy.linalg.lstsq(A, b, rcond=None)[0]
        for T, E in blocks:
            if abs(numpy.einsum('cx,xa->ac', C, E) - T).max() > PROBE_TOL * scale:
                raise Declined('integrand is not a constant-coefficient linear form in (N, grad N)')
        C[abs(C) < 1e-13 * abs(C).max()] = 0.
        coef = C
    spec = dict(bases=bases, rules=info.rules, nodes=nodes, ncomp=nc, key=(id(basis), id(info.sample), nc, info.nodes_digest()))
    if face is not None:
        # the adjacent volume elements as an element set with points on their faces (b2_elemset_set_faces)
        vshape = tuple(b.nelems for b in bases)
        full = numpy.zeros((nd,) + tuple(n + 1 for n in vshape))
        lo = face['index']
        sl = (slice(None),) * (1 + face['dim']) + (slice(lo, lo + 2),)
        full[sl] = nodes
        grids = [numpy.arange(n) for n in vshape]
        grids[face['dim']] = numpy.array([lo])
        elem_ids = numpy.sort(numpy.ravel_multi_index(numpy.stack(numpy.meshgrid(*grids, indexing='ij'), -1).reshape(-1, nd).T, vshape)).astype(numpy.int64)
        spec.update(nodes=full, face=dict(face), elem_ids=elem_ids, xi=numpy.array(info.coords), weights=numpy.array(info.weights))
    if info.elemset is not None:
        spec.update(elemset=info.elemset, renumber=renumber, nbasis_new=len(basis))
    plan = dict(spec=spec, kind='matrix' if k == 2 else 'vector', coef=coef, nc=nc, nbasis=len(basis), gaxes=gaxes, keep=(basis, info), id=next(_COUNTER))   # nbasis: functions of the (pruned) basis
    _PLANS[plan['id']] = plan
    return plan


def _plan(node):
    d = node.__dict__
    if '_b200_plan' not in d:
        try:
            d['_b200_plan'] = _recognise(node)
            STATS['accelerated'] += 1
        except Declined as e:
            d['_b200_plan'] = None
            _note_declined(str(e))
    return d['_b200_plan']


def _replace_index(value, index, new):
    return _NT['replace_index'](value, index, new)


# ---- the classes that subclass the reference's (created once the reference is imported) --------------------------------

def _define_classes():
    ev = _NT['evaluable']
    F = _NT['function']
    util = _NT['_util']

    @util.shallow_replace
    def replace_index(value, index, new):
        if value is index:
            return new
    _NT['replace_index'] = replace_index

    class Integral(F.Array):
        'the counterpart of nutils.sample._Integral (sample.py:944-956) whose lowering is a B200Sum instead of a LoopSum'

        def __init__(self, integrand, sample, token):
            self._integrand = integrand
            self._sample = sample
            self._token = token
            super().__init__(shape=integrand.shape, dtype=float if integrand.dtype in (bool, int) else integrand.dtype,
                             spaces=integrand.spaces - frozenset(sample.spaces), arguments=integrand.arguments)

        def lower(self, args):
            ielem = ev.loop_index('_sample_{}'.format(len(args.args)), self._sample.nelems)
            weights = ev.astype(self._sample.get_evaluable_weights(ielem), self.dtype)
            integrand = ev.astype(self._integrand.lower(args * self._sample.get_lower_args(ielem)), self.dtype)
            elem_integral = ev.einsum('B,ABC->AC', weights, integrand, B=weights.ndim, C=self.ndim)
            if len(args.points_shape):   # an integral evaluated inside another sample: not ours
                return ev.loop_sum(elem_integral, ielem)
            return B200Sum(elem_integral, elem_integral.shape, ielem.loop_id, ielem.length, self._token, (1,) * self.ndim)

    class B200Sum(ev.Array):
        '''sum over the elements of `func` (which depends on the loop index): the semantics of evaluable.LoopSum
        (evaluable.py:5234-5343), evaluated by the CUDA path when the integrand is recognised, by the stock loop otherwise'''

        func: ev.Array
        shape: typing.Tuple[ev.Array, ...]
        loop_id: ev._LoopId
        length: ev.Array
        token: int
        groups: typing.Tuple[int, ...]

        @property
        def dtype(self):
            return self.func.dtype

        @property
        def index(self):
            return ev._LoopIndex(self.loop_id, self.length)

        @property
        def dependencies(self):
            args = sorted((a for a in self.func.arguments if isinstance(a, ev.Argument)), key=lambda a: a.name)
            return (*self.shape, self.length, *args)

        @property
        def stock(self):
            return ev.loop_sum(self.func, self.index)

        def _derivative(self, var, seen):
            if var.dtype in (bool, int) or var not in self.arguments:
                return ev.Zeros(self.shape + var.shape, dtype=self.dtype)
            return B200Sum(ev.derivative(self.func, var, seen), self.shape + var.shape, self.loop_id, self.length, self.token, self.groups + (len(var.shape),))

        def _simplified(self):
            if ev.iszero(self.func):
                return ev.zeros_like(self)

        def _optimized_for_numpy(self):
            # the last rewriting stage of evaluable.compile (evaluable.py:6691-6694): decide who evaluates this integral
            if _plan(self) is None:
                return self.stock

        def _argument_degree(self, argument):
            return self.func.argument_degree(argument)

        @property
        def _assparse(self):
            plan = _plan(self)
            if plan is None or plan['kind'] != 'matrix':
                return self.stock._assparse if plan is None else super()._assparse
            rowptr, colidx = BACKEND.pattern(plan['spec'])
            rows = numpy.repeat(numpy.arange(len(rowptr) - 1, dtype=numpy.int64), numpy.diff(rowptr))
            nc = plan['nc']
            idx = []
            for ax, dof in zip(plan['gaxes'], (rows, colidx)):
                if len(ax) == 2:
                    idx += [dof // nc, dof % nc]
                else:
                    idx.append(dof)
            return (*(ev.constant(i) for i in idx), B200Values(plan['id'])),

        def _evalf(self, *args):
            plan = _plan(self)
            if plan is None:
                raise RuntimeError('B200Sum evaluated without a plan (evaluable.compile(_optimize=False) is not supported)')
            shape = tuple(int(n) for n in args[:len(self.shape)])
            if plan['kind'] == 'vector':
                (), (rhs,) = BACKEND.assemble(plan['spec'], [], [plan['coef']])
                return rhs.reshape(shape)
            (values,), () = BACKEND.assemble(plan['spec'], [plan['coef']], [])
            rowptr, colidx = BACKEND.pattern(plan['spec'])
            n = len(rowptr) - 1
            dense = numpy.zeros((n, n))
            dense[numpy.repeat(numpy.arange(n), numpy.diff(rowptr)), colidx] = values
            nb, nc = plan['nbasis'], plan['nc']
            if nc > 1 or any(len(ax) == 2 for ax in plan['gaxes']):
                dense = dense.reshape(nb, nc, nb, nc)
            return dense.reshape(shape)

        def _compile(self, builder):
            args = builder.compile(self.dependencies)
            evalf = builder.add_constant(self._evalf)
            out = builder.get_variable_for_evaluable(self)
            builder.get_block_for_evaluable(self).assign_to(out, evalf.call(*args))
            return out

    class B200Values(ev.Array):
        'the stored values (CSR order of the analytic pattern) of a recognised, argument-free matrix-valued B200Sum'

        plan_id: int

        dtype = float
        dependencies = ()

        @property
        def shape(self):
            rowptr, colidx = BACKEND.pattern(_PLANS[self.plan_id]['spec'])
            return ev.constant(len(colidx)),

        def _evalf(self):
            plan = _PLANS[self.plan_id]
            (values,), () = BACKEND.assemble(plan['spec'], [plan['coef']], [])
            return values

        def _compile(self, builder):
            evalf = builder.add_constant(self._evalf)
            out = builder.get_variable_for_evaluable(self)
            builder.get_block_for_evaluable(self).assign_to(out, evalf.call())
            return out

    _NT['Integral'] = Integral
    _NT['B200Sum'] = B200Sum
    _NT['B200Values'] = B200Values
