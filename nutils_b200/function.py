'''Minimal lazy-array layer for integrands on the accelerated path.

The reference represents integrands as ``nutils.function.Array`` trees that are
lowered to evaluables and compiled to a Python loop (function.py:257-650,
sample.py:944-956).  The B200 path covers the integrands of the north-star
forms -- (bi)linear forms in a tensor-product basis with constant coefficients
on a multilinear geometry -- and represents every such array in closed form:

    value[plain axes..., dof axes...] = sum over slots of
        C[plain..., (c1, x1), (c2, x2)] * d_x1 N^(c1)_i * d_x2 N^(c2)_j   [* |det J|]

with d_0 the value and d_k the derivative to physical coordinate k-1.  ``C`` is
a small dense numpy tensor; multiplication, indexing with ``None``, summation,
gradients etc. act on ``C``.  At integration time ``C`` IS the coefficient
tensor ``D`` (two dof axes) or ``C`` (one dof axis) of the C ABI
(include/b200fem.h, b2_assemble_*).  Anything outside this closed form raises
``NotImplementedError`` -- the reference's own dispatch convention for "not
mine" (``_util.nutils_dispatch`` returning NotImplemented, _util.py:813-832).

API names follow the reference: ``Array.grad``, ``Array.sum``, ``outer``,
``J``, ``eval``, ``as_csr``, ``as_coo`` (function.py:2050-2316, 2408-2452).
'''

import numpy


class Geometry:
    '''Multilinear nodal geometry of a structured topology (what ``mesh.rectilinear`` returns next to the topology).'''

    def __init__(self, topo, nodes):
        self.topo = topo
        # own, READ-ONLY copy: device-resident plans are cached per geometry object, so the coordinates must not change
        # behind their back -- a deformed mesh is a new geometry (topo.nodal_geometry(new_nodes))
        self.nodes = numpy.array(nodes, dtype=float, order='C')
        self.nodes.setflags(write=False)
        self.ndims = self.nodes.shape[0]
        self.shape = (self.ndims,)

    def __repr__(self):
        return 'Geometry<{}>'.format('x'.join(str(n - 1) for n in self.nodes.shape[1:]))

    def __getitem__(self, i):
        'coordinate function x_i: a PointFunction usable as a coefficient, e.g. ``numpy.cosh(geom[0]) * basis * J(geom)``'
        i = int(i)
        if not -self.ndims <= i < self.ndims:
            raise IndexError(i)
        return PointFunction(self, lambda x, i=i: x[..., i])

    def __iter__(self):
        return (self[i] for i in range(self.ndims))

    def __len__(self):
        return self.ndims


class PointFunction:
    '''Scalar function of the physical coordinates, evaluated on the host at the quadrature points and handed to the
    kernels as a per-point coefficient (b2_elemset_set_coefficient).  Built from ``geom[i]`` with arithmetic and numpy
    ufuncs -- the spelling of the reference's examples (``numpy.cosh(x[0])``, examples/laplace.py:60-70).'''

    __array_priority__ = 90.

    def __init__(self, geom, f):
        self.geom = geom
        self.f = f

    def __call__(self, x):
        return numpy.broadcast_to(self.f(x), x.shape[:-1])

    def _lift(self, other):
        if isinstance(other, PointFunction):
            if other.geom is not self.geom:
                raise NotImplementedError('functions of different geometries')
            return other.f
        if isinstance(other, (Array, Geometry, _Jacobian)) or hasattr(other, '_array'):
            return None
        c = float(other)
        return lambda x: c

    def _binary(self, other, op, swap=False):
        g = self._lift(other)
        if g is None:
            return NotImplemented
        f = self.f
        return PointFunction(self.geom, (lambda x: op(g(x), f(x))) if swap else (lambda x: op(f(x), g(x))))

    def __add__(self, o): return self._binary(o, numpy.add)
    def __radd__(self, o): return self._binary(o, numpy.add, True)
    def __sub__(self, o): return self._binary(o, numpy.subtract)
    def __rsub__(self, o): return self._binary(o, numpy.subtract, True)
    def __truediv__(self, o): return self._binary(o, numpy.divide)
    def __rtruediv__(self, o): return self._binary(o, numpy.divide, True)
    def __pow__(self, o): return self._binary(o, numpy.power)
    def __neg__(self): return PointFunction(self.geom, lambda x, f=self.f: -f(x))

    def __mul__(self, o):
        if isinstance(o, (Array, _Jacobian)) or hasattr(o, '_array'):
            return Array.cast(o) * self if not isinstance(o, _Jacobian) else NotImplemented
        return self._binary(o, numpy.multiply)

    __rmul__ = __mul__

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__' or kwargs:
            return NotImplemented
        fs = [self._lift(a) for a in inputs]
        if any(f is None for f in fs):
            return NotImplemented
        return PointFunction(self.geom, lambda x: ufunc(*[f(x) for f in fs]))


class Array:
    '''Closed-form integrand: see the module docstring.

    shape    : full shape, dof axes have length ndofs
    dofaxes  : tuple of axis positions that are dof axes (at most two, increasing)
    space    : the Basis the dof axes refer to (one space per array)
    C        : float64[plainshape + (ncomp, na) * len(dofaxes)], plainshape = shape with dof axes set to 1
    jac      : geometry whose |det J| multiplies the array, or None
    '''

    __array_priority__ = 100.  # numpy defers to our __rmul__ etc.

    def __init__(self, shape, dofaxes, space, C, jac=None, coef=None):
        self.shape = tuple(int(n) for n in shape)
        self.dofaxes = tuple(dofaxes)
        self.space = space
        self.C = numpy.asarray(C, dtype=float)
        self.jac = jac
        self.coef = coef  # PointFunction multiplying the whole array, or None
        assert self.C.ndim == len(self.shape) + 2 * len(self.dofaxes)

    ndim = property(lambda self: len(self.shape))

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return 'Array<{}>'.format(','.join(map(str, self.shape)))

    # -- helpers -----------------------------------------------------------------------------------

    @property
    def _plainshape(self):
        return tuple(1 if i in self.dofaxes else n for i, n in enumerate(self.shape))

    @staticmethod
    def cast(value):
        if isinstance(value, Array):
            return value
        arr = getattr(value, '_array', None)  # topology.Basis
        if isinstance(arr, Array):
            return arr
        if isinstance(value, Geometry):
            raise NotImplementedError('geometry-dependent coefficients are outside the accelerated path')
        a = numpy.asarray(value, dtype=float)
        return Array(a.shape, (), None, a)

    def _with(self, **kw):
        d = dict(shape=self.shape, dofaxes=self.dofaxes, space=self.space, C=self.C, jac=self.jac, coef=self.coef)
        d.update(kw)
        return Array(**d)

    # -- structure ---------------------------------------------------------------------------------

    def __getitem__(self, item):
        if not isinstance(item, tuple):
            item = item,
        if Ellipsis in item:
            k = item.index(Ellipsis)
            nfill = self.ndim - sum(1 for it in item if it is not None and it is not Ellipsis)
            item = item[:k] + (slice(None),) * nfill + item[k + 1:]
        item = item + (slice(None),) * (self.ndim - sum(1 for it in item if it is not None))
        shape, dofaxes, cidx = [], [], []
        ax = 0
        for it in item:
            if it is None:
                shape.append(1)
                cidx.append(None)
                continue
            if ax in self.dofaxes:
                if not (isinstance(it, slice) and it == slice(None)):
                    raise NotImplementedError('slicing a dof axis')
                dofaxes.append(len(shape))
                shape.append(self.shape[ax])
                cidx.append(slice(None))
            elif isinstance(it, slice):
                n = len(range(*it.indices(self.shape[ax])))
                shape.append(n)
                cidx.append(it)
            else:
                cidx.append(int(it))
            ax += 1
        C = self.C[tuple(cidx) + (Ellipsis,)]
        return Array(shape, dofaxes, self.space, C, self.jac, self.coef)

    def sum(self, axis=None):
        axes = range(self.ndim) if axis is None else ([axis] if numpy.ndim(axis) == 0 else list(axis))
        axes = sorted({a % self.ndim for a in axes}, reverse=True)
        out = self
        for a in axes:
            if a in out.dofaxes:
                raise NotImplementedError('summation over a dof axis')
            shape = out.shape[:a] + out.shape[a + 1:]
            dofaxes = tuple(d - (d > a) for d in out.dofaxes)
            C = out.C
            if C.shape[a] == 1 and out.shape[a] != 1:
                C = C * out.shape[a]
            out = Array(shape, dofaxes, out.space, C.sum(a), out.jac, out.coef)
        return out

    def swapaxes(self, a, b):
        a %= self.ndim
        b %= self.ndim
        perm = list(range(self.ndim))
        perm[a], perm[b] = perm[b], perm[a]
        return self.transpose(perm)

    def transpose(self, axes=None):
        axes = list(range(self.ndim))[::-1] if axes is None else [a % self.ndim for a in axes]
        shape = tuple(self.shape[a] for a in axes)
        newdof = [axes.index(d) for d in self.dofaxes]
        nslot = 2 * len(self.dofaxes)
        C = self.C.transpose(list(axes) + [self.ndim + k for k in range(nslot)])
        if len(newdof) == 2 and newdof[0] > newdof[1]:
            C = C.transpose(list(range(self.ndim)) + [self.ndim + 2, self.ndim + 3, self.ndim, self.ndim + 1])
            newdof = newdof[::-1]
        return Array(shape, newdof, self.space, C, self.jac, self.coef)

    T = property(lambda self: self.transpose())

    def trace(self, axis1=-2, axis2=-1):
        a, b = sorted((axis1 % self.ndim, axis2 % self.ndim))
        if a in self.dofaxes or b in self.dofaxes or self.shape[a] != self.shape[b]:
            raise NotImplementedError('trace over these axes')
        n = self.shape[a]
        acc = None
        for k in range(n):
            idx = [slice(None)] * self.ndim
            idx[a] = k
            idx[b] = k
            term = self[tuple(idx)]
            acc = term if acc is None else acc + term
        return acc

    # -- calculus ----------------------------------------------------------------------------------

    def grad(self, geom, ndims=0):
        '''gradient to the geometry: appends an axis of length ndims (function.py:2050-2080).'''
        if not isinstance(geom, Geometry):
            raise NotImplementedError('gradient to a non-nodal geometry')
        if len(self.dofaxes) != 1:
            raise NotImplementedError('gradient of an array with {} dof axes'.format(len(self.dofaxes)))
        if numpy.any(self.C[..., 1:] != 0):
            raise NotImplementedError('second derivatives')
        nd = geom.ndims
        C = numpy.zeros(self.C.shape[:self.ndim] + (nd,) + self.C.shape[self.ndim:])
        for k in range(nd):
            C[(slice(None),) * self.ndim + (k, slice(None), 1 + k)] = self.C[..., 0]
        return Array(self.shape + (nd,), self.dofaxes, self.space, C, self.jac, self.coef)

    def div(self, geom):
        return self.grad(geom).trace(-2, -1)

    def symgrad(self, geom):
        g = self.grad(geom)
        return .5 * (g + g.swapaxes(-2, -1))

    # -- arithmetic --------------------------------------------------------------------------------

    def _binary_linear(self, other, sign):
        other = Array.cast(other)
        if other.coef is not self.coef:
            raise NotImplementedError('sum of arrays with different coefficient functions (integrate them separately)')
        if other.dofaxes != self.dofaxes or other.shape != self.shape or (self.dofaxes and other.space is not self.space) or other.jac is not self.jac:
            a, b = _broadcast_plain(self, other)
            if a.dofaxes != b.dofaxes or (a.dofaxes and a.space is not b.space) or a.jac is not b.jac:
                raise NotImplementedError('sum of arrays with different dof structure')
            shape = tuple(max(m, n) for m, n in zip(a.shape, b.shape))
            return Array(shape, a.dofaxes, a.space, a.C + sign * b.C, a.jac, a.coef)
        return self._with(C=self.C + sign * other.C)

    def __add__(self, other):
        return self._binary_linear(other, 1.)

    __radd__ = __add__

    def __sub__(self, other):
        return self._binary_linear(other, -1.)

    def __rsub__(self, other):
        return (-self)._binary_linear(other, 1.)

    def __neg__(self):
        return self._with(C=-self.C)

    def __truediv__(self, other):
        if isinstance(other, (Array, Geometry)):
            raise NotImplementedError('division by a function')
        return self * (1. / numpy.asarray(other, dtype=float))

    def __mul__(self, other):
        if isinstance(other, _Jacobian):
            if self.jac is not None:
                raise NotImplementedError('product of two jacobians')
            return self._with(jac=other.geom)
        if isinstance(other, PointFunction):
            return self._with(coef=other if self.coef is None else self.coef * other)
        other = Array.cast(other)
        a, b = _broadcast_plain(self, other)
        coef = a.coef if b.coef is None else b.coef if a.coef is None else a.coef * b.coef
        if a.jac is not None and b.jac is not None:
            raise NotImplementedError('product of two jacobians')
        jac = a.jac if a.jac is not None else b.jac
        if a.dofaxes and b.dofaxes:
            if a.space is not b.space:
                raise NotImplementedError('products of different bases (mixed forms)')
            if len(a.dofaxes) + len(b.dofaxes) > 2 or set(a.dofaxes) & set(b.dofaxes):
                raise NotImplementedError('nonlinear product of basis functions')
            nd = a.ndim
            (da,), (db,) = a.dofaxes, b.dofaxes
            Ca = a.C[(Ellipsis, None, None)]                       # plain, c1, x1, 1, 1
            Cb = b.C[(slice(None),) * nd + (None, None)]           # plain, 1, 1, c2, x2
            C = Ca * Cb
            if da > db:  # slot groups follow the order of the dof axes
                C = C.transpose(list(range(nd)) + [nd + 2, nd + 3, nd, nd + 1])
            shape = tuple(max(m, n) for m, n in zip(a.shape, b.shape))
            return Array(shape, sorted((da, db)), a.space, C, jac, coef)
        if b.dofaxes:
            a, b = b, a
        nslot = 2 * len(a.dofaxes)
        C = a.C * b.C[(Ellipsis,) + (None,) * nslot]
        shape = tuple(max(m, n) for m, n in zip(a.shape, b.shape))
        return Array(shape, a.dofaxes, a.space, C, jac, coef)

    __rmul__ = __mul__

    def dot(self, other, axes=None):
        '''Inner product with the reference's semantics (function.py:484-524): without `axes` the second argument must be a vector
        and is contracted with the FIRST axis of self.'''
        other = Array.cast(other)
        if axes is None:
            if other.ndim != 1 or other.shape[0] != self.shape[0]:
                raise ValueError('dot without axes: the second argument must be a vector matching the first axis')
            if self.ndim > 1:
                other = other[(slice(None),) + (None,) * (self.ndim - 1)]
            axes = 0
        return (self * other).sum(axes)

    # -- integration-time views --------------------------------------------------------------------

    def coefficient_tensor(self):
        '''(kind, tensor): ('matrix', D[nc,na,nc,na]) or ('vector', C[nc,na]) for arrays whose only axes are dof axes.'''
        if self.ndim != len(self.dofaxes) or not self.dofaxes:
            raise NotImplementedError('integrand must have shape (ndofs,) or (ndofs, ndofs), got {}'.format(self.shape))
        C = self.C.reshape(self.C.shape[self.ndim:])
        return ('matrix' if len(self.dofaxes) == 2 else 'vector'), numpy.ascontiguousarray(C)


def _broadcast_plain(a, b):
    'align the number of axes (prepending) of two arrays; plain broadcasting is left to numpy on C'
    nd = max(a.ndim, b.ndim)

    def pad(x):
        k = nd - x.ndim
        if not k:
            return x
        return Array((1,) * k + x.shape, tuple(d + k for d in x.dofaxes), x.space, x.C[(None,) * k], x.jac, x.coef)
    a, b = pad(a), pad(b)
    for m, n, i in zip(a.shape, b.shape, range(nd)):
        if m != n and m != 1 and n != 1:
            raise ValueError('shapes {} and {} do not broadcast'.format(a.shape, b.shape))
    return a, b


class _Jacobian:
    'the factor |det J| of a geometry (function.J, function.py:2291-2316)'

    def __init__(self, geom):
        self.geom = geom

    def __mul__(self, other):
        if isinstance(other, PointFunction):
            return Array((), (), None, numpy.ones(()), self.geom, other)
        return Array.cast(other) * self

    __rmul__ = __mul__


def J(geom, ndims=None):
    if not isinstance(geom, Geometry):
        raise NotImplementedError('jacobian of a non-nodal geometry')
    return _Jacobian(geom)


def outer(arg1, arg2=None, axis=0):
    'outer product over `axis` (function.outer): arg1[:,None] * arg2[None,:] for axis 0'
    if arg2 is None:
        arg2 = arg1
    arg1, arg2 = Array.cast(arg1), Array.cast(arg2)
    i1 = (slice(None),) * (axis + 1) + (None,)
    i2 = (slice(None),) * axis + (None,)
    return arg1[i1] * arg2[i2]


def grad(arg, geom, ndims=0):
    return Array.cast(arg).grad(geom, ndims)


def eye(n):
    return numpy.eye(n)


def trace(arg, axis1=-2, axis2=-1):
    return Array.cast(arg).trace(axis1, axis2)


# ---- evaluation of integrals (function.eval / as_csr / as_coo, function.py:2408-2452) -------------------

class Integral:
    '''Postponed integral of an Array over a Sample (sample._Integral, sample.py:944-956).'''

    def __init__(self, sample, func):
        self.sample = sample
        self.func = func
        self.kind, self.tensor = func.coefficient_tensor()
        if func.jac is None:
            raise NotImplementedError('integrands without a geometry jacobian (function.J) are outside the accelerated path')
        self.shape = func.shape

    def eval(self, **arguments):
        return eval(self, arguments)

    def __add__(self, other):
        if not isinstance(other, Integral) or other.sample is not self.sample:
            raise NotImplementedError('sum of integrals over different samples')
        return Integral(self.sample, self.func + other.func)

    def __sub__(self, other):
        if not isinstance(other, Integral) or other.sample is not self.sample:
            raise NotImplementedError('difference of integrals over different samples')
        return Integral(self.sample, self.func - other.func)

    def __mul__(self, scalar):
        return Integral(self.sample, self.func * float(scalar))

    __rmul__ = __mul__

    def __neg__(self):
        return Integral(self.sample, -self.func)


class _Sparse:
    def __init__(self, integral, fmt):
        if not isinstance(integral, Integral) or integral.kind != 'matrix':
            raise NotImplementedError('as_{} needs a 2-D integral'.format(fmt))
        self.integral = integral
        self.fmt = fmt


def as_csr(integral):
    'evaluates to (values, rowptr, colidx): float64, int64, int64; sorted, unique, structural zeros kept'
    return _Sparse(integral, 'csr')


def as_coo(integral):
    'evaluates to (values, rowidx, colidx), lexicographically ordered and unique'
    return _Sparse(integral, 'coo')


def eval(funcs, arguments=None, **kwargs):
    '''Evaluate one integral or a (nested) tuple of integrals / as_csr / as_coo objects.

    All integrals over the same sample and basis are assembled by ONE kernel launch
    (the reference merges equal-length loops into one, evaluable.py:6841-6895).'''
    if arguments:
        raise NotImplementedError('integrands with arguments are outside the accelerated path')
    single = not isinstance(funcs, (tuple, list))
    flat = []

    def flatten(f):
        if isinstance(f, (tuple, list)):
            return tuple(flatten(g) for g in f)
        flat.append(f)
        return len(flat) - 1
    tree = flatten((funcs,) if single else tuple(funcs))
    results = [None] * len(flat)
    groups = {}
    for i, f in enumerate(flat):
        integral = f.integral if isinstance(f, _Sparse) else f
        if not isinstance(integral, Integral):
            raise NotImplementedError('cannot evaluate {!r}'.format(f))
        key = id(integral.sample), id(integral.func.space)
        groups.setdefault(key, []).append(i)
    for idxs in groups.values():
        integrals = [flat[i].integral if isinstance(flat[i], _Sparse) else flat[i] for i in idxs]
        sample = integrals[0].sample
        outs = sample._evaluate(integrals)
        for i, integral, out in zip(idxs, integrals, outs):
            f = flat[i]
            if integral.kind == 'vector':
                results[i] = out
            elif isinstance(f, _Sparse):
                values, rowptr, colidx = out
                if f.fmt == 'csr':
                    results[i] = values, rowptr, colidx
                else:
                    rowidx = numpy.repeat(numpy.arange(len(rowptr) - 1, dtype=numpy.int64), numpy.diff(rowptr))
                    results[i] = values, rowidx, colidx
            else:
                # like the reference, a plain 2-D integral evaluates to a dense array (sample.py:160-175)
                values, rowptr, colidx = out
                n = len(rowptr) - 1
                if n * n > 2**28:
                    raise MemoryError('dense {}x{} result requested; use function.as_csr'.format(n, n))
                dense = numpy.zeros((n, n))
                rowidx = numpy.repeat(numpy.arange(n), numpy.diff(rowptr))
                dense[rowidx, colidx] = values
                results[i] = dense

    def unflatten(t):
        return tuple(unflatten(u) for u in t) if isinstance(t, tuple) else results[t]
    out = unflatten(tree)
    return out[0] if single else out
