'''Structured topology: the host-side producer of the tables the CUDA kernels consume.

Mirrors the part of ``nutils.topology.StructuredTopology`` (src/nutils/topology.py:1982-2420)
that the integration path touches: ``basis``, ``field``-style vector bases, ``sample``,
``integrate`` and ``integral`` (topology.py:365-440, 1566-1575), with the same argument
meaning.  Mesh algebra (boundaries, refinement, trimming) is out of scope.
'''

import numpy
from . import bspline, points, function, sample as _sample


class Basis:
    '''Tensor-product spline basis (function.StructuredBasis, function.py:3040-3100).

    Behaves like an array of shape (ndofs,) -- or (ndofs, ncomp) for vector-valued
    bases with the dof numbering ibasis*ncomp + comp of ``function.field``
    (function.py:2623-2626).'''

    def __init__(self, topo, bases1d, ncomp=1):
        self.topo = topo
        self.bases1d = tuple(bases1d)
        self.ncomp = int(ncomp)
        self.nbasis = int(numpy.prod([b.ndofs for b in bases1d]))
        self.ndofs = self.nbasis * self.ncomp
        self.degree = tuple(b.degree for b in bases1d)
        na = topo.ndims + 1
        if self.ncomp == 1:
            C = numpy.zeros((1, 1, na))
            C[0, 0, 0] = 1.
            self._array = function.Array((self.ndofs,), (0,), self, C)
        else:
            C = numpy.zeros((1, self.ncomp, self.ncomp, na))
            C[0, :, :, 0] = numpy.eye(self.ncomp)
            self._array = function.Array((self.ndofs, self.ncomp), (0,), self, C)
        self.shape = self._array.shape

    def __len__(self):
        return self.ndofs

    def __getitem__(self, item):
        return self._array[item]

    def __getattr__(self, name):
        # array behaviour: grad, sum, dot, ...
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self._array, name)

    def __mul__(self, other):
        return self._array * other

    __rmul__ = __mul__

    def __add__(self, other):
        return self._array + other

    def __neg__(self):
        return -self._array

    def get_dofs(self, ielem):
        'dofs of the basis functions with support on element `ielem` (function.py:2795-2814)'
        idx = numpy.unravel_index(ielem, self.topo.shape)
        dofs = numpy.zeros(1, dtype=numpy.int64)
        for b, i in zip(self.bases1d, idx):
            dofs = (dofs[:, None] * b.ndofs + (b.start[i] + numpy.arange(b.degree + 1))[None, :]).ravel()
        if self.ncomp > 1:
            dofs = (dofs[:, None] * self.ncomp + numpy.arange(self.ncomp)[None, :]).ravel()
        return dofs

    def vector(self, ncomp):
        '''Vector-valued basis of shape (ndofs*ncomp, ncomp) with the interleaved numbering of ``function.field``.

        Note: ``nutils`` ``Array.vector`` numbers component-major; solver.System / field use the
        interleaved numbering reproduced here (function.py:2623-2626).'''
        if self.ncomp != 1:
            raise ValueError('basis is already vector-valued')
        return self.topo._vector_basis(self, ncomp)


BOUNDARY_NAMES = {1: ('left', 'right'), 2: ('left', 'right', 'bottom', 'top'), 3: ('left', 'right', 'bottom', 'top', 'front', 'back')}


class _Boundary:
    "``topo.boundary['left']``, ``topo.boundary['left,top']`` (topology.py:2049-2057, names of StructuredTopology.boundary)"

    def __init__(self, topo):
        self.topo = topo

    def __getitem__(self, names):
        faces = []
        for name in names.split(','):
            try:
                k = BOUNDARY_NAMES[self.topo.ndims].index(name.strip())
            except ValueError:
                raise KeyError(name)
            faces.append((k // 2, k % 2))
        return BoundaryTopology(self.topo, tuple(faces))


class BoundaryTopology:
    '''One or more sides of a structured topology: integrals over it are boundary integrals (Neumann terms, the boundary
    projections of solve_constraints).  The points are the face Gauss points of the adjacent volume elements; the basis
    and its dof numbering are the volume ones.'''

    def __init__(self, parent, faces):
        self.parent = parent
        self.faces = faces
        self.ndims = parent.ndims - 1
        self._samples = {}

    def sample(self, ischeme, degree):
        if ischeme != 'gauss':
            raise NotImplementedError('only gauss samples are on the accelerated path')
        key = tuple(numpy.ravel(degree).tolist())
        s = self._samples.get(key)
        if s is None:
            s = self._samples[key] = _sample.Sample(self.parent, points.tensor_gauss(self.parent.ndims, degree), faces=self.faces)
        return s

    def integrate(self, funcs, ischeme='gauss', degree=None, *, arguments=None):
        ischeme, degree = StructuredTopology._parse(ischeme, degree)
        return self.sample(ischeme, degree).integrate(funcs, arguments=arguments or {})

    def integral(self, func, ischeme='gauss', degree=None):
        ischeme, degree = StructuredTopology._parse(ischeme, degree)
        return self.sample(ischeme, degree).integral(func)


class StructuredTopology:
    'structured grid of shape `shape` (number of elements per dimension)'

    def __init__(self, shape, space='X'):
        self.shape = tuple(int(n) for n in shape)
        self.ndims = len(self.shape)
        self.space = space
        self._samples = {}
        self._bases = {}

    def __len__(self):
        return int(numpy.prod(self.shape))

    @property
    def boundary(self):
        return _Boundary(self)

    # -- bases ---------------------------------------------------------------------------------------

    def basis(self, name, degree, shape=(), **kwargs):
        '''``basis('spline'|'std', degree=p)`` (topology.py:344-362, 2209-2365).

        `shape=(n,)` gives a vector-valued basis with field numbering.'''
        if name == 'std':
            kwargs.setdefault('continuity', 0)
        elif name != 'spline':
            raise NotImplementedError('basis type {!r} is outside the accelerated path'.format(name))
        degrees = [degree] * self.ndims if numpy.ndim(degree) == 0 else list(degree)
        cont = kwargs.pop('continuity', -1)
        conts = [cont] * self.ndims if numpy.ndim(cont) == 0 else list(cont)
        knotvalues = kwargs.pop('knotvalues', None)
        knotmult = kwargs.pop('knotmultiplicities', None)
        if kwargs.pop('periodic', None):
            raise NotImplementedError('periodic bases are outside the accelerated path')
        if kwargs.pop('removedofs', None):
            raise NotImplementedError('removedofs is outside the accelerated path')
        if kwargs:
            raise TypeError('unexpected arguments: {}'.format(', '.join(kwargs)))
        if knotvalues is None or numpy.ndim(knotvalues[0]) == 0:
            knotvalues = [knotvalues] * self.ndims
        if knotmult is None or numpy.ndim(knotmult[0]) == 0:
            knotmult = [knotmult] * self.ndims
        key = tuple(degrees), tuple(conts), repr(knotvalues), repr(knotmult)
        b = self._bases.get(key)
        if b is None:
            b1 = [bspline.spline_basis_1d(n, p, continuity=c, knotvalues=k, knotmultiplicities=m)
                  for n, p, c, k, m in zip(self.shape, degrees, conts, knotvalues, knotmult)]
            b = self._bases[key] = Basis(self, b1)
        shape = tuple(shape)
        if not shape:
            return b
        if len(shape) != 1:
            raise NotImplementedError('tensor-valued bases')
        return self._vector_basis(b, shape[0])

    def _vector_basis(self, scalar, ncomp):
        key = id(scalar), int(ncomp)
        b = self._bases.get(key)
        if b is None:
            b = self._bases[key] = Basis(self, scalar.bases1d, ncomp=ncomp)
        return b

    def nodal_geometry(self, nodes):
        'geometry from explicit nodal coordinates float64[ndims, n0+1, ...]'
        nodes = numpy.asarray(nodes, dtype=float)
        if nodes.shape != (self.ndims,) + tuple(n + 1 for n in self.shape):
            raise ValueError('nodes must have shape {}'.format((self.ndims,) + tuple(n + 1 for n in self.shape)))
        return function.Geometry(self, nodes)

    # -- integration ---------------------------------------------------------------------------------

    def sample(self, ischeme, degree):
        "``sample('gauss', degree)`` (topology.py:1566-1575)"
        if ischeme != 'gauss':
            raise NotImplementedError('only gauss samples are on the accelerated path')
        key = ischeme, tuple(numpy.ravel(degree).tolist())
        s = self._samples.get(key)
        if s is None:
            s = self._samples[key] = _sample.Sample(self, points.tensor_gauss(self.ndims, degree))
        return s

    @staticmethod
    def _parse(ischeme, degree):
        if degree is None:  # legacy 'gauss4' spelling (element.parse_legacy_ischeme)
            name = ischeme.rstrip('0123456789')
            return name, int(ischeme[len(name):])
        return ischeme, degree

    def integrate(self, funcs, ischeme='gauss', degree=None, *, arguments=None):
        'integrate functions (topology.py:427-431)'
        ischeme, degree = self._parse(ischeme, degree)
        return self.sample(ischeme, degree).integrate(funcs, arguments=arguments or {})

    def integral(self, func, ischeme='gauss', degree=None):
        'postponed integral (topology.py:433-440)'
        ischeme, degree = self._parse(ischeme, degree)
        return self.sample(ischeme, degree).integral(func)
