'''Matrix objects on the output side of the path, with the interface of ``nutils.matrix.Matrix``.

``nutils.matrix.backend(obj)`` accepts any object with ``.assemble(values, rowptr, colidx, ncols)`` returning a matrix
(src/nutils/matrix/__init__.py:20-27, 70); this module is such an object, so

    with nutils.matrix.backend(nutils_b200.matrix): ...          # or NUTILS_MATRIX-style selection by the caller

makes ``solver.System`` build :class:`Matrix` objects.  The interface restated here is what the reference's solvers use
(matrix/_base.py:33-302; call sites solver.py:332, 386, 600-604, 634, 661, 935): ``shape``, ``+ - * / neg``, ``@``, ``T``,
``submatrix(rows, cols)`` (cached like _base.py:261-285), ``export('csr' | 'coo' | 'dense')``, ``diagonal``, ``rowsupp``,
``solve(rhs, lhs0=, constrain=, rconstrain=, solver=, atol=, rtol=)`` with the constraint semantics of _base.py:100-173,
``solve_leniently``; errors are :class:`MatrixError` / :class:`ToleranceNotReached` (``.best``).

Two classes:

* :class:`Matrix` -- CSR arrays on the host (what ``assemble_csr`` receives) with a lazily created DEVICE mirror: an explicit
  pattern (``b2_pattern_create_csr``) plus the values in HBM.  ``@`` and ``solve(solver='cg')`` run there (``b2_spmv_device``,
  ``b2_cg_device``); index manipulation (sums of matrices with different patterns, transposes, submatrices) is host-side
  integer work on the CSR arrays; ``solve(solver='direct')`` -- the solver is outside the scope of this repository, SURVEY.md
  section 8 -- factorises with scipy like the reference's scipy backend (matrix/_scipy.py:86-93).
* :class:`DeviceMatrix` -- values that never left HBM, on the analytic or element-set pattern of an ``engine`` plan
  (``Sample.integrate_device``): products, diagonal and constrained CG solves on the device; everything else
  goes through :meth:`DeviceMatrix.tohost`.
'''

import numpy

from . import _lib


class MatrixError(Exception):
    'invalid matrix data or operation (nutils.matrix.MatrixError)'


class ToleranceNotReached(MatrixError, _lib.ToleranceNotReached):
    'the linear solver stopped above the requested tolerance; ``.best`` is the non-conforming solution (matrix/_base.py:21-30)'

    def __init__(self, best):
        _lib.ToleranceNotReached.__init__(self, 'solver failed to reach tolerance', best)


def _asboolean(index, size):
    'bool mask from a bool array or an int index array (numeric.asboolean)'
    index = numpy.asarray(index)
    if index.dtype == bool:
        if index.shape != (size,):
            raise MatrixError('mask has the wrong shape')
        return index
    mask = numpy.zeros(size, dtype=bool)
    mask[index] = True
    return mask


class Matrix:
    'CSR matrix: host arrays (values float64[nnz], rowptr int64[nrows+1], colidx int64[nnz]) + a device mirror on demand'

    def __init__(self, values, rowptr, colidx, ncols):
        self.values = numpy.ascontiguousarray(values, dtype=numpy.float64)
        self.rowptr = numpy.ascontiguousarray(rowptr, dtype=numpy.int64)
        self.colidx = numpy.ascontiguousarray(colidx, dtype=numpy.int64)
        self.shape = len(self.rowptr) - 1, int(ncols)
        self.dtype = float
        self._dev = None
        self._cached_submatrix = None

    def __repr__(self):
        return '{}<{}x{}>'.format(type(self).__qualname__, *self.shape)

    @property
    def size(self):
        return self.shape[0] * self.shape[1]

    # -- conversions ---------------------------------------------------------------------------------------------------------

    def _sp(self):
        import scipy.sparse
        return scipy.sparse.csr_matrix((self.values, self.colidx, self.rowptr), shape=self.shape)

    @classmethod
    def _from_sp(cls, A):
        A = A.tocsr()
        A.sum_duplicates()
        A.sort_indices()
        return cls(A.data, A.indptr.astype(numpy.int64), A.indices.astype(numpy.int64), A.shape[1])

    def _convert(self, other):
        if isinstance(other, DeviceMatrix):
            other = other.tohost()
        if not isinstance(other, Matrix):
            if hasattr(other, 'export') and hasattr(other, 'shape'):   # a matrix of another backend
                data, indices, indptr = other.export('csr')
                other = Matrix(data, indptr, indices, other.shape[1])
            else:
                raise TypeError('cannot convert {} to Matrix'.format(type(other).__name__))
        if other.shape != self.shape:
            raise MatrixError('non-matching shapes')
        return other

    def export(self, form):
        "'csr' -> (data, indices, indptr); 'coo' -> (data, (rows, cols)); 'dense' (matrix/_base.py:291-302)"
        if form == 'csr':
            return self.values, self.colidx, self.rowptr
        rows = numpy.repeat(numpy.arange(self.shape[0], dtype=numpy.int64), numpy.diff(self.rowptr))
        if form == 'coo':
            return self.values, (rows, self.colidx)
        if form == 'dense':
            dense = numpy.zeros(self.shape)
            dense[rows, self.colidx] = self.values
            return dense
        raise NotImplementedError('cannot export {} to {!r}'.format(type(self).__name__, form))

    def export_reference(self):
        'argument tuple of nutils.matrix.assemble_csr'
        return self.values, self.rowptr, self.colidx, self.shape[1]

    def todevice(self, device=0):
        'the device mirror: explicit pattern + values in HBM (created once)'
        if self._dev is None:
            from . import engine
            ctx = engine.Context.get(device)
            pattern = engine.CsrPattern(ctx, self.rowptr, self.colidx, self.shape[1])
            buf = ctx.device_alloc(8 * max(len(self.values), 1))
            if len(self.values):
                buf.from_host(self.values)
            self._dev = DeviceMatrix(pattern, buf)
        return self._dev

    # -- arithmetic (matrix/_base.py:41-86) ------------------------------------------------------------------------------------

    def __add__(self, other):
        other = self._convert(other)
        if other.rowptr is self.rowptr or (numpy.array_equal(other.rowptr, self.rowptr) and numpy.array_equal(other.colidx, self.colidx)):
            return Matrix(self.values + other.values, self.rowptr, self.colidx, self.shape[1])
        return Matrix._from_sp(self._sp() + other._sp())

    def __sub__(self, other):
        return self.__add__(-self._convert(other))

    def __mul__(self, other):
        if not isinstance(other, (int, float, numpy.integer, numpy.floating)):
            raise TypeError('a matrix can be multiplied with a scalar only')
        return Matrix(self.values * other, self.rowptr, self.colidx, self.shape[1])

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self.__mul__(1 / other)

    def __neg__(self):
        return Matrix(-self.values, self.rowptr, self.colidx, self.shape[1])

    def __matmul__(self, other):
        other = numpy.asarray(other)
        if other.ndim == 0 or other.shape[0] != self.shape[1]:
            raise MatrixError('shape mismatch')
        dev = self._device_or_none()
        if dev is not None and other.dtype != complex:
            cols = other.reshape(other.shape[0], -1)
            out = numpy.stack([dev @ numpy.ascontiguousarray(cols[:, k], dtype=float) for k in range(cols.shape[1])], axis=1) if cols.shape[1] else numpy.zeros((self.shape[0], 0))
            return out.reshape((self.shape[0],) + other.shape[1:])
        # no device in this process: the container still multiplies (host, like the reference's numpy backend)
        prod = self.values.reshape((-1,) + (1,) * (other.ndim - 1)) * other[self.colidx]
        nonempty = numpy.diff(self.rowptr) > 0
        out = numpy.zeros((self.shape[0],) + other.shape[1:], dtype=prod.dtype)
        if len(self.values):
            out[nonempty] = numpy.add.reduceat(prod, self.rowptr[:-1][nonempty], axis=0)
        return out

    def _device_or_none(self):
        if self._dev is not None:
            return self._dev
        try:
            return self.todevice()
        except _lib.BackendNotAvailable:
            return None

    @property
    def T(self):
        return Matrix._from_sp(self._sp().T)

    def diagonal(self):
        if self.shape[0] != self.shape[1]:
            raise MatrixError('failed to extract diagonal: matrix is not square')
        return self._sp().diagonal()

    def rowsupp(self, tol=0):
        'rows with entries larger than tol in absolute value (matrix/_base.py:91-98)'
        data, (row, col) = self.export('coo')
        supp = numpy.zeros(self.shape[0], dtype=bool)
        supp[row[abs(data) > tol]] = True
        return supp

    def submatrix(self, rows, cols):
        'matrix of the selected rows and columns (bool masks or index arrays), cached for repeated selections (matrix/_base.py:261-285)'
        rows = _asboolean(rows, self.shape[0])
        cols = _asboolean(cols, self.shape[1])
        if rows.all() and cols.all():
            return self
        if self._cached_submatrix is None or (rows != self._cached_rows).any() or (cols != self._cached_cols).any():
            self._cached_rows, self._cached_cols = rows, cols
            self._cached_submatrix = Matrix._from_sp(self._sp()[rows, :][:, cols])
        return self._cached_submatrix

    # -- solving (matrix/_base.py:100-213) -----------------------------------------------------------------------------------

    def solve(self, rhs=None, *, lhs0=None, constrain=None, rconstrain=None, solver='direct', atol=0., rtol=0., **solverargs):
        '''Solve A x = rhs with initial value and constraints as ``nutils.matrix.Matrix.solve``: constrain is a float array (NaN =
        free, number = constrained to it) or a bool array (True = constrained to lhs0); rconstrain optionally selects the rows.
        solver: 'direct' (also for the reference's default 'arnoldi': scipy sparse LU on the host), 'cg' (Jacobi-preconditioned CG
        on the device, b2_cg_device), or a scipy.sparse.linalg iterative method name.'''
        nrows, ncols = self.shape
        if lhs0 is None and constrain is None and rconstrain is None:
            if rhs is None:
                raise MatrixError('nothing to solve')
            return self._solver(numpy.asarray(rhs, dtype=float), solver, atol=atol, rtol=rtol, **solverargs)
        rhs = numpy.zeros(nrows) if rhs is None else numpy.asarray(rhs, dtype=float)
        if lhs0 is None:
            lhs = numpy.zeros((ncols,) + rhs.shape[1:])
        else:
            lhs = numpy.array(lhs0, dtype=float)
            while lhs.ndim < rhs.ndim:
                lhs = lhs[..., numpy.newaxis].repeat(rhs.shape[lhs.ndim], axis=lhs.ndim)
            if lhs.shape != (ncols,) + rhs.shape[1:]:
                raise MatrixError('lhs0 has the wrong shape')
        if constrain is None:
            J = numpy.ones(ncols, dtype=bool)
        else:
            constrain = numpy.asarray(constrain)
            if constrain.shape != (ncols,):
                raise MatrixError('constrain has the wrong shape')
            if constrain.dtype == bool:
                J = ~constrain
            else:
                J = numpy.isnan(constrain)
                lhs[~J] = constrain[~J]
        if rconstrain is None:
            if nrows != ncols:
                raise MatrixError('constrained matrix is not square: {}x{}'.format(nrows, ncols))
            I = J
        else:
            rconstrain = numpy.asarray(rconstrain)
            if rconstrain.shape != (nrows,) or rconstrain.dtype != bool:
                raise MatrixError('rconstrain must be a bool array with one entry per row')
            I = ~rconstrain
        try:
            lhs[J] += self.submatrix(I, J)._solver((rhs - self @ lhs)[I], solver, atol=atol, rtol=rtol, **solverargs)
        except ToleranceNotReached as e:
            lhs[J] += e.best
            raise ToleranceNotReached(lhs) from None
        return lhs

    def solve_leniently(self, *args, **kwargs):
        'like solve, but a missed tolerance only warns and returns the best solution (matrix/_base.py:177-188)'
        import warnings
        try:
            return self.solve(*args, **kwargs)
        except ToleranceNotReached as e:
            warnings.warn(str(e))
            return e.best

    def _solver(self, rhs, solver, *, atol, rtol, **solverargs):
        if self.shape[0] != self.shape[1]:
            raise MatrixError('constrained matrix is not square: {}x{}'.format(*self.shape))
        if rhs.shape[0] != self.shape[0]:
            raise MatrixError('right-hand size shape does not match matrix shape')
        rhsnorm = numpy.linalg.norm(rhs, axis=0).max() if rhs.size else 0.
        atol = max(atol, rtol * rhsnorm)
        if rhsnorm <= atol:
            return numpy.zeros_like(rhs)
        solverargs.pop('symmetric', None)
        precon = solverargs.pop('precon', None)
        solverargs.pop('preconargs', None)
        solverargs.pop('truncate', None)
        if solver in ('direct', 'arnoldi'):
            import scipy.sparse.linalg
            try:
                lhs = scipy.sparse.linalg.splu(self._sp().tocsc()).solve(rhs)
            except Exception as e:
                raise MatrixError('solver failed with error: {}'.format(e)) from e
        elif solver == 'cg':
            if precon not in (None, 'diag'):
                raise MatrixError("invalid preconditioner {!r} for the device CG (available: 'diag')".format(precon))
            cols = rhs.reshape(rhs.shape[0], -1)
            dev = self.todevice()
            lhs = numpy.stack([dev.solve(cols[:, k], atol=atol, maxiter=solverargs.get('maxiter', 0), _raise=False) for k in range(cols.shape[1])], axis=1).reshape(rhs.shape)
        else:
            import scipy.sparse.linalg
            fun = getattr(scipy.sparse.linalg, solver, None)
            if fun is None:
                raise MatrixError('invalid solver {!r} for {}'.format(solver, type(self).__name__))
            lhs, status = fun(self._sp(), rhs, atol=atol, rtol=0., **solverargs)
            if status != 0:
                raise MatrixError('solver failed with status {}'.format(status))
        if not numpy.isfinite(lhs).all():
            raise MatrixError('solver returned non-finite left hand side')
        resnorm = numpy.linalg.norm(rhs - self @ lhs, axis=0).max()
        if resnorm > atol > 0:
            raise ToleranceNotReached(lhs)
        return lhs


class DeviceMatrix:
    '''Matrix whose values stay in HBM: the device-resident counterpart of ``nutils.matrix.Matrix`` for what
    ``solver.System`` does with an assembled matrix -- products and constrained solves (solver.py:318-425) -- so that a
    multi-GB result never crosses PCIe.  ``plan`` owns the pattern (engine.Plan: analytic; engine.ElemSetPlan: element set;
    engine.CsrPattern: explicit), ``values`` is an engine.DeviceBuffer (or torch tensor) of nnz float64.  Operations that
    change the pattern (``T`` of a non-symmetric matrix, ``submatrix``, sums with other patterns) return host :class:`Matrix`
    objects.'''

    def __init__(self, plan, values):
        self.plan = plan
        self.values = values
        self.shape = plan.ndofs, getattr(plan, 'ncols', plan.ndofs)
        self.dtype = float
        self._host = None

    def __repr__(self):
        return '{}<{}x{}>'.format(type(self).__qualname__, *self.shape)

    @property
    def size(self):
        return self.shape[0] * self.shape[1]

    def _vec(self, n, host=None):
        buf = self.plan.ctx.device_alloc(8 * max(n, 1))
        if host is not None and n:
            buf.from_host(numpy.ascontiguousarray(host, dtype=numpy.float64))
        return buf

    def __matmul__(self, other):
        other = numpy.asarray(other, dtype=float)
        if other.ndim != 1:
            return numpy.stack([self @ other[:, k] for k in range(other.shape[1])], axis=1) if other.ndim == 2 else self.tohost() @ other
        if other.shape != (self.shape[1],):
            raise MatrixError('shape mismatch')
        if self.shape[0] == 0:
            return numpy.zeros(0)
        x, y = self._vec(self.shape[1], other), self._vec(self.shape[0])
        self.plan.spmv_device(self.values, x, y)
        return y.to_host()[:self.shape[0]]

    def diagonal(self):
        if self.shape[0] != self.shape[1]:
            raise MatrixError('failed to extract diagonal: matrix is not square')
        d = self._vec(self.shape[0])
        self.plan.diagonal_device(self.values, d)
        return d.to_host()[:self.shape[0]]

    def solve(self, rhs=None, *, lhs0=None, constrain=None, solver='cg', atol=0., rtol=0., precon='diag', maxiter=0, _raise=True, **ignored):
        '''``Matrix.solve`` (matrix/_base.py:100-173) with the one solver this class carries: Jacobi-preconditioned CG on the
        device.  constrain: float array (NaN = free, number = constrained to that value) or bool array (True = constrained to
        lhs0).  Raises ToleranceNotReached (with ``.best``) if an explicit atol/rtol is not reached, or if atol = rtol = 0
        ("machine precision") and the iteration limit is hit far above it.  Other solvers: ``self.tohost().solve(...)``.'''
        if solver != 'cg':
            return self.tohost().solve(rhs, lhs0=lhs0, constrain=constrain, solver=solver, atol=atol, rtol=rtol, **ignored)
        if precon != 'diag':
            raise MatrixError('invalid preconditioner {!r} for DeviceMatrix'.format(precon))
        n = self.shape[0]
        if n != self.shape[1]:
            raise MatrixError('matrix is not square')
        lhs = numpy.zeros(n) if lhs0 is None else numpy.array(lhs0, dtype=float)
        mask = None
        if constrain is not None:
            constrain = numpy.asarray(constrain)
            if constrain.shape != (n,):
                raise MatrixError('constrain has the wrong shape')
            if constrain.dtype == bool:
                fixed = constrain
            else:
                fixed = ~numpy.isnan(constrain)
                lhs[fixed] = constrain[fixed]
            mask = self.plan.ctx.device_alloc(max(n, 1))
            mask.from_host(fixed.astype(numpy.uint8))
        x = self._vec(n, lhs)
        b = None if rhs is None else self._vec(n, rhs)
        try:
            self.plan.cg_device(self.values, b, x, constrained=mask, atol=atol, rtol=rtol, maxiter=maxiter)
        except _lib.ToleranceNotReached:
            best = x.to_host()[:n]
            if _raise:
                raise ToleranceNotReached(best) from None
            return best
        return x.to_host()[:n]

    def solve_leniently(self, *args, **kwargs):
        import warnings
        try:
            return self.solve(*args, **kwargs)
        except ToleranceNotReached as e:
            warnings.warn(str(e))
            return e.best

    def tohost(self):
        'the host Matrix with the same entries (one device-to-host copy of the values, cached)'
        if self._host is None:
            values = self.values.to_host() if hasattr(self.values, 'to_host') else self.values.cpu().numpy()
            rowptr, colidx = self.plan.csr_pattern()
            self._host = Matrix(values[:len(colidx)], rowptr, colidx, self.shape[1])
        return self._host

    def export(self, form):
        return self.tohost().export(form)

    def rowsupp(self, tol=0):
        return self.tohost().rowsupp(tol)

    def submatrix(self, rows, cols):
        rows = _asboolean(rows, self.shape[0])
        cols = _asboolean(cols, self.shape[1])
        if rows.all() and cols.all():
            return self
        return self.tohost().submatrix(rows, cols)

    @property
    def T(self):
        return self.tohost().T

    def __add__(self, other):
        return self.tohost() + other

    def __sub__(self, other):
        return self.tohost() - other

    def __mul__(self, other):
        return self.tohost() * other

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self.tohost() / other

    def __neg__(self):
        return -self.tohost()


def assemble_csr(values, rowptr, colidx, ncols):
    'validate and wrap, with the checks of matrix/__init__.py:50-70'
    values = numpy.asarray(values)
    rowptr = numpy.asarray(rowptr)
    colidx = numpy.asarray(colidx)
    if values.ndim != 1 or rowptr.ndim != 1 or colidx.ndim != 1:
        raise MatrixError('values, rowptr and colidx must be one-dimensional')
    if len(rowptr) == 0 or rowptr[0] != 0 or rowptr[-1] != len(values) or len(colidx) != len(values):
        raise MatrixError('rowptr does not match the number of values')
    if (numpy.diff(rowptr) < 0).any():
        raise MatrixError('rowptr is not monotonically increasing')
    if len(colidx) and (colidx.min() < 0 or colidx.max() >= ncols):
        raise MatrixError('column index out of bounds')
    if len(colidx) > 1:
        inc = numpy.diff(colidx) > 0
        inc[rowptr[1:-1][(rowptr[1:-1] > 0) & (rowptr[1:-1] < len(colidx))] - 1] = True
        if not inc.all():
            raise MatrixError('column indices are not strictly increasing within rows')
    return Matrix(values, rowptr, colidx, ncols)


# the backend protocol of the reference: ``nutils.matrix.backend(obj)`` accepts any object with
# ``.assemble(values, rowptr, colidx, ncols)`` (matrix/__init__.py:20-27) -- this module is such an object
assemble = assemble_csr
