'''CSR matrix container on the output side of the path.

Mirrors the boundary ``nutils.matrix.assemble_csr(values, rowptr, colidx, ncols)``
(src/nutils/matrix/__init__.py:30-70): validates the CSR triplet the same way and wraps
it in an object with the ``export`` forms of ``nutils.matrix.Matrix``
(matrix/_base.py:291-302).  Solving is out of scope; ``export`` hands
the data to scipy or back to nutils (``nutils.matrix.assemble_csr(*m.export_reference())``).
'''

import numpy


class MatrixError(Exception):
    'invalid matrix data (nutils.matrix.MatrixError)'


class Matrix:
    def __init__(self, values, rowptr, colidx, ncols):
        self.values = values
        self.rowptr = rowptr
        self.colidx = colidx
        self.shape = len(rowptr) - 1, int(ncols)

    def export(self, form):
        "export('csr') -> (data, indices, indptr); 'coo' -> (data, (rows, cols)); 'dense' (matrix/_base.py:291-302)"
        if form == 'csr':
            return self.values, self.colidx, self.rowptr
        if form == 'coo':
            rows = numpy.repeat(numpy.arange(self.shape[0], dtype=numpy.int64), numpy.diff(self.rowptr))
            return self.values, (rows, self.colidx)
        if form == 'dense':
            dense = numpy.zeros(self.shape)
            rows = numpy.repeat(numpy.arange(self.shape[0]), numpy.diff(self.rowptr))
            dense[rows, self.colidx] = self.values
            return dense
        raise NotImplementedError('cannot export matrix to {!r}'.format(form))

    def export_reference(self):
        'argument tuple of nutils.matrix.assemble_csr'
        return self.values, self.rowptr, self.colidx, self.shape[1]

    def __matmul__(self, other):
        other = numpy.asarray(other)
        if other.shape[0] != self.shape[1]:
            raise MatrixError('shape mismatch')
        prod = self.values.reshape((-1,) + (1,) * (other.ndim - 1)) * other[self.colidx]
        nonempty = numpy.diff(self.rowptr) > 0
        out = numpy.zeros((self.shape[0],) + other.shape[1:])
        if len(self.values):
            out[nonempty] = numpy.add.reduceat(prod, self.rowptr[:-1][nonempty], axis=0)
        return out

    @property
    def T(self):
        import scipy.sparse
        A = scipy.sparse.csr_matrix((self.values, self.colidx, self.rowptr), shape=self.shape).T.tocsr()
        A.sort_indices()
        return Matrix(A.data, A.indptr.astype(numpy.int64), A.indices.astype(numpy.int64), self.shape[0])


class DeviceMatrix:
    '''Matrix whose values stay in HBM: the device-resident counterpart of ``nutils.matrix.Matrix``
    (matrix/_base.py:33-302) for what ``solver.System`` does with an assembled matrix -- products and constrained
    solves (solver.py:318-425) -- so that a multi-GB result never crosses PCIe.  ``plan`` is the engine.Plan /
    engine.ElemSetPlan that owns the pattern, ``values`` an engine.DeviceBuffer (or torch tensor) of nnz float64.'''

    def __init__(self, plan, values):
        self.plan = plan
        self.values = values
        self.shape = plan.ndofs, plan.ndofs

    def _vec(self, host=None):
        buf = self.plan.ctx.device_alloc(8 * self.shape[0])
        if host is not None:
            buf.from_host(numpy.ascontiguousarray(host, dtype=numpy.float64))
        return buf

    def __matmul__(self, other):
        other = numpy.asarray(other, dtype=float)
        if other.shape != (self.shape[1],):
            raise MatrixError('shape mismatch')
        x, y = self._vec(other), self._vec()
        self.plan.spmv_device(self.values, x, y)
        return y.to_host()

    def diagonal(self):
        d = self._vec()
        self.plan.diagonal_device(self.values, d)
        return d.to_host()

    def solve(self, rhs=None, *, lhs0=None, constrain=None, solver='cg', atol=0., rtol=0., precon='diag', maxiter=0):
        '''``Matrix.solve`` (matrix/_base.py:100-173) with the one solver this backend carries: Jacobi-preconditioned CG on the
        device.  constrain: float array (NaN = free, number = constrained to that value) or bool array (True = constrained to
        lhs0).  Raises ToleranceNotReached (with ``.best``) if an explicit atol/rtol is not reached.'''
        from ._lib import ToleranceNotReached
        if solver != 'cg' or precon != 'diag':
            raise MatrixError('invalid solver {!r}/{!r} for DeviceMatrix'.format(solver, precon))
        n = self.shape[0]
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
            mask = self.plan.ctx.device_alloc(n)
            mask.from_host(fixed.astype(numpy.uint8))
        x = self._vec(lhs)
        b = None if rhs is None else self._vec(rhs)
        try:
            self.plan.cg_device(self.values, b, x, constrained=mask, atol=atol, rtol=rtol, maxiter=maxiter)
        except ToleranceNotReached as e:
            raise ToleranceNotReached(str(e), best=x.to_host()) from None
        return x.to_host()

    def export(self, form):
        values = self.values.to_host() if hasattr(self.values, 'to_host') else self.values.cpu().numpy()
        rowptr, colidx = self.plan.csr_pattern()
        return Matrix(values, rowptr, colidx, self.shape[1]).export(form)


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
