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
