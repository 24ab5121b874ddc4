'''ctypes front-end of oracle/fem_oracle.c (TEST INFRASTRUCTURE ONLY; see the header of that file).'''

import ctypes
import os
import subprocess
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libfem_oracle.so')

FORM_MASS, FORM_STIFFNESS, FORM_ELASTICITY, FORM_GENERIC, FORM_LOAD = range(5)
_KIND = dict(mass=FORM_MASS, stiffness=FORM_STIFFNESS, elasticity=FORM_ELASTICITY, generic=FORM_GENERIC, load=FORM_LOAD)

_dp = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


class _Problem(ctypes.Structure):
    _fields_ = [('ndims', ctypes.c_int32), ('ncomp', ctypes.c_int32), ('nelems', ctypes.c_int64 * 3), ('degree', ctypes.c_int32 * 3),
                ('nq', ctypes.c_int32 * 3), ('ndofs_d', ctypes.c_int64 * 3), ('coeffs', _dp * 3), ('setidx', _i32p * 3),
                ('start', _i64p * 3), ('qpts', _dp * 3), ('qwts', _dp * 3), ('nodes', _dp)]


def build(force=False):
    'compile the shared library with the committed Makefile'
    src = os.path.join(HERE, 'fem_oracle.c')
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(['make', '-C', HERE, '-s'] + (['-B'] if force else []), check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_sort_unique.restype = ctypes.c_int64
        _lib.oracle_element_loop.restype = ctypes.c_int
    return _lib


def _ptr(a, T):
    return a.ctypes.data_as(T)


class CProblem:
    'keeps the numpy arrays of an oracle.fem_oracle.Problem alive next to the C struct'

    def __init__(self, prob):
        self.prob = prob
        self.keep = []
        s = _Problem()
        s.ndims = prob.ndims
        s.ncomp = prob.ncomp
        for d in range(prob.ndims):
            s.nelems[d] = prob.nelems[d]
            s.degree[d] = prob.degree[d]
            s.nq[d] = len(prob.qpts[d])
            s.ndofs_d[d] = prob.ndofs_d[d]
            arrs = (numpy.ascontiguousarray(prob.coeffs[d], dtype=numpy.float64), numpy.ascontiguousarray(prob.setidx[d], dtype=numpy.int32),
                    numpy.ascontiguousarray(prob.start[d], dtype=numpy.int64), numpy.ascontiguousarray(prob.qpts[d], dtype=numpy.float64),
                    numpy.ascontiguousarray(prob.qwts[d], dtype=numpy.float64))
            self.keep.append(arrs)
            s.coeffs[d] = _ptr(arrs[0], _dp)
            s.setidx[d] = _ptr(arrs[1], _i32p)
            s.start[d] = _ptr(arrs[2], _i64p)
            s.qpts[d] = _ptr(arrs[3], _dp)
            s.qwts[d] = _ptr(arrs[4], _dp)
        nodes = numpy.ascontiguousarray(prob.nodes, dtype=numpy.float64)
        self.keep.append(nodes)
        s.nodes = _ptr(nodes, _dp)
        self.struct = s


def _forms(forms):
    kinds = numpy.array([_KIND[f[0]] for f in forms], dtype=numpy.int32)
    params = [numpy.ascontiguousarray(numpy.ravel(numpy.asarray(f[1:] if f[0] == 'elasticity' else (f[1] if len(f) > 1 else 0.), dtype=numpy.float64))) for f in forms]
    ptrs = (_dp * max(len(forms), 1))(*[_ptr(p, _dp) for p in params])
    return kinds, params, ptrs


def element_loop(prob, matrix_forms=(), vector_forms=(), elem_range=None, nthreads=0, want_indices=True):
    '''COO output of the element loop: ([values...], rows, cols, [rhs...]).'''
    cp = CProblem(prob)
    e0, e1 = elem_range or (0, prob.ntotal)
    ne = int(numpy.prod([p + 1 for p in prob.degree])) * prob.ncomp
    n = (e1 - e0) * ne * ne
    vals = [numpy.empty(n) for _ in matrix_forms]
    rows = numpy.empty(n if want_indices else 0, dtype=numpy.int64)
    cols = numpy.empty(n if want_indices else 0, dtype=numpy.int64)
    rhs = [numpy.zeros(prob.ndofs) for _ in vector_forms]
    mk, mp, mpp = _forms(matrix_forms)
    vk, vp, vpp = _forms(vector_forms)
    vptr = (_dp * max(len(vals), 1))(*[_ptr(v, _dp) for v in vals])
    rptr = (_dp * max(len(rhs), 1))(*[_ptr(r, _dp) for r in rhs])
    status = lib().oracle_element_loop(ctypes.byref(cp.struct), ctypes.c_int64(e0), ctypes.c_int64(e1),
                                       len(vals), _ptr(mk, _i32p), mpp, vptr,
                                       _ptr(rows, _i64p) if want_indices else None, _ptr(cols, _i64p) if want_indices else None,
                                       len(rhs), _ptr(vk, _i32p), vpp, rptr, int(nthreads))
    if status:
        raise RuntimeError('oracle_element_loop failed')
    return vals, rows, cols, rhs


def coo_to_csr(values_list, rows, cols, nrows, ncols):
    n = len(rows)
    inverse = numpy.empty(n, dtype=numpy.int64)
    ukeys = numpy.empty(n, dtype=numpy.int64)
    nnz = lib().oracle_sort_unique(ctypes.c_int64(n), _ptr(rows, _i64p), _ptr(cols, _i64p), ctypes.c_int64(ncols), _ptr(inverse, _i64p), _ptr(ukeys, _i64p))
    if nnz < 0:
        raise MemoryError('oracle_sort_unique')
    rowptr = numpy.empty(nrows + 1, dtype=numpy.int64)
    colidx = numpy.empty(nnz, dtype=numpy.int64)
    lib().oracle_compress_rows(ctypes.c_int64(nnz), _ptr(ukeys, _i64p), ctypes.c_int64(nrows), ctypes.c_int64(ncols), _ptr(rowptr, _i64p), _ptr(colidx, _i64p))
    out = []
    for v in values_list:
        data = numpy.empty(nnz)
        lib().oracle_accumulate(ctypes.c_int64(n), _ptr(inverse, _i64p), _ptr(v, _dp), ctypes.c_int64(nnz), _ptr(data, _dp))
        out.append(data)
    return out, rowptr, colidx


def assemble(prob, matrix_forms=(), vector_forms=(), nthreads=0):
    '''Same contract as oracle.fem_oracle.assemble: ([(values,rowptr,colidx),...], [rhs,...]).'''
    vals, rows, cols, rhs = element_loop(prob, matrix_forms, vector_forms, nthreads=nthreads)
    datas, rowptr, colidx = coo_to_csr(vals, rows, cols, prob.ndofs, prob.ndofs)
    return [(d, rowptr, colidx) for d in datas], rhs


def max_threads():
    return int(lib().oracle_max_threads())
