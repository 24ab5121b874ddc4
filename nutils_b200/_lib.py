'''ctypes binding of the C ABI in include/b200fem.h (libb200fem.so, built in-tree by csrc/build.sh).

Follows the reference's one ctypes boundary, the MKL matrix backend
(src/nutils/matrix/_mkl.py:10-12, 90-97; _util.loadlib :195-237): the library is
located next to the package, arrays are passed as C-contiguous numpy buffers,
status codes are mapped to exceptions, and a missing library raises
:class:`BackendNotAvailable` at load time.  There is no CPU fallback.
'''

import ctypes
import os
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, 'libb200fem.so')


class B200Error(Exception):
    'error reported by libb200fem (the MatrixError of this backend)'


class ToleranceNotReached(B200Error):
    'iterative solve stopped above the requested tolerance; ``best`` holds the last iterate (nutils.matrix.ToleranceNotReached)'

    def __init__(self, msg, best=None):
        super().__init__(msg)
        self.best = best


class BackendNotAvailable(B200Error):
    'libb200fem.so or a usable sm_100 device is missing (cf. nutils.matrix.BackendNotAvailable)'


c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_vp = ctypes.c_void_p
p_vp = ctypes.POINTER(ctypes.c_void_p)
p_f64 = ctypes.POINTER(ctypes.c_double)
p_i32 = ctypes.POINTER(ctypes.c_int32)
p_i64 = ctypes.POINTER(ctypes.c_int64)
pp_f64 = ctypes.POINTER(p_f64)
pp_i32 = ctypes.POINTER(p_i32)
pp_i64 = ctypes.POINTER(p_i64)

# name -> (restype, argtypes); must list every function declared in include/b200fem.h
SIGNATURES = {
    'b2_strerror': (ctypes.c_char_p, [ctypes.c_int]),
    'b2_last_error': (ctypes.c_char_p, [c_vp]),
    'b2_version': (ctypes.c_int, []),
    'b2_device_count': (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    'b2_ctx_create': (ctypes.c_int, [ctypes.c_int, p_vp]),
    'b2_ctx_destroy': (ctypes.c_int, [c_vp]),
    'b2_ctx_set_stream': (ctypes.c_int, [c_vp, c_vp]),
    'b2_ctx_synchronize': (ctypes.c_int, [c_vp]),
    'b2_ctx_timer_start': (ctypes.c_int, [c_vp]),
    'b2_ctx_timer_stop': (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_float)]),
    'b2_ctx_launch_count': (c_i64, [c_vp]),
    'b2_ctx_kernel_time': (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_double), p_i64]),
    'b2_pattern_row_offset': (c_i64, [c_vp, c_i64]),
    'b2_ctx_set_option': (ctypes.c_int, [c_vp, ctypes.c_char_p, c_i64]),
    'b2_host_alloc': (ctypes.c_int, [c_vp, c_i64, p_vp]),
    'b2_host_free': (ctypes.c_int, [c_vp, c_vp]),
    'b2_device_alloc': (ctypes.c_int, [c_vp, c_i64, p_vp]),
    'b2_device_free': (ctypes.c_int, [c_vp, c_vp]),
    'b2_memcpy_d2h': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64]),
    'b2_memcpy_h2d': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64]),
    'b2_memset_zero': (ctypes.c_int, [c_vp, c_vp, c_i64]),
    'b2_flush_l2': (ctypes.c_int, [c_vp]),
    'b2_basis_create': (ctypes.c_int, [c_vp, ctypes.c_int, p_i64, p_i32, p_i32, pp_f64, pp_i32, pp_i64, p_i64, ctypes.c_int, p_vp]),
    'b2_basis_destroy': (ctypes.c_int, [c_vp]),
    'b2_basis_ndofs': (c_i64, [c_vp]),
    'b2_quad_create_tensor': (ctypes.c_int, [c_vp, ctypes.c_int, p_i32, pp_f64, pp_f64, p_vp]),
    'b2_quad_destroy': (ctypes.c_int, [c_vp]),
    'b2_geom_create_nodal': (ctypes.c_int, [c_vp, ctypes.c_int, p_i64, c_vp, p_vp]),
    'b2_geom_update_nodal': (ctypes.c_int, [c_vp, c_vp]),
    'b2_geom_destroy': (ctypes.c_int, [c_vp]),
    'b2_pattern_create': (ctypes.c_int, [c_vp, c_vp, p_vp]),
    'b2_pattern_destroy': (ctypes.c_int, [c_vp]),
    'b2_pattern_nnz': (c_i64, [c_vp]),
    'b2_pattern_nrows': (c_i64, [c_vp]),
    'b2_pattern_export_host': (ctypes.c_int, [c_vp, c_vp, c_vp]),
    'b2_pattern_export_device': (ctypes.c_int, [c_vp, c_vp, c_vp]),
    'b2_assemble_device': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, pp_f64, p_vp, ctypes.c_int, pp_f64, p_vp]),
    'b2_assemble_rows_device': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, pp_f64, p_vp, ctypes.c_int, pp_f64, p_vp]),
    'b2_elemset_create': (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, ctypes.c_int, p_vp]),
    'b2_elemset_destroy': (ctypes.c_int, [c_vp]),
    'b2_elemset_set_faces': (ctypes.c_int, [c_vp, c_vp]),
    'b2_elemset_set_normals': (ctypes.c_int, [c_vp, c_vp, c_i64]),
    'b2_elemset_set_coefficient': (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_i64]),
    'b2_elemset_set_coefficient_field': (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_double]),
    'b2_elemset_ndofs': (c_i64, [c_vp]),
    'b2_elemset_npoints': (c_i64, [c_vp]),
    'b2_geom_create_spline': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, p_vp]),
    'b2_pattern_create_elemset': (ctypes.c_int, [c_vp, c_vp, p_vp]),
    'b2_assemble_elemset_device': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, pp_f64, p_vp, ctypes.c_int, pp_f64, p_vp]),
    'b2_assemble_elemset_host': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, pp_f64, p_vp, ctypes.c_int, pp_f64, p_vp]),
    'b2_evaluate_elemset_device': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'b2_spmv_device': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    'b2_diagonal_device': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    'b2_cg_device': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)]),
    'b2_pattern_create_csr': (ctypes.c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, p_vp]),
    'b2_pattern_general': (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, p_vp]),
    'b2_assemble_general_host': (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_i64, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, pp_f64, pp_f64, pp_f64, pp_f64, pp_f64,
                                                c_vp, c_vp, c_vp, c_vp, ctypes.c_int, pp_f64, p_vp, ctypes.c_int, pp_f64, p_vp]),
    'b2_assemble_host': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, pp_f64, p_vp, ctypes.c_int, pp_f64, p_vp]),
}

_lib = None


def load():
    '''Load libb200fem.so and declare the signatures; raises BackendNotAvailable if it is missing.'''
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise BackendNotAvailable('{} not found; build it with nutils_b200/csrc/build.sh (or __graft_entry__.build())'.format(LIBPATH))
        try:
            lib = ctypes.CDLL(LIBPATH)
        except OSError as e:
            raise BackendNotAvailable('cannot load {}: {}'.format(LIBPATH, e))
        for name, (restype, argtypes) in SIGNATURES.items():
            f = getattr(lib, name)
            f.restype = restype
            f.argtypes = argtypes
        _lib = lib
    return _lib


def check(status, ctx=None):
    if status == 0:
        return
    lib = load()
    msg = lib.b2_strerror(status).decode()
    if ctx:
        detail = lib.b2_last_error(ctx).decode()
        if detail:
            msg = '{}: {}'.format(msg, detail)
    if status == -4:
        raise BackendNotAvailable(msg)
    if status == -6:
        raise ToleranceNotReached(msg)
    raise B200Error(msg)


def as_f64(a):
    return numpy.ascontiguousarray(a, dtype=numpy.float64)


def ptr_array(arrays, ptype):
    'C array of pointers to the given numpy arrays (which the caller keeps alive)'
    return (ptype * max(len(arrays), 1))(*[a.ctypes.data_as(ptype) for a in arrays])
