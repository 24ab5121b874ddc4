'''The C-ABI library loads, exports every symbol include/b200fem.h declares, and refuses to work
without a device (no CPU fallback).  No compute calls here.'''

import ctypes
import os
import re
import pytest

from nutils_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'b200fem.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(b2_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIBPATH)
    for name in names:
        assert hasattr(lib, name), name


def test_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == _declared()


def test_load_and_strerror():
    lib = _lib.load()
    assert lib.b2_version() >= 100
    assert lib.b2_strerror(0) == b'ok'
    assert b'invalid' in lib.b2_strerror(-1)


def test_null_arguments_rejected():
    lib = _lib.load()
    assert lib.b2_ctx_create(0, None) == -1
    assert lib.b2_pattern_nnz(None) == 0
    assert lib.b2_ctx_synchronize(None) == -1


@pytest.mark.skipif(os.path.exists('/dev/nvidiactl'), reason='GPU present')
def test_no_device_no_fallback():
    from nutils_b200 import engine
    with pytest.raises(_lib.BackendNotAvailable):
        engine.Context(0)
