'''Minimal stand-in for the third-party ``treelog`` package (absent from this image).

TEST INFRASTRUCTURE ONLY.  It exists so that the unmodified reference
(/root/reference/src/nutils) can be imported inside this container to generate
golden vectors (oracle/make_golden.py).  Logging calls are no-ops; iterators are
plain pass-throughs.  Nothing under nutils_b200/ imports this.
'''

import contextlib
import functools
import io
import os
import tempfile
from . import proto, iter  # noqa: F401


def _noop(*args, **kwargs):
    pass


debug = info = user = warning = error = _noop


@contextlib.contextmanager
def context(title, *args, **kwargs):
    yield lambda *a, **k: None


def withcontext(f):
    @functools.wraps(f)
    def wrapped(*args, **kwargs):
        return f(*args, **kwargs)
    return wrapped


@contextlib.contextmanager
def _devnull_file(name, mode='w', **kwargs):
    with open(os.devnull, mode) as f:
        yield f


debugfile = infofile = userfile = warningfile = errorfile = _devnull_file


class _Log:
    def __init__(self, *args, **kwargs):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def pushcontext(self, title):
        pass

    def popcontext(self):
        pass

    def recontext(self, title):
        pass

    def write(self, msg, level):
        pass

    def replay(self, log=None):
        pass

    @property
    def filename(self):
        return os.path.join(tempfile.gettempdir(), 'log.html')


NullLog = StdoutLog = RichOutputLog = LoggingLog = HtmlLog = RecordLog = DataLog = TeeLog = _Log


class FilterLog(_Log):
    pass


current = _Log()


@contextlib.contextmanager
def set(log):
    global current
    old, current = current, log
    try:
        yield log
    finally:
        current = old


@contextlib.contextmanager
def add(log):
    yield log


def disable():
    return set(_Log())
