'''Stand-in for ``ags`` (test infrastructure; lets the reference import here).

Only ``ucsl.loads(str, type)`` is functional (used by nutils'
``defaults_from_env`` to parse NUTILS_* environment variables).
'''
from . import ucsl, yaml  # noqa: F401


def load(path, sig):
    raise NotImplementedError('ags.load is not available in the oracle shim')
