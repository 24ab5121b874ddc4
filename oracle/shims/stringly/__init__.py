'''Stand-in for ``stringly`` (test infrastructure; lets the reference import here).'''
from . import util  # noqa: F401
