'''Part of the ags stand-in.'''


def dumps(obj, T=None):
    return repr(obj)


def loads(s, T):
    from .ucsl import loads as _l
    return _l(s, T)
