'''Part of the ags stand-in: parse simple scalars from strings.'''
import typing


def loads(s, T):
    origin = typing.get_origin(T)
    if origin is typing.Union:
        for t in typing.get_args(T):
            if t is type(None):
                continue
            return loads(s, t)
    if T is bool:
        if s.lower() in ('true', 'yes', '1', 'on'):
            return True
        if s.lower() in ('false', 'no', '0', 'off'):
            return False
        raise ValueError(s)
    if T in (int, float, str, complex):
        return T(s)
    return s


def dumps(obj, T=None):
    return str(obj)
