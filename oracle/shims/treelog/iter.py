'''Part of the treelog stand-in (test infrastructure, see __init__): iterator wrappers.'''
import builtins


class _Wrap:
    def __init__(self, iterable):
        self._it = builtins.iter(iterable)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)

    def close(self):
        pass


def _zipped(args):
    return args[0] if len(args) == 1 else zip(*args)


def plain(title, *args):
    return _Wrap(_zipped(args))


def fraction(title, *args, length=None):
    return _Wrap(_zipped(args))


def percentage(title, *args, length=None):
    return _Wrap(_zipped(args))


def wrap(titles, iterable):
    return _Wrap(iterable)
