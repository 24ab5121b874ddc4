'''Stand-in for ``appdirs`` (test infrastructure; lets the reference import here).'''
import os
import tempfile


def user_cache_dir(appname=None, appauthor=None, *args, **kwargs):
    return os.path.join(tempfile.gettempdir(), 'b200fem-refcache', appname or 'app')
