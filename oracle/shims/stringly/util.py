'''Part of the stringly stand-in: only what nutils._util.cli touches.'''


class DocString:
    def __init__(self, f):
        self.presets = {}
        self.defaults = {}
        self.text = getattr(f, '__doc__', '') or ''
        self.argdocs = {}
