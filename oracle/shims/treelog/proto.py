'''Part of the treelog stand-in (test infrastructure, see __init__).'''
import enum


class Level(enum.IntEnum):
    debug = 0
    info = 1
    user = 2
    warning = 3
    error = 4
