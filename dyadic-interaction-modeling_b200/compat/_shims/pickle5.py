"""Stand-in for the `pickle5` backport (the reference imports it as `pickle`): Python >= 3.8 has protocol 5 built in."""
from pickle import *  # noqa: F401,F403
from pickle import HIGHEST_PROTOCOL, dump, dumps, load, loads  # noqa: F401
