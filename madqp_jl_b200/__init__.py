"""Import shim.

The product package lives in the directory ``madqp.jl_b200/`` (the layout the
project brief names).  A dot is not legal in a Python package name, so this
shim makes the same files importable as ``madqp_jl_b200.<module>`` by putting
that directory on this package's search path.  No code lives here.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "madqp.jl_b200"))

from madqp_jl_b200._version import __version__  # noqa: E402,F401
