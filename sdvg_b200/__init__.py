"""Import shim: the product package lives in ``sd-video-gen_b200/`` (a directory name that is not a valid
Python identifier); ``import sdvg_b200`` resolves its sub-modules from there."""
import os as _os

_PKG = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "sd-video-gen_b200")
__path__.insert(0, _PKG)

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
