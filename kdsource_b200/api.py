"""``import kdsource_b200.api as kds`` -- the import facade of the reference
(``python/kdsource/api.py.in:42-50``) for the kernel-density path."""

__version__ = "0.1.0"
__author__ = "kdsource-b200"

from . import geom  # noqa: F401
from . import kde  # noqa: F401
from . import plist  # noqa: F401
from .geom import Geometry, Metric  # noqa: F401
from .kde import bw_silv  # noqa: F401
from .kdsource import KDSource, load  # noqa: F401
from .plist import PList, appendssv, convert2mcpl, join2mcpl, savessv  # noqa: F401
from .resample import resample_to_mcpl  # noqa: F401
from .writemcpl import write_mcpl  # noqa: F401
