"""factorizer_b200: B200-native (sm_100a) drop-in for the context-modelling hot path of
pashtari/factorizer -- ``import factorizer_b200 as ft`` exposes the same names as ``factorizer``."""
from .helpers import *  # noqa: F401,F403
from .layers import *  # noqa: F401,F403
from .operations import *  # noqa: F401,F403
from .matrix_factorization import *  # noqa: F401,F403
from .factorizer import *  # noqa: F401,F403
from .segmentation import *  # noqa: F401,F403
from . import distributed  # noqa: F401

__version__ = "0.1.0"
