"""hipims_ocl_b200 -- B200-native executor for the HiPIMS explicit shallow-water hot path."""
from .config import *  # noqa: F401,F403
from .config import SchemeConfig  # noqa: F401
