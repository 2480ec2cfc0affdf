from .cyclic_scheduler import CyclicLRWithRestarts  # noqa: F401
from .ranger import Ranger  # noqa: F401
from .fused_adam import FusedAdam  # noqa: F401
