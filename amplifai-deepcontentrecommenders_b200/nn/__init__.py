from .dcue import DCUE  # noqa: F401
from .trainer import Trainer  # noqa: F401
