"""amplifai-deepcontentrecommenders_b200: B200-native (sm_100a) DCUE training + scoring hot path
behind the reference's DCUENet / DCUE API  (reference: `from dcrecommend import DCUE`)."""
from . import _lib, eval, graph, ops  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401
from .dcue.dcue import DCUENet  # noqa: F401
from .nn.dcue import DCUE  # noqa: F401

__all__ = ["DCUE", "DCUENet", "GraphedTrainStep"]
