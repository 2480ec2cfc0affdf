"""B200 drop-in for the reference's dcrecommend/dcue/audiomodels/truedcuemel1d.py."""
from ._tower import TowerBase


class TrueDcueNetMel1D(TowerBase):

    """ConvNet used on data prepared with melspectogram transform (B200 kernels)."""

    _has_bn = False
    _res = False
