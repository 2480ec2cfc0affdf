"""B200 drop-in for the reference's dcrecommend/dcue/audiomodels/truedcuemel1dresbn.py."""
from ._tower import TowerBase


class TrueDcueNetMel1DResBn(TowerBase):

    """ConvNet used on data prepared with melspectogram transform (B200 kernels)."""

    _has_bn = True
    _res = True
