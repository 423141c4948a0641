"""genome-nas_b200: B200-native (sm_100a) implementation of the genomeDARTS supernet
search step behind the reference's module API (GenomeNet/Genome-NAS).

Public names mirror the reference: MixedOp, CNN_Cell_search (darts_tools/model_discCNN.py),
DARTSCellSearch, RNNModelSearch (model_search.py), Architect (darts_tools/architect.py).
"""
from . import _lib  # noqa: F401
from .cnn import CNN_Cell_search, MixedOp  # noqa: F401
from .operations import OPS, PRIMITIVES_cnn, PRIMITIVES_rnn  # noqa: F401
from .rnn import DARTSCellSearch  # noqa: F401
from .model import Genotype, RNNModel, RNNModelSearch, encode_onehot, expand_onehot  # noqa: F401
from .architect import Architect  # noqa: F401
from .train import Hyper, SearchStep, infer, train  # noqa: F401
from .head import bce_loss, tar_loss  # noqa: F401
from . import mask_supernet  # noqa: F401  (OSP-NAS / Hyperband / random-search mask supernet)

__version__ = "0.1.0"
