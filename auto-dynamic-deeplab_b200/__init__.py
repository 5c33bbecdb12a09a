"""add_b200 — B200-native (sm_100a) implementation of Auto-Dynamic-DeepLab's dense-segmentation
forward path behind the reference's own Python operator API.

The directory is named `auto-dynamic-deeplab_b200/`; import it as `add_b200` (repo-root shim
`add_b200.py`).  Importing requires the in-tree `libadd_b200.so` (built by
`__graft_entry__.build()`); there is no CPU / PyTorch fallback."""
from ._lib import lib, AddError, EXPORTED_SYMBOLS
from .runtime import (set_default_precision, default_precision, set_tc_enabled, tc_available, Builder, Plan, View,
                      ConvWeights)
from .genotypes import PRIMITIVES, AUTODEEPLAB_CELL, NETWORKS
from .operations import (OPS, ReLUConvBN, DilConv, SepConv, Identity, Zero, FactorizedReduce,
                         DoubleFactorizedReduce, SynchronizedBatchNorm2d, ASPP, normalized_shannon_entropy,
                         confidence_max)
from .aspp_train import ASPP_train
from .decoder import Decoder
from .ADD import ADD, Cell, EDM
from .baseline_model import Baselin_Model, AutoDeepLab, Cell_baseline, Cell_AutoDeepLab
from .cell_level_search import MixedOp, softmax_rows
from . import cell_level_search
from . import training
from .metrics import Evaluator
from .factory import build_add, Args, synthetic_batch, synthetic_batch_u8, normalize_u8_hwc_host
from .pipeline import HostPipeline, ResidentPipeline
from .io_edges import (encode_segmap, decode_segmap, full_image_eval_preprocess, load_checkpoint, pad_labels,
                       get_cityscapes_labels, cityscapes_label_lut)
from .parallel import shard_range, env_rank_world, all_reduce_confusion

__version__ = "0.1.0"
