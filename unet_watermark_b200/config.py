"""Configuration surface of the path (mirror of reference src/configs/config.py).

The reference uses ``yacs.config.CfgNode``; yacs is not a dependency here, so ``CfgNode`` below is a
small attribute-dict with the subset of the yacs API the path touches: attribute access,
``clone``/``defrost``/``freeze``/``merge_from_file``/``merge_from_other_cfg``/``merge_from_list``/``dump``.
Defaults are the reference's (src/configs/config.py:8-86); the keys the hot path reads are
``MODEL.*``, ``DATA.IMG_SIZE``, ``PREDICT.THRESHOLD`` and ``DEVICE``.

``install_yacs_shim()`` registers this class as ``yacs.config.CfgNode`` when yacs is missing so that
checkpoints written by the reference trainer — which pickle the live CfgNode under ``'config'``
(reference src/train.py:433) — can be unpickled by ``torch.load(..., weights_only=False)``.
"""
from __future__ import annotations

import copy
import logging
import sys
import types
from typing import Any

import yaml


logger = logging.getLogger(__name__)


class CfgNode(dict):
    IMMUTABLE = "__immutable__"

    def __init__(self, init_dict=None):
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # attribute access -------------------------------------------------------------------------
    def __getattr__(self, name: str) -> Any:
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name: str, value: Any):
        if self.is_frozen():
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        self[name] = value

    def __setstate__(self, state):          # pickles written by yacs carry {'__immutable__': ...}
        self.__dict__.update(state if isinstance(state, dict) else {})
        self.__dict__.setdefault(CfgNode.IMMUTABLE, False)

    # yacs API subset --------------------------------------------------------------------------
    def is_frozen(self) -> bool:
        return self.__dict__.get(CfgNode.IMMUTABLE, False)

    def _immutable(self, flag: bool):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._immutable(flag)

    def freeze(self):
        self._immutable(True)

    def defrost(self):
        self._immutable(False)

    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode.IMMUTABLE] = self.is_frozen()
        return out

    def merge_from_other_cfg(self, other: "CfgNode"):
        _merge(other, self, [])

    def merge_from_file(self, path: str, allow_new: bool = False):
        with open(path, "r", encoding="utf-8") as f:
            loaded = yaml.safe_load(f) or {}
        _merge(CfgNode(loaded), self, [], allow_new)

    def merge_from_list(self, kv):
        if len(kv) % 2:
            raise ValueError("merge_from_list expects KEY VALUE pairs")
        for key, val in zip(kv[0::2], kv[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    raise KeyError(f"Non-existent key: {key}")
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(f"Non-existent key: {key}")
            if isinstance(val, str):
                try:
                    val = yaml.safe_load(val)
                except yaml.YAMLError:
                    pass
            dict.__setitem__(node, parts[-1], val)

    def to_dict(self) -> dict:
        return {k: (v.to_dict() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def dump(self, **kwargs) -> str:
        return yaml.safe_dump(self.to_dict(), **kwargs)


def _merge(src: CfgNode, dst: CfgNode, path, allow_new: bool = False):
    for k, v in src.items():
        full = ".".join(path + [k])
        if k not in dst:
            if not allow_new:
                raise KeyError(f"Non-existent config key: {full}")      # yacs behaviour
            # reference YAMLs such as unet_text_watermark.yaml carry keys the reference's own schema does not
            # declare (DATA.TEXT_ENHANCEMENT, TEXT_WATERMARK.*): the mask path never reads them
            logger.warning("config key %s is not in the schema; kept as given", full)
            dict.__setitem__(dst, k, CfgNode(v) if isinstance(v, dict) else copy.deepcopy(v))
            continue
        if isinstance(v, dict) and isinstance(dst[k], CfgNode):
            _merge(v if isinstance(v, CfgNode) else CfgNode(v), dst[k], path + [k], allow_new)
        else:
            dict.__setitem__(dst, k, copy.deepcopy(v))


CN = CfgNode

_C = CN()
_C.DEVICE = "cpu"

_C.MODEL = CN()
_C.MODEL.NAME = "UnetPlusPlus"          # reference default (config.py:15); this path needs "Unet"
_C.MODEL.ENCODER_NAME = "resnet34"
_C.MODEL.ENCODER_WEIGHTS = "imagenet"
_C.MODEL.ENCODER_DEPTH = 5
_C.MODEL.DECODER_CHANNELS = [256, 128, 64, 32, 16]
_C.MODEL.IN_CHANNELS = 3
_C.MODEL.CLASSES = 1
_C.MODEL.ACTIVATION = None

_C.DATA = CN()
_C.DATA.ROOT_DIR = "data/train"
_C.DATA.ADDITIONAL_ROOT_DIRS = []
_C.DATA.IMG_SIZE = 512
_C.DATA.GENERATE_MASK_THRESHOLD = 30
_C.DATA.TRAIN_RATIO = 0.8
_C.DATA.VAL_RATIO = 0.2
_C.DATA.SHUFFLE = True
_C.DATA.SEED = 42
_C.DATA.NUM_WORKERS = 4
_C.DATA.CACHE_IMAGES = False
_C.DATA.PREFETCH_FACTOR = 2
_C.DATA.AUGMENTATION_TYPE = "transparent_watermark"

_C.TRAIN = CN()
_C.TRAIN.BATCH_SIZE = 16
_C.TRAIN.EPOCHS = 300
_C.TRAIN.LR = 0.0001
_C.TRAIN.WEIGHT_DECAY = 0.0001
_C.TRAIN.OUTPUT_DIR = "logs/output"
_C.TRAIN.MODEL_SAVE_PATH = "models/unet_watermark.pth"
_C.TRAIN.LOG_INTERVAL = 10
_C.TRAIN.SAVE_INTERVAL = 50
_C.TRAIN.USE_EARLY_STOPPING = True
_C.TRAIN.EARLY_STOPPING_PATIENCE = 10
_C.TRAIN.CHECKPOINT_DIR = "models/checkpoints"
_C.TRAIN.SAVE_BEST_ONLY = False
_C.TRAIN.USE_AMP = False
_C.TRAIN.GRADIENT_CLIP = 1.0

_C.LOSS = CN()
_C.LOSS.NAME = "DiceLoss"
_C.LOSS.MODE = "binary"
_C.LOSS.SMOOTH = 1e-5
_C.LOSS.BCE_WEIGHT = 0.5
_C.LOSS.DICE_WEIGHT = 0.5
_C.LOSS.DICE_SMOOTH = 1e-5
_C.LOSS.FOCAL_ALPHA = 0.25
_C.LOSS.FOCAL_GAMMA = 2.0

_C.OPTIMIZER = CN()
_C.OPTIMIZER.NAME = "Adam"
_C.OPTIMIZER.LR_SCHEDULER = "ReduceLROnPlateau"
_C.OPTIMIZER.SCHEDULER_PATIENCE = 5
_C.OPTIMIZER.SCHEDULER_FACTOR = 0.5

_C.PREDICT = CN()
_C.PREDICT.INPUT_PATH = "data/input"
_C.PREDICT.OUTPUT_DIR = "data/output"
_C.PREDICT.BATCH_SIZE = 8
_C.PREDICT.AUTO_BATCH_SIZE = True
_C.PREDICT.MAX_BATCH_SIZE = 32
_C.PREDICT.THRESHOLD = 0.5
_C.PREDICT.POST_PROCESS = True

_C.VAL = CN()
_C.VAL.METRICS = ["dice", "iou", "accuracy"]


def get_cfg_defaults() -> CfgNode:
    """Copy of the defaults (reference src/configs/config.py:88-90)."""
    return _C.clone()


def update_config(cfg: CfgNode, config_file: str, strict: bool = True):
    """Merge a YAML file and freeze (reference src/configs/config.py:92-96).  ``strict=True`` is yacs' behaviour
    (KeyError on keys outside the schema); the CLI passes ``strict=False`` so that every YAML shipped with the
    reference loads (unknown keys are kept and logged)."""
    cfg.defrost()
    cfg.merge_from_file(config_file, allow_new=not strict)
    cfg.freeze()


def install_yacs_shim():
    """Make ``yacs.config.CfgNode`` importable (as this class) if yacs itself is not installed."""
    try:
        import yacs.config  # noqa: F401
        return False
    except ImportError:
        pkg = types.ModuleType("yacs")
        mod = types.ModuleType("yacs.config")
        mod.CfgNode = CfgNode
        pkg.config = mod
        sys.modules.setdefault("yacs", pkg)
        sys.modules.setdefault("yacs.config", mod)
        return True
