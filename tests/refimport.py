"""Import the reference's image_generation modules read-only from /root/reference (only
available in the build container, never on the GPU box).  Three in-memory stubs stand in for
packages the reference imports but the hot path never uses."""
import os
import sys
import types

import numpy as np

REF_ROOT = "/root/reference/image_generation"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model.py"))


class _EasyDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


_loaded = None


def load(words_num: int = 18):
    """Returns a namespace with the reference modules: model, GlobalAttention, losses, utils, cfg."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference not mounted")
    np.int, np.float = int, float  # removed numpy aliases the reference still uses
    stubs = {"easydict": {"EasyDict": _EasyDict}, "skimage": {}, "skimage.transform": {},
             "models.roi_align._ext": {}, "models.roi_align._ext.roi_align": {}}
    for name, attrs in stubs.items():
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    sys.path.insert(0, REF_ROOT)
    from miscc.config import cfg
    cfg.CUDA = False
    cfg.TEXT.WORDS_NUM = words_num
    import model
    import GlobalAttention
    from miscc import losses, utils
    _loaded = types.SimpleNamespace(model=model, GlobalAttention=GlobalAttention, losses=losses, utils=utils,
                                    cfg=cfg)
    return _loaded
