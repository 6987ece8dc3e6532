"""Model zoo on the fused step program.  ``get_model`` mirrors ``main.py:37-68`` of the reference
(case-insensitive names); families not yet ported raise ``NotImplementedError`` by name."""
from .esmm import ESMM
from .mmoe import MMOE
from .ple import PLE
from .sharedbottom import SharedBottom

_REGISTRY = {"mmoe": MMOE, "ple": PLE, "sharedbottom": SharedBottom, "esmm": ESMM}
try:  # families added after the first milestone
    from .star import STAR
    _REGISTRY["star"] = STAR
except ImportError:
    pass
try:
    from .pepnet import PepNet
    _REGISTRY["pepnet"] = PepNet
except ImportError:
    pass

REFERENCE_NAMES = ("mmoe", "esmm", "sharedbottom", "ple", "snr_trans", "mssm", "star", "pcg", "apg", "mlp",
                   "cross_stitch", "aitm", "escm", "hmoe", "pepnet")


def get_model_class(model_name: str):
    name = model_name.lower()
    if name in _REGISTRY:
        return _REGISTRY[name]
    if name in REFERENCE_NAMES:
        raise NotImplementedError(f"model '{name}' is outside the B200 hot-path scope (see DESIGN.md)")
    raise ValueError(f"unknown model name {model_name!r}")


def get_model(model_name, df_columns=None, config=None, device="cuda"):
    return get_model_class(model_name)(df_columns, device=device, config=config)
