"""Model zoo on the fused step program.  ``get_model`` mirrors ``main.py:37-68`` of the reference
(case-insensitive names): all fifteen names the reference's factory knows are on the fused step."""
from .aitm import AITM
from .apg import APG
from .cross_stitch import CrossStitch
from .escm import ESCM
from .esmm import ESMM
from .hmoe import HMOE
from .mlp import MLP
from .mmoe import MMOE
from .mssm import MSSM
from .pepnet import PepNet
from .ple import PLE
from .sharedbottom import SharedBottom
from .snr_trans import SNR_trans
from .star import STAR

_REGISTRY = {"mmoe": MMOE, "ple": PLE, "sharedbottom": SharedBottom, "esmm": ESMM, "star": STAR, "pepnet": PepNet,
             "mlp": MLP, "cross_stitch": CrossStitch, "hmoe": HMOE, "escm": ESCM, "aitm": AITM, "snr_trans": SNR_trans,
             "mssm": MSSM, "apg": APG,
             # main.py:53-54 builds an MMOE for 'pcg' and wraps its optimizer in PCGrad (basemodel.py:564-565).  The loop
             # hands pc_backward ONE objective -- the summed loss (basemodel.py:309-310) -- so the projection is the
             # identity and the step is MMoE's (pinned by the golden case pcg_kuairec_adam)
             "pcg": MMOE}

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
