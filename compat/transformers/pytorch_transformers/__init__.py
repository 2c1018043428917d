"""`transformers.pytorch_transformers` as the reference scripts import it (run_retrieval.py:20-21,
run_pretrain_ml.py:26-31, run_vqa.py:26-27): config, checkpoint names, fused AdamW and the LR schedules come
from mvp_pytorch_b200; tokenizers (CPU text preparation, out of the hot path) are the reference's own modules,
found through the extended package path and imported lazily."""
import os
import sys

from mvp_pytorch_b200.modeling_utils import BertConfig, PretrainedConfig, PreTrainedModel, CONFIG_NAME, WEIGHTS_NAME  # noqa: F401
from mvp_pytorch_b200.optimization import (AdamW, ConstantLRSchedule, WarmupConstantSchedule, WarmupCosineSchedule,  # noqa: F401
                                           WarmupLinearSchedule)

_here = os.path.abspath(__path__[0])
for _root in list(sys.path):
    _cand = os.path.abspath(os.path.join(_root or ".", "transformers", "pytorch_transformers"))
    if _cand != _here and os.path.isfile(os.path.join(_cand, "tokenization_bert.py")) and _cand not in __path__:
        __path__.append(_cand)

_LAZY = {"BertTokenizer": "tokenization_bert", "BasicTokenizer": "tokenization_bert",
         "WordpieceTokenizer": "tokenization_bert", "PreTrainedTokenizer": "tokenization_utils"}


def __getattr__(name):
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError(f"transformers.pytorch_transformers (mvp_pytorch_b200 compat) has no attribute {name!r}")
    import importlib
    try:
        m = importlib.import_module(f"{__name__}.{mod}")
    except ImportError as exc:
        raise ImportError(f"{name} is the reference's CPU tokenizer ({mod}.py): put the reference checkout on "
                          "PYTHONPATH after compat/ (see compat/README.md)") from exc
    return getattr(m, name)
