"""Import shim for the REAL reference (``/root/reference``).  TEST INFRASTRUCTURE ONLY.

Exists only in the authoring container; the GPU box has no ``/root/reference``
so nothing in ``-m gpu`` tests / ``smoke()`` / ``bench.py`` may call this.
Used by ``oracle/make_golden.py`` and by CPU tests that are skipped when the
reference is absent.  Recipe: SURVEY.md Appendix A (none of the shims touch
arithmetic).
"""
import os
import sys
import types

REF_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "oscar", "modeling"))


def load():
    """Returns the reference's ``oscar.modeling.modeling_vlbert`` module."""
    if not available():
        raise RuntimeError("reference not present at " + REF_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if not isinstance(sys.modules.get("transformers"), types.ModuleType) or \
            getattr(sys.modules.get("transformers"), "__path__", None) != [REF_ROOT + "/transformers"]:
        m = types.ModuleType("transformers")
        m.__path__ = [REF_ROOT + "/transformers"]
        sys.modules["transformers"] = m
    for name in ("boto3", "botocore", "botocore.exceptions", "anytree"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["botocore.exceptions"].ClientError = Exception
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from transformers.pytorch_transformers.modeling_utils import PreTrainedModel

    def _tie(self, first, second, only_vocab=False, only_word_size=None):
        # modern torch refuses to assign a tensor slice to a registered Parameter
        # (reference modeling_utils.py:282); make it a true view so grads reach
        # word_embeddings exactly as in the reference.
        if only_vocab:
            first._parameters.pop("weight", None)
            first.weight = second.weight[:only_word_size, :]
        else:
            first.weight = second.weight

    PreTrainedModel._tie_or_clone_weights = _tie
    import oscar.modeling.modeling_vlbert as mv
    return mv


def make_config(mv, cfg, **extra):
    """Reference BertConfig from an oracle Cfg."""
    from transformers.pytorch_transformers.modeling_bert import BertConfig
    c = BertConfig(vocab_size_or_config_json_file=cfg.vocab_size, hidden_size=cfg.hidden_size,
                   num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                   intermediate_size=cfg.intermediate_size, max_position_embeddings=cfg.max_position_embeddings,
                   type_vocab_size=cfg.type_vocab_size, layer_norm_eps=cfg.layer_norm_eps,
                   hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    c.only_word_size = cfg.only_word_size
    c.qa_answer_size = cfg.qa_answer_size
    c.img_feature_dim = cfg.img_feature_dim
    c.img_feature_type = "faster_r-cnn"
    c.use_img_layernorm = cfg.use_img_layernorm
    c.img_layer_norm_eps = cfg.img_layer_norm_eps
    c.loss_type = cfg.loss_type
    c.num_labels = cfg.num_labels
    c.num_contrast_classes = cfg.num_contrast_classes
    for k, v in extra.items():
        setattr(c, k, v)
    return c
