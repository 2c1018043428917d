"""BERT building blocks with the reference's module tree (transformers/pytorch_transformers/
modeling_bert.py:229-533), so ``state_dict()`` keys/shapes are identical to the reference's.

These modules only OWN parameters.  They have no ``forward``: the arithmetic of the whole
stack runs in the CUDA engine (engine.py -> libmvptr_b200.so), driven by the model classes in
modeling_vlbert.py.  Calling one directly raises, so nothing can silently fall back to eager
PyTorch.
"""
import torch
from torch import nn

from .modeling_utils import BertConfig, PreTrainedModel  # noqa: F401  (re-exported like the reference)


class _ParamOnly(nn.Module):
    def forward(self, *a, **kw):
        raise RuntimeError(f"{type(self).__name__} is a parameter container; the computation runs in the "
                           "mvp_pytorch_b200 CUDA engine through the model's forward()")


class BertLayerNorm(_ParamOnly):
    """TF-style LayerNorm parameters (modeling_bert.py:233-246)."""

    def __init__(self, hidden_size, eps=1e-12):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.bias = nn.Parameter(torch.zeros(hidden_size))
        self.variance_epsilon = eps


class _Linear(nn.Linear):
    """nn.Linear used purely as a (weight, bias) holder with the reference's names."""

    def forward(self, *a, **kw):
        raise RuntimeError("parameter container; see mvp_pytorch_b200.engine")


class _Embedding(nn.Embedding):
    def forward(self, *a, **kw):
        raise RuntimeError("parameter container; see mvp_pytorch_b200.engine")


class BertEmbeddings(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = _Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = _Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = _Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class BertSelfAttention(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError(
                "The hidden size (%d) is not a multiple of the number of attention "
                "heads (%d)" % (config.hidden_size, config.num_attention_heads))
        self.output_attentions = config.output_attentions
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = _Linear(config.hidden_size, self.all_head_size)
        self.key = _Linear(config.hidden_size, self.all_head_size)
        self.value = _Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class BertSelfOutput(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        self.dense = _Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class BertAttention(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        self.self = BertSelfAttention(config)
        self.output = BertSelfOutput(config)


class BertIntermediate(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        if config.hidden_act != "gelu":
            raise NotImplementedError("the fused FFN epilogue implements the reference default erf-GELU only")
        self.dense = _Linear(config.hidden_size, config.intermediate_size)


class BertOutput(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        self.dense = _Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class BertLayer(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        self.attention = BertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)


class BertEncoder(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        self.output_attentions = config.output_attentions
        self.output_hidden_states = config.output_hidden_states
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(config.num_hidden_layers)])


class BertPooler(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        self.dense = _Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()


class BertPredictionHeadTransform(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        self.dense = _Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class _TiedDecoder(_ParamOnly):
    """``decoder`` of an ``only_vocab`` LM head: no parameter of its own -- it reads
    ``word_embeddings.weight[:only_word_size]`` (tie of modeling_utils.py:275-287 /
    modeling_vlbert.py:1208-1216), so gradients of both MLM heads accumulate into the
    embedding table's rows [0, only_word_size)."""


class BertLMPredictionHead(_ParamOnly):
    def __init__(self, config, only_vocab=False, tied=True):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        n_out = config.only_word_size if only_vocab else config.vocab_size
        self.n_out = n_out
        if tied:
            self.decoder = _TiedDecoder()
        else:
            self.decoder = _Linear(config.hidden_size, n_out, bias=False)
        self.bias = nn.Parameter(torch.zeros(n_out))


class BertQAPredictionHead(_ParamOnly):
    def __init__(self, config, only_vocab=False):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.n_out = config.num_labels
        self.decoder = _Linear(config.hidden_size, config.num_labels, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.num_labels))


class BertPreTrainedModel(PreTrainedModel):
    config_class = BertConfig
    base_model_prefix = "bert"
