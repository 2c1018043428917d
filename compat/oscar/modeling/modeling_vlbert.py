"""Drop-in for /root/reference/oscar/modeling/modeling_vlbert.py: the same class names, constructor /
forward signatures, output tuples and state_dict keys, computed by hand-written sm_100a CUDA
(mvp_pytorch_b200.modeling_vlbert).  Bound by run_retrieval.py:19, run_pretrain_ml.py:25, run_vqa.py:25,
run_ve.py, run_re.py:28 and modeling_pipeline.py:3."""
from mvp_pytorch_b200.modeling_vlbert import *  # noqa: F401,F403
from mvp_pytorch_b200.modeling_vlbert import (BiBertImgModel, BiBertImgForPreTraining, BiImageBertForRetrieval,  # noqa: F401
                                              BiImageBertForSequenceClassification,
                                              BiImageBertForSequenceClassificationPlus, BiImageBertForVQA,
                                              BiImageBertForRE, BiImageBertRep, BiBertImgForMLM, BertConfig)
