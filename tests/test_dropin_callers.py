"""The drop-in seam: the reference's run scripts import their classes from
`oscar.modeling.modeling_vlbert` / `transformers.pytorch_transformers` (run_retrieval.py:19-21,
run_pretrain_ml.py:25-31, run_vqa.py:25-27, run_ve.py, run_re.py:28, modeling_pipeline.py:3).  With `compat/` first
on PYTHONPATH those lines bind the B200 classes and the scripts stay unedited.

* CPU (authoring container only -- needs /root/reference): every in-scope reference script / module IMPORTS
  unchanged under the compat path and the names it binds are this repo's classes; the tokenizer stays the
  reference's own module.  Third-party packages that are absent from this image (deepspeed, tensorboardX,
  jsonlines, boto3, anytree) are stubbed -- they are not on the hot path.
* GPU: tests/dropin_replay.py replays the scripts' call sequences (from_pretrained, grouped AdamW without model=,
  WarmupLinearSchedule, clip_grad_norm_, scheduler.step / optimizer.step / model.zero_grad, save_pretrained,
  .half() + fp16 features, test_coarse / test_fine_i2t loops) through the same import lines and checks the
  3-step trajectory against the oracle.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

_IMPORT_CHECK = r'''
import sys, types
for name in ("boto3", "botocore", "botocore.exceptions", "anytree", "deepspeed", "deepspeed.utils",
             "deepspeed.utils.zero_to_fp32", "tensorboardX", "jsonlines"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["botocore.exceptions"].ClientError = Exception
sys.modules["deepspeed.utils.zero_to_fp32"].get_fp32_state_dict_from_zero_checkpoint = None
sys.modules["tensorboardX"].SummaryWriter = object
import json
import oscar.run_retrieval as rr, oscar.run_vqa as rv, oscar.run_pretrain_ml as rp, oscar.run_ve as ve, oscar.run_re as re_
import oscar.modeling.modeling_pipeline as mp
import oscar.utils.tsv_file as tsv
out = {
  "retrieval": [rr.BiImageBertForRetrieval.__module__, rr.AdamW.__module__, rr.WarmupLinearSchedule.__module__,
                rr.BertConfig.__module__, rr.BertTokenizer.__module__, rr.WEIGHTS_NAME],
  "vqa": [rv.BiImageBertForVQA.__module__, rv.BiImageBertForSequenceClassification.__module__, rv.AdamW.__module__],
  "pretrain": [rp.BiBertImgForPreTraining.__module__, rp.AdamW.__module__, rp.BertConfig.__module__],
  "ve": [ve.BiImageBertForSequenceClassificationPlus.__module__],
  "re": [re_.BiImageBertForRE.__module__],
  "pipeline": [mp.BiImageBertRep.__module__, mp.BiBertImgForMLM.__module__],
  "script_files": [rr.__file__, rv.__file__, rp.__file__, tsv.__file__],
}
print("IMPORTS " + json.dumps(out))
'''


def _env(with_ref):
    env = dict(os.environ)
    paths = [os.path.join(ROOT, "compat"), ROOT] + ([REF] if with_ref else [])
    env["PYTHONPATH"] = os.pathsep.join(paths)
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    return env


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "oscar")), reason="needs the reference checkout (authoring container)")
def test_reference_scripts_import_unchanged_and_bind_the_b200_classes():
    res = subprocess.run([sys.executable, "-B", "-c", _IMPORT_CHECK], env=_env(True), cwd="/tmp", capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("IMPORTS ")][0][8:])
    ours = "mvp_pytorch_b200."
    for key in ("retrieval", "vqa", "pretrain", "ve", "re", "pipeline"):
        for mod in out[key]:
            if mod in ("pytorch_model.bin",):
                continue
            if mod == "transformers.pytorch_transformers.tokenization_bert":
                continue  # the tokenizer is the reference's own (CPU text prep)
            assert mod.startswith(ours), (key, mod)
    assert out["retrieval"][4] == "transformers.pytorch_transformers.tokenization_bert"
    assert all(f.startswith(REF) for f in out["script_files"])  # the scripts themselves are the reference's files, unedited


def test_compat_packages_resolve_without_the_reference():
    """On a box without the reference checkout the compat packages still serve every model / optimizer name."""
    code = ("from oscar.modeling.modeling_vlbert import BiImageBertForRetrieval, BiBertImgForPreTraining, BiImageBertForVQA, "
            "BiImageBertRep, BiBertImgForMLM, BiImageBertForSequenceClassification, BiBertImgModel\n"
            "from transformers.pytorch_transformers import BertConfig, WEIGHTS_NAME, AdamW, WarmupLinearSchedule, WarmupConstantSchedule\n"
            "import transformers.pytorch_transformers as t\n"
            "try:\n    t.BertTokenizer\n    print('TOKENIZER present')\nexcept ImportError as e:\n    print('TOKENIZER absent:', str(e)[:40])\n"
            "print('OK', BiImageBertForRetrieval.__module__, AdamW.__module__)")
    res = subprocess.run([sys.executable, "-B", "-c", code], env=_env(False), cwd="/tmp", capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    assert "OK mvp_pytorch_b200.modeling_vlbert mvp_pytorch_b200.optimization" in res.stdout
    assert "TOKENIZER absent" in res.stdout


@pytest.mark.gpu
def test_reference_call_sequences_through_the_compat_imports():
    import conftest
    res = subprocess.run([sys.executable, "-B", os.path.join(ROOT, "tests", "dropin_replay.py")], env=_env(False),
                         cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("DROPIN_REPLAY ")][0][14:])
    conftest.PARITY_LINES.append("drop-in replay (reference train/eval call sequence via compat imports): " + json.dumps(out))
    for got, ref in zip(out["losses"], out["oracle_losses"]):
        assert abs(got - ref) <= 1e-2 * abs(ref), (out["losses"], out["oracle_losses"])
    for gn, ref in out["grad_norm"]:  # torch.nn.utils.clip_grad_norm_ saw the arena's gradients
        assert abs(gn - ref) <= 2e-2 * ref
    # Adam moves every element by ~lr whatever the size of its gradient, so elements whose gradient is below the
    # bf16 noise floor move in a noise-determined direction: the update vector is compared with the fp32 reference
    # update like every other quantity -- within max(1e-2, 2 x what bf16 storage alone does to it on the same state)
    for err_cuda, err_floor in out["param_update_err_vs_fp32_and_bf16_floor_per_step"]:
        assert err_cuda <= max(1e-2, 2 * err_floor), out
    assert out["half_eval_sims_err"] < 3e-2 and out["half_eval_prob_err"] < 3e-2
