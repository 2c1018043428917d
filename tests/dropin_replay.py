"""Replay of the reference scripts' call sequences against the drop-in, bound through the SAME import lines
the scripts use (PYTHONPATH=compat:<repo>; see compat/README.md).  Run as a subprocess by
tests/test_dropin_callers.py so that `transformers` resolves to the compat package, never to site-packages.

Mirrors (call for call, not copied): run_retrieval.py:1037-1049 (config + from_pretrained + half), :560-640
(grouped AdamW, WarmupLinearSchedule, train loop with clip_grad_norm_ / scheduler.step / optimizer.step /
model.zero_grad), :694-741 (test_coarse), :788-826 (test_fine_i2t), prepare_inputs (fp16 features under
--half_evaluation).  Prints one JSON line with the observed numbers.
"""
import json
import os
import sys
import tempfile

import torch
import torch.nn as nn

# ---- the reference's import lines (run_retrieval.py:19-21), unchanged ------------------------------------
from oscar.modeling.modeling_vlbert import BiImageBertForRetrieval
from transformers.pytorch_transformers import BertConfig, WEIGHTS_NAME
from transformers.pytorch_transformers import AdamW, WarmupLinearSchedule, WarmupConstantSchedule  # noqa: F401

from oracle import mvptr_oracle as O  # checker only

TINY = dict(vocab_size=1500, only_word_size=1000, hidden_size=128, num_hidden_layers=4, num_attention_heads=2,
            intermediate_size=256, max_position_embeddings=64, img_feature_dim=70, qa_answer_size=37, num_labels=2)


def main():
    dev = torch.device("cuda")
    cfg = O.Cfg(**TINY)
    sd = O.random_state_dict(cfg, "retrieval", seed=2)
    B, La, Lt, R = 6, 12, 5, 9
    cpu_b = O.synthetic_batch(cfg, B, La, Lt, R, seed=12, ragged=True)
    out = {}
    with tempfile.TemporaryDirectory() as ckpt:
        # a checkpoint directory in the reference's format: config.json + pytorch_model.bin with the reference's keys
        c0 = BertConfig(vocab_size_or_config_json_file=cfg.vocab_size, hidden_size=cfg.hidden_size,
                        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                        intermediate_size=cfg.intermediate_size, max_position_embeddings=cfg.max_position_embeddings)
        c0.only_word_size, c0.qa_answer_size = cfg.only_word_size, cfg.qa_answer_size
        c0.to_json_file(os.path.join(ckpt, "config.json"))
        torch.save(sd, os.path.join(ckpt, WEIGHTS_NAME))

        # run_retrieval.py:1024-1038
        config = BertConfig.from_pretrained(ckpt, num_labels=2, finetuning_task="ir")
        config.img_feature_dim, config.img_feature_type = cfg.img_feature_dim, "frcnn"
        config.hidden_dropout_prob = 0.0               # args.drop_out (run_retrieval.py:1033)
        config.attention_probs_dropout_prob = 0.0      # (the script leaves 0.1; 0 here so that the oracle can follow)
        config.loss_type, config.img_layer_norm_eps, config.use_img_layernorm = "sfmx", 1e-12, 1
        model = BiImageBertForRetrieval.from_pretrained(ckpt, from_tf=bool(".ckpt" in ckpt), config=config)
        model.to(dev)

        # run_retrieval.py:560-575
        no_decay = ["bias", "LayerNorm.weight"]
        grouped_parameters = [
            {"params": [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": 0.05},
            {"params": [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
        optimizer = AdamW(grouped_parameters, lr=2e-4, eps=1e-8)
        scheduler = WarmupLinearSchedule(optimizer, warmup_steps=1, t_total=4)

        # Per-step parity: before every step the oracle takes the CUDA model's CURRENT weights and the optimizer's
        # CURRENT moments (optimizer.state_dict(), the reference's exp_avg / exp_avg_sq layout), computes the same
        # step on the CPU (bf16-store forward/backward, clip, per-tensor AdamW of optimization.py:130-189) and the
        # results are compared -- so Adam's amplification of near-zero gradients cannot accumulate across steps.
        dices = [torch.randperm(B, generator=torch.Generator().manual_seed(100 + i)) for i in range(3)]
        import mvp_pytorch_b200.engine as E
        orig_rp, orig_hn = torch.randperm, E.hard_negatives
        losses, ref_losses, upd = [], [], []
        names = [n for n, _ in model.named_parameters()]
        order = [n for n in names if not any(nd in n for nd in no_decay)] + [n for n in names if any(nd in n for nd in no_decay)]
        model.zero_grad()
        for step in range(3):
            ref = {k: v.detach().cpu().clone() for k, v in model.named_parameters()}
            st = optimizer.state_dict()["state"]
            mom = {k: ((st[i]["exp_avg"].cpu().clone(), st[i]["exp_avg_sq"].cpu().clone()) if i in st
                       else (torch.zeros_like(ref[k]), torch.zeros_like(ref[k]))) for i, k in enumerate(order)}
            before = {k: v.clone() for k, v in ref.items()}
            # ---- oracle steps on the same state: (a) plain fp32 = the reference's arithmetic, (b) with the CUDA path's
            #      bf16 stores AND its bf16 compute copy of the (no longer bf16-exact) master weights ----
            rb = lambda t: t.to(torch.bfloat16).to(torch.float32)
            leaf32 = {k: v.clone().requires_grad_(True) for k, v in ref.items()}
            r32, _, _, _, _ = O.retrieval_train_forward(
                leaf32, cfg, cpu_b["input_ids_a"], cpu_b["token_type_ids_a"], cpu_b["attention_mask_a"],
                cpu_b["input_ids_b"], cpu_b["token_type_ids_b"], cpu_b["attention_mask_b"], cpu_b["img_feats"],
                max_tag_length=Lt, dice_index=dices[step])
            r32.backward()
            leaf = {k: (rb(v) if k != "logit_scale" else v.clone()).requires_grad_(True) for k, v in ref.items()}
            with O.bf16_stores():
                r_total, _, _, _, _ = O.retrieval_train_forward(
                    leaf, cfg, cpu_b["input_ids_a"], cpu_b["token_type_ids_a"], cpu_b["attention_mask_a"],
                    cpu_b["input_ids_b"], cpu_b["token_type_ids_b"], cpu_b["attention_mask_b"], cpu_b["img_feats"],
                    max_tag_length=Lt, dice_index=dices[step])
                r_total.backward()
            with torch.no_grad():
                txt, vis, _, _ = O.stage1(ref, cfg, *[cpu_b[k] for k in ("input_ids_a", "token_type_ids_a", "attention_mask_a",
                                                                            "input_ids_b", "token_type_ids_b", "attention_mask_b",
                                                                            "img_feats")])
                gt, gi = O.global_embeddings(ref, txt, vis)
                h_img, h_txt = O.hard_negative_indexes(gt @ gi.t())
            # ---- the reference's train-loop body, run_retrieval.py:597-640 ----
            model.train()
            model.forward_mod = "train"
            batch = tuple(cpu_b[k].to(dev) for k in ("input_ids_a", "attention_mask_a", "token_type_ids_a", "input_ids_b",
                                                     "attention_mask_b", "token_type_ids_b", "img_feats"))
            inputs = {"input_ids_a": batch[0], "attention_mask_a": batch[1], "token_type_ids_a": batch[2],
                      "input_ids_b": batch[3], "attention_mask_b": batch[4], "token_type_ids_b": batch[5],
                      "img_feats": batch[6], "max_tag_length": Lt}
            try:
                torch.randperm = lambda n, **kw: dices[step].to(kw.get("device", "cpu"))
                E.hard_negatives = lambda rt, sim: (h_img.to(sim.device), h_txt.to(sim.device))
                loss, logits, r_loss, f_loss, pseudo_labels = model(**inputs)
            finally:
                torch.randperm, E.hard_negatives = orig_rp, orig_hn
            loss.backward()
            gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            scheduler.step()
            optimizer.step()
            model.zero_grad()
            losses.append(float(loss.detach()))
            ref_losses.append(float(r_total.detach()))
            # oracle updates: clip (max_norm 1.0) + per-tensor AdamW with the schedule's lr, from both gradient sets
            lr = optimizer.param_groups[0]["lr"]

            def updated(leaves):
                new = {k: v.clone() for k, v in ref.items()}
                total_norm = torch.sqrt(sum((v.grad ** 2).sum() for v in leaves.values() if v.grad is not None))
                coef = min(1.0, 1.0 / (float(total_norm) + 1e-6))
                for k, v in leaves.items():
                    if v.grad is None:
                        continue
                    wd = 0.0 if any(nd in k for nd in no_decay) else 0.05
                    O.adamw_step(new[k], v.grad * coef, mom[k][0].clone(), mom[k][1].clone(), step + 1, lr, eps=1e-8,
                                 weight_decay=wd)
                return new, float(total_norm)

            w32, _ = updated(leaf32)
            w16, total_norm = updated(leaf)
            out.setdefault("grad_norm", []).append([float(gn), total_norm])
            # the whole update vector of this step against the fp32 reference update (global relative L2): the CUDA
            # path's, and -- the storage floor -- the bf16-store oracle's
            den = sum(float((before[k] - w32[k]).norm()) ** 2 for k in ref) ** 0.5
            err_cuda = sum(float((p.detach().cpu() - w32[k]).norm()) ** 2 for k, p in model.named_parameters()) ** 0.5 / den
            err_floor = sum(float((w16[k] - w32[k]).norm()) ** 2 for k in ref) ** 0.5 / den
            upd.append([err_cuda, err_floor])
        out["losses"], out["oracle_losses"] = losses, ref_losses
        out["param_update_err_vs_fp32_and_bf16_floor_per_step"] = upd
        ref = {k: v.detach().cpu().clone() for k, v in model.named_parameters()}

        # save_pretrained / from_pretrained round trip + --half_evaluation (run_retrieval.py:1040-1049)
        model.save_pretrained(ckpt)
        config = BertConfig.from_pretrained(ckpt)
        model2 = BiImageBertForRetrieval.from_pretrained(ckpt, config=config)
        model2 = model2.half()
        dtype = torch.float16
        model2.to(dev)
        model2.eval()
        # test_coarse (:694-741) with prepare_inputs' cast of the float inputs to args.dtype
        model2.forward_mod = "coarse"
        with torch.no_grad():
            inputs = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in inputs.items() if torch.is_tensor(v)}
            inputs["max_tag_length"] = Lt
            global_txt, global_img = model2(**inputs)[:2]
            full_sims = global_img @ global_txt.t()
            # test_fine_i2t (:788-826)
            model2.forward_mod = "fine"
            logits = model2(**inputs)
            probs = nn.Softmax(dim=1)(logits)[:, 1]
            o_gt, o_gi = O.forward_single(ref, cfg, **cpu_b)
            o_fine = O.retrieval_fine_forward(ref, cfg, cpu_b["input_ids_a"], cpu_b["token_type_ids_a"],
                                              cpu_b["attention_mask_a"], max_tag_length=Lt,
                                              input_ids_b=cpu_b["input_ids_b"], token_type_ids_b=cpu_b["token_type_ids_b"],
                                              attention_mask_b=cpu_b["attention_mask_b"], img_feats=cpu_b["img_feats"])
        out["half_eval_sims_err"] = float((full_sims.float().cpu() - o_gi @ o_gt.t()).abs().max())
        out["half_eval_prob_err"] = float((probs.float().cpu() - torch.softmax(o_fine, 1)[:, 1]).abs().max())
    print("DROPIN_REPLAY " + json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
