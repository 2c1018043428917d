"""GPU: the CUDA-graph training step (mvp_pytorch_b200/graphs.py) must walk the same
trajectory as the eager step: same losses step by step on the same batches (dropout 0, device RNG
re-seeded before every step so hard-negative dice and WRA draws coincide), LR schedule followed,
masked-LM capacity overflow reported instead of silently dropping labels."""
import pytest
import torch

from oracle import mvptr_oracle as O
import mvptr_parity_utils as P

pytestmark = pytest.mark.gpu

KEYS = ("input_ids_a", "token_type_ids_a", "attention_mask_a", "masked_lm_labels_a", "input_ids_b",
        "token_type_ids_b", "attention_mask_b", "masked_lm_labels_b", "img_feats", "img_index", "phrase_index")


def _setup(dropout=0.0):
    from mvp_pytorch_b200.optimization import AdamW
    cfg = O.Cfg(vocab_size=1500, only_word_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                intermediate_size=256, max_position_embeddings=64, img_feature_dim=70)
    sd = O.random_state_dict(cfg, "pretrain", seed=3)
    B, La, Lt, R = 8, 14, 5, 9
    batches = []
    for s in range(4):
        b = O.synthetic_batch(cfg, B, La, Lt, R, seed=40 + s, ragged=True, with_labels=True)
        batches.append({k: b[k].cuda() for k in KEYS})
    model = P.build("BiBertImgForPreTraining", cfg, sd, dropout=dropout, train=True, max_text_seq_length=La)
    opt = AdamW.for_model(model, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    return model, opt, batches, Lt


def _eager(model, opt, batch, Lt):
    model.zero_grad()
    out = model(max_tag_length=Lt, **batch)
    out[0].backward()
    opt.step()
    return torch.stack([o.detach().float() for o in out])


def test_graphed_step_follows_eager_trajectory():
    from mvp_pytorch_b200.graphs import GraphedTrainStep
    lrs = [1e-3, 5e-4, 2e-3, 1e-3, 7e-4]
    m1, o1, batches, Lt = _setup()
    eager = []
    for i, lr in enumerate(lrs):
        o1.param_groups[0]["lr"] = lr
        torch.cuda.manual_seed(100 + i)
        eager.append(_eager(m1, o1, batches[i % 4], Lt).cpu())

    m2, o2, batches, Lt = _setup()
    sd0 = {k: v.clone() for k, v in m2.state_dict().items()}
    step = GraphedTrainStep(m2, o2, batches[0], forward_kwargs=dict(max_tag_length=Lt), warmup=2)
    # the capture warm-up trained the model: rewind weights, moments and step count
    m2.load_state_dict(sd0, strict=True)
    o2._m.zero_(); o2._v.zero_(); o2._step = 0
    got = []
    for i, lr in enumerate(lrs):
        o2.param_groups[0]["lr"] = lr
        torch.cuda.manual_seed(100 + i)
        got.append(step(batches[i % 4]).clone().cpu())
    torch.cuda.synchronize()
    step.check_overflow()
    for i, (e, g) in enumerate(zip(eager, got)):
        assert torch.isfinite(g).all()
        # vis_mlm, vsc, mlm do not depend on the device RNG at all; total / itm / wra depend on the
        # re-seeded draws (identical call sequence -> identical philox offsets)
        # (not bit-identical: fp32 atomics / split-K reductions are unordered, and five optimizer steps amplify it)
        P.close(g, e, 1e-2, 1e-2 * (i + 1), f"step {i} losses (graph vs eager)")
    assert o2._step == len(lrs)


def test_graphed_step_dropout_varies_and_overflow_is_loud():
    from mvp_pytorch_b200 import _lib
    from mvp_pytorch_b200.graphs import GraphedTrainStep
    m, o, batches, Lt = _setup(dropout=0.1)
    o.param_groups[0]["lr"] = 0.0  # weights frozen: only the dropout masks can change the loss
    step = GraphedTrainStep(m, o, batches[0], forward_kwargs=dict(max_tag_length=Lt), warmup=1,
                            mlm_capacity=(32, 48))
    a = step(batches[0]).clone()
    b = step(batches[0]).clone()
    torch.cuda.synchronize()
    assert float((a[3] - b[3]).abs()) > 0, "replays must draw fresh dropout masks (MLM loss identical)"
    # overflow: label every position -> more rows than the captured capacity
    full = dict(batches[1])
    full["masked_lm_labels_a"] = torch.full_like(full["masked_lm_labels_a"], 7)
    step(full)
    with pytest.raises(_lib.MvptrError):
        step.check_overflow()


def test_replays_enqueued_ahead_keep_their_own_lr_step_and_dropout_epoch():
    """The host enqueues all replays WITHOUT synchronising (it runs far ahead of the device): replay n must
    still see the learning rate / AdamW step of step n (mvptr_step_params reads a ring slot selected by a
    device-side replay counter) -- the trajectory has to match eager steps taken one by one."""
    from mvp_pytorch_b200.graphs import GraphedTrainStep
    lrs = [2e-3, 1e-4, 3e-3, 5e-4, 2e-3, 1e-3, 4e-3, 2e-4]
    m1, o1, batches, Lt = _setup()
    eager = []
    for i, lr in enumerate(lrs):
        o1.param_groups[0]["lr"] = lr
        torch.cuda.manual_seed(100 + i)
        eager.append(_eager(m1, o1, batches[i % 4], Lt).cpu())
    w_eager = m1.state_dict()["bert.txt_encoder.layer.0.attention.self.query.weight"].float().cpu()

    m2, o2, batches, Lt = _setup()
    sd0 = {k: v.clone() for k, v in m2.state_dict().items()}
    step = GraphedTrainStep(m2, o2, batches[0], forward_kwargs=dict(max_tag_length=Lt), warmup=2)
    m2.load_state_dict(sd0, strict=True)
    o2._m.zero_(); o2._v.zero_(); o2._step = 0
    torch.cuda.synchronize()
    got = []
    for i, lr in enumerate(lrs):  # no host-device synchronisation inside this loop
        o2.param_groups[0]["lr"] = lr
        torch.cuda.manual_seed(100 + i)
        got.append(step(batches[i % 4]).clone())
    torch.cuda.synchronize()
    assert int(step.counter) == step.n == 2 + len(lrs)
    for i, (e, g) in enumerate(zip(eager, got)):
        P.close(g.cpu(), e, 1e-2, 1e-2 * (i + 1), f"step {i} losses (graph, host running ahead, vs eager)")
    w_graph = m2.state_dict()["bert.txt_encoder.layer.0.attention.self.query.weight"].float().cpu()
    # the weights integrate every step's lr: a replay that read a neighbour's lr would show up here
    w0 = sd0["bert.txt_encoder.layer.0.attention.self.query.weight"].float().cpu()
    assert (w_graph - w_eager).abs().mean() < 0.1 * (w_eager - w0).abs().mean()

    # dropout epochs: two replays enqueued back to back (no sync) must draw different masks
    m3, o3, batches, Lt = _setup(dropout=0.1)
    o3.param_groups[0]["lr"] = 0.0
    step3 = GraphedTrainStep(m3, o3, batches[0], forward_kwargs=dict(max_tag_length=Lt), warmup=1)
    a = step3(batches[0]).clone()
    b = step3(batches[0]).clone()
    c = step3(batches[0]).clone()
    torch.cuda.synchronize()
    assert float((a[3] - b[3]).abs()) > 0 and float((b[3] - c[3]).abs()) > 0
