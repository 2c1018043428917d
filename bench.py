#!/usr/bin/env python
"""Headline benchmark: MVP base 2-stage pre-training step (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over one synthetic batch of 256 image-text pairs per GPU:
BiBertImgForPreTraining forward (text / visual / cross-modal encoders incl. the hard-negative
pass, MLM + tag-MLM + VSC + ITM + WRA losses) + backward + gradient all-reduce (N>1) + fused AdamW,
bf16 compute with fp32 master weights, dropout 0.1.  Prints ONE JSON line (contract in the task
statement): `value` = pairs/s with inputs resident in HBM, `e2e` = the same through the public
model API with pinned-host inputs copied every step and the six losses read back every step.

`--impl reference` times the reference algorithm's CPU implementation (the oracle port,
oracle/mvptr_oracle.py -- the reference itself is Python and cannot travel to the GPU box) on the
host cores for the same metric on a bounded batch.
"""
import argparse
import json
import math
import os
import statistics
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# grow allocator segments by virtual-memory mapping instead of cudaMalloc/cudaFree (which synchronise):
# the number of masked-LM rows is data dependent, so buffer sizes vary a little from step to step
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
# stdout carries the ONE JSON line and nothing else: keep a private handle on the real stdout for it and point
# descriptor 1 at stderr, so that NCCL's own lines (it logs to stdout: version, "ncclCommInitRank ... nranks N",
# the NVLS / ring / tree channel setup) and anything a library prints land on stderr, in order, for every rank
# (done in _claim_stdout(), only when bench.py is the program: tools and tests import this module)
_JSON_OUT = sys.stdout
os.environ.setdefault("NCCL_DEBUG", "INFO")
os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")

import torch  # noqa: E402


def _claim_stdout():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

WORK = dict(B=256, La=40, Lt=20, R=50, n_phrase=5, H=768, I=3072, layers=6, heads=12, vocab=86051,
            only_word=30522, img_dim=2054, p_drop=0.1, mlm_prob=0.15)


# ------------------------------------------------------------------------------------------
# algorithmic work: multiply-add = 2 FLOPs, training step = 3 x forward
# ------------------------------------------------------------------------------------------
def flops_per_step(B, La, Lt, R, n_mask_txt, n_mask_tag, H=768, I=3072, nl=6, img_dim=2054, vocab=30522):
    f_lin = 2 * (4 * H * H + 2 * H * I)

    def enc(L):
        return nl * (L * f_lin + 4 * L * L * H)

    per_pair = 2 * R * img_dim * H + enc(La) + enc(Lt + R) + 2 * enc(La + R) + 4 * H * H + 2 * 2 * H * H
    heads = (n_mask_txt + n_mask_tag) * (2 * H * H + 2 * H * vocab)
    return 3.0 * (B * per_pair + heads)


def synthetic_batch(seed, B, La, Lt, R, n_phrase, vocab, only_word, img_dim, mlm_prob, img_dtype):
    """All-valid synthetic pre-training batch with the 13 tensors of oscar_tsv4.py:364-377."""
    g = torch.Generator().manual_seed(seed)
    ids_a = torch.randint(1000, only_word, (B, La), generator=g)
    ids_a[:, La - n_phrase:] = torch.randint(only_word, vocab, (B, n_phrase), generator=g)  # phrase concepts
    ids_b = torch.randint(1000, only_word, (B, Lt), generator=g)
    img = torch.randn(B, R, img_dim, generator=g).to(img_dtype)
    lab_a = torch.full((B, La), -1, dtype=torch.long)
    pick = torch.rand(B, La - n_phrase, generator=g) < mlm_prob
    lab_a[:, : La - n_phrase][pick] = torch.randint(1000, only_word, (int(pick.sum()),), generator=g)
    lab_b = torch.full((B, Lt + R), -1, dtype=torch.long)
    pick_b = torch.rand(B, Lt, generator=g) < mlm_prob
    lab_b[:, :Lt][pick_b] = torch.randint(1000, only_word, (int(pick_b.sum()),), generator=g)
    return dict(
        input_ids_a=ids_a, token_type_ids_a=torch.zeros(B, La, dtype=torch.long),
        attention_mask_a=torch.ones(B, La, dtype=torch.long), masked_lm_labels_a=lab_a,
        input_ids_b=ids_b, token_type_ids_b=torch.ones(B, Lt, dtype=torch.long),
        attention_mask_b=torch.ones(B, Lt + R, dtype=torch.long), masked_lm_labels_b=lab_b, img_feats=img,
        phrase_index=torch.tensor([[La - n_phrase, La]] * B), img_index=torch.tensor([[La, La + R]] * B))


# region features as the reference's loaders deliver them: float32 decoded from the TSV's base64 (oscar_tsv4.py:696-727);
# the region-projection input kernel (pad/cast) converts to bf16 on the device.  bf16 / fp16: A/B only.
IMG_DTYPES = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


def h2d_probe_gbps(dev, world, nbytes):
    """Pinned host -> device bandwidth of THIS box (all ranks copy at once, as they do in the end-to-end leg): median of
    three copies of one step's region features; reported next to the end-to-end number it bounds."""
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    ts = []
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        dst.copy_(src, non_blocking=True)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    gbps = nbytes / (sorted(ts)[1] * 1e-3) / 1e9
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([gbps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)  # every rank takes the same decision
        gbps = float(t)
    return gbps



def make_config(drop):
    from mvp_pytorch_b200.modeling_utils import BertConfig
    c = BertConfig(vocab_size_or_config_json_file=WORK["vocab"], hidden_dropout_prob=drop,
                   attention_probs_dropout_prob=drop)
    c.only_word_size, c.qa_answer_size, c.img_feature_dim = WORK["only_word"], 3129, WORK["img_dim"]
    c.img_feature_type, c.use_img_layernorm, c.img_layer_norm_eps = "faster_r-cnn", 1, 1e-12
    c.loss_type, c.num_contrast_classes, c.max_text_seq_length = "sfmx", 2, 35
    return c


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region through NVML (the same
    counters as the recipe's `nvidia-smi --query-gpu=clocks.sm,...` line, B200_PROFILING.md).  An
    in-process NVML thread is used instead of an `nvidia-smi -lms` child: the child's start-up takes
    driver-wide locks for ~0.5 s and measurably stalled kernel launches inside the timed region."""

    def __init__(self, gpu_index):
        import threading
        self.samples, self.active, self.stop_flag = [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # NVML missing: report nothing rather than fail the bench
            self.nv = None
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                         nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                         nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
                except Exception:
                    pass
            time.sleep(0.05)

    def start(self):
        self.active = True

    def stop(self):
        self.active = False
        self.stop_flag = True
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.nv is None or not self.samples:
            return out
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r & bit for _, _, r in self.samples))
        return {"sm_mhz": statistics.median(c for c, _, _ in self.samples), "sm_max_mhz": float(self.max_sm),
                "reasons": reasons, "power_w_max": max(p for _, p, _ in self.samples), "samples": len(self.samples)}


def retrieval_c3(dev, world, rank, n_img=5000, caps_per_img=5, k_i2t=128, k_t2i=64, pair_batch=2048, check_rows=16):
    """Second half of BASELINE.json's metric at FULL size (configs[2]): COCO-5k-shaped retrieval -- 5 000 images x
    25 000 captions (55 tokens; 20 tags + 50 regions), stage 1 once per caption / image, 5k x 25k similarities,
    top-128 captions per image and top-64 images per caption on the GPU, then the cross-modal encoder + ITM head for
    ALL 640 k + 1.6 M candidate pairs (run_retrieval.py:694-826, 429-522) through retrieval.RetrievalScorer, with
    captions, images and pairs sharded contiguously over the ranks.  Times are CUDA events, max over ranks.
    On rank 0 the candidate lists of `check_rows` sampled rows are re-derived on the CPU from the gathered fp32
    embeddings (np.argsort-order top-k of run_retrieval.py:487, :506) and must match at fp32 tie resolution."""
    from mvp_pytorch_b200.modeling_vlbert import BiImageBertForRetrieval
    from mvp_pytorch_b200.retrieval import RetrievalScorer, rank_of_first_positive
    import torch.distributed as dist
    cfg = make_config(0.0)
    cfg.num_labels = 2
    torch.manual_seed(3)
    model = BiImageBertForRetrieval(cfg).to(dev).eval()
    if world > 1:
        dist.broadcast(model.runtime().arena.master, 0)
        model.runtime().arena.refresh_shadow(force=True)
    n_cap, La, Lt, R = n_img * caps_per_img, 55, 20, 50
    g = torch.Generator(device=dev).manual_seed(2)  # identical inputs on every rank
    caps = dict(input_ids_a=torch.randint(1000, WORK["only_word"], (n_cap, La), generator=g, device=dev),
                token_type_ids_a=torch.zeros(n_cap, La, dtype=torch.long, device=dev),
                attention_mask_a=torch.ones(n_cap, La, dtype=torch.long, device=dev))
    imgs = dict(input_ids_b=torch.randint(1000, WORK["only_word"], (n_img, Lt), generator=g, device=dev),
                token_type_ids_b=torch.ones(n_img, Lt, dtype=torch.long, device=dev),
                attention_mask_b=torch.ones(n_img, Lt + R, dtype=torch.long, device=dev),
                img_feats=torch.randn(n_img, R, WORK["img_dim"], generator=g, device=dev, dtype=torch.float32))  # the loader dtype (2 GB at 5 000 images)
    sc = RetrievalScorer(model, max_tag_length=Lt, stage1_batch=512, pair_batch=pair_batch)
    # warm-up on a sliver (allocator pools, tensor-map entry point)
    sc.encode({k: v[:64] for k, v in caps.items()}, {k: v[:16] for k, v in imgs.items()})
    z = torch.zeros(64 * world, dtype=torch.long, device=dev)
    sc.fine(z, z)

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return out, ms

    _, ms_enc = timed(lambda: sc.encode(caps, imgs))
    (i2t, t2i), ms_coarse = timed(lambda: sc.coarse(k_i2t, k_t2i))
    img_of = torch.arange(n_img, device=dev).repeat_interleave(i2t.shape[1])
    cap_of = torch.arange(n_cap, device=dev).repeat_interleave(t2i.shape[1])
    p_i2t, ms_f1 = timed(lambda: sc.fine(i2t.reshape(-1), img_of))
    p_t2i, ms_f2 = timed(lambda: sc.fine(cap_of, t2i.reshape(-1)))

    def ranks():
        pos_i = (i2t // caps_per_img) == torch.arange(n_img, device=dev)[:, None]
        pos_t = t2i == (torch.arange(n_cap, device=dev) // caps_per_img)[:, None]
        return (rank_of_first_positive(p_i2t.view(n_img, -1), pos_i), rank_of_first_positive(p_t2i.view(n_cap, -1), pos_t))

    (r_i, r_t), ms_rank = timed(ranks)
    assert torch.isfinite(p_i2t).all() and torch.isfinite(p_t2i).all()
    checked = None
    if rank == 0 and check_rows:
        gi, gt = sc.global_img.float().cpu(), sc.global_txt.float().cpu()
        rows_i = torch.linspace(0, n_img - 1, check_rows).long()
        rows_c = torch.linspace(0, n_cap - 1, check_rows).long()
        worst, exact = 0.0, []
        for rows, q, c, got in ((rows_i, gi, gt, i2t), (rows_c, gt, gi, t2i)):
            sims = q[rows] @ c.t()
            k = got.shape[1]
            order = torch.sort(sims.flip(-1), dim=-1, descending=True, stable=True)[1]
            ref = (sims.shape[1] - 1 - order)[:, :k]  # descending score, ties -> larger index (np.argsort(x)[::-1])
            g_idx = got[rows.to(got.device)].cpu()
            exact.append(float((g_idx == ref).float().mean()))
            worst = max(worst, float((torch.gather(sims, 1, g_idx) - torch.gather(sims, 1, ref)).abs().max()))
        assert worst <= 1e-5, f"candidate lists differ from the CPU ranking beyond fp32 ties: {worst}"
        checked = {"rows": 2 * check_rows, "identical_positions": min(exact), "worst_score_gap_at_a_difference": worst}
    n_pairs = i2t.numel() + t2i.numel()
    L = La + R
    flops_pair = 6 * (L * 2 * (4 * 768 * 768 + 2 * 768 * 3072) + 4 * L * L * 768) + 2 * 768 * 768
    fine_ms = ms_f1 + ms_f2
    return {"value": n_pairs / (fine_ms / 1e3), "unit": "pairs/s", "pairs": n_pairs, "ms": fine_ms,
            "achieved_tflops": n_pairs * flops_pair / (fine_ms / 1e3) / 1e12,
            "workload": f"COCO-5k-shaped retrieval (BASELINE.json configs[2]): {n_img} images x {n_cap} captions, top-{i2t.shape[1]} / "
                        f"top-{t2i.shape[1]} coarse candidates, cross-modal ITM re-rank of every candidate pair "
                        "(stage 2 only, on cached stage-1 outputs), sharded over the ranks",
            "stage1_encode_ms": ms_enc, "coarse_sim_topk_ms": ms_coarse, "fine_i2t_ms": ms_f1, "fine_t2i_ms": ms_f2,
            "ranks_ms": ms_rank, "total_s": (ms_enc + ms_coarse + fine_ms + ms_rank) / 1e3,
            "end_to_end_pairs_per_s": n_pairs / ((ms_enc + ms_coarse + fine_ms + ms_rank) / 1e3),
            "i2t_R@1": float((r_i < 1).float().mean()), "t2i_R@1": float((r_t < 1).float().mean()),
            "candidate_lists_checked_on_cpu": checked,
            "note": "random-init weights: recall is chance level; timing, scale (2.24 M pairs), finiteness and ranking order are what is measured"}


def cross_modal_encoder_leg(model, dev, B, L, steps=10, warmup=3):
    """north_star's target kernel chain on its own: the cross-modal (stage-2) encoder forward+backward over the
    2B x (text+phrases+regions) joint sequences of the pre-training step -- 6 CaptionBertLayers, dropout 0.1,
    weight gradients reduce-added into the arena.  Timed with CUDA events; algorithmic FLOPs per SURVEY 8d:
    3 x 6 x (M x F_lin + 2B x 4 L^2 H)."""
    from mvp_pytorch_b200 import engine as E
    bert = model.bert
    rt, pf = bert._ctx()
    rt.begin_forward(True)
    sync = getattr(rt, "grad_sync", None)  # a single-GPU kernel measurement: no gradient collectives
    sync_was = sync.enabled if sync is not None else False
    if sync is not None:
        sync.enabled = False
    H, I = rt.H, rt.I
    nl = model.config.num_hidden_layers // 2
    g = torch.Generator(device=dev).manual_seed(11)
    x = torch.randn(2 * B, L, H, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
    gy = (0.01 * torch.randn(2 * B, L, H, device=dev, generator=g)).to(torch.bfloat16)
    maskadd = torch.zeros(2 * B, L, device=dev, dtype=torch.float32)

    def one():
        anchor = rt.anchor(bert.txt_proj)
        y = E.encoder(rt, pf + "mul_encoder", x, maskadd, nl, anchor)
        y.backward(gy)
        x.grad = None

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        one()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    M = 2 * B * L
    fl = 3.0 * nl * (M * 2 * (4 * H * H + 2 * H * I) + 2 * B * 4 * L * L * H)
    model.zero_grad()
    if sync is not None:
        sync.enabled = sync_was
    return ms, fl


def measured_gemm_traffic():
    """DRAM bytes per gemm_kernel launch from the committed ncu pass over one step (profiles/gemm_traffic.json,
    written by tools/summarize_launches.py); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    try:
        d = json.load(open(p))["gemm_kernel"]
        return d["dram_bytes_per_launch"], d["launches"]
    except Exception:  # noqa: BLE001
        return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_pretrain_pairs_per_s(B_cpu, steps, warmup):
    from oracle import mvptr_oracle as O  # the only product-side use of oracle/: reported CPU baseline
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.Cfg()
    sd = {k: v.requires_grad_(k != "qa_head.weight" and k != "qa_head.bias")
          for k, v in O.random_state_dict(cfg, "pretrain", seed=0, bf16_exact=False).items()}
    W = WORK
    b = O.synthetic_batch(cfg, B_cpu, W["La"], W["Lt"], W["R"], seed=2, ragged=False, with_labels=True)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        losses = O.pretrain_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                    b["masked_lm_labels_a"], b["input_ids_b"], b["token_type_ids_b"],
                                    b["attention_mask_b"], b["masked_lm_labels_b"], b["img_feats"],
                                    max_tag_length=W["Lt"], img_index=b["img_index"], phrase_index=b["phrase_index"],
                                    dice_index=b["dice_index"], neg_img=b["neg_img"], rand_pos=b["rand_pos"],
                                    rand_neg=b["rand_neg"])
        losses[0].backward()
        with torch.no_grad():  # plain SGD-free AdamW-equivalent memory pass is negligible on CPU; zero grads
            for v in sd.values():
                v.grad = None
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = statistics.median(times)
    return B_cpu / t, cores, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B_cpu = 8
    v, cores, t = cpu_pretrain_pairs_per_s(B_cpu, args.steps, args.warmup)
    W = WORK
    line = {
        "impl": "reference", "metric": "image-text pairs/sec (pretrain step)", "value": v, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MVP base 2-stage pre-training step (MLM+ITM+VSC+WRA), fwd+bwd, CPU fp32",
                   "batch_per_step": B_cpu, "text_phrase_len": W["La"], "tags": W["Lt"], "regions": W["R"]},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"oracle port of BiBertImgForPreTraining fwd+bwd, batch {B_cpu}, median of {args.steps}"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from mvp_pytorch_b200 import _lib
    from mvp_pytorch_b200.modeling_vlbert import BiBertImgForPreTraining
    from mvp_pytorch_b200.optimization import AdamW

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    W = WORK
    B = args.batch or W["B"]
    torch.manual_seed(1234 + rank)
    p_drop = W["p_drop"] if args.p_drop is None else args.p_drop  # --p-drop: A/B experiments only (with --quick)
    model = BiBertImgForPreTraining(make_config(p_drop)).to(dev).train()
    if world > 1:  # identical replicas: broadcast rank-0 weights through the flat arena
        rt = model.runtime()
        dist.broadcast(rt.arena.master, 0)
    opt = AdamW.for_model(model, lr=1e-4, weight_decay=0.01, max_grad_norm=10.0)
    rt = model.runtime()
    arena = rt.arena
    if world > 1:  # bucketed gradient all-reduce on a side stream, overlapped with backward
        from mvp_pytorch_b200.parallel import allreduce_gradients, enable_overlapped_allreduce
        # default "tail-bf16": overlapped per-layer buckets fp32 in place, the exposed tail as bf16 (parallel.GradientSync);
        # MVPTR_DP_REDUCE=fp32 | bf16 and MVPTR_DP_MIN_BUCKET are A/B knobs
        rd = {"fp32": torch.float32, "bf16": torch.bfloat16}.get(os.environ.get("MVPTR_DP_REDUCE", "tail-bf16"), "tail-bf16")
        enable_overlapped_allreduce(model, reduce_dtype=rd, min_bucket=int(os.environ.get("MVPTR_DP_MIN_BUCKET", str(1 << 16))))

    fp32_feat_bytes = B * W["R"] * W["img_dim"] * 4
    h2d_gbps = h2d_probe_gbps(dev, world, fp32_feat_bytes)
    img_dtype = args.img_dtype
    n_batches = 4
    host = []
    for i in range(n_batches):
        b = synthetic_batch(100 * rank + i, B, W["La"], W["Lt"], W["R"], W["n_phrase"], W["vocab"], W["only_word"],
                            W["img_dim"], W["mlm_prob"], IMG_DTYPES[img_dtype])
        host.append({k: v.pin_memory() for k, v in b.items()})
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    n_mask_txt = float(sum((b["masked_lm_labels_a"] > -1).sum() for b in host)) / n_batches
    n_mask_tag = float(sum((b["masked_lm_labels_b"] > -1).sum() for b in host)) / n_batches
    flops = flops_per_step(B, W["La"], W["Lt"], W["R"], n_mask_txt, n_mask_tag)

    def eager_step(batch):
        model.zero_grad()
        out = model(max_tag_length=W["Lt"], **batch)
        out[0].backward()
        if world > 1:
            allreduce_gradients(model)
        opt.step()
        return out

    # The public fast path: the whole step captured as one CUDA graph (mvp_pytorch_b200/graphs.py).
    # --no-graph (or a failed capture, reported on stderr) falls back to the eager step.
    graphed = None
    if not args.no_graph and not args.profile_step:
        try:
            from mvp_pytorch_b200.graphs import GraphedTrainStep
            for i in range(3):
                eager_step(resident[i % n_batches])
            graphed = GraphedTrainStep(model, opt, resident[0], forward_kwargs=dict(max_tag_length=W["Lt"]),
                                       allreduce=world > 1)
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] CUDA-graph capture failed, using the eager step: {exc!r}", file=sys.stderr)
            graphed = None

    def train_step(batch):
        if graphed is not None:
            return graphed(batch)
        return eager_step(batch)

    def timed(run_one, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = rt.launches
        t0 = time.perf_counter()
        s.record()
        for i in range(steps):
            run_one(i)
        e.record()
        timed.host_ms = (time.perf_counter() - t0) * 1e3 / steps  # enqueue time per step (host side)
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, rt.launches - l0

    if world > 1:
        # leave a few SMs to the NCCL kernels that reduce gradients during backward (see mvptr_gemm_set_max_ctas)
        _lib.set_gemm_max_ctas(int(os.environ.get("MVPTR_GEMM_MAX_CTAS", str(DP_GEMM_CTAS))))
    sampler = ClockSampler(local) if rank == 0 else None  # NVML initialised outside the timed region
    # ---- warm-up (also builds the arena, cuTensorMap entry point, allocator pools)
    for i in range(max(args.warmup, 3)):
        out = train_step(resident[i % n_batches])
    torch.cuda.synchronize()
    assert all(math.isfinite(float(x)) for x in out), "non-finite loss in warm-up"
    n_kernels_per_step = None

    if args.profile_step:  # for ncu --profile-from-start off: exactly one steady-state step
        torch.cuda.profiler.start()
        train_step(resident[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    # ---- (1) device-resident throughput
    if sampler:
        sampler.start()
    ms, launches = timed(lambda i: train_step(resident[i % n_batches]), args.steps)
    host_ms = timed.host_ms
    clocks = sampler.stop() if sampler else None
    # The library counts its own launches (mvptr_launch_count, AdamW and the gradient norm included).  Under a
    # CUDA graph the host does not launch kernels one by one: count one (eager) step -- the graph replays it.
    if graphed is not None:
        l0 = rt.launches
        eager_step(resident[0])
        launches = (rt.launches - l0) * args.steps
        graphed.check_overflow()
    launches_per_step = launches / args.steps

    # ---- (2) end to end through the public API: pinned host batches -> data.PinnedPrefetcher (H2D on a copy
    #          stream, double buffered) -> the training step -> six losses read back into pinned memory, every step
    from mvp_pytorch_b200.data import PinnedPrefetcher
    loss_host = torch.zeros(args.steps, 6, dtype=torch.float32).pin_memory()
    e2e_mode = set(filter(None, (args.e2e_debug or "").split(",")))  # A/B only: "noh2d", "nod2h"

    def run_e2e(steps):
        """Enqueues `steps` end-to-end steps; returns the host->device bytes copied."""
        if "noh2d" in e2e_mode:
            feed, pf = (resident[i % n_batches] for i in range(steps)), None
        else:
            pf = PinnedPrefetcher((host[i % n_batches] for i in range(steps)), dev, depth=2)
            feed = pf
        for i, batch in enumerate(feed):
            out = train_step(batch)
            if "nod2h" not in e2e_mode:
                loss_host[i].copy_(out if torch.is_tensor(out) else torch.stack([x.detach().float() for x in out]),
                                   non_blocking=True)
        return pf.h2d_bytes if pf is not None else 0

    def timed_e2e(steps):
        # untimed warm-up THROUGH the end-to-end path (the W steps above only exercised the resident one): the copy
        # stream, the prefetcher's device buffers (2 x 106 MB out of the caching allocator) and the pinned result
        # buffer's first use cost tens of ms once -- 5.7 ms per step of a 10-step run at N=2 before this
        run_e2e(max(1, min(args.warmup, 4, steps)))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        nbytes = run_e2e(steps)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, nbytes

    if args.quick:  # A/B runs: the resident step only (plus the end-to-end loop when --e2e-debug is given)
        res = {"quick": True, "n_gpus": world, "ms_per_step": ms / args.steps,
               "pairs_per_s": B * world * args.steps / (ms / 1e3)}
        if args.e2e_debug is not None:
            ms_q, _ = timed_e2e(args.steps)
            res["e2e_ms_per_step"], res["e2e_debug"] = ms_q / args.steps, args.e2e_debug
            ms_r, _ = timed(lambda i: train_step(resident[i % n_batches]), args.steps)
            res["resident_again_ms_per_step"] = ms_r / args.steps
        if rank == 0:
            print(json.dumps(res), file=_JSON_OUT, flush=True)
        if graphed is not None:
            torch.cuda.synchronize()
            graphed.release()
        _shutdown(world)
        return

    ms_e2e, e2e_bytes = timed_e2e(args.steps)
    assert torch.isfinite(loss_host).all(), "non-finite loss in the end-to-end run"
    assert e2e_bytes == h2d_bytes * args.steps, (e2e_bytes, h2d_bytes)  # every step's inputs crossed PCIe

    # The captured graph holds NCCL work: release it before anything tears the communicator down
    # (destroying a communicator under a live graph hangs).
    if graphed is not None:
        torch.cuda.synchronize()
        graphed.release()
        graphed_was_used = True
    else:
        graphed_was_used = False

    # ---- (3) one instrumented step: CUDA events around every kernel launch -> per-kernel roofline
    #          (every rank runs the step -- it contains collectives -- only rank 0 records)
    prof = None
    if rank == 0:
        _lib.profile_enable(True)
    eager_step(resident[0])
    torch.cuda.synchronize()
    if rank == 0:
        prof = _lib.profile_collect()
        _lib.profile_enable(False)

    # ---- (3b) the cross-modal encoder alone (north_star's >= 50 % target is quoted on it)
    xenc_ms, xenc_fl = cross_modal_encoder_leg(model, dev, B, W["La"] + W["R"])

    # ---- (4) ITM scoring throughput (the other half of the metric); frees the training state first
    del model, opt, resident
    torch.cuda.empty_cache()
    itm = retrieval_c3(dev, world, rank, n_img=args.retrieval_images)

    if rank != 0:
        _shutdown(world)
        return

    sustained, burst, hbm, how = peaks()
    pairs = B * world * args.steps
    value = pairs / (ms / 1e3)
    e2e = pairs / (ms_e2e / 1e3)
    gemm_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("gemm["))
    gemm_fl = sum(v["work"] for k, v in prof.items() if k.startswith("gemm["))
    gemm_n = sum(v["launches"] for k, v in prof.items() if k.startswith("gemm["))
    total_prof_ms = sum(v["ms"] for v in prof.values())
    achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    top = sorted(((v["ms"], k, v["launches"]) for k, v in prof.items()), reverse=True)[:8]
    traffic, traffic_launches = measured_gemm_traffic()
    line = {
        "metric": "image-text pairs/sec (pretrain step)", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "MVP base 2-stage pre-training step (MLM+tag-MLM+ITM+VSC+WRA), fwd+bwd+AdamW, "
                               "dropout 0.1, BASELINE.json configs[1]",
                   "batch_per_gpu": B, "global_batch": B * world, "text_phrase_len": W["La"], "tags": W["Lt"],
                   "regions": W["R"], "img_dim": W["img_dim"], "parallelism": f"dp{world}",
                   "gradient_allreduce": None if world == 1 else os.environ.get("MVPTR_DP_REDUCE", "tail-bf16") + " (per-layer buckets overlapped with backward; fp32 arena)",
                   "master_weights": "fp32", "cuda_graph": graphed_was_used,
                   "img_feats": img_dtype + " [B, 50, 2054] on the host and in HBM (float32 = the reference loaders' dtype), cast inside the region-"
                                "projection input kernel; pinned host->device link measured %.1f GB/s with all ranks copying" % h2d_gbps, "l2": "inputs larger than L2: per-step working set (~11 GB activations, "
                                                    "2.4 GB weights/grads/moments) >> 126 MB L2, 4 rotating batches"},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                     "frac": achieved / sustained, "traffic": traffic,
                     "traffic_note": (f"DRAM read+write bytes per gemm_kernel launch, ncu pass over the {traffic_launches} "
                                      "GEMM launches of one step (profiles/gemm_traffic.json)") if traffic else None,
                     "kernel": "gemm_kernel (tcgen05/TMEM/TMA), all launches of one step",
                     "launches_per_step": gemm_n, "gemm_ms_per_step": gemm_ms,
                     "gemm_share_of_kernel_time": gemm_ms / total_prof_ms if total_prof_ms else None,
                     "peak_source": how + " bf16_tflops_sustained (kernel timed inside a long step)",
                     "whole_step": {"algorithmic_tflop": flops / 1e12,
                                    "achieved_tflops": flops / (ms / args.steps / 1e3) / 1e12,
                                    "frac_of_sustained": flops / (ms / args.steps / 1e3) / 1e12 / sustained,
                                    "frac_of_burst": flops / (ms / args.steps / 1e3) / 1e12 / burst},
                     "top_kernels_ms": [{"kernel": k, "ms": round(m, 3), "launches": n} for m, k, n in top]},
        "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 24,
                "h2d_probe_gbps": h2d_gbps,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(round(launches_per_step * args.steps)),
        "host_enqueue_ms_per_step": host_ms,
        "clocks": clocks,
    }
    line["itm_scoring"] = itm
    xt = xenc_fl / (xenc_ms / 1e3) / 1e12
    line["cross_modal_encoder"] = {
        "workload": f"cross-modal encoder fwd+bwd, {2 * B} joint sequences x {W['La'] + W['R']} tokens "
                    "(joint + hard-negative pairs of one step), 6 layers, dropout 0.1, eager launches",
        "ms": xenc_ms, "algorithmic_tflop": xenc_fl / 1e12, "achieved_tflops": xt,
        "frac_of_sustained": xt / sustained, "frac_of_burst": xt / burst, "target_frac": 0.5}
    if world == 1 and not args.no_cpu:
        v, cores, t = cpu_pretrain_pairs_per_s(8, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                                "sample": f"oracle port of BiBertImgForPreTraining fwd+bwd (fp32), batch 8, "
                                          f"median of 3 after 1 warm-up ({t:.2f} s/step)"}
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    _shutdown(world)


DP_GEMM_CTAS = 0  # persistent GEMM CTAs in data-parallel runs (0 = all SMs); tuned in profiles/README.md


def _shutdown(world):
    """Leave a multi-rank run without ever blocking the launcher: barrier, then tear the process group down
    on a helper thread and hard-exit if NCCL has not finished within a few seconds (the result line is
    already printed and flushed at this point)."""
    if world <= 1:
        return
    import threading
    import torch.distributed as dist
    torch.cuda.synchronize()
    done = threading.Event()

    def _destroy():
        try:
            dist.barrier()
            dist.destroy_process_group()
        finally:
            done.set()

    threading.Thread(target=_destroy, daemon=True).start()
    if not done.wait(20.0):
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="pairs per GPU (default: the 256 of BASELINE.json)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="time the resident step only (A/B experiments)")
    ap.add_argument("--e2e-debug", default=None,
                    help="with --quick: also time the end-to-end loop; comma list of noh2d / nod2h ('' = the real loop)")
    ap.add_argument("--p-drop", type=float, default=None, help="override dropout (only with --quick; the bench line uses 0.1)")
    ap.add_argument("--img-dtype", default="fp32", choices=["fp32", "bf16", "fp16"],
                    help="dtype of the region features fed to the step (default: float32, what the reference's loaders deliver)")
    ap.add_argument("--no-graph", action="store_true", help="run the eager step instead of the CUDA-graph step")
    ap.add_argument("--retrieval-images", type=int, default=5000,
                    help="images of the configs[2] retrieval leg (x5 captions); 5000 = the full COCO-5k shape")
    ap.add_argument("--profile-step", action="store_true",
                    help="warm up, then run ONE step between cudaProfilerStart/Stop (for ncu) and exit")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.p_drop is not None and not args.quick:
        raise SystemExit("--p-drop is an A/B knob: use it with --quick (the bench line is defined at dropout 0.1)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    run_b200(args)


if __name__ == "__main__":
    _claim_stdout()
    main()
