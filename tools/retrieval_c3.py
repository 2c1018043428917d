"""BASELINE.json configs[2] at FULL size: COCO-5k-shaped retrieval -- synthetic 5 000 images x 25 000 captions,
uni-modal similarity top-k (128 captions per image, 64 images per caption), then the cross-modal ITM re-rank
of all 640 k + 1.6 M candidate pairs (run_retrieval.py:694-826, 429-522), through retrieval.RetrievalScorer.

    python tools/retrieval_c3.py [--images 5000 --caps-per-img 5]          # 1 GPU
    torchrun --nproc-per-node N tools/retrieval_c3.py                       # pairs / captions / images sharded

Prints one JSON line: per-stage milliseconds (CUDA events, max over ranks), pairs/s of the fine stage and the
achieved TFLOP/s (stage 2 = 9.12 GFLOP per pair)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=5000)
    ap.add_argument("--caps-per-img", type=int, default=5)
    ap.add_argument("--k-i2t", type=int, default=128)
    ap.add_argument("--k-t2i", type=int, default=64)
    ap.add_argument("--pair-batch", type=int, default=2048)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from mvp_pytorch_b200.modeling_vlbert import BiImageBertForRetrieval
    from mvp_pytorch_b200.retrieval import RetrievalScorer, rank_of_first_positive
    W = bench.WORK
    cfg = bench.make_config(0.0)
    cfg.num_labels = 2
    torch.manual_seed(3)
    model = BiImageBertForRetrieval(cfg).to(dev).eval()
    if world > 1:
        dist.broadcast(model.runtime().arena.master, 0)
        model.runtime().arena.refresh_shadow(force=True)
    n_img, n_cap, La, Lt, R = args.images, args.images * args.caps_per_img, 55, 20, 50
    g = torch.Generator(device=dev).manual_seed(2)  # identical inputs on every rank
    caps = dict(input_ids_a=torch.randint(1000, W["only_word"], (n_cap, La), generator=g, device=dev),
                token_type_ids_a=torch.zeros(n_cap, La, dtype=torch.long, device=dev),
                attention_mask_a=torch.ones(n_cap, La, dtype=torch.long, device=dev))
    imgs = dict(input_ids_b=torch.randint(1000, W["only_word"], (n_img, Lt), generator=g, device=dev),
                token_type_ids_b=torch.ones(n_img, Lt, dtype=torch.long, device=dev),
                attention_mask_b=torch.ones(n_img, Lt + R, dtype=torch.long, device=dev),
                img_feats=torch.randn(n_img, R, W["img_dim"], generator=g, device=dev, dtype=torch.bfloat16))
    sc = RetrievalScorer(model, max_tag_length=Lt, stage1_batch=512, pair_batch=args.pair_batch)
    # warm-up on a sliver (allocator pools, tensor-map entry point)
    warm_c = {k: v[:64] for k, v in caps.items()}
    warm_i = {k: v[:16] for k, v in imgs.items()}
    sc.encode(warm_c, warm_i)
    sc.fine(torch.zeros(64 * world, dtype=torch.long, device=dev), torch.zeros(64 * world, dtype=torch.long, device=dev))

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
    (i2t, t2i), ms_coarse = timed(lambda: sc.coarse(args.k_i2t, args.k_t2i))
    img_of = torch.arange(n_img, device=dev).repeat_interleave(i2t.shape[1])
    cap_of = torch.arange(n_cap, device=dev).repeat_interleave(t2i.shape[1])
    p_i2t, ms_f1 = timed(lambda: sc.fine(i2t.reshape(-1), img_of))
    p_t2i, ms_f2 = timed(lambda: sc.fine(cap_of, t2i.reshape(-1)))

    def ranks():
        pos_i = (i2t // args.caps_per_img) == torch.arange(n_img, device=dev)[:, None]
        pos_t = t2i == (torch.arange(n_cap, device=dev) // args.caps_per_img)[:, None]
        return (rank_of_first_positive(p_i2t.view(n_img, -1), pos_i),
                rank_of_first_positive(p_t2i.view(n_cap, -1), pos_t))

    (r_i, r_t), ms_rank = timed(ranks)
    assert torch.isfinite(p_i2t).all() and torch.isfinite(p_t2i).all()
    n_pairs = i2t.numel() + t2i.numel()
    L = La + R
    flops_pair = 6 * (L * 2 * (4 * 768 * 768 + 2 * 768 * 3072) + 4 * L * L * 768) + 2 * 768 * 768
    fine_ms = ms_f1 + ms_f2
    if rank == 0:
        print(json.dumps({
            "workload": f"COCO-5k-shaped retrieval: {n_img} images x {n_cap} captions, top-{i2t.shape[1]} / top-{t2i.shape[1]} "
                        "coarse candidates, cross-modal ITM re-rank of every candidate pair (BASELINE.json configs[2])",
            "n_gpus": world, "pairs": n_pairs, "stage1_encode_ms": ms_enc, "coarse_sim_topk_ms": ms_coarse,
            "fine_i2t_ms": ms_f1, "fine_t2i_ms": ms_f2, "ranks_ms": ms_rank,
            "total_s": (ms_enc + ms_coarse + fine_ms + ms_rank) / 1e3,
            "fine_pairs_per_s": n_pairs / (fine_ms / 1e3),
            "fine_achieved_tflops": n_pairs * flops_pair / (fine_ms / 1e3) / 1e12,
            "i2t_R@1": float((r_i < 1).float().mean()), "t2i_R@1": float((r_t < 1).float().mean()),
            "note": "random-init weights: recall is chance level; the run checks scale (2.24 M pairs), finiteness and timing"}),
            flush=True)
    if world > 1:
        bench._shutdown(world)


if __name__ == "__main__":
    main()
