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
    res = bench.retrieval_c3(dev, world, rank, n_img=args.images, caps_per_img=args.caps_per_img, k_i2t=args.k_i2t,
                             k_t2i=args.k_t2i, pair_batch=args.pair_batch)
    if rank == 0:
        res["n_gpus"] = world
        res["fine_pairs_per_s"] = res["value"]
        print(json.dumps(res), flush=True)
    if world > 1:
        bench._shutdown(world)


if __name__ == "__main__":
    main()
