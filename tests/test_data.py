"""Host -> device input staging (mvp_pytorch_b200/data.py): every batch arrives intact and in order,
whatever the mix of pinned / pageable tensors and changing shapes; on the GPU the copies run on a side stream and
slot reuse is ordered by events (checked by letting the consumer lag behind the producer)."""
import pytest
import torch

from mvp_pytorch_b200.data import PinnedPrefetcher


def _batches(n, pin=False, vary=False):
    g = torch.Generator().manual_seed(0)
    out = []
    for i in range(n):
        rows = 4 + (i % 3 if vary else 0)
        b = {"ids": torch.randint(0, 1000, (rows, 7), generator=g), "feats": torch.randn(rows, 5, 6, generator=g),
             "tag": torch.full((rows,), i, dtype=torch.int64)}
        if pin and torch.cuda.is_available():
            b = {k: v.pin_memory() for k, v in b.items()}
        out.append(b)
    return out


def test_prefetcher_cpu_passthrough_keeps_order_and_values():
    src = _batches(5, vary=True)
    got = list(PinnedPrefetcher(src, "cpu"))
    assert len(got) == 5
    for a, b in zip(src, got):
        assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert list(PinnedPrefetcher([], "cpu")) == []


@pytest.mark.gpu
@pytest.mark.parametrize("pin,vary,depth", [(False, False, 2), (True, False, 2), (False, True, 3), (True, True, 2)])
def test_prefetcher_gpu_batches_arrive_intact_while_the_consumer_lags(pin, vary, depth):
    src = _batches(9, pin=pin, vary=vary)
    pf = PinnedPrefetcher(src, "cuda", depth=depth)
    burn = torch.randn(2048, 2048, device="cuda")
    sums = []
    n = 0
    for i, b in enumerate(pf):
        # slow consumer work queued on the compute stream BEFORE the batch is read: a premature overwrite of the
        # slot by a later copy would corrupt what the delayed reads below see
        for _ in range(6):
            burn = burn @ burn * 1e-3
        assert b["ids"].is_cuda and int(b["tag"][0]) == i
        sums.append((b["ids"].sum() + b["tag"].sum(), b["feats"].double().sum()))
        n += 1
    torch.cuda.synchronize()
    assert n == len(src)
    for (si, sf), a in zip(sums, src):
        assert int(si) == int(a["ids"].sum() + a["tag"].sum())
        assert abs(float(sf) - float(a["feats"].double().sum())) < 1e-6
    assert pf.h2d_bytes == sum(v.numel() * v.element_size() for a in src for v in a.values())
