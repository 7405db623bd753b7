"""N > 1 host logic on CPU: two gloo ranks shard a batch by image, compute their shard's loss scalars (with the
oracle, since there is no GPU here), all-reduce them, and must reproduce the single-process totals."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import PKG, ROOT  # noqa: E402


def _worker(rank, world, port, boxes, nbox, seed, out_path):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from densehead import distributed as D
    from oracle import dense_head_ref as O
    from oracle import synth
    my_boxes, my_nbox = D.shard_batch([boxes, nbox], rank, world)
    lo, hi = D.shard_range(len(nbox), rank, world)
    pred = synth.fcos_predictions(len(nbox), 256, 6, seed)
    total = torch.zeros(4, dtype=torch.float64)
    for i, b in enumerate(range(lo, hi)):
        tg, _ = O.fcos_format_data(my_boxes[i, :my_nbox[i]], [256, 256], 6)
        cls, reg, cen = O.fcos_model_loss(tg, [p[b] for p in pred])
        npos = sum(int((t[..., 5:].max(-1) >= 1).sum()) for t in tg)
        total += torch.tensor([float(cls), float(reg), float(cen), float(npos)], dtype=torch.float64)
    D.allreduce_losses(total)
    counts = D.gather_counts(hi - lo)
    if rank == 0:
        np.save(out_path, np.concatenate([total.numpy(), np.array(counts, dtype=np.float64)]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loss_allreduce_matches_single_process(tmp_path):
    from oracle import dense_head_ref as O
    from oracle import synth
    boxes, nbox = synth.make_boxes(5, 256, 8, 6, 8.0, 150.0, 31)   # 5 images over 2 ranks: ragged shards (3 + 2)
    seed = 99
    out = str(tmp_path / "tot.npy")
    mp.spawn(_worker, args=(2, 29533, boxes, nbox, seed, out), nprocs=2, join=True)
    got = np.load(out)
    pred = synth.fcos_predictions(5, 256, 6, seed)
    want = np.zeros(4)
    for b in range(5):
        tg, _ = O.fcos_format_data(boxes[b, :nbox[b]], [256, 256], 6)
        cls, reg, cen = O.fcos_model_loss(tg, [p[b] for p in pred])
        want += [float(cls), float(reg), float(cen), sum(int((t[..., 5:].max(-1) >= 1).sum()) for t in tg)]
    assert np.allclose(got[:4], want, rtol=1e-12)
    assert got[4:].tolist() == [3.0, 2.0]


def test_shard_range_is_a_partition():
    from densehead.distributed import shard_range
    for n in (0, 1, 7, 256, 257):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
