#!/usr/bin/env python
"""Multi-GPU checks of the loss-scalar exchange (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

  1. dh_allreduce_loss over the peer mailboxes == the float64 sum of the per-rank vectors, bit-identical on every rank,
     2000 back-to-back calls with changing values (double-buffer / ordering stress), and as CUDA-graph replays;
  2. the same through NCCL (DH_OPT_ALLREDUCE = 1);
  3. ONE 256-image RetinaNet-COCO batch sharded with `shard_batch`: the all-reduced total of the fused encode+loss
     (exchange inside the loss kernel, then as a separate launch, then NCCL) equals the single-GPU total of the whole
     batch (1e-6 relative on the sums, the positive count exactly) and per-image rows equal the single-GPU rows.
Prints one JSON line on rank 0 and exits non-zero on any mismatch.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from densehead import _capi, distributed, retinanet
    from oracle import synth
    info = distributed.init_comm(fuse=False)
    report = {"world": world, "comm": {k: info[k] for k in ("peer", "nccl", "errors")}, "checks": {}}
    ok_all = True

    def allreduce_check(name, iters):
        nonlocal ok_all
        base = torch.arange(4, device=dev, dtype=torch.float32)
        bad = 0
        t0 = time.perf_counter()
        for it in range(iters):
            v = (base + 1.0) * (rank + 1) * 0.37 + it * 0.001
            check = it % 97 == 0 or it == iters - 1
            if check:  # the expected sum from the very inputs: float64 accumulation in rank order, rounded once
                parts = [torch.zeros_like(v) for _ in range(world)]
                dist.all_gather(parts, v)
                want = sum(p_.double() for p_ in parts).float().cpu().numpy()
            distributed.allreduce_losses(v)
            # (the mailboxes add in float64 in rank order: exact; NCCL adds float32 in its own order: one rounding per rank)
            got = v.cpu().numpy()
            if check and not (np.array_equal(got, want) if name != "nccl" else np.allclose(got, want, rtol=1e-6 * world, atol=0)):
                bad += 1
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # graph replays
        v = torch.ones(4, device=dev) * (rank + 1)
        distributed.allreduce_losses(v)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        w = torch.ones(4, device=dev)
        with torch.cuda.graph(g):
            w.fill_(float(rank + 1))
            distributed.allreduce_losses(w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(200):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 200 * 1e3
        good = bad == 0 and bool((w == world * (world + 1) / 2).all().item())
        report["checks"][name] = {"ok": good, "mismatches": bad, "iters": iters, "eager_us_per_call": dt / iters * 1e6,
                                  "graph_us_per_step (fill + all-reduce)": us}
        ok_all = ok_all and good

    if info["peer"]:
        _capi.set_option(local, _capi.DH_OPT_ALLREDUCE, 2)
        allreduce_check("peer_mailboxes", 2000)
    if info["nccl"]:
        _capi.set_option(local, _capi.DH_OPT_ALLREDUCE, 1)
        allreduce_check("nccl", 500)
    _capi.set_option(local, _capi.DH_OPT_ALLREDUCE, 0 if info["peer"] else 1)

    # ---- one global batch, sharded -----------------------------------------------------------------
    side, classes, B = 640, 80, int(os.environ.get("DH_CHECK_BATCH", "256"))
    boxes, nbox = synth.config_boxes("retina_coco", B, synth.seed_for(5, 99))
    gen = torch.Generator(device=dev)
    gen.manual_seed(999)
    pred = []
    for s in (8, 16, 32, 64, 128):
        h = side // s
        p = torch.empty((B, 9, h, h, classes + 4), device=dev)
        p[..., :4].uniform_(-1, 2, generator=gen)
        p[..., 4:].normal_(-4.595, 1.0, generator=gen)
        pred.append(p)
    dims = np.tile(np.array([[side, side]], dtype=np.float32), (B, 1))
    lo, hi = distributed.shard_range(B, rank, world)
    sb, sn, sd = distributed.shard_batch([boxes, nbox, dims], rank, world)
    spred = [p[lo:hi] for p in pred]
    full_pi, full_tot, _ = retinanet.encode_loss_batch(boxes, nbox, dims, classes, [side, side], pred)
    torch.cuda.synchronize()
    modes = []
    if info["peer"]:
        modes += [("fused_in_kernel", True, 2), ("peer_kernel", False, 2)]
    if info["nccl"]:
        modes += [("nccl", False, 1)]
    if not modes:
        modes = [("torch_distributed", False, None)]
    for name, fused, transport in modes:
        if transport is not None:
            _capi.set_option(local, _capi.DH_OPT_ALLREDUCE, transport)
        distributed.set_fused(fused)
        pi, tot, _ = retinanet.encode_loss_batch(sb, sn, sd, classes, [side, side], spred)
        if not fused:
            if transport is None:
                dist.all_reduce(tot)
            else:
                distributed.allreduce_losses(tot)
        torch.cuda.synchronize()
        a, b = tot.double().cpu().numpy(), full_tot.double().cpu().numpy()
        rel = float(np.max(np.abs(a[:3] - b[:3]) / np.maximum(1.0, np.abs(b[:3]))))
        rows = float((pi - full_pi[lo:hi]).abs().max().item() / max(1.0, float(full_pi.abs().max().item())))
        gathered = [torch.zeros_like(tot) for _ in range(world)]
        dist.all_gather(gathered, tot)
        same = all(bool(torch.equal(g_, gathered[0])) for g_ in gathered)
        good = rel <= 1e-6 and a[3] == b[3] and rows <= 1e-6 and same
        report["checks"]["shard_sum_" + name] = {"ok": bool(good), "allreduced": a.tolist(), "single_gpu": b.tolist(), "max_rel_diff": rel,
                                                 "rows_max_rel_diff": rows, "identical_on_all_ranks": same}
        ok_all = ok_all and bool(good)
    distributed.set_fused(False)
    st = _capi.status(local)
    report["status_bits"] = st
    ok_all = ok_all and st == 0
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok_all))
    report["ok"] = all(flags)
    if rank == 0:
        print(json.dumps(report), flush=True)
    distributed.destroy_comm()
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
