"""Per-CTA time line of the fused encode+loss kernel (development aid): dh_set_trace makes every CTA record when it
started, when its first chunk was staged, when its chunk loop ended and how many chunks it processed (globaltimer).
Prints, per case: launch ramp (spread of the start stamps), prologue, spread of the end stamps (the tail), time the
last CTA spent in the in-kernel reduction, and chunks per CTA."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import densehead as dh  # noqa: E402
from densehead import _capi  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    batches = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "16,32,64,256").split(",")]
    buf = torch.zeros(12 * 2048 + 8, dtype=torch.int64, device="cuda")
    h = _capi.handle(0)
    for B in batches:
        boxes, nbox = synth.config_boxes("retina_coco", B, 3)
        bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
        dims = torch.tensor([[640., 640.]] * B, device="cuda")
        gen = torch.Generator(device="cuda")
        gen.manual_seed(5)
        pred = []
        for hh in (80, 40, 20, 10, 5):
            p = torch.empty((B, 9, hh, hh, 84), device="cuda")
            p[..., :4].uniform_(-1, 2, generator=gen)
            p[..., 4:].normal_(-4.595, 1.0, generator=gen)
            pred.append(p)
        for tag, n, tail in (("boxes tail=1", nd, 1), ("boxes tail=0", nd, 0), ("no boxes tail=1", torch.zeros_like(nd), 1)):
            dh.set_option(0, _capi.DH_OPT_FUSED_TAIL, tail)
            for _ in range(3):
                dh.retinanet.encode_loss_batch(bd, n, dims, 80, [640, 640], pred)
            torch.cuda.synchronize()
            _capi.check(_capi.lib().dh_set_trace(h, buf.data_ptr(), buf.numel() * 8), "dh_set_trace")
            buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dh.retinanet.encode_loss_batch(bd, n, dims, 80, [640, 640], pred)
            e1.record()
            torch.cuda.synchronize()
            _capi.check(_capi.lib().dh_set_trace(h, None, 0), "dh_set_trace")
            t = buf.cpu().numpy()
            grid = int(np.count_nonzero(t[0:12 * 2048:12])) - 1  # (the stamp of the in-kernel reduction sits right behind the last CTA row)
            rows = t[:12 * grid].reshape(grid, 12)
            t0 = rows[:, 0].min()
            start, ready, end, chunks = rows[:, 0] - t0, rows[:, 1] - t0, rows[:, 2] - t0, rows[:, 3]
            fin = t[12 * grid] - t0
            phases = rows[:, 4:10].mean(axis=0) / 1e3
            q = lambda v, p_: float(np.percentile(v, p_)) / 1e3  # noqa: E731
            print(json.dumps({
                "B": B, "case": tag, "grid": grid, "event_us": round(e0.elapsed_time(e1) * 1e3, 1),
                "start_us p50/p99/max": [round(q(start, 50), 1), round(q(start, 99), 1), round(q(start, 100), 1)],
                "prologue_us p50/max": [round(q(ready - start, 50), 1), round(q(ready - start, 100), 1)],
                "end_us p1/p10/p50/p90/max": [round(q(end, x), 1) for x in (1, 10, 50, 90, 100)],
                "finalize_done_us": round(float(fin) / 1e3, 1), "chunks min/mean/max": [int(chunks.min()), round(float(chunks.mean()), 2), int(chunks.max())],
                "busy_frac (mean end / max end)": round(float(end.mean() / end.max()), 3),
                "mean us per CTA in {stage, mark, stream, barrier, visit, chunk end}": [round(float(v), 1) for v in phases]}), flush=True)
        dh.set_option(0, _capi.DH_OPT_FUSED_TAIL, 1)
        del pred
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
