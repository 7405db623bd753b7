"""RetinaNet detect pipeline (C4 shape, 64 images) under the NMS options, same process, same inputs: slab filter off / on
with several dense-walk thresholds, chain walk serial / parallel rounds; kept indices asserted identical."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import torch  # noqa: E402
import densehead as dh  # noqa: E402
from densehead import _capi, retinanet  # noqa: E402

dev = torch.device("cuda", 0)
B = 64
LEVELS = [80, 40, 20, 10, 5]


def heads(proper):
    gen = torch.Generator(device=dev)
    gen.manual_seed(6)
    out = []
    for h in LEVELS:
        p = torch.empty((B, 9, h, h, 84), device=dev)
        if proper:
            p[..., :2].uniform_(-0.5, 0.5, generator=gen)
            p[..., 2:4].uniform_(0.5, 1.5, generator=gen)
        else:
            p[..., :4].uniform_(-0.5, 1.5, generator=gen)
        p[..., 4:].normal_(-4.595, 2.5, generator=gen)
        out.append(p)
    return out


def timed(hd, reps=10):
    for _ in range(3):
        retinanet.detect_batch(hd, 80, [640, 640], pre_nms_topk=1000)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        r = retinanet.detect_batch(hd, 80, [640, 640], pre_nms_topk=1000)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3, r


for proper in (True, False):
    hd = heads(proper)
    base = None
    for filt, chain in ((0, 1), (0, 0), (1, 0), (16, 0), (24, 0), (32, 0), (48, 0), (64, 0)):
        dh.set_option(0, _capi.DH_OPT_NMS_FILTER, filt)
        dh.set_option(0, _capi.DH_OPT_NMS_CHAIN, chain)
        us, (cand, keep, n_keep) = timed(hd)
        if base is None:
            base = (keep.clone(), n_keep.clone())
        same = torch.equal(n_keep, base[1]) and all(torch.equal(keep[b, :int(n_keep[b])], base[0][b, :int(n_keep[b])]) for b in range(B))
        print(json.dumps({"heads": "proper" if proper else "r01", "filter": filt, "serial_chain": chain, "us": round(us, 1),
                          "kept_mean": float(n_keep.float().mean()), "identical": same}), flush=True)
    dh.set_option(0, _capi.DH_OPT_NMS_FILTER, 1)
    dh.set_option(0, _capi.DH_OPT_NMS_CHAIN, 0)
    del hd
