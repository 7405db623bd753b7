"""Where does the fused encode+loss kernel spend its time?  (development aid; bench.py is the contract)
Times dh_retina_encode_loss on COCO-shaped batches with the real GT boxes, with no boxes at all (pure
streaming pass) and with the shared-memory target-tile kernel (DH_OPT_FUSED_LOSS_KERNEL = 1)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import densehead as dh  # noqa: E402
from oracle import synth  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
boxes, nbox = synth.config_boxes("retina_coco", B, 3)
bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
zero = torch.zeros_like(nd)
dims = torch.tensor([[640., 640.]] * B, device="cuda")
gen = torch.Generator(device="cuda")
gen.manual_seed(5)
pred = []
for h in (80, 40, 20, 10, 5):
    p = torch.empty((B, 9, h, h, 84), device="cuda")
    p[..., :4].uniform_(-1, 2, generator=gen)
    p[..., 4:].normal_(-4.595, 1.0, generator=gen)
    pred.append(p)
nbytes = sum(p.numel() for p in pred) * 4
res = {}
for tag, n, opt in (("real boxes", nd, 0), ("no boxes (stream only)", zero, 0), ("smem-tile kernel", nd, 1)):
    dh.set_option(0, 5, opt)
    t = timeit(lambda: dh.retinanet.encode_loss_batch(bd, n, dims, 80, [640, 640], pred))
    _, tot, pairs = dh.retinanet.encode_loss_batch(bd, n, dims, 80, [640, 640], pred)
    res[tag] = tot.cpu().numpy().tolist()
    print(json.dumps({"case": tag, "B": B, "us": round(t * 1e6, 1), "GBps": round(nbytes / t / 1e9, 1),
                      "total": res[tag], "pairs": int(pairs.sum())}), flush=True)
dh.set_option(0, 5, 0)
for cpc in (2, 3, 4, 6, 8, 12, 16):
    dh.set_option(0, 8, cpc)
    t = timeit(lambda: dh.retinanet.encode_loss_batch(bd, nd, dims, 80, [640, 640], pred))
    print(json.dumps({"chunks_per_cta": cpc, "B": B, "us": round(t * 1e6, 1), "GBps": round(nbytes / t / 1e9, 1)}), flush=True)
dh.set_option(0, 8, 12)
a, b = np.array(res["real boxes"]), np.array(res["smem-tile kernel"])
print("stream+correct vs smem-tile kernel, relative difference:", (np.abs(a - b) / np.maximum(np.abs(b), 1e-30)).tolist())
