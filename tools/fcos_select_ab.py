"""A/B timing of dh_fcos_detect's candidate selection (DH_OPT_FCOS_SELECT): 3 = thread-block cluster per long level,
2 = one CTA per (image, level), 4 = estimate + one streaming pass + finish, 1 = score every pair, 0 = the default choice.  C4 shape: batch 64, 640x640, 80 classes, top-1000 per level."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import torch
import densehead as dh
from densehead import fcos
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(6)
for B in [int(b) for b in os.environ.get("B", "64,32,8,1").split(",")]:
    hf = []
    for h in [80, 40, 20, 10, 5]:
        p = torch.empty((B, h, h, 85), device=dev)
        p[..., :4].uniform_(0.5, 6.0, generator=gen)
        p[..., 4:].normal_(-4.595, 2.5, generator=gen)
        hf.append(p)
    ref = None
    for mode in (3, 2, 4, 0):
        dh.set_option(0, 7, mode)
        for _ in range(3):
            r = fcos.detect_batch(hf, 80, [640, 640], pre_nms_topk=1000, with_candidates=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            r = fcos.detect_batch(hf, 80, [640, 640], pre_nms_topk=1000, with_candidates=True)
        e1.record(); torch.cuda.synchronize()
        same = True if ref is None else all(torch.equal(a, b) for a, b in zip(r, ref))
        ref = ref or r
        print("B=%d mode %d: %.1f us per detect_batch, identical to mode 3: %s" % (B, mode, e0.elapsed_time(e1) / n * 1e3, same), flush=True)
    dh.set_option(0, 7, 0)
