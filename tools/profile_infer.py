"""C4 workload for ncu: one warm-up + one measured detect_batch call for the FCOS and RetinaNet heads (batch 64)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import torch
import densehead as dh
from densehead import fcos, retinanet
dev = torch.device("cuda", 0)
B = int(os.environ.get("B", "64"))
gen = torch.Generator(device=dev); gen.manual_seed(6)
LEVELS = [80, 40, 20, 10, 5]
def heads(shape_fn, lo, hi):
    out = []
    for h in LEVELS:
        p = torch.empty(shape_fn(h), device=dev)
        p[..., :4].uniform_(lo, hi, generator=gen)
        p[..., 4:].normal_(-4.595, 2.5, generator=gen)
        out.append(p)
    return out
hf = heads(lambda h: (B, h, h, 85), 0.5, 6.0)
for _ in range(2):
    r = fcos.detect_batch(hf, 80, [640, 640], pre_nms_topk=1000)
torch.cuda.synchronize(); print("fcos valid", r[3][:8].tolist())
del hf
hr = heads(lambda h: (B, 9, h, h, 84), -0.5, 0.5)   # centre offsets; sizes below: proper boxes, as a trained head emits
for p in hr:
    p[..., 2:4].uniform_(0.5, 1.5, generator=gen)
for _ in range(2):
    c, k, n = retinanet.detect_batch(hr, 80, [640, 640], pre_nms_topk=1000)
torch.cuda.synchronize(); print("retina kept", n[:8].tolist())
