"""ncu target: dh_fcos_detect on the C4 shape (batch B, 640x640, 80 classes, top-1000 per level), one warm-up and one
measured call per DH_OPT_FCOS_SELECT mode listed in MODES."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import torch
import densehead as dh
from densehead import fcos
dev = torch.device("cuda", 0)
B = int(os.environ.get("B", "64"))
gen = torch.Generator(device=dev); gen.manual_seed(6)
hf = []
for h in [80, 40, 20, 10, 5]:
    p = torch.empty((B, h, h, 85), device=dev)
    p[..., :4].uniform_(0.5, 6.0, generator=gen)
    p[..., 4:].normal_(-4.595, 2.5, generator=gen)
    hf.append(p)
for mode in [int(m) for m in os.environ.get("MODES", "0,2").split(",")]:
    dh.set_option(0, 7, mode)
    for _ in range(2):
        r = fcos.detect_batch(hf, 80, [640, 640], pre_nms_topk=1000)
    torch.cuda.synchronize()
    print("mode", mode, "valid", r[3][:4].tolist())
dh.set_option(0, 7, 0)
