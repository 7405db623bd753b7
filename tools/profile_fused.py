"""ncu target: the fused loss kernel on 64 COCO-shaped images, once without boxes (pure streaming pass) and once with."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import torch
import densehead as dh
from oracle import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
boxes, nbox = synth.config_boxes("retina_coco", B, 3)
bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
dims = torch.tensor([[640., 640.]] * B, device="cuda")
gen = torch.Generator(device="cuda"); gen.manual_seed(5)
pred = []
for h in (80, 40, 20, 10, 5):
    p = torch.empty((B, 9, h, h, 84), device="cuda")
    p[..., :4].uniform_(-1, 2, generator=gen); p[..., 4:].normal_(-4.595, 1.0, generator=gen)
    pred.append(p)
for n in (torch.zeros_like(nd), nd, torch.zeros_like(nd), nd):
    dh.retinanet.encode_loss_batch(bd, n, dims, 80, [640, 640], pred)
torch.cuda.synchronize()
