"""Small deterministic workload for ncu: a few launches of each hot kernel (no timing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import numpy as np, torch
import densehead as dh
from oracle import synth
SC = [32, 64, 128, 256, 512]
dev = torch.device("cuda", 0)
def boxes(cfg, B, seed, side):
    b, n = synth.config_boxes(cfg, B, seed)
    return torch.from_numpy(b).to(dev), torch.from_numpy(n).to(dev), torch.tensor([[float(side)] * 2] * B, device=dev)
REPS = int(os.environ.get("REPS", "3"))
b, n, d = boxes("retina_coco", 64, 3, 640)
outs, pr = dh.retinanet.format_data_batch(b, n, d, 80, [640, 640])
g = torch.Generator(device=dev); g.manual_seed(1)
pred = []
for o in outs:  # the bench's distribution: regs ~ U(-1, 2), class logits ~ N(-4.595, 1)
    p = torch.empty(o.shape, device=dev)
    p[..., :4].uniform_(-1, 2, generator=g)
    p[..., 4:].normal_(-4.595, 1.0, generator=g)
    pred.append(p)
for _ in range(REPS):
    dh.retinanet.format_data_batch(b, n, d, 80, [640, 640], out=outs, num_pairs=pr)
for _ in range(REPS):
    dh.retinanet.encode_loss_batch(b, n, d, 80, [640, 640], pred)
for _ in range(REPS):
    dh.retinanet.loss_batch(outs, pred)
del outs, pred
b, n, d = boxes("fcos_voc", 256, 1, 512)
outs, cnt = dh.fcos.format_data_batch(b, n, d, 20, [512, 512])
for _ in range(REPS):
    dh.fcos.format_data_batch(b, n, d, 20, [512, 512], out=outs, num_targets=cnt)
b, n, d = boxes("centernet_crowdhuman", 256, 2, 512)
o, st = dh.centernet.format_data_batch(b, n, d, 1, [512, 512], stride=4, mode="s8", box_scales=SC)
for _ in range(REPS):
    dh.centernet.format_data_batch(b, n, d, 1, [512, 512], stride=4, mode="s8", box_scales=SC, out=o, status=st)
torch.cuda.synchronize()
print("profile target done; launches:", dh.launch_count(0))
