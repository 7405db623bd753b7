"""ncu target: the small named configs (C1 FCOS VOC batch 8, C2 CenterNet batch 32)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import torch
import densehead as dh
from oracle import synth
dev = torch.device("cuda", 0)
def boxes(cfg, B, seed, side):
    b, n = synth.config_boxes(cfg, B, seed)
    return torch.from_numpy(b).to(dev), torch.from_numpy(n).to(dev), torch.tensor([[float(side)] * 2] * B, device=dev)
b, n, d = boxes("fcos_voc", 8, 1, 512)
outs, cnt = dh.fcos.format_data_batch(b, n, d, 20, [512, 512])
for _ in range(4):
    dh.fcos.format_data_batch(b, n, d, 20, [512, 512], out=outs, num_targets=cnt)
b, n, d = boxes("centernet_crowdhuman", 32, 2, 512)
SC = [32, 64, 128, 256, 512]
o, st = dh.centernet.format_data_batch(b, n, d, 1, [512, 512], stride=4, mode="s8", box_scales=SC)
for _ in range(4):
    dh.centernet.format_data_batch(b, n, d, 1, [512, 512], stride=4, mode="s8", box_scales=SC, out=o, status=st)
torch.cuda.synchronize()
