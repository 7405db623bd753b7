import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import numpy as np, torch
import densehead as dh
from densehead import _capi
from oracle import synth
SC = [32, 64, 128, 256, 512]
def probe(tag, fn):
    fn(); torch.cuda.synchronize()
    dh.set_option(0, 4, 1)
    fn(); torch.cuda.synchronize()
    t = _capi.phase_timing(0)
    n = max(t[5], 1)
    print(tag, "tiles(cta0)=%d" % t[5], "cyc/tile:", [round(x / n) for x in t[:4]], "drain", t[4],
          "cta0 ns=%d  => SM clock %.2f GHz, %.0f ns/tile" % (t[6], sum(t[:5]) / max(t[6], 1), t[6] / n), flush=True)
    dh.set_option(0, 4, 0)
for tb, cps in ((16384, 4), (28672, 4), (49152, 2)):
    dh.set_option(0, 2, tb); dh.set_option(0, 3, cps)
    B = 64
    boxes, nbox = synth.config_boxes("retina_coco", B, 3)
    bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
    dims = torch.tensor([[640., 640.]] * B, device="cuda")
    outs, pr = dh.retinanet.format_data_batch(bd, nd, dims, 80, [640, 640])
    probe("retina tile=%d ctas=%d" % (tb, cps), lambda: dh.retinanet.format_data_batch(bd, nd, dims, 80, [640, 640], out=outs, num_pairs=pr))
    nd0 = torch.zeros_like(nd)
    probe("retina NO BOXES tile=%d ctas=%d" % (tb, cps), lambda: dh.retinanet.format_data_batch(bd, nd0, dims, 80, [640, 640], out=outs, num_pairs=pr))
    B = 256
    boxes, nbox = synth.config_boxes("fcos_voc", B, 1)
    bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
    dims = torch.tensor([[512., 512.]] * B, device="cuda")
    outs, cnt = dh.fcos.format_data_batch(bd, nd, dims, 20, [512, 512])
    probe("fcos tile=%d ctas=%d" % (tb, cps), lambda: dh.fcos.format_data_batch(bd, nd, dims, 20, [512, 512], out=outs, num_targets=cnt))
    boxes, nbox = synth.config_boxes("centernet_crowdhuman", B, 2)
    bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
    out, st = dh.centernet.format_data_batch(bd, nd, dims, 1, [512, 512], stride=4, mode="s8", box_scales=SC)
    probe("centernet tile=%d ctas=%d" % (tb, cps), lambda: dh.centernet.format_data_batch(bd, nd, dims, 1, [512, 512], stride=4, mode="s8", box_scales=SC, out=out, status=st))
