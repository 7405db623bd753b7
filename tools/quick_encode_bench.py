"""Quick device-side timing of the encoders (development aid; bench.py is the contract)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import numpy as np, torch
import densehead as dh
from oracle import synth

def timeit(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3

def run(tag, fn, nbytes):
    t = timeit(fn)
    print(json.dumps({"case": tag, "us": round(t * 1e6, 2), "GBps": round(nbytes / t / 1e9, 1)}), flush=True)

SC = [32, 64, 128, 256, 512]
for mc in (2,):
    dh.set_option(0, 9, mc)
    for cfg, B, side, C in (("fcos_voc", 8, 512, 20), ("fcos_voc", 32, 512, 20), ("fcos_voc", 256, 512, 20)):
        boxes, nbox = synth.config_boxes(cfg, B, 1)
        bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
        dims = torch.tensor([[float(side)] * 2] * B, device="cuda")
        outs, cnt = dh.fcos.format_data_batch(bd, nd, dims, C, [side, side])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            dh.fcos.format_data_batch(bd, nd, dims, C, [side, side], out=outs, num_targets=cnt)
        run("min_chunk=%d fcos B=%d (graph)" % (mc, B), g.replay, sum(o.numel() for o in outs) * 4)
    for B in (32, 256):
        boxes, nbox = synth.config_boxes("centernet_crowdhuman", B, 2)
        bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
        dims = torch.tensor([[512., 512.]] * B, device="cuda")
        out, st = dh.centernet.format_data_batch(bd, nd, dims, 1, [512, 512], stride=4, mode="s8", box_scales=SC)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            dh.centernet.format_data_batch(bd, nd, dims, 1, [512, 512], stride=4, mode="s8", box_scales=SC, out=out, status=st)
        run("min_chunk=%d centernet B=%d (graph)" % (mc, B), g.replay, out.numel() * 4)
dh.set_option(0, 9, 2)
sys.exit(0)
SC = [32, 64, 128, 256, 512]
for tma in (1, 0):
    dh.set_option(0, 1, tma)
    for B in (8, 256):
        boxes, nbox = synth.config_boxes("fcos_voc", B, 1)
        bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
        dims = torch.tensor([[512., 512.]] * B, device="cuda")
        outs, cnt = dh.fcos.format_data_batch(bd, nd, dims, 20, [512, 512])
        run("fcos_voc B=%d tma=%d" % (B, tma), lambda: dh.fcos.format_data_batch(bd, nd, dims, 20, [512, 512], out=outs, num_targets=cnt), sum(o.numel() for o in outs) * 4)
    for B in (32, 256):
        boxes, nbox = synth.config_boxes("centernet_crowdhuman", B, 2)
        bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
        dims = torch.tensor([[512., 512.]] * B, device="cuda")
        out, st = dh.centernet.format_data_batch(bd, nd, dims, 1, [512, 512], stride=4, mode="s8", box_scales=SC)
        run("centernet_s8 B=%d tma=%d" % (B, tma), lambda: dh.centernet.format_data_batch(bd, nd, dims, 1, [512, 512], stride=4, mode="s8", box_scales=SC, out=out, status=st), out.numel() * 4)
    for B in (8, 64):
        boxes, nbox = synth.config_boxes("retina_coco", B, 3)
        bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
        dims = torch.tensor([[640., 640.]] * B, device="cuda")
        outs, pr = dh.retinanet.format_data_batch(bd, nd, dims, 80, [640, 640])
        run("retina_coco B=%d tma=%d" % (B, tma), lambda: dh.retinanet.format_data_batch(bd, nd, dims, 80, [640, 640], out=outs, num_pairs=pr), sum(o.numel() for o in outs) * 4)
        if B == 64:
            big = torch.empty(sum(o.numel() for o in outs), device="cuda")
            run("cudaMemset same bytes", lambda: big.zero_(), big.numel() * 4)
            src = torch.empty_like(big)
            run("copy same bytes (r+w)", lambda: big.copy_(src), big.numel() * 8)
dh.set_option(0, 1, 1)
for tb in (16384, 24576, 28672, 32768, 40960, 49152):
    for cps in (2, 3, 4):
        dh.set_option(0, 2, tb); dh.set_option(0, 3, cps)
        try:
            run("retina B=64 tile=%d ctas=%d" % (tb, cps), lambda: dh.retinanet.format_data_batch(bd, nd, dims, 80, [640, 640], out=outs, num_pairs=pr), sum(o.numel() for o in outs) * 4)
        except Exception as e:
            print("skip", tb, cps, e)

# ---- losses
dh.set_option(0, 2, 28672); dh.set_option(0, 3, 4)
B = 64
boxes, nbox = synth.config_boxes("retina_coco", B, 3)
bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
dims = torch.tensor([[640., 640.]] * B, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(1)
pred = [torch.randn((B, 9, h, h, 84), device="cuda", generator=g) - 4.0 for h in (80, 40, 20, 10, 5)]
nb = sum(p.numel() for p in pred) * 4
run("retina fused encode+loss B=64", lambda: dh.retinanet.encode_loss_batch(bd, nd, dims, 80, [640, 640], pred), nb)
lab, _ = dh.retinanet.format_data_batch(bd, nd, dims, 80, [640, 640])
run("retina unfused loss B=64 (2x bytes)", lambda: dh.retinanet.loss_batch(lab, pred), 2 * nb)
for tb in (8192, 16384, 32768, 65536):
    dh.set_option(0, 2, tb)
    run("retina fused tile=%d" % (tb // 2), lambda: dh.retinanet.encode_loss_batch(bd, nd, dims, 80, [640, 640], pred), nb)
dh.set_option(0, 2, 28672)
B = 256
boxes, nbox = synth.config_boxes("fcos_voc", B, 1)
bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
dims = torch.tensor([[512., 512.]] * B, device="cuda")
pred = [torch.randn((B, h, h, 25), device="cuda", generator=g) - 4.0 for h in (64, 32, 16, 8, 4)]
nb = sum(p.numel() for p in pred) * 4
run("fcos fused encode+loss B=256", lambda: dh.fcos.encode_loss_batch(bd, nd, dims, 20, [512, 512], pred), nb)
