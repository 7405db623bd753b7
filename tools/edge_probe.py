import sys, os, numpy as np, torch
sys.path[:0] = ['/root/repo', '/root/repo/cv-lite-object-detection_b200', '/root/repo/tests']
import densehead as dh
from oracle import synth
case = sys.argv[1]
C = 20
if case == "empty":
    r = dh.fcos.encode_loss_batch(np.zeros((0, 4, 5), np.float32), np.zeros((0,), np.int32), np.zeros((0, 2), np.float32),
                                  C, [256, 256], [torch.zeros((0, 256 // s, 256 // s, C + 5), device="cuda") for s in (8, 16, 32, 64, 128)])
elif case == "noboxes":
    pred = synth.fcos_predictions(2, 256, C, 5)
    r = dh.fcos.encode_loss_batch(np.zeros((2, 4, 5), np.float32), np.zeros((2,), np.int32), [256, 256], C, [256, 256], pred)
elif case in ("cap0", "cap1"):
    boxes, nbox = synth.make_boxes(1, 512, 256, C, 4.0, 200.0, synth.seed_for(5, 90), full=True)
    pred = synth.fcos_predictions(1, 512, C, 6)
    dh.set_option(0, 5, int(case[-1]))
    r = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], C, [512, 512], pred)
elif case == "cap_encode":
    boxes, nbox = synth.make_boxes(1, 512, 256, C, 4.0, 200.0, synth.seed_for(5, 90), full=True)
    pred = synth.fcos_predictions(1, 512, C, 6)
    tg, _ = dh.fcos.format_data_batch(boxes, nbox, [512, 512], C, [512, 512])
    torch.cuda.synchronize(); print("encode done")
    r = dh.fcos.model_loss_batch(tg, pred)
elif case == "c150_full":
    boxes, nbox = synth.make_boxes(2, 256, 12, 150, 8.0, 150.0, synth.seed_for(5, 91))
    pred = synth.fcos_predictions(2, 256, 150, 7)
    tg, _ = dh.fcos.format_data_batch(boxes, nbox, [256, 256], 150, [256, 256])
    torch.cuda.synchronize(); print("encode done")
    r = dh.fcos.model_loss_batch(tg, pred)
    torch.cuda.synchronize(); print("unfused done")
    try:
        dh.fcos.encode_loss_batch(boxes, nbox, [256, 256], 150, [256, 256], pred, weights=(1, 1, 1))
    except ValueError as e:
        print("raised", str(e)[:60])
elif case == "c150":
    boxes, nbox = synth.make_boxes(2, 256, 12, 150, 8.0, 150.0, synth.seed_for(5, 91))
    pred = synth.fcos_predictions(2, 256, 150, 7)
    r = dh.fcos.encode_loss_batch(boxes, nbox, [256, 256], 150, [256, 256], pred)
elif case == "misaligned":
    boxes, nbox = synth.make_boxes(2, 256, 12, 80, 8.0, 150.0, synth.seed_for(5, 92))
    pred = synth.retina_predictions(2, 256, 80, 8)
    shifted = []
    for p in [torch.from_numpy(p).cuda() for p in pred]:
        buf = torch.empty(p.numel() + 1, device="cuda"); buf[1:].copy_(p.reshape(-1)); shifted.append(buf[1:].view(p.shape))
    r = dh.retinanet.encode_loss_batch(boxes, nbox, [256, 256], 80, [256, 256], shifted)
torch.cuda.synchronize()
print(case, "ok", r[1].tolist() if len(r) > 1 else None)
