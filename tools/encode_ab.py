"""Tile streamer (DH_OPT_ENCODE_KERNEL = 1) against the direct-store kernel (= 2) for every encoder family at several
batch sizes: CUDA-graph-timed single launches, bit-identical outputs asserted.  One JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cv-lite-object-detection_b200")]
import torch  # noqa: E402
import densehead as dh  # noqa: E402
from densehead import _capi  # noqa: E402
from oracle import synth  # noqa: E402

PEAK = 6553.3
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
SC = [32, 64, 128, 256, 512]


def graph_time(fn, reps=20, per_graph=10):
    """Per-launch time with `per_graph` launches captured in one graph: a graph of ONE small kernel measures the
    graph-launch latency (6-8 us per replay on this box), not the kernel."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per_graph):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * per_graph) * 1e-3


def case(tag, B, make):
    res = {}
    outs_by_kernel = {}
    for kern in (1, 2):
        dh.set_option(0, _capi.DH_OPT_ENCODE_KERNEL, kern)
        fn, outs = make()
        for o in outs:
            o.fill_(float("nan"))
        fn()
        torch.cuda.synchronize()
        outs_by_kernel[kern] = [o.clone() for o in outs]
        t = graph_time(fn)
        nbytes = sum(o.numel() for o in outs) * 4
        res[kern] = (t, nbytes)
    same = all(torch.equal(a, b) for a, b in zip(outs_by_kernel[1], outs_by_kernel[2]))
    dh.set_option(0, _capi.DH_OPT_ENCODE_KERNEL, 0)
    print(json.dumps({"case": tag, "B": B, "MB": round(res[1][1] / 1e6, 1),
                      "streamer_us": round(res[1][0] * 1e6, 1), "streamer_frac": round(res[1][1] / res[1][0] / 1e9 / PEAK, 3),
                      "direct_us": round(res[2][0] * 1e6, 1), "direct_frac": round(res[2][1] / res[2][0] / 1e9 / PEAK, 3),
                      "identical": same}), flush=True)
    assert same, tag


def main():
    for B in (1, 8, 32, 64, 256):
        boxes, nbox = synth.config_boxes("fcos_voc", B, 1)
        bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
        dims = torch.tensor([[512., 512.]] * B, device="cuda")
        for mode in ("fcos", "center"):
            def make(mode=mode):
                outs, cnt = dh.fcos.format_data_batch(bd, nd, dims, 20, [512, 512], mode=mode)
                return (lambda: dh.fcos.format_data_batch(bd, nd, dims, 20, [512, 512], mode=mode, out=outs, num_targets=cnt)), outs
            case("fcos_voc %s" % mode, B, make)
    for B in (1, 8, 32, 64, 256):
        boxes, nbox = synth.config_boxes("centernet_crowdhuman", B, 2)
        bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
        dims = torch.tensor([[512., 512.]] * B, device="cuda")
        for mode, kw in (("s8", dict(stride=4, mode="s8", box_scales=SC)), ("falloff", dict(stride=4, mode="falloff"))):
            def make(kw=kw):
                out, st = dh.centernet.format_data_batch(bd, nd, dims, 1, [512, 512], **kw)
                return (lambda: dh.centernet.format_data_batch(bd, nd, dims, 1, [512, 512], out=out, status=st, **kw)), [out]
            case("centernet %s stride 4" % mode, B, make)
    for B in (1, 4, 16, 64):
        boxes, nbox = synth.config_boxes("retina_coco", B, 3)
        bd, nd = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
        dims = torch.tensor([[640., 640.]] * B, device="cuda")

        def make():
            outs, pr = dh.retinanet.format_data_batch(bd, nd, dims, 80, [640, 640])
            return (lambda: dh.retinanet.format_data_batch(bd, nd, dims, 80, [640, 640], out=outs, num_pairs=pr)), outs
        case("retina_coco", B, make)


if __name__ == "__main__":
    main()
