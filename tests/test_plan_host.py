"""The host-side chunk planner of the fused encode + loss kernel (dh_plan_fused_chunks: no device needed).  Whatever the
batch, the plan covers every tile of every image exactly once with chunks the kernel can hold; the plans the measurements in
DESIGN.md section 4.2 were taken with are pinned."""
import ctypes

import numpy as np
import pytest


def _plan(batch, tpi, ch, grid=592, grad=0, tail=1, max_chunk=16, rows_per_tile=256):
    from densehead import _capi
    out = (ctypes.c_int32 * 34)()
    n = _capi.lib().dh_plan_fused_chunks(batch, tpi, rows_per_tile, ch, grid, grad, tail, max_chunk, out)
    assert n >= 1 and out[0] == n
    tiers = [dict(image0=out[2 + 4 * k], chunk_tiles=out[3 + 4 * k], cpi=out[4 + 4 * k], chunk0=out[5 + 4 * k]) for k in range(n)]
    return tiers, int(out[1])


@pytest.mark.parametrize("tpi,ch", [(324, 84), (23, 25), (320, 5), (1, 84), (7, 10)])
@pytest.mark.parametrize("grad", [0, 1])
@pytest.mark.parametrize("tail", [0, 1])
def test_every_plan_covers_the_batch(tpi, ch, grad, tail):
    for batch in [1, 2, 3, 7, 8, 16, 24, 31, 32, 33, 40, 48, 64, 96, 100, 128, 200, 256, 1000]:
        tiers, n_chunks = _plan(batch, tpi, ch, grad=grad, tail=tail)
        assert tiers[0]["image0"] == 0 and tiers[0]["chunk0"] == 0
        chunk0 = 0
        for k, t in enumerate(tiers):
            end = tiers[k + 1]["image0"] if k + 1 < len(tiers) else batch
            images = end - t["image0"]
            assert images >= 1, (batch, tiers)
            assert 1 <= t["chunk_tiles"] <= 18
            assert t["cpi"] * t["chunk_tiles"] >= tpi > (t["cpi"] - 1) * t["chunk_tiles"]  # an image is exactly cpi chunks
            assert t["chunk0"] == chunk0
            chunk0 += images * t["cpi"]
        assert chunk0 == n_chunks


def test_plans_behind_the_measurements():
    # RetinaNet-COCO (324 tiles of 86 KB per image) on 592 resident CTAs
    t, n = _plan(32, 324, 84)
    assert len(t) == 1 and t[0]["chunk_tiles"] == 18 and n == 576                 # one chunk per CTA
    t, n = _plan(64, 324, 84)
    assert len(t) == 1 and t[0]["chunk_tiles"] == 18 and n == 1152                # 1.95 waves
    t, n = _plan(16, 324, 84)
    assert len(t) == 1 and n <= 592 and n / 592 >= 0.85
    for batch in (128, 256):                                                     # tiers of 16, 8, 4, 2: none below 150 KB
        t, n = _plan(batch, 324, 84)
        assert [x["chunk_tiles"] for x in t] == [16, 8, 4, 2]
    t, _ = _plan(32, 324, 84, grad=1)                                            # the gradient kernel keeps its tiers
    assert len(t) > 1 and max(x["chunk_tiles"] for x in t) <= 16
    # FCOS-VOC (23 tiles of 25.6 KB): equal 12-tile chunks at 256 images; nothing under 6 tiles in a tiered plan
    t, n = _plan(256, 23, 25)
    assert len(t) == 1 and t[0]["chunk_tiles"] == 12 and n == 512
    t, _ = _plan(1000, 23, 25)
    assert min(x["chunk_tiles"] for x in t) >= 6
    # CenterNet-s8 stride 4 (320 tiles of 5 KB): always the largest chunks
    t, _ = _plan(256, 320, 5)
    assert min(x["chunk_tiles"] for x in t) >= 16
    # uniform plan on request (DH_OPT_FUSED_TAIL = 0): 4..8 tiles
    t, _ = _plan(256, 324, 84, tail=0)
    assert len(t) == 1 and 4 <= t[0]["chunk_tiles"] <= 8


def test_bad_arguments_are_refused():
    from densehead import _capi
    out = (ctypes.c_int32 * 34)()
    assert _capi.lib().dh_plan_fused_chunks(4, 324, 256, 84, 0, 0, 1, 16, out) < 0
    assert _capi.lib().dh_plan_fused_chunks(4, 324, 256, 84, 592, 0, 1, 3, out) < 0
    assert _capi.lib().dh_plan_fused_chunks(4, 324, 256, 84, 592, 0, 1, 16, None) < 0
