"""f-2 on the GPU: the 4-scale hourglass encoder (CenterNet/train_hourglass_voc.py:99-153) bit-exact against the frozen
reference output, and its losses (CenterNet/tf_hourglass_net.py:347-388: sigmoid cross-entropy or focal + masked L1),
unfused and fused, with gradients."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import dense_head_ref as O  # noqa: E402
from oracle import synth  # noqa: E402
from conftest import assert_close  # noqa: E402


def _dh():
    import densehead
    return densehead


def test_encoder_golden_and_batch(golden):
    dh = _dh()
    z = golden("hourglass4")
    for t in range(2):
        raw_dims, img_dims = (int(v) for v in z["hg4_%d_dims" % t])
        got = dh.centernet.format_data_hourglass4(z["hg4_%d_labels" % t], raw_dims, img_dims, 5)
        assert np.array_equal(got.cpu().numpy(), z["hg4_%d_map" % t])
    # ragged batch incl. an empty image, padded frame
    boxes, nbox = synth.make_boxes(4, 288, 30, 7, 6.0, 270.0, synth.seed_for(6, 30))
    nbox[1] = 0
    out, _ = dh.centernet.format_data_batch(boxes, nbox, [288, 288], 7, [320, 320], stride=8, mode="hourglass4")
    for b in range(4):
        assert np.array_equal(out[b].cpu().numpy(), O.hourglass4_format_data(boxes[b, :nbox[b]], 288, 320, 7))


@pytest.mark.parametrize("loss_type", ["sigmoid", "focal"])
def test_losses_unfused_fused_and_gradients(loss_type):
    dh = _dh()
    B, C = 3, 7
    boxes, nbox = synth.make_boxes(B, 288, 30, C, 6.0, 270.0, synth.seed_for(6, 31))
    rng = np.random.default_rng(32)
    yp = rng.normal(-2.0, 1.5, size=(B, 40, 40, 4, C + 5)).astype(np.float32)
    yt, _ = dh.centernet.format_data_batch(boxes, nbox, [288, 288], C, [320, 320], stride=8, mode="hourglass4")
    ytn = yt.cpu().numpy()
    masks = ytn[..., 4].copy()                       # the objectness channel is the regression mask (train loop)
    want = O.hourglass4_model_loss(ytn, masks, yp, loss_type)
    got = dh.centernet.model_loss_hourglass4(yt, masks, yp, loss_type)
    assert_close([float(got[0]), float(got[1])], [float(want[0]), float(want[1])], 1e-5, what="unfused")
    W = (1.0, 0.1, 0.0)
    pi, tot, _, grad = dh.centernet.encode_loss_batch(boxes, nbox, [288, 288], C, [320, 320], yp, stride=8, mode="hourglass4",
                                                     cls_type=loss_type, delta=0.0, weights=W)
    assert_close(tot[:2].cpu().numpy(), [float(want[0]), float(want[1])], 1e-5, what="fused")
    assert int(tot[3]) == int(masks.sum())
    # analytic gradient: BCE d/dx = sigmoid(x) - z on channels 4:, L1 sign(x - y) * mask on :4
    x = yp.astype(np.float64)
    if loss_type == "sigmoid":
        gc = 1.0 / (1.0 + np.exp(-x[..., 4:])) - ytn[..., 4:]
    else:
        gc = O._focal_grad64(ytn[..., 4:].astype(np.float64), x[..., 4:], 0.25, 2.0)
    gr = np.sign(x[..., :4] - ytn[..., :4]) * masks[..., None]
    wantg = np.concatenate([W[1] * gr, W[0] * gc], axis=-1)
    err = np.abs(grad.cpu().numpy() - wantg)
    assert err.max() <= 2e-5 * max(1.0, np.abs(wantg).max()), err.max()
