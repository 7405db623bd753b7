"""The offline COCO -> sparse FCOS target formatter (/format_COCO_annotations_fcos.py:66-183).

The golden fixture (`tests/golden/sparse_fcos.npz`) holds what the UNMODIFIED reference script produced on a small
annotation table (oracle/make_golden.py: sparse_fcos).  CPU: the oracle's restatement reproduces it entry for entry.
GPU: `densehead.fcos.sparse_format_batch` reproduces the oracle bit for bit (NaN centerness included: the script takes
square roots of negative ratios when NumPy's slice rules wrap a negative corner), batched and ragged."""
import hashlib

import numpy as np
import pytest

from oracle import dense_head_ref as O


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=a.dtype.kind == "f")


def _sha(idx, val):
    return hashlib.sha256(np.ascontiguousarray(idx).tobytes() + np.ascontiguousarray(val).tobytes()).hexdigest()


def test_oracle_reproduces_the_reference_script(golden):
    z = golden("sparse_fcos")
    for f in z["files"]:
        key = str(f).split(".")[0]
        idx, val = O.fcos_sparse_format(z[key + "_objects"], z[key + "_src_dims"], tuple(z[key + "_img_dims"]))
        assert len(val) == int(z[key + "_nnz"])
        if key + "_indices" in z.files:
            assert _same(idx, z[key + "_indices"]) and _same(val, z[key + "_values"])
        assert _sha(idx, val) == str(z[key + "_sha"])
    assert O.fcos_sparse_scales((448, 448)) == [28, 56, 112, 224, 448]


def test_oracle_entry_layout():
    idx, val = O.fcos_sparse_format([[10.0, 20.0, 3.0, 2.0, 5]], (448, 448))
    assert idx.shape == (3 * 2 * 7, 4) and idx.dtype == np.int32 and val.dtype == np.float32
    cells = idx.reshape(6, 7, 4)
    assert [tuple(c[0, :2]) for c in cells] == [(20, 10), (21, 10), (20, 11), (21, 11), (20, 12), (21, 12)]  # x-major
    assert np.all(cells[:, :6, 3] == np.arange(6)) and np.all(cells[:, 6, 3] == 9) and np.all(cells[:, :, 2] == 0)
    v = val.reshape(6, 7)
    assert np.array_equal(v[0], [0, 2, 0, 3, 0, 1, 1]) and np.array_equal(v[3, :4], [1, 1, 1, 2])
    assert v[3, 4] == np.float32(np.sqrt(1 / 2) * np.sqrt(1 / 1))


def _random_objects(rng, batch, nmax, side_w, side_h):
    n = rng.integers(0, nmax + 1, size=batch).astype(np.int32)
    n[0] = nmax
    obj = np.zeros((batch, nmax, 5))
    obj[..., 0] = rng.uniform(-20, side_w, (batch, nmax))
    obj[..., 1] = rng.uniform(-20, side_h, (batch, nmax))
    obj[..., 2] = rng.uniform(-2, 1, (batch, nmax)) ** 2 * rng.choice([20, 60, 150, 400], (batch, nmax))
    obj[..., 3] = rng.uniform(-2, 1, (batch, nmax)) ** 2 * rng.choice([20, 60, 150, 400], (batch, nmax))
    obj[..., 2:4] *= np.where(rng.random((batch, nmax, 1)) < 0.03, -1.0, 1.0)  # a few negative sizes: skipped
    obj[..., 4] = rng.integers(1, 81, (batch, nmax))
    return obj, n


@pytest.mark.gpu
def test_gpu_matches_golden_and_oracle(golden):
    torch = pytest.importorskip("torch")
    import densehead as dh
    z = golden("sparse_fcos")
    keys = [str(f).split(".")[0] for f in z["files"]]
    nmax = max(len(z[k + "_objects"]) for k in keys)
    obj = np.zeros((len(keys), nmax, 5))
    nb = np.zeros(len(keys), np.int32)
    for b, k in enumerate(keys):
        nb[b] = len(z[k + "_objects"])
        obj[b, :nb[b]] = z[k + "_objects"]
    idx, val, off = dh.fcos.sparse_format_batch(obj, nb, np.stack([z[k + "_src_dims"] for k in keys]))
    idx, val, off = idx.cpu().numpy(), val.cpu().numpy(), off.cpu().numpy()
    assert off[0] == 0 and off[-1] == len(val) == sum(int(z[k + "_nnz"]) for k in keys)
    for b, k in enumerate(keys):
        assert off[b + 1] - off[b] == int(z[k + "_nnz"])
        assert _sha(idx[off[b]:off[b + 1]], val[off[b]:off[b + 1]]) == str(z[k + "_sha"])


@pytest.mark.gpu
@pytest.mark.parametrize("canvas", [(448, 448), (320, 512)])
def test_gpu_random_batch_matches_the_oracle(canvas):
    torch = pytest.importorskip("torch")
    import densehead as dh
    rng = np.random.default_rng(77)
    B = 12
    obj, n = _random_objects(rng, B, 9, 640, 480)
    src = np.stack([rng.choice([640.0, 500.0, 448.0, 1024.0], B), rng.choice([480.0, 375.0, 448.0, 333.0], B)], axis=1)
    idx, val, off = dh.fcos.sparse_format_batch(obj, n, src, canvas)
    idx, val, off = idx.cpu().numpy(), val.cpu().numpy(), off.cpu().numpy()
    for b in range(B):
        wi, wv = O.fcos_sparse_format(obj[b, :n[b]], src[b], canvas)
        assert off[b + 1] - off[b] == len(wv)
        assert _same(idx[off[b]:off[b + 1]], wi) and _same(val[off[b]:off[b + 1]], wv)
    # without counts every row is an object; an empty batch is empty
    idx2, val2, off2 = dh.fcos.sparse_format_batch(obj[:1], None, src[:1], canvas)
    assert _same(idx2.cpu().numpy(), idx[:off[1]]) and int(off2[-1]) == off[1]
    e = dh.fcos.sparse_format_batch(np.zeros((0, 3, 5)), np.zeros(0, np.int32), np.zeros((0, 2)))
    assert e[0].shape == (0, 4) and e[1].shape == (0,) and e[2].tolist() == [0]


@pytest.mark.gpu
def test_gpu_full_batch_matches_the_oracle():
    """256 images x up to 30 objects on the script's 448 x 448 canvas (about 10 M entries): every image against the oracle."""
    torch = pytest.importorskip("torch")
    import densehead as dh
    rng = np.random.default_rng(78)
    B = 256
    obj, n = _random_objects(rng, B, 30, 640, 480)
    src = np.stack([rng.choice([640.0, 500.0, 448.0, 1024.0], B), rng.choice([480.0, 375.0, 448.0, 333.0], B)], axis=1)
    idx, val, off = dh.fcos.sparse_format_batch(obj, n, src)
    idx, val, off = idx.cpu().numpy(), val.cpu().numpy(), off.cpu().numpy()
    assert off[0] == 0 and off[-1] == len(val) and np.all(np.diff(off) >= 0)
    for b in range(B):
        wi, wv = O.fcos_sparse_format(obj[b, :n[b]], src[b])
        assert off[b + 1] - off[b] == len(wv), b
        assert _same(idx[off[b]:off[b + 1]], wi) and _same(val[off[b]:off[b + 1]], wv), b


@pytest.mark.gpu
def test_gpu_capacity_is_respected():
    torch = pytest.importorskip("torch")
    import densehead as dh
    from densehead import _capi
    obj = torch.tensor([[[10.0, 20.0, 30.0, 40.0, 3.0]]], dtype=torch.float64, device="cuda")
    src = torch.tensor([[448.0, 448.0]], dtype=torch.float64, device="cuda")
    off = torch.zeros(2, dtype=torch.int64, device="cuda")
    idx = torch.full((1000 + 64, 4), -7, dtype=torch.int32, device="cuda")
    val = torch.full((1000 + 64,), -7.0, device="cuda")
    _capi.check(_capi.lib().dh_fcos_sparse_encode(_capi.handle(0), obj.data_ptr(), None, src.data_ptr(), 1, 1, 448, 448, 5, 1000,
                                                  idx.data_ptr(), val.data_ptr(), off.data_ptr(), None), "dh_fcos_sparse_encode")
    torch.cuda.synchronize()
    assert off.tolist() == [0, 30 * 40 * 7]
    wi, wv = O.fcos_sparse_format([[10.0, 20.0, 30.0, 40.0, 3.0]], (448, 448))
    assert np.array_equal(idx[:1000].cpu().numpy(), wi[:1000]) and np.array_equal(val[:1000].cpu().numpy(), wv[:1000])
    assert bool((idx[1000:] == -7).all()) and bool((val[1000:] == -7).all())
