"""Import the reference's library modules UNMODIFIED from /root/reference under the TF stub.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` (run in the build container,
where /root/reference exists) and by the `-m "not gpu"` tests that cross-check the oracle
against the live reference when it is present (they skip otherwise).  Nothing in the
product path, `-m gpu` tests, `smoke()` or `bench.py` may import this module: the
reference tree does not exist on the GPU box.

No reference source is copied: modules are executed from where they lie.
"""
import ast
import importlib.util
import os
import sys

import numpy as np

REF_ROOT = os.environ.get("DENSEHEAD_REFERENCE", "/root/reference")
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tfstub")
_SIBLINGS = ("utils", "tf_bias_layer", "fcos", "data_preprocess")
_STUB_PKGS = ("tensorflow", "matplotlib", "classification_models")
_cache = {}
_stub_mods = {}  # one shared instance of the stub packages, so Tensor classes match across loads


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "FCOS"))


def _enter(folder):
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                  if k.split(".")[0] in _SIBLINGS + _STUB_PKGS}
    sys.modules.update(_stub_mods)
    sys.path[:0] = [_STUB, os.path.join(REF_ROOT, folder)]
    if not hasattr(np, "int"):
        np.int = int  # RetinaNet/retinanet_module.py:303 uses the alias removed in NumPy 1.24
    return saved_path, saved_mods


def _leave(state):
    saved_path, saved_mods = state
    for k in list(sys.modules):
        root = k.split(".")[0]
        if root in _STUB_PKGS:
            _stub_mods[k] = sys.modules.pop(k)
        elif root in _SIBLINGS:
            sys.modules.pop(k)
    sys.modules.update(saved_mods)
    sys.path[:] = saved_path


def load(folder, module):
    """Return reference module `<folder>/<module>.py` executed under the stub."""
    key = (folder, module)
    if key in _cache:
        return _cache[key]
    if not available():
        raise FileNotFoundError("reference tree not found at %s" % REF_ROOT)
    state = _enter(folder)
    try:
        path = os.path.join(REF_ROOT, folder, module + ".py")
        spec = importlib.util.spec_from_file_location("_ref_%s_%s" % (folder, module), path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        _leave(state)
    _cache[key] = mod
    return mod


def tf():
    """The stub `tensorflow` module (for building Tensor inputs)."""
    if "tf" not in _cache:
        state = _enter("FCOS")
        try:
            import tensorflow  # noqa: the stub
            _cache["tf"] = tensorflow
        finally:
            _leave(state)
    return _cache["tf"]


class float64_mode:
    """Evaluate the reference's formulas in double precision: inside the block the stub's `tf.float32` IS float64 and
    Python floats become float64 tensors, so `tf.cast(x, tf.float32)` and literals such as `mask=1.0` no longer round
    to single precision.  Used to take finite differences through the reference's loss functions (make_golden.grad)."""

    def __enter__(self):
        stub = tf()
        self.stub, self.saved = stub, (stub.float32, stub._default_dtype)
        inner = stub._default_dtype

        def default_dtype(x):
            dt = inner(x)
            return np.dtype(np.float64) if dt == np.float32 and not isinstance(x, (np.ndarray, np.generic)) else dt
        stub.float32, stub._default_dtype = np.float64, default_dtype
        return self

    def __exit__(self, *exc):
        self.stub.float32, self.stub._default_dtype = self.saved
        return False


def retinanet(n_classes, **kwargs):
    """Instantiate the reference RetinaNet class with its Keras graph builder patched out
    (RetinaNet/retinanet_module.py:194-196 builds a backbone in __init__)."""
    mod = load("RetinaNet", "retinanet_module")
    mod.build_model = lambda *a, **k: None
    return mod.RetinaNet(n_classes, {i: str(i) for i in range(n_classes)}, **kwargs)


def functions_only(folder, module, names, extra_globals=None):
    """Exec only the named top-level FunctionDefs (+ imports) of a script whose module-level
    code cannot run here (e.g. FCOS/infer_fcos.py loads pickles from C:/ at import time)."""
    path = os.path.join(REF_ROOT, folder, module + ".py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    keep = [n for n in tree.body
            if isinstance(n, (ast.Import, ast.ImportFrom))
            or (isinstance(n, ast.FunctionDef) and n.name in names)]
    state = _enter(folder)
    try:
        ns = dict(extra_globals or {})
        exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    finally:
        _leave(state)
    return ns


def hourglass_inline_encoder(samples, n_classes, raw_dims, img_dims, pad_dims):
    """Run the 4-scale target encoder that CenterNet/train_hourglass_voc.py keeps INLINE in train() (lines 95-153, not a
    function) on `samples` = list of {"objects": {"bbox": [n,4] (xmin,ymin,xmax,ymax) normalised, "label": [n]}}.
    The statements are sliced out of the reference file by their text markers, dedented and exec'd unmodified under the
    stub with the loop variables the surrounding code would have set.  Returns the list of [H/8, W/8, 4, C+5] maps."""
    path = os.path.join(REF_ROOT, "CenterNet", "train_hourglass_voc.py")
    with open(path) as f:
        lines = f.read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip() == "img_boxes = []")
    end = next(i for i in range(start, len(lines)) if lines[i].strip() == "del img_bbox")
    body = "\n".join(l[8:] if l.startswith(" " * 8) else l for l in lines[start:end + 1])
    utils = load("CenterNet", "utils")
    ns = {"np": np, "convert_to_xywh": utils.convert_to_xywh, "train_data": samples, "batch_sample": list(range(len(samples))),
          "n_classes": n_classes, "raw_dims": raw_dims, "img_dims": img_dims, "pad_dims": pad_dims}
    state = _enter("CenterNet")
    try:
        exec(compile(body, path, "exec"), ns)
    finally:
        _leave(state)
    return ns["img_boxes"]


def offline_fcos_formatter(objects, labels):
    """Run the reference's offline COCO -> sparse FCOS target script, `/format_COCO_annotations_fcos.py`, UNMODIFIED from
    where it lies.  It is a top-level script that reads two CSV files from a Windows path: `pandas.read_csv` is patched
    for the duration to hand it `objects` (columns filename, img_width, img_height, id, x_lower, y_lower, box_width,
    box_height) and `labels` (columns id, name); `tf.sparse.SparseTensor` is a recorder (the script appends one 5-long
    index per cell among the 4-long ones, :166, so the real constructor would reject its input); `print` is silenced and
    `open` refuses to write.  The script's last statements index `train_objects[1]` with one image scale configured
    (:192) -- an IndexError after all the work is done, caught here.  Returns `train_objects[0]`: a list of
    (filename, (img_width, img_height), recorder) with `.indices`, `.values`, `.dense_shape` as the script built them."""
    import pandas as pd

    class SparseRecorder:
        def __init__(self, indices, values, dense_shape):
            self.indices, self.values, self.dense_shape = indices, values, dense_shape

    def read_csv(path, *a, **k):
        return (labels if str(path).endswith("labels.csv") else objects).copy()

    def no_open(*a, **k):
        raise PermissionError("the offline formatter must not touch the file system here")

    path = os.path.join(REF_ROOT, "format_COCO_annotations_fcos.py")
    with open(path) as f:
        code = compile(f.read(), path, "exec")
    stub = tf()
    saved_read, had_sparse = pd.read_csv, getattr(stub, "sparse", None)
    state = _enter("FCOS")
    ns = {"__name__": "_ref_offline_fcos", "print": lambda *a, **k: None, "open": no_open}
    try:
        import types
        pd.read_csv = read_csv
        stub.sparse = types.SimpleNamespace(SparseTensor=SparseRecorder)
        try:
            exec(code, ns)
        except IndexError:
            pass
    finally:
        pd.read_csv = saved_read
        if had_sparse is None:
            del stub.sparse
        else:
            stub.sparse = had_sparse
        _leave(state)
    return ns["train_objects"][0]

