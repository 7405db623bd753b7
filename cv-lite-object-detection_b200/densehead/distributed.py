"""Multi-GPU plumbing for the dense-head path: images are independent, so the batch is sharded by image
across ranks (one process per GPU) and the only exchange is one all-reduce of the loss scalars
{cls, reg, cen, n_pos} per step.

The exchange itself lives in libdensehead.so (`dh_comm_*`, `dh_allreduce_loss`, include/densehead.h): NVLink peer
mailboxes -- optionally inside the fused loss kernel's last CTA -- or NCCL on the caller's stream.  `init_comm`
only needs an out-of-band channel to hand the bootstrap blobs around (64 B per rank for the mailboxes, 128 B for
NCCL); here that channel is whatever `torch.distributed` group is up (NCCL on the B200 box, gloo in the CPU
tests), a TensorFlow-side caller would use MPI or a file.
"""
import ctypes

import torch

from . import _capi

_comm = {}  # device index -> {"world", "rank", "peer", "nccl", "fused"}


def shard_range(n_images, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of the batch owned by `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(int(n_images), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(arrays, rank, world_size):
    """Slice every [B, ...] array / tensor of `arrays` down to this rank's images."""
    n = len(arrays[0])
    lo, hi = shard_range(n, rank, world_size)
    return [a[lo:hi] for a in arrays]


def _all_ok(ok, group):
    import torch.distributed as dist
    flags = [None] * dist.get_world_size(group)
    dist.all_gather_object(flags, bool(ok), group=group)
    return all(flags)


def init_comm(group=None, transports=("peer", "nccl"), fuse=True):
    """Collective: attach a communicator to this process's densehead handle (current CUDA device).  Returns a dict
    {"world", "rank", "peer": bool, "nccl": bool, "fused": bool}.  With `fuse` (and the peer mailboxes up) the
    `total` returned by every `*encode_loss_batch` call is already summed over all ranks -- the exchange happens in
    the last CTA of the loss kernel, there is no second launch (DH_OPT_LOSS_ALLREDUCE)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.cuda.current_device()
    info = {"world": world, "rank": rank, "peer": False, "nccl": False, "fused": False, "errors": []}
    if world == 1:
        _comm[dev] = info
        return info
    L, h = _capi.lib(), _capi.handle(dev)
    if "peer" in transports:
        blob = ctypes.create_string_buffer(_capi.DH_IPC_HANDLE_BYTES)
        ok = True
        try:
            _capi.check(L.dh_comm_peer_export(h, blob), "dh_comm_peer_export")
        except Exception as exc:  # noqa: BLE001 -- any failure means "no mailboxes", decided collectively below
            ok = False
            info["errors"].append(repr(exc))
        blobs = [None] * world
        dist.all_gather_object(blobs, (ok, blob.raw), group=group)
        ok = all(b[0] for b in blobs)
        if ok:
            try:
                _capi.check(L.dh_comm_peer_import(h, world, rank, b"".join(b[1] for b in blobs)), "dh_comm_peer_import")
            except Exception as exc:  # noqa: BLE001
                ok = False
                info["errors"].append(repr(exc))
        info["peer"] = _all_ok(ok, group)
    if "nccl" in transports:
        uid = ctypes.create_string_buffer(_capi.DH_UNIQUE_ID_BYTES)
        box = [None]
        if rank == 0:
            try:
                _capi.check(L.dh_comm_get_unique_id(uid), "dh_comm_get_unique_id")
                box = [uid.raw]
            except Exception as exc:  # noqa: BLE001
                info["errors"].append(repr(exc))
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ok = box[0] is not None
        if ok:
            try:
                _capi.check(L.dh_comm_init_rank(h, world, rank, box[0]), "dh_comm_init_rank")
            except Exception as exc:  # noqa: BLE001
                ok = False
                info["errors"].append(repr(exc))
        info["nccl"] = _all_ok(ok, group)
    if not info["peer"]:  # a rank whose import worked must not use mailboxes the others do not have
        _capi.set_option(dev, _capi.DH_OPT_ALLREDUCE, 1)
    info["fused"] = bool(fuse and info["peer"])
    _capi.set_option(dev, _capi.DH_OPT_LOSS_ALLREDUCE, 1 if info["fused"] else 0)
    _comm[dev] = info
    return info


def comm_info(device_index=None):
    dev = torch.cuda.current_device() if device_index is None else device_index
    return _comm.get(dev)


def set_fused(on, device_index=None):
    """Switch the in-kernel exchange of the fused loss calls on or off (every rank must do the same)."""
    dev = torch.cuda.current_device() if device_index is None else device_index
    info = _comm.get(dev)
    on = bool(on and info and info["peer"])
    _capi.set_option(dev, _capi.DH_OPT_LOSS_ALLREDUCE, 1 if on else 0)
    if info:
        info["fused"] = on
    return on


def allreduce_losses(total, group=None, stream=None):
    """Sum the per-rank loss vector across ranks in place.  Device tensors go through the library's communicator when
    `init_comm` set one up (one tiny kernel over the peer mailboxes, or ncclAllReduce; both capturable in a CUDA
    graph); otherwise through torch.distributed (no-op when that is not initialised)."""
    if isinstance(total, torch.Tensor) and total.is_cuda:
        info = _comm.get(total.device.index)
        if info and info["world"] > 1 and (info["peer"] or info["nccl"]):
            if total.dtype != torch.float32 or not total.is_contiguous():
                raise ValueError("the loss vector must be contiguous float32")
            s = torch.cuda.current_stream(total.device) if stream is None else stream
            _capi.check(_capi.lib().dh_allreduce_loss(_capi.handle(total.device.index), total.data_ptr(), int(total.numel()),
                                                      int(s.cuda_stream)), "dh_allreduce_loss")
            return total
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total


def destroy_comm(group=None):
    """Collective: barrier, then drop this process's communicator."""
    import torch.distributed as dist
    dev = torch.cuda.current_device()
    if dev not in _comm:
        return
    torch.cuda.synchronize()
    if dist.is_available() and dist.is_initialized():
        dist.barrier(group=group)
    _capi.set_option(dev, _capi.DH_OPT_LOSS_ALLREDUCE, 0)
    _capi.check(_capi.lib().dh_comm_destroy(_capi.handle(dev)), "dh_comm_destroy")
    del _comm[dev]


def gather_counts(per_rank_count, group=None):
    """All ranks learn how many images every rank processed (ragged last shard)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [int(per_rank_count)]
    t = torch.tensor([int(per_rank_count)], dtype=torch.int64, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return [int(x[0]) for x in out]
