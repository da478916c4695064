"""Batched and multi-GPU drivers above the C ABI.

The registered single-template signature (tcw:533) cannot batch by itself and the reference's
search loops are serial Python (grid_based_searches.py:1186, mcmc_based_searches.py:3511-3516),
so throughput comes from here: one ``tcw_map_batch`` call per batch of templates, and
template-level sharding over GPUs -- contiguous blocks ``[r*T/G, (r+1)*T/G)`` per rank, no
traffic inside a template, one ``all_gather`` of the 80-byte result records at the end
(SURVEY 8e).  ``torch.distributed`` is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

import numpy as np

from . import _lib
from .atoms import AtomBatch
from .backend import default_flags, get_handle
from .window import TransientWindowRange


def map_batch(batch: AtomBatch, window, BtSG: bool = False, want_fmn: bool = False, *, device: int = -1,
              flags: int | None = None, raise_on_degenerate: bool = True):
    """All templates of ``batch`` through one window range on one GPU.

    Returns ``(records, F_mn)``: ``records`` is a structured array (``_lib.RESULT_DTYPE``:
    ``maxF, m_ML, n_ML, t0_ML, tau_ML, lnBtSG, t0_MP, tau_MP, ...``), ``F_mn`` is
    ``[T, N_t0, N_tau]`` float32 or ``None``.
    """
    if flags is None:
        flags = default_flags()
    flags |= (_lib.WANT_BTSG if BtSG else 0) | (_lib.WANT_FMN if want_fmn else 0)
    w = TransientWindowRange.from_any(window)
    return get_handle(device).map_batch(batch, w, flags, raise_on_degenerate=raise_on_degenerate)


def submit_batch(batch: AtomBatch, window, BtSG: bool = False, *, device: int = -1, flags: int | None = None):
    """Asynchronous :func:`map_batch`: enqueue the copies and kernels (``tcw_submit``) and return a
    ticket; the host is free until :func:`wait_batch`.  One batch in flight per device."""
    if flags is None:
        flags = default_flags()
    flags |= _lib.WANT_BTSG if BtSG else 0
    h = get_handle(device)
    h.submit(batch, TransientWindowRange.from_any(window), flags)
    return h


def wait_batch(ticket, *, raise_on_degenerate: bool = True) -> np.ndarray:
    """Records of the batch submitted with :func:`submit_batch` (``tcw_wait``)."""
    return ticket.wait(raise_on_degenerate=raise_on_degenerate)[0]


def map_again(window, batch: AtomBatch, *, BtSG: bool = False, device: int = -1, flags: int | None = None,
              raise_on_degenerate: bool = True) -> np.ndarray:
    """Another window range over the batch that the LAST :func:`map_batch` call on this device
    uploaded -- its atoms are still resident, so nothing is copied host-to-device
    (``tcw_map_resident`` + ``tcw_fetch_results``).  ``batch`` is only used to check that the
    caller means the same batch; returns the records."""
    if flags is None:
        flags = default_flags()
    flags |= _lib.WANT_BTSG if BtSG else 0
    h = get_handle(device)
    if getattr(h, "_last_batch", None) is not batch or not hasattr(h, "map_resident"):
        return h.map_batch(batch, TransientWindowRange.from_any(window), flags,
                           raise_on_degenerate=raise_on_degenerate)[0]
    w = TransientWindowRange.from_any(window)
    w.check_type()
    h.map_resident(w, flags)
    return h.fetch_results(raise_on_degenerate)


def shard_range(T: int, rank: int, world_size: int):
    """Contiguous block of templates owned by ``rank`` (all templates cost the same for a
    fixed window range, so no dynamic balancing)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return (rank * T) // world_size, ((rank + 1) * T) // world_size


_gather_cache = {}


def gather_records(local: np.ndarray, T: int, group=None, device_src=None) -> np.ndarray:
    """``all_gather`` of per-template records over the ranks of ``group``; every rank returns
    the ``T`` records in template order.  Falls back to identity without a process group.

    One collective per call on preallocated buffers.  With NCCL and ``device_src`` (the records
    where the kernels left them in device memory, ``Handle.results_device()``) the shard goes from
    device memory straight into the collective -- no host-to-device hop; the gathered records come
    back in one device-to-host copy."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if len(local) != T:
            raise ValueError("no process group: the local shard must hold all templates")
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(T, rank, world)
    if len(local) != hi - lo:
        raise ValueError(f"rank {rank} holds {len(local)} records, expected {hi - lo}")
    itemsize = local.dtype.itemsize
    max_shard = max(shard_range(T, r, world)[1] - shard_range(T, r, world)[0] for r in range(world))
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    key = (world, max_shard * itemsize, str(dev))
    bufs = _gather_cache.get(key)
    if bufs is None:
        send = torch.zeros(max_shard * itemsize, dtype=torch.uint8, device=dev)
        recv = torch.empty(world * max_shard * itemsize, dtype=torch.uint8, device=dev)
        stage = torch.zeros(max_shard * itemsize, dtype=torch.uint8, pin_memory=use_cuda)
        host = torch.empty(world * max_shard * itemsize, dtype=torch.uint8, pin_memory=use_cuda)
        bufs = _gather_cache[key] = (send, recv, stage, host)
    send, recv, stage, host = bufs
    nb = len(local) * itemsize
    if use_cuda and device_src is not None and device_src.nbytes >= nb:
        if nb:
            send[:nb].copy_(torch.as_tensor(device_src, device=dev)[:nb], non_blocking=True)  # device to device
    else:
        if nb:
            stage[:nb] = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1))
        send.copy_(stage, non_blocking=use_cuda)
    dist.all_gather_into_tensor(recv, send, group=group)
    host.copy_(recv, non_blocking=use_cuda)
    if use_cuda:
        torch.cuda.current_stream().synchronize()
    raw = host.numpy().reshape(world, max_shard * itemsize)
    out = np.zeros(T, dtype=local.dtype)
    for r in range(world):
        a, b = shard_range(T, r, world)
        if b > a:
            out[a:b] = raw[r, : (b - a) * itemsize].view(local.dtype)
    return out


def map_sharded(make_shard, T: int, window, BtSG: bool = False, *, group=None, compute=None,
                device: int = -1, chunk: int = 0) -> np.ndarray:
    """Template-parallel map over all ranks of the (initialised) process group.

    ``make_shard(lo, hi) -> AtomBatch`` produces the atoms of templates ``[lo, hi)`` on the
    calling rank (each rank runs its own atom producer).  ``compute`` defaults to
    :func:`map_batch` on this rank's GPU and exists so that the host logic can be exercised
    on CPU (gloo) with an injected checker.  ``chunk`` > 0 bounds the templates per call.
    Returns the ``T`` gathered records on every rank.
    """
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(T, rank, world)
    default_compute = compute is None
    if compute is None:
        def compute(batch):  # noqa: E306
            return map_batch(batch, window, BtSG, device=device, raise_on_degenerate=False)[0]
    parts = []
    step = chunk if chunk > 0 else max(hi - lo, 1)
    for a in range(lo, hi, step):
        b = min(a + step, hi)
        parts.append(np.asarray(compute(make_shard(a, b))))
    local = np.concatenate(parts) if parts else np.zeros(0, dtype=_lib.RESULT_DTYPE)
    # one launch computed the whole shard: its records still sit in device memory, hand them to the
    # collective from there
    device_src = get_handle(device).results_device() if (default_compute and len(parts) == 1 and world > 1) else None
    return gather_records(local, T, group, device_src=device_src)
