"""Host-side partition of independent streams across the GPUs of one box
(SURVEY section 8e): entries / members are self-contained deflate streams, so
each GPU gets one CONTIGUOUS range of descriptors (in input-offset order) of
about equal weight - one contiguous compressed byte range to copy in, one
contiguous output range to copy back, and no collective on the data path."""
from __future__ import annotations


def partition_contiguous(descs, parts: int):
    """The C partitioner of the library (b2i_partition_contiguous, csrc/b2i_partition.c):
    [(lo, hi)] * parts, contiguous descriptor ranges of ~equal in+out bytes."""
    from . import capi
    return capi.partition_contiguous(descs, parts)


def rebase(descs, idx):
    """Descriptors `idx` (indices into descs) re-based so that their input span and
    output range start at 0: -> (sub_descs, in_lo, in_hi, out_bytes)."""
    from . import capi
    items = [capi.StreamDesc.from_buffer_copy(descs[i]) for i in idx]
    if not items:
        return capi.make_descs([]), 0, 0, 0
    in_lo = min(int(d.in_off) for d in items) & ~15
    in_hi = max(int(d.in_off + d.in_len) for d in items)
    out = 0
    for d in items:
        d.in_off -= in_lo
        d.out_off = out
        out = (out + int(d.out_cap) + 15) & ~15
    return capi.make_descs(items), in_lo, in_hi, out


def shard_descs(descs, rank: int, world: int):
    """The descriptors rank `rank` of `world` decodes (contiguous partition), re-based:
    -> (sub_descs, in_lo, in_hi, out_bytes, (lo, hi))."""
    lo, hi = partition_contiguous(descs, world)[rank]
    sub, in_lo, in_hi, out = rebase(descs, range(lo, hi))
    return sub, in_lo, in_hi, out, (lo, hi)


def shard_descs_lpt(descs, rank: int, world: int):
    """LPT partition (batches dominated by a few huge streams, BASELINE config 4):
    -> (sub_descs, in_lo, in_hi, out_bytes, indices)."""
    from . import capi
    owner, _ = capi.partition_lpt(descs, world)
    idx = [i for i, o in enumerate(owner) if o == rank]
    sub, in_lo, in_hi, out = rebase(descs, idx)
    return sub, in_lo, in_hi, out, idx


def bind_to_gpu_numa(device_index: int):
    """Pin this process to the CPUs next to GPU `device_index` (NVML's ideal CPU set) BEFORE
    pinned host buffers are allocated: with one process per GPU, staging memory then lives on
    the GPU's own NUMA node and host<->device copies do not cross the socket interconnect.
    Returns the number of CPUs bound to, or 0 when nothing was changed (no NVML, one node...)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus or len(cpus) >= len(os.sched_getaffinity(0)):
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0
