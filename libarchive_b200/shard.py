"""Host-side partition of independent streams across the GPUs of one box
(SURVEY section 8e): entries / members are self-contained deflate streams, so
each GPU gets one CONTIGUOUS range of descriptors (in input-offset order) of
about equal weight - one contiguous compressed byte range to copy in, one
contiguous output range to copy back, and no collective on the data path."""
from __future__ import annotations


def partition_contiguous(weights, parts: int):
    """Cut `weights` (one per stream, input-offset order) into `parts` contiguous
    ranges of ~equal total weight.  Returns [(lo, hi)] * parts, possibly empty ranges."""
    n = len(weights)
    total = float(sum(weights))
    cuts, acc, k = [0], 0.0, 1
    for i, w in enumerate(weights):
        # close range k-1 when the running sum passes k/parts of the total
        while k < parts and acc + w / 2.0 >= total * k / parts and len(cuts) == k:
            cuts.append(i)
            k += 1
        acc += w
    while len(cuts) < parts:
        cuts.append(n)
    cuts.append(n)
    return [(cuts[i], max(cuts[i], cuts[i + 1])) for i in range(parts)]


def shard_descs(descs, rank: int, world: int):
    """The descriptors rank `rank` of `world` decodes, re-based so that its input
    span and output range start at 0: -> (sub_descs, in_lo, in_hi, out_bytes)."""
    from . import capi
    w = [int(d.in_len) + int(d.out_cap) for d in descs]
    lo, hi = partition_contiguous(w, world)[rank]
    items = [capi.StreamDesc.from_buffer_copy(descs[i]) for i in range(lo, hi)]
    if not items:
        return capi.make_descs([]), 0, 0, 0, (lo, hi)
    in_lo = min(int(d.in_off) for d in items) & ~15
    in_hi = max(int(d.in_off + d.in_len) for d in items)
    out = 0
    for d in items:
        d.in_off -= in_lo
        d.out_off = out
        out = (out + int(d.out_cap) + 15) & ~15
    return capi.make_descs(items), in_lo, in_hi, out, (lo, hi)
