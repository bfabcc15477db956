"""N>1 host path on CPU: two gloo ranks partition one archive's descriptors by
byte range (shard.py), each decodes only its shard (device code under the host SIMT
emulator), and the gathered per-entry results equal zlib's for the whole archive."""
import os
import socket
import sys
import zlib

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import ctypes as C
    from emul_ctx import EmulContext
    from libarchive_b200 import capi, reader, shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = synth.split_text(24 * 9000, 9000, 8)
    z = synth.make_zip([synth.ZipMember("g%02d" % i, p) for i, p in enumerate(parts)], threads=1)
    entries, _, _ = capi.zip_index(z)
    descs, _, _ = reader.plan_zip(entries)
    sub, in_lo, in_hi, out_bytes, (lo, hi) = shard.shard_descs(descs, rank, world)
    span = z[in_lo:in_hi]                                     # the only input bytes this rank touches
    inbuf = C.create_string_buffer(span, len(span) + 48)
    outbuf = C.create_string_buffer(out_bytes + 48)
    res = EmulContext().decode_host(inbuf, len(span), sub, outbuf, out_bytes)
    mine = torch.zeros(len(descs), 3, dtype=torch.int64)
    for k in range(hi - lo):
        assert outbuf.raw[sub[k].out_off:sub[k].out_off + res[k].out_bytes] == parts[lo + k]
        mine[lo + k] = torch.tensor([res[k].crc, res[k].out_bytes, 1 + res[k].status + res[k].flags])
    dist.all_reduce(mine)                                     # test-side gather only: not a data-path collective
    if rank == 0:
        ok = all(int(mine[i, 0]) == (zlib.crc32(p) & 0xFFFFFFFF) and int(mine[i, 1]) == len(p) and int(mine[i, 2]) == 1
                 for i, p in enumerate(parts))
        q.put((ok, [lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_decode():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    ok, rng0 = q.get(timeout=5)
    assert ok and rng0[0] == 0 and 8 <= rng0[1] <= 16
