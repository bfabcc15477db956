"""GPU: the bench's own consumption pattern of the pipe (get every 64th stream + release)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from libarchive_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else "tiny4k"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
archive, kind = bench.build_workload(name, 0, scale)
descs, out_bytes, usize, csize = bench.plan_for(archive, kind)
n = len(descs)
L = capi.lib()
ctx = capi.Context(0)
h_in = L.b2i_host_alloc(len(archive) + 64)
C.memmove(h_in, archive, len(archive))
for rep in range(3):
    t0 = time.perf_counter()
    p = capi.Pipe([ctx], descs, mem=h_in, mem_size=len(archive))
    t1 = time.perf_counter()
    tg = 0.0
    for i in range(0, n, 64):
        a = time.perf_counter()
        p.get(min(i + 63, n - 1))
        tg += time.perf_counter() - a
        p.release(i)
    _, _, r = p.get(n - 1)
    t2 = time.perf_counter()
    p.close()
    t3 = time.perf_counter()
    print("rep %d: open %.1f ms, consume %.1f ms (in get %.1f ms), close %.1f ms, windows %d" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, tg * 1e3, (t3 - t2) * 1e3, p.n and 0), flush=True)
