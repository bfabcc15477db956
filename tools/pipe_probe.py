"""GPU: b2i_pipe throughput on one workload under tuning knobs (windows, slices)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from libarchive_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else "mixed"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[name][3]
archive, kind = bench.build_workload(name, 0, scale)
descs, out_bytes, usize, csize = bench.plan_for(archive, kind)
n = len(descs)
L = capi.lib()
ctx = capi.Context(0)
h_in = L.b2i_host_alloc(len(archive) + 64)
C.memmove(h_in, archive, len(archive))
def run(window_mb, depth):
    p = capi.Pipe([ctx], descs, mem=h_in, mem_size=len(archive), window_out=window_mb << 20, depth=depth)
    nw = p.windows
    t0 = time.perf_counter()
    step = max(1, n // 200)
    for i in range(0, n, step):
        p.get(i)
    _, _, r = p.get(n - 1)
    dt = time.perf_counter() - t0
    p.close()
    return dt, nw
for slices in (None, "1", "2"):
    if slices: os.environ["B2I_PIPE_SLICES"] = slices
    else: os.environ.pop("B2I_PIPE_SLICES", None)
    for window_mb, depth in ((0, 0), (64, 4), (256, 4), (512, 4), (1024, 3)):
        run(window_mb, depth)
        dt, nw = run(window_mb, depth)
        print("%s slices=%s window=%s MiB depth=%d windows=%d: %.1f ms = %.1f GB/s" % (name, slices, window_mb or "auto", depth, nw, dt * 1e3, usize / dt / 1e9), flush=True)
