"""GPU: timeline of the streaming engine under the public API (B2I_PIPE_TRACE)."""
import os, subprocess, sys
sys.path.insert(0, os.getcwd())
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "zip64k"
archive, kind = bench.build_workload(name, 0, bench.CONFIGS[name][3])
path = "/dev/shm/b2i_trace.bin"
open(path, "wb").write(archive)
exe = os.path.join("libarchive_b200", "api_bench")
cmd = [exe, path, "--mode", "block", "--steps", "1", "--warmup", "3"] + (["--raw"] if kind != "zip" else [])
r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, B2I_PIPE_TRACE="1"))
lines = r.stderr.splitlines()
# the last pass only
starts = [i for i, l in enumerate(lines) if "window   0  stage" in l]
print("\n".join(lines[starts[-1]:]))
print(r.stdout.strip().splitlines()[-1][:300])
os.unlink(path)
