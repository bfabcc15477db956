"""GPU: where the public-API time goes (api_bench on the drop-in): headers only, blocks, data,
memory and file sources, with the pass's own phase stamps."""
import json, os, subprocess, sys
sys.path.insert(0, os.getcwd())
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "zip64k"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[name][3]
archive, kind = bench.build_workload(name, 0, scale)
path = "/dev/shm/b2i_probe.bin"
open(path, "wb").write(archive)
exe = os.path.join("libarchive_b200", "api_bench")
for mode, extra in (("headers", []), ("block", []), ("data", []), ("block", ["--file"]), ("block", ["--check"])):
    cmd = [exe, path, "--mode", mode, "--steps", "5", "--warmup", "3"] + extra + (["--raw"] if kind != "zip" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        print(name, mode, extra, "%.2f GB/s mean %.2f ms best %.2f ms" % (j["gbps_mean"], j["seconds_mean"] * 1e3, j["seconds_best"] * 1e3), j["last_pass"], "errors", j["errors"], flush=True)
    except Exception as ex:
        print(name, mode, extra, "FAILED", r.returncode, r.stderr[-400:], r.stdout[-200:])
os.unlink(path)
