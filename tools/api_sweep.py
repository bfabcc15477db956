"""GPU: public-API throughput of config 1 under the streaming engine's knobs."""
import json, os, subprocess, sys
sys.path.insert(0, os.getcwd())
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "zip64k"
archive, kind = bench.build_workload(name, 0, bench.CONFIGS[name][3])
path = "/dev/shm/b2i_sweep.bin"
open(path, "wb").write(archive)
exe = os.environ.get("API_BENCH", os.path.join("libarchive_b200", "api_bench"))
SWEEP = json.loads(os.environ["SWEEP"]) if os.environ.get("SWEEP") else None
for env in SWEEP or ({}, {"B2I_PIPE_JOBS": "4", "B2I_PIPE_DEPTH": "5"}, {"B2I_PIPE_JOBS": "4", "B2I_PIPE_DEPTH": "5", "B2I_PIPE_WINDOW_MB": "64", "B2I_PIPE_FIRST_MB": "16"},
            {"B2I_PIPE_JOBS": "4", "B2I_PIPE_DEPTH": "5", "B2I_PIPE_WINDOW_MB": "64", "B2I_PIPE_FIRST_MB": "64"},
            {"B2I_PIPE_JOBS": "4", "B2I_PIPE_DEPTH": "6", "B2I_PIPE_WINDOW_MB": "32", "B2I_PIPE_FIRST_MB": "16"},
            {"B2I_PIPE_JOBS": "3", "B2I_PIPE_DEPTH": "4", "B2I_PIPE_WINDOW_MB": "96", "B2I_PIPE_FIRST_MB": "32"},
            {"B2I_PIPE_JOBS": "2", "B2I_PIPE_DEPTH": "3", "B2I_PIPE_WINDOW_MB": "128", "B2I_PIPE_FIRST_MB": "128"}):
    cmd = [exe, path, "--mode", "block", "--steps", "6", "--warmup", "3"] + (["--raw"] if kind != "zip" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, **env))
    j = json.loads(r.stdout.strip().splitlines()[-1])
    print(name, env, "%.2f GB/s mean %.2f ms best %.2f ms first block %.2f ms" % (j["gbps_mean"], j["seconds_mean"] * 1e3, j["seconds_best"] * 1e3, j["last_pass"]["first_block_s"] * 1e3), flush=True)
os.unlink(path)
