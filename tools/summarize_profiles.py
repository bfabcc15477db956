"""profiles/README_r02.md from the bench records under profiles/ (numbers quoted in DESIGN.md
come from here)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")

def load(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None

def f(x, nd=1):
    return "-" if x is None else ("%.*f" % (nd, x))

out = ["# profiles - round 2 records", "",
       "All numbers: B200, clocks 1 965 MHz, no throttle reasons (see `clocks` in each record); GB/s of decompressed",
       "output unless noted.  `bench.py` writes one JSON line; the files below are those lines.", ""]
for tag, title in (("r02_bench_n1.json", "Default line, 1 GPU (`python bench.py`)"),
                   ("r02_bench_full.json", "Full BASELINE sizes, 1 GPU (`python bench.py --full`)"),
                   ("r02_bench_n2.json", "2 GPUs (`torchrun --nproc-per-node 2 bench.py --gpus 2`)"),
                   ("r02_bench_n4.json", "4 GPUs"), ("r02_bench_n8.json", "8 GPUs")):
    j = load(tag)
    if j is None:
        continue
    out += ["## %s  - `%s`" % (title, tag), ""]
    e = j["e2e"]
    out += ["Headline config 1: **%s GB/s** device-resident (%s ms/step, roofline frac %s of %s GB/s HBM); "
            "e2e through libarchive's public API on the drop-in: **%s GB/s** `archive_read_data_block`, %s GB/s "
            "`archive_read_data` (64 KiB); C ABI: `b2i_decode_host` %s, `b2i_submit/b2i_wait` %s, `b2i_pipe` %s GB/s; "
            "host link per GPU (all ranks copying) H2D %s / D2H %s GB/s; CPU baseline %s." %
            (f(j["value"]), f(j["ms_per_step"], 3), f(j["roofline"]["frac"], 4), f(j["roofline"]["peak"], 0),
             f(e.get("value")), f(e.get("archive_read_data_64KiB")), f(e.get("b2i_decode_host")),
             f(e.get("b2i_submit_wait_two_jobs")), f(e.get("b2i_pipe")),
             f(e["host_link"]["h2d_GBps_per_gpu_concurrent"]), f(e["host_link"]["d2h_GBps_per_gpu_concurrent"]),
             ("%s GB/s on %d cores (%s)" % (f(j["cpu_baseline"]["value"], 2), j["cpu_baseline"]["cores"], j["cpu_baseline"]["kind"]))
             if j.get("cpu_baseline") else "not run at N > 1"), ""]
    if j.get("configs"):
        out += ["| config | size | device-resident GB/s | ms/step | roofline frac | `b2i_decode_host` | `b2i_submit/wait` | `b2i_pipe` | reference CPU GB/s |",
                "|---|---|---:|---:|---:|---:|---:|---:|---:|"]
        for k, v in j["configs"].items():
            if "error" in v:
                out.append("| %s | error: %s |" % (k, v["error"]))
                continue
            ee = v.get("e2e", {})
            cb = v.get("cpu_baseline") or {}
            out.append("| %d %s | scale %s, %d streams, %.2f GB out | %s | %s | %s | %s | %s | %s | %s |" %
                       (v["baseline_config"], k, v["config"]["scale"], v["streams_per_gpu"], v["out_bytes_per_gpu"] / 1e9,
                        f(v["value"]), f(v["ms_per_step"], 2), f(v["roofline"]["frac"], 4), f(ee.get("b2i_decode_host")),
                        f(ee.get("b2i_submit_wait_two_jobs")), f(ee.get("b2i_pipe")), f(cb.get("value"), 2)))
        out.append("")
    if j.get("strong"):
        out += ["Strong scaling (ONE archive split over the ranks by the C partitioner, device-resident):", "",
                "| config | archive | GB/s | ms/step | partition | largest stream's share of the bytes |", "|---|---|---:|---:|---|---:|"]
        for k, v in j["strong"].items():
            if "error" in v:
                out.append("| %s | error: %s |" % (k, v["error"]))
                continue
            out.append("| %d %s | %d streams, %.2f GB out | %s | %s | %s | %s |" %
                       (v["baseline_config"], k, v["streams"], v["out_bytes"] / 1e9, f(v["value"]), f(v["ms_per_step"], 2),
                        v["partition"], f(v["largest_stream_share"], 4)))
        out.append("")
for name, title in (("r02_team_stream.txt", "One large stream alone on the GPU (team kernel, `tools/team_prof.py`)"),
                    ("r02_api_probe.txt", "Public API of the drop-in, phases of a pass (`tools/api_probe.py`)"),
                    ("r02_bigfile.json", "16 GiB archive through `archive_read_open_filename` (`tools/bigfile_test.py`)"),
                    ("r02_sass_summary.txt", "SASS of the final kernels")):
    path = os.path.join(P, name)
    if os.path.exists(path):
        out += ["## %s - `%s`" % (title, name), "", "```", open(path).read().strip()[:6000], "```", ""]
open(os.path.join(P, "README_r02.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
