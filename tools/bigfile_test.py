"""GPU box: a ZIP of 16 GiB (uncompressed) read through archive_read_open_filename on the
drop-in library: throughput, peak resident memory, and per-entry reports against the
unmodified reference (file-backed sources; the archive is read in windows
through __archive_read_seek / __archive_read_ahead, never as one image).

  python tools/bigfile_test.py [out_gib=16] [dir=/tmp]
"""
import json, os, resource, subprocess, sys, time, zlib
sys.path.insert(0, os.getcwd())
from libarchive_b200 import synth

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
where = sys.argv[2] if len(sys.argv) > 2 else "/tmp"
path = os.path.join(where, "b2i_big.zip")
each = 1 << 20
n = int(gib * (1 << 30)) // each
t0 = time.time()
base = synth.text_corpus_parts(256 * each, each, 77)[:256]
comps = [synth.deflate_raw(p, 6) for p in base]
crcs = [zlib.crc32(p) & 0xFFFFFFFF for p in base]
members = [synth.ZipMember("big/%06d.txt" % i, base[i % 256], comp=comps[i % 256], crc=crcs[i % 256]) for i in range(n)]
blob = synth.make_zip(members, zip64=True, threads=1)
with open(path, "wb") as f:
    f.write(blob)
size = len(blob)
del blob, members
print("archive: %d entries, %.2f GiB out, %.2f GB file, built in %.0f s" % (n, gib, size / 1e9, time.time() - t0), flush=True)

def run(cmd, timeout=1800):
    t = time.time()
    r0 = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    rss = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    return p, time.time() - t, rss

out = {"entries": n, "out_bytes": n * each, "file_bytes": size}
p, dt, rss = run(["libarchive_b200/api_bench", path, "--file", "--mode", "block", "--steps", "1", "--warmup", "1"])   # the warm-up pass pays the CUDA context (seconds in a cold process)
j = json.loads(p.stdout.strip().splitlines()[-1])
out["dropin_open_filename_block"] = {"GBps": j["gbps_mean"], "seconds": j["seconds_mean"], "errors": j["errors"],
                                     "peak_resident_MiB": j["vm_hwm_kb"] / 1024.0,
                                     "resident_at_exit_MiB": {"anon": j["rss_anon_kb"] / 1024.0, "file": j["rss_file_kb"] / 1024.0,
                                                              "shmem_pinned": j["rss_shmem_kb"] / 1024.0},
                                     "phases": j["last_pass"]}
print(json.dumps(out["dropin_open_filename_block"]), flush=True)
p, dt, rss2 = run(["libarchive_b200/api_bench", path, "--file", "--mode", "data", "--steps", "1", "--warmup", "1"])
j = json.loads(p.stdout.strip().splitlines()[-1])
out["dropin_open_filename_data64k"] = {"GBps": j["gbps_mean"], "seconds": j["seconds_mean"], "errors": j["errors"]}
print(json.dumps(out["dropin_open_filename_data64k"]), flush=True)
# reports: drop-in vs the unmodified reference, entry by entry (names, sizes, return codes, CRC of the delivered bytes)
def listing(binary):
    p, dt, _ = run([binary, "list", path, "--file"])
    rows = [json.loads(l) for l in p.stdout.splitlines() if l.strip()]
    for r in rows:
        r.pop("nblk", None); r.pop("blocks", None)
    return rows, dt
if os.path.exists("oracle/_ref/oracle_extract") and "--no-ref" not in sys.argv:
    a, ta = listing("oracle/_ref/oracle_extract")
    b, tb = listing("libarchive_b200/dropin_extract")
    out["reports_identical_to_reference"] = (a == b)
    out["entries_compared"] = len(a)
    out["reference_seconds"], out["dropin_list_seconds"] = ta, tb
    out["reference_GBps_one_thread"] = n * each / ta / 1e9
print(json.dumps(out))
os.unlink(path)
