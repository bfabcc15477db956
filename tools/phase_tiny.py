"""Phase clocks of the inflate kernel on config-5-shaped input (4 KiB text entries).
B2I_LIB=libarchive_b200/libb200inflate_prof.so python tools/phase_tiny.py"""
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = [sys.argv[0]]
import importlib.util
spec = importlib.util.spec_from_file_location("pp", os.path.join(os.path.dirname(__file__), "phase_probe.py"))
src = open(spec.origin).read().split("txt = synth.synth_text")[0]
g = {"__name__": "pp"}
exec(compile(src, spec.origin, "exec"), g)
synth = g["synth"]
n = int(os.environ.get("N", "65536"))
parts = synth.split_text(2048 * 4096, 4096, 5)
e = [synth.deflate_raw(p, 6) for p in parts]
g["run"]("tiny4k", [e[i % 2048] for i in range(n)], [4096] * n, reps=3)
parts = synth.split_text(256 * 65536, 65536, 5)
e = [synth.deflate_raw(p, 6) for p in parts]
g["run"]("text64k", [e[i % 256] for i in range(4096)], [65536] * 4096, reps=3)
