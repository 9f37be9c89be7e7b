"""Runs the 64^3 refine loop once warm and once more; prints the library's kernel-launch
counter before and after the last loop so `ncu -s <skip> -c <count>` can target exactly it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from omega_h_b200 import AdaptOpts, Lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = Lib(device=0).init()
base = bench.build_input(n, lib)
opts = AdaptOpts(base)
m = base.copy()
bench.run_loop(m, opts)
lib.sync()
a = lib.launch_count()
m = base.copy()
bench.run_loop(m, opts)
lib.sync()
b = lib.launch_count()
print("SKIP %d COUNT %d" % (a, b - a))
