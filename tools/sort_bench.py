"""sort_by_keys microbench: n keys of `width` int32 words with `bits` random bits each, timed with CUDA
events on the library stream. Usage: python tools/sort_bench.py [n] [width] [bits] [--profile]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omega_h_b200 import Lib  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if len(args) > 0 else 50_000_000
    width = int(args[1]) if len(args) > 1 else 3
    bits = int(args[2]) if len(args) > 2 else 25
    lib = Lib(device=0).init()
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 1 << bits, size=n * width, dtype=np.int64).astype(np.int32)
    d_k = lib.to_device(keys)
    d_p = lib.empty_device(n, np.int32)
    fn = lambda: lib.check(lib.c.oshb_sort_by_keys_i32(d_k.ptr, C.c_int64(n), C.c_int(width), d_p.ptr))
    fn()
    lib.sync()
    best = 1e30
    for _ in range(3):
        lib.timer_start()
        fn()
        best = min(best, lib.timer_stop())
    if "--profile" in sys.argv:
        lib.profile_begin(None)
        fn()
        for name, ms in lib.profile_end():
            print("   %-28s %.3f ms" % (name.split("\t")[0], ms))
    npass = width * ((bits + 7) // 8)
    print("sort_by_keys n=%d width=%d bits=%d: %.3f ms, %.3f ms/pass (%d passes), %.1f GB/s per pass (16 B/key)" % (
        n, width, bits, best, best / npass, npass, 16.0 * n / 1e9 / (best / npass / 1e3)))
    if n <= 20_000_000:
        k2 = keys.reshape(n, width)
        want = np.lexsort([k2[:, k] for k in range(width - 1, -1, -1)]).astype(np.int32)
        assert np.array_equal(d_p.to_host(), want)
        print("   matches numpy lexsort")


if __name__ == "__main__":
    main()
