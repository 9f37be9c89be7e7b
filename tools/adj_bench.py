"""Config 5 of BASELINE.json: adjacency-derivation microbench on a synthetic box mesh
(default 256x256x254 hexes x 6 = 99.9 M tets): invert_adj (R->V, R->F->up, F->E->up),
reflect_down (R->F), transit, offset_scan, sort_by_keys, each timed with CUDA events on the
library stream, reported as algorithmic GB/s (SURVEY.md section 8d byte counts) against the
measured HBM peak. Usage: python tools/adj_bench.py [nx ny nz] [--json]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omega_h_b200 import Lib, build_box  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    nx, ny, nz = (int(args[0]), int(args[1]), int(args[2])) if len(args) >= 3 else (256, 256, 254)
    lib = Lib(device=0).init()
    out = run(lib, nx, ny, nz, profile="--profile" in sys.argv, log=sys.stderr)
    if "--json" in sys.argv:
        print(json.dumps(out))


def run(lib, nx=256, ny=256, nz=254, profile=False, log=None):
    """the microbench as a function (bench.py emits its result as `also.adj_100M`)"""
    class _Null:
        def write(self, *_):
            pass
    if log is None:
        log = _Null()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    lib.timer_start()
    m = build_box(1.0, 1.0, 1.0, nx, ny, nz, lib=lib)
    t_build = lib.timer_stop()
    nv, ne, nf, nr = (m.nents(d) for d in range(4))
    print("build_box %dx%dx%d: V %d E %d F %d R %d in %.1f ms" % (nx, ny, nz, nv, ne, nf, nr, t_build), file=log)
    c = lib.c
    out = {"mesh": {"nverts": nv, "nedges": ne, "nfaces": nf, "ntets": nr, "build_box_ms": t_build}, "peak_gbs": peak,
           "kernels": {}}

    def dev_of_adj(from_dim, to_dim):
        deg = {(3, 2): 4, (3, 0): 4, (2, 1): 3, (3, 1): 6, (2, 0): 3, (1, 0): 2}[(from_dim, to_dim)]
        n = m.nents(from_dim) * deg
        d_ab = lib.empty_device(n, np.int32)
        d_co = lib.empty_device(n, np.int8) if to_dim > 0 else None
        lib.check(c.oshb_mesh_ask_down(m.h, C.c_int(from_dim), C.c_int(to_dim), d_ab.ptr, d_co.ptr if d_co else None, C.c_int(0)))
        return d_ab, d_co, deg

    def timeit(fn, reps=5):
        fn()
        lib.sync()
        best = 1e30
        for _ in range(reps):
            lib.timer_start()
            fn()
            best = min(best, lib.timer_stop())
        if profile:
            # per-kernel split of one more call (CUDA events around every launch)
            lib.profile_begin(None)
            fn()
            agg = {}
            for name, ms in lib.profile_end():
                k = name.split("\t")[0]
                agg[k] = agg.get(k, 0.0) + ms
            print("    " + "  ".join("%s %.2f" % kv for kv in sorted(agg.items(), key=lambda kv: -kv[1])), file=log)
        return best

    def report(name, ms, nbytes):
        gbs = nbytes / 1e9 / (ms / 1e3)
        out["kernels"][name] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": gbs, "frac_of_peak": gbs / peak}
        print("%-34s %9.3f ms %9.1f GB/s  %5.1f%% of %.0f" % (name, ms, gbs, 100 * gbs / peak, peak), file=log)

    # ---- invert_adj -------------------------------------------------------------------------
    for (hd, ld) in ((3, 0), (3, 2), (2, 1), (3, 1)):
        d_ab, d_co, deg = dev_of_adj(hd, ld)
        nh, nl = m.nents(hd), m.nents(ld)
        d_off = lib.empty_device(nl + 1, np.int32)
        d_up = lib.empty_device(nh * deg, np.int32)
        d_uc = lib.empty_device(nh * deg, np.int8)
        ms = timeit(lambda: lib.check(c.oshb_invert_adj(d_ab.ptr, d_co.ptr if d_co else None, C.c_int64(nh), C.c_int(deg),
                                                        C.c_int32(nl), d_off.ptr, d_up.ptr, d_uc.ptr)))
        nbytes = (9 + (1 if d_co else 0)) * nh * deg + 4 * nl
        report("invert_adj %d->%d (N=%d d=%d)" % (hd, ld, nh, deg), ms, nbytes)
        del d_ab, d_co, d_off, d_up, d_uc
    # ---- reflect_down R->F ---------------------------------------------------------------------
    d_rv, _, _ = dev_of_adj(3, 0)
    d_fv, _, _ = dev_of_adj(2, 0)
    d_hl = lib.empty_device(nr * 4, np.int32)
    d_hc = lib.empty_device(nr * 4, np.int8)
    ms = timeit(lambda: lib.check(c.oshb_reflect_down(d_rv.ptr, C.c_int64(nr), C.c_int(3), d_fv.ptr, C.c_int64(nf), C.c_int(2),
                                                      C.c_int32(nv), d_hl.ptr, d_hc.ptr)), reps=3)
    report("reflect_down 3->2", ms, 4 * nr * 4 + 4 * nf * 3 + 5 * nr * 4)
    del d_hl, d_hc
    # ---- offset_scan / sort_by_keys ---------------------------------------------------------------
    n = nr * 4
    d_in8 = lib.to_device(np.ones(n, dtype=np.int8))
    d_out = lib.empty_device(n + 1, np.int32)
    ms = timeit(lambda: lib.check(c.oshb_offset_scan_i8(d_in8.ptr, C.c_int64(n), d_out.ptr)))
    report("offset_scan i8 (n=%d)" % n, ms, n * 1 + (n + 1) * 4)
    ms = timeit(lambda: lib.check(c.oshb_offset_scan_i32(d_rv.ptr, C.c_int64(n), d_out.ptr)))
    report("offset_scan i32 (n=%d)" % n, ms, n * 4 + (n + 1) * 4)
    del d_in8, d_out
    d_perm = lib.empty_device(nf, np.int32)
    ms = timeit(lambda: lib.check(c.oshb_sort_by_keys_i32(d_fv.ptr, C.c_int64(nf), C.c_int(3), d_perm.ptr)), reps=2)
    report("sort_by_keys i32 width 3 (n=%d)" % nf, ms, nf * (3 * 4 + 4))
    # a radix sort moves (word, index) pairs once per 8-bit digit: also report the rate per pass
    nbits = max(1, int(nv - 1).bit_length())
    npass = 3 * ((nbits + 7) // 8)
    k = out["kernels"]["sort_by_keys i32 width 3 (n=%d)" % nf]
    k["passes"] = npass
    k["per_pass_gbs_16B_per_key"] = 16.0 * nf / 1e9 / (ms / npass / 1e3)
    k["per_pass_frac_of_peak"] = k["per_pass_gbs_16B_per_key"] / peak
    return out


if __name__ == "__main__":
    main()
