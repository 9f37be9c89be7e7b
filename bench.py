#!/usr/bin/env python
"""bench.py -- new tets/s of the metric-driven refine loop (BASELINE.json metric).

One "step" = the whole `while (refine_by_size(&mesh, opts))` loop on one freshly built box
mesh: config[1] of BASELINE.json, `build_box 64^3 (x6 tets)` with the uniform isotropic
metric h = 1/(2n) (4 doubling passes, 1,572,864 -> 25,165,824 tets, 23,592,960 new tets).

  value : new tets / device time of the loops, input mesh resident in HBM (CUDA events on the
          library's stream, max over ranks)
  e2e   : the same loop driven through the public host-buffer API: every step uploads the
          input mesh from pinned host arrays (Mesh.set_ents/add_tag), runs the loop and reads
          the whole refined mesh back into pinned host arrays
  roofline : the dominant kernel, timed live with CUDA events inside the timed region,
          algorithmic bytes declared by its call site / measured HBM peak
  cpu_baseline / --impl reference : the UNMODIFIED reference (oracle/_ref, OpenMP build) on the
          box's host cores, bounded sample of the same workload

N > 1: ONE box of N x n^3 cells partitioned over the ranks (omega_h_b200/dist.py, DESIGN.md
"Multi-GPU"): weak scaling, value = global new tets / max-over-ranks time. --replicas: N
independent boxes instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "new tets/sec, adapt() refine pass"
UNIT = "tets/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"],
                    help="reference = the unmodified reference on the host cores (OpenMP); reference-cuda = the reference's "
                         "own CUDA backend (OMEGA_H_USE_CUDA, built by `make -C oracle refcuda`) on this GPU")
    ap.add_argument("--n", "--box-n", dest="n", type=int, default=64,
                    help="box cells per axis and GPU (config[1] = 64); under torchrun spell it --box-n")
    ap.add_argument("--workload", default="iso", choices=["iso", "aniso"],
                    help="iso = BASELINE config[1] (the metric's configuration); aniso = config[2] tanh shock layer")
    ap.add_argument("--profile", action="store_true", help="print a per-kernel time table to stderr and exit")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--parity-halo", type=int, default=2,
                    help="N > 1: halo of the N-rank == serial self-check (default 2: below its 4 passes, so the library's "
                         "re-ghosting over NCCL is part of what is checked; 0 = --halo)")
    ap.add_argument("--no-time-reghost", dest="time_reghost", action="store_false",
                    help="N > 1: do not time the library's re-ghosting (oshb_dist_reghost) of the refined anisotropic part "
                         "(twice: first call, warm) after the `also` loop")
    ap.add_argument("--time-reghost", dest="time_reghost", action="store_true", help=argparse.SUPPRESS)
    ap.set_defaults(time_reghost=True)
    ap.add_argument("--no-also", action="store_true", help="skip the `also` block (N = 1: aniso n=64 loop + 100 M-tet adjacency microbench; N > 1: the "
                         "anisotropic partitioned loop)")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the N-rank == serial self-check")
    ap.add_argument("--halo", type=int, default=4, help="N > 1: element layers each part keeps of its neighbours")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent boxes instead of one partitioned box")
    ap.add_argument("--parting", default="rib", choices=["rib", "hilbert"],
                    help="N > 1: parts by recursive inertial bisection (Mesh::balance, the reference's partitioner) or by "
                         "contiguous ranges of the Hilbert element order")
    return ap.parse_args()


def workload_config(n, workload="iso"):
    if workload == "aniso":
        return {
            "workload": "3D tet unit cube build_box %d^3 (x6 tets), anisotropic tanh shock-layer metric (3x3, "
                        "hx=1/n, hy=0.7/n, hz=(1/n)(1-0.75 sech^2(20(z-1/2)))), while(refine_by_size) loop with "
                        "field transfer" % n,
            "box_n": n, "metric": "anisotropic ncomps=6",
            "l2": "inputs larger than L2; no explicit flush",
        }
    return {
        "workload": "3D tet unit cube build_box %d^3 (x6 tets), uniform isotropic metric h=1/(2n), "
                    "while(refine_by_size) loop (4 doubling passes, x16 elements)" % n,
        "box_n": n,
        "metric": "isotropic ncomps=1",
        "l2": "inputs larger than L2 (%.0f MB input mesh, GB-scale outputs); no explicit flush" % (n ** 3 * 6 * 145 / 1e6),
    }


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        # "under load": the upper half of the samples (the loop is short, idle samples drag the median)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference on the host cores
# ---------------------------------------------------------------------------------------------
def run_reference_cuda_loops(n, kind, reps, budget_s, device=0):
    """The reference's OWN CUDA backend (generic kernels + Thrust/CUB, src/Omega_h_for.hpp:16-58) compiled for sm_100
    from the unmodified sources (oracle/Makefile `refcuda`), running the same complete loop on this GPU."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver_cuda")
    if not os.path.exists(exe):
        return None
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = str(device)
    out = subprocess.run([exe, "timeloops", "3", str(n), str(kind), str(reps), str(budget_s)], capture_output=True,
                         text=True, env=env)
    if out.returncode != 0:
        return None
    recs = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]
    return recs or None


def gpu_baseline_sample(n, workload="iso", device=0):
    """gpu_baseline of the N = 1 line: 1 warm-up + 2 timed complete loops by the reference's own CUDA backend"""
    recs = run_reference_cuda_loops(n, 2 if workload == "aniso" else 0, 3, 120, device)
    if not recs:
        return None
    timed = recs[1:] if len(recs) > 1 else recs
    new = timed[0]["nelems_after"] - timed[0]["nelems_before"]
    secs = sum(r["seconds"] for r in timed) / len(timed)
    return {"value": new / secs, "unit": UNIT, "kind": "reference-cuda", "ms_per_step": secs * 1e3,
            "sample": "%d complete %d^3 loops (%d passes, %d -> %d tets) after one warm-up, the reference's own CUDA backend "
                      "(OMEGA_H_USE_CUDA, unmodified sources, nvcc sm_100, oracle/_ref/ref_driver_cuda) on this GPU" % (
                          len(timed), n, timed[0]["passes"], timed[0]["nelems_before"], timed[0]["nelems_after"])}


def run_reference_loops(n, kind, reps, budget_s, omp=True):
    """The UNMODIFIED reference (oracle/_ref, built by oracle/Makefile) running the WHOLE
    `while (refine_by_size)` loop `reps` times on copies of one input mesh, all host cores; one record
    per repetition. Stops early after budget_s seconds of wall time."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver_omp" if omp else "ref_driver")
    if not os.path.exists(exe):
        return None
    env = dict(os.environ)
    cores = os.cpu_count() or 1
    env["OMP_NUM_THREADS"] = str(cores)
    env["OMP_PROC_BIND"] = "false"
    out = subprocess.run([exe, "timeloops", "3", str(n), str(kind), str(reps), str(budget_s)], capture_output=True,
                         text=True, env=env)
    if out.returncode != 0:
        return None
    recs = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]
    return recs or None


def reference_sample(n, workload="iso"):
    """cpu_baseline of the N = 1 line: ONE complete loop of the same workload (same box, same metric,
    all passes) by the unmodified reference on all host cores."""
    recs = run_reference_loops(n, 2 if workload == "aniso" else 0, 1, 1e9)
    if not recs:
        return None
    r = recs[-1]
    new = r["nelems_after"] - r["nelems_before"]
    return {"value": new / r["seconds"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
            "sample": "one complete %d^3 loop (%d passes, %d -> %d tets) in %.2f s, OpenMP build of the unmodified "
                      "reference (oracle/_ref), %d threads" % (n, r["passes"], r["nelems_before"], r["nelems_after"],
                                                               r["seconds"], r["threads"]),
            "seconds": r["seconds"]}


def main_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, the same config, metric and
    unit; a step = the complete loop (all passes), W warm-up + K timed repetitions as asked, cut short
    only if that would exceed REF_BUDGET_S of wall time (then `steps` says how many were timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    budget = float(os.environ.get("OSHB_REF_BUDGET_S", "330"))
    kind = 2 if args.workload == "aniso" else 0
    cuda = args.impl == "reference-cuda"
    if cuda:
        recs = run_reference_cuda_loops(args.n, kind, args.warmup + args.steps, budget, int(os.environ.get("LOCAL_RANK", "0")))
    else:
        recs = run_reference_loops(args.n, kind, args.warmup + args.steps, budget)
    if not recs:
        emit({"impl": args.impl, "unavailable": "oracle/_ref/%s missing or failed" %
              ("ref_driver_cuda (make -C oracle refcuda)" if cuda else "ref_driver_omp")})
        return 0
    warm = min(args.warmup, max(len(recs) - 1, 0))
    timed = recs[warm:]
    secs = sum(r["seconds"] for r in timed) / len(timed)
    new = timed[0]["nelems_after"] - timed[0]["nelems_before"]
    value = new * len(timed) / sum(r["seconds"] for r in timed)
    cfg = workload_config(args.n, args.workload)
    cfg["passes_per_step"] = timed[0]["passes"]
    cfg["tets_per_step"] = "%d -> %d" % (timed[0]["nelems_before"], timed[0]["nelems_after"])
    cfg["parallelism"] = ("1 GPU, the reference's own CUDA backend (OMEGA_H_USE_CUDA)" if cuda else
                          "host CPU, OpenMP, %d threads" % timed[0]["threads"])
    cfg["input_caches"] = "length + quality tags measured before the timed loop (ask_lengths/ask_qualities), as in our arm"
    line = {
        "impl": args.impl, "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(timed), "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": secs * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": timed[0]["threads"], "kind": "reference",
                         "sample": "%d complete loops (%d passes each, %d -> %d tets), %.2f s per loop, %s build of the "
                                   "unmodified reference (oracle/_ref)" % (len(timed), timed[0]["passes"],
                                                                          timed[0]["nelems_before"], timed[0]["nelems_after"], secs,
                                                                          "CUDA (sm_100)" if cuda else "OpenMP")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def build_input(n, lib, workload="iso"):
    import numpy as np
    from omega_h_b200 import VERT, build_box
    m = build_box(1.0, 1.0, 1.0, n, n, n, lib=lib)
    if workload == "aniso":
        # BASELINE config[2]: tanh shock layer at z = 1/2, hx = 1/n, hy = 0.7/n (distinct
        # eigenvalues), hz = (1/n) (1 - 0.75 sech^2(20 (z - 1/2))); metric = diag(1/h^2)
        x = m.coords().reshape(-1, 3)
        t = np.tanh(20.0 * (x[:, 2] - 0.5))
        hz = (1.0 / n) * (1.0 - 0.75 * (1.0 - t * t))
        met = np.zeros((m.nverts(), 6))
        met[:, 0] = 1.0 / (1.0 / n) ** 2
        met[:, 1] = 1.0 / (0.7 / n) ** 2
        met[:, 2] = 1.0 / hz ** 2
        m.add_tag(VERT, "metric", 6, met.reshape(-1))
    else:
        h = 1.0 / n / 2.0
        m.add_tag(VERT, "metric", 1, np.full(m.nverts(), 1.0 / (h * h)))
    if os.environ.get("OSHB_BENCH_EXTRA_TAG"):
        # diagnosis only: what one more inherited int32 tag on every dimension costs the rebuild
        # (the partitioned path carries "own:part")
        for d in range(4):
            m.add_tag(d, "own:part", 1, np.zeros(m.nents(d), dtype=np.int32))
    m.ask_lengths()
    m.ask_qualities()
    lib.sync()
    return m


def run_loop(m, opts):
    from omega_h_b200 import refine_by_size
    passes = 0
    while refine_by_size(m, opts):
        passes += 1
    return passes


def host_copy_of(m, lib):
    """Pinned host image of a mesh: what a caller that keeps its mesh on the host hands over."""
    dim = m.dim()
    img = {"dim": dim, "nverts": m.nverts(), "down": {}, "tags": {}}
    nbytes = 0
    for d in range(1, dim + 1):
        ab2b, codes = m.ask_down(d, d - 1)
        pa = lib.pinned_empty(ab2b.size, ab2b.dtype)
        pa[:] = ab2b
        pc = None
        if codes is not None:
            pc = lib.pinned_empty(codes.size, codes.dtype)
            pc[:] = codes
            nbytes += pc.nbytes
        img["down"][d] = (pa, pc)
        nbytes += pa.nbytes
    for d in range(dim + 1):
        for name, _, nc in m.tags(d):
            a = m.get_array(d, name)
            pa = lib.pinned_empty(a.size, a.dtype)
            pa[:] = a
            img["tags"][(d, name)] = (nc, pa)
            nbytes += pa.nbytes
    img["nbytes"] = nbytes
    return img


def upload(img, lib):
    from omega_h_b200 import Mesh
    m = Mesh(img["dim"], lib=lib)
    m.set_verts(img["nverts"])
    for d, (pa, pc) in img["down"].items():
        m.set_ents(d, pa, pc)
    for (d, name), (nc, pa) in img["tags"].items():
        m.add_tag(d, name, nc, pa, internal=True)
    return m


class Download:
    """Reads the whole refined mesh (downward adjacencies + every tag) into reusable pinned buffers."""

    def __init__(self, lib):
        self.lib = lib
        self.bufs = {}

    def buf(self, key, n, dtype):
        b = self.bufs.get(key)
        if b is None or b.size < n:
            b = self.lib.pinned_empty(int(n * 1.05) + 16, dtype)
            self.bufs[key] = b
        return b

    def __call__(self, m):
        import numpy as np
        from omega_h_b200 import simplex_degree
        from omega_h_b200._lib import NP_OF
        nbytes = 0
        dim = m.dim()
        for d in range(1, dim + 1):
            n = m.nents(d) * simplex_degree(d, d - 1)
            ob = self.buf(("down", d), n, np.int32)
            oc = self.buf(("codes", d), n, np.int8) if d > 1 else None
            m.ask_down(d, d - 1, out=ob, out_codes=oc)
            nbytes += n * 4 + (n if d > 1 else 0)
        for d in range(dim + 1):
            for name, t, nc in m.tags(d):
                n = m.nents(d) * nc
                ob = self.buf(("tag", d, name), n, NP_OF[t])
                m.get_array(d, name, out=ob)
                nbytes += n * np.dtype(NP_OF[t]).itemsize
        return nbytes


def also_aniso(lib, n=64, steps=3):
    """BASELINE config[2]'s shape on one GPU (anisotropic tanh-layer 3x3 metric, field transfer), n^3 box:
    device-timed loops after the headline region, so the anisotropic rate is in the driver's record too."""
    from omega_h_b200 import AdaptOpts
    base = build_input(n, lib, "aniso")
    opts = AdaptOpts(base)
    n0 = base.nelems()
    for _ in range(2):
        m = base.copy()
        run_loop(m, opts)
    lib.sync()
    ms = 0.0
    for _ in range(steps):
        m = base.copy()
        lib.timer_start()
        passes = run_loop(m, opts)
        ms += lib.timer_stop()
    n1 = m.nelems()
    return {"workload": workload_config(n, "aniso")["workload"], "value": (n1 - n0) * steps / (ms / 1e3), "unit": UNIT,
            "ms_per_step": ms / steps, "steps": steps, "passes_per_step": passes, "tets_per_step": "%d -> %d" % (n0, n1),
            "timing": "CUDA events on the library stream, input resident in HBM"}


def also_adj(lib):
    """BASELINE config[4]: adjacency-derivation microbench on a 100 M-tet box (tools/adj_bench.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import adj_bench
    out = adj_bench.run(lib, 256, 256, 254)
    return {"mesh": out["mesh"], "peak_gbs": out["peak_gbs"],
            "bytes": "algorithmic (SURVEY.md 8d): compulsory reads + writes of each primitive",
            "kernels": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                        for k, v in out["kernels"].items()}}


def aniso_metric_of_box(m, n):
    """BASELINE config[2]/[3]'s tanh shock layer, one layer per unit cube of the box (z taken modulo 1) so that every
    rank's share of an N-cube box is the N = 1 workload: hx = 1/n, hy = 0.7/n, hz = (1/n)(1 - 0.75 sech^2(20(z' - 1/2)))"""
    import numpy as np
    x = m.coords().reshape(-1, 3)
    z = x[:, 2] - np.floor(x[:, 2])
    z[x[:, 2] == np.floor(x[:, 2])] = 0.0
    t = np.tanh(20.0 * (z - 0.5))
    hz = (1.0 / n) * (1.0 - 0.75 * (1.0 - t * t))
    met = np.zeros((m.nverts(), 6))
    met[:, 0] = 1.0 / (1.0 / n) ** 2
    met[:, 1] = 1.0 / (0.7 / n) ** 2
    met[:, 2] = 1.0 / hz ** 2
    return met.reshape(-1)


def also_aniso_partitioned(lib, device, n, halo, parting, steps=3, time_reghost=False):
    """BASELINE config[3] (the anisotropic cube partitioned over the ranks, metric + coordinates + classification
    transferred every pass): N x n^3 cells, one shock layer per unit cube, RIB parts. The loop has 8 passes; it runs on
    an 8-layer halo so that no re-ghosting falls into it; a re-ghosting (DistMesh.reghost = oshb_dist_reghost) of the
    refined part is timed separately (first call + warm): a loop on a thinner halo pays that once per `halo` passes."""
    import torch
    import torch.distributed as dist
    from omega_h_b200 import VERT, AdaptOpts, build_box
    from omega_h_b200 import dist as D
    world = dist.get_world_size()
    shape = box_shape(world)
    base = build_box(float(shape[0]), float(shape[1]), float(shape[2]), shape[0] * n, shape[1] * n, shape[2] * n, lib=lib)
    base.add_tag(VERT, "metric", 6, aniso_metric_of_box(base, n))
    base.ask_lengths()
    base.ask_qualities()
    nglobal0 = base.nelems()
    halo = max(halo, 8)
    part0 = D.distribute(base, halo, device, parting=parting)
    del base
    torch.cuda.empty_cache()
    opts = AdaptOpts(part0.mesh)

    def barrier():
        lib.sync()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def loop(part):
        passes = 0
        while part.refine_by_size(opts):
            passes += 1
        return passes
    for _ in range(2):
        part = part0.clone()
        npasses = loop(part)
    t = torch.tensor([float(part.owned_nelems())], device=device, dtype=torch.float64)
    dist.all_reduce(t)
    nglobal1 = int(t.item())
    reghosts = int(getattr(part, "nreghosts", 0))
    del part
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        part = part0.clone()
        loop(part)
    lib.sync()
    ev1.record()
    barrier()
    tm = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm[0]) / steps
    # one re-ghosting of the refined part (what a loop on a halo thinner than its pass count pays per `halo` passes)
    local_tets = part.mesh.nelems()
    reghost_ms = [None, None]
    for k in range(2 if time_reghost else 0):   # the first call also sets up NCCL's point-to-point connections
        barrier()
        t0 = time.perf_counter()
        part.reghost()
        barrier()
        tr = torch.tensor([(time.perf_counter() - t0) * 1e3], device=device, dtype=torch.float64)
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        reghost_ms[k] = float(tr[0])
    del part, part0
    torch.cuda.empty_cache()
    lib.trim()
    return {"workload": "3D tet box %dx%dx%d cells (x6 tets) = %d ranks x %d^3, anisotropic tanh shock-layer metric (3x3; one "
                        "layer per unit cube), while(refine_by_size) loop on the partitioned mesh with field transfer" %
                        (shape[0] * n, shape[1] * n, shape[2] * n, world, n),
            "value": (nglobal1 - nglobal0) / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "passes_per_step": npasses, "tets_per_step": "%d -> %d" % (nglobal0, nglobal1), "halo": halo,
            "reghosts_per_step": reghosts, "parting": parting,
            "reghost_ms": reghost_ms[1], "reghost_first_call_ms": reghost_ms[0], "reghost_local_tets_rank0": int(local_tets),
            "timing": "CUDA events between two barriers, max over ranks; reghost_ms: wall clock of one DistMesh.reghost of "
                      "the refined part between two barriers, max over ranks, outside the loop's timed region"}


ALSO_TIMEOUT_S = 300


def run_guarded(seconds, on_timeout, fn):
    """fn() with a dead-man's switch: if it has not returned after `seconds`, on_timeout() runs on a timer thread
    (bench.py uses it to print the headline line and end the process when the optional `also` block hangs)"""
    import threading
    done = threading.Event()

    def fire():
        if not done.is_set():
            on_timeout()
    timer = threading.Timer(seconds, fire)
    timer.daemon = True
    timer.start()
    try:
        return fn()
    finally:
        done.set()
        timer.cancel()


def partition_parity_check(lib, device, halo, n=16):
    """N-rank == serial, checked on the GPUs of this very run before anything is timed: an n^3-per-rank box
    goes through the partitioned loop (NCCL exchanges and all), is assembled on rank 0 (DistMesh.gather) and
    compared with the serial loop's mesh ARRAY BY ARRAY -- connectivity, codes, global numbers, every tag.
    A mismatch aborts the benchmark."""
    import numpy as np
    import torch.distributed as dist
    from omega_h_b200 import VERT, build_box, refine_by_size, AdaptOpts, simplex_degree
    from omega_h_b200 import dist as D
    world, rank = dist.get_world_size(), dist.get_rank()
    shape = box_shape(world)
    base = build_box(float(shape[0]), float(shape[1]), float(shape[2]), shape[0] * n, shape[1] * n, shape[2] * n, lib=lib)
    h = 1.0 / n / 2.0
    base.add_tag(VERT, "metric", 1, np.full(base.nverts(), 1.0 / (h * h)))
    serial = None
    spasses = 0
    if rank == 0:
        serial = base.copy()
        sopts = AdaptOpts(serial)
        while refine_by_size(serial, sopts):
            spasses += 1
    part = D.distribute(base, halo, device, parting="rib" if (world & (world - 1)) == 0 else "hilbert")
    opts = AdaptOpts(part.mesh)
    passes = 0
    while part.refine_by_size(opts):
        passes += 1
    whole = part.gather(0)
    res = {"ok": True, "ranks": world, "box": "%dx%dx%d cells" % (shape[0] * n, shape[1] * n, shape[2] * n),
           "halo": halo, "passes": passes, "reghosts": int(getattr(part, "nreghosts", 0)), "arrays_compared": 0}
    if rank == 0:
        bad = []
        if passes != spasses:
            bad.append("passes %d vs %d" % (passes, spasses))
        dim = serial.dim()
        for d in range(dim + 1):
            if whole.nents(d) != serial.nents(d):
                bad.append("nents(%d)" % d)
                continue
            for name, _, _ in serial.tags(d):
                res["arrays_compared"] += 1
                if not np.array_equal(serial.get_array(d, name), whole.get_array(d, name)):
                    bad.append("tag %d:%s" % (d, name))
            if d >= 1:
                a, ac = serial.ask_down(d, d - 1)
                b, bc = whole.ask_down(d, d - 1)
                res["arrays_compared"] += 1
                if not np.array_equal(a, b):
                    bad.append("down %d" % d)
                if d >= 2:
                    res["arrays_compared"] += 1
                    if not np.array_equal(ac, bc):
                        bad.append("codes %d" % d)
        res["tets"] = int(serial.nelems())
        res["compared"] = "every tag (reals bit for bit), downward adjacency and alignment codes of every dimension"
        if bad:
            res["ok"] = False
            res["mismatch"] = bad
    flag = __import__("torch").tensor([1 if res["ok"] else 0], device=device)
    dist.broadcast(flag, 0)
    if int(flag.item()) != 1:
        if rank == 0:
            emit({"parity_check": res})
        raise SystemExit("partitioned result differs from the serial loop: refusing to time it")
    return res


def summarize_profile(recs):
    agg = {}
    for name, ms in recs:
        parts = name.split("\t")
        nm = parts[0]
        b = int(parts[1]) if len(parts) > 1 else 0
        a = agg.setdefault(nm, [0, 0.0, 0])
        a[0] += 1
        a[1] += ms
        a[2] += b
    return agg


def main_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from omega_h_b200 import AdaptOpts, Lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = Lib(device=local_rank).init()
    if lib.is_emulation:
        raise SystemExit("bench.py refuses to run on the emulation build")

    def barrier():
        lib.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    base = build_input(args.n, lib, args.workload)
    opts = AdaptOpts(base)
    nelems0 = base.nelems()

    if args.profile:
        for _ in range(2):
            m = base.copy()
            run_loop(m, opts)
        lib.sync()
        import ctypes as C

        def host_stats():
            a, f, s_, n = C.c_double(), C.c_double(), C.c_double(), C.c_uint64()
            lib.c.oshb_host_time_stats(C.byref(a), C.byref(f), C.byref(s_), C.byref(n))
            return a.value, f.value, s_.value, n.value, lib.sync_count(), lib.launch_count()
        # un-profiled loop with host-side accounting
        m = base.copy()
        h0 = host_stats()
        t0 = time.perf_counter()
        lib.timer_start()
        run_loop(m, opts)
        tot = lib.timer_stop()
        t1 = time.perf_counter()
        h1 = host_stats()
        print("plain loop: %.3f ms device, %.3f ms wall; host time in alloc %.3f ms (%d allocs), free %.3f ms, "
              "blocking read-backs %.3f ms (%d), launches %d" %
              (tot, (t1 - t0) * 1e3, (h1[0] - h0[0]) * 1e3, h1[3] - h0[3], (h1[1] - h0[1]) * 1e3,
               (h1[2] - h0[2]) * 1e3, h1[4] - h0[4], h1[5] - h0[5]), file=sys.stderr)
        m = base.copy()
        lib.profile_begin(None)
        lib.timer_start()
        run_loop(m, opts)
        total = lib.timer_stop()
        recs = lib.profile_end()
        agg = summarize_profile(recs)
        if os.environ.get("OSHB_PROFILE_LAUNCHES"):
            want = os.environ["OSHB_PROFILE_LAUNCHES"].split(",")
            for name, ms in recs:
                nm = name.split("\t")[0]
                if nm in want:
                    print("  launch %-28s %9.3f ms  algo bytes %s" % (nm, ms, name.split("\t")[1]), file=sys.stderr)
        ksum = sum(v[1] for v in agg.values())
        print("loop %.3f ms (profiled), kernels %.3f ms, %d launches" % (total, ksum, sum(v[0] for v in agg.values())),
              file=sys.stderr)
        for nm, (cnt, ms, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            gbs = (b / 1e9) / (ms / 1e3) if b and ms > 0 else 0.0
            print("%-36s n=%4d %9.3f ms %5.1f%% %8.1f GB/s(algo)" % (nm, cnt, ms, 100 * ms / ksum, gbs), file=sys.stderr)
        return 0

    # ---- pick the dominant kernel (outside the timed region) --------------------------------
    for _ in range(max(args.warmup, 3)):
        m = base.copy()
        run_loop(m, opts)
    lib.sync()
    m = base.copy()
    lib.profile_begin(None)
    run_loop(m, opts)
    agg = summarize_profile(lib.profile_end())
    top = max(agg.items(), key=lambda kv: kv[1][1])[0]
    del m

    # ---- timed region: device-resident input -----------------------------------------------
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = lib.launch_count()
    lib.profile_begin(top)
    new_tets = 0
    dev_ms = 0.0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        m = base.copy()
        lib.timer_start()
        npasses = run_loop(m, opts)
        dev_ms += lib.timer_stop()
        new_tets += m.nelems() - nelems0
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    top_recs = lib.profile_end()
    launches = lib.launch_count() - launches0
    clocks = sampler.stop()
    nelems1 = m.nelems()
    del m

    if world > 1:
        t = torch.tensor([dev_ms, float(new_tets)], device="cuda", dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms_max = float(tmax[0])
        total_new = float(tsum[1])
    else:
        dev_ms_max, total_new = dev_ms, float(new_tets)
    value = total_new / (dev_ms_max / 1e3)

    # roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    tagg = summarize_profile(top_recs)
    cnt, tms, tbytes = tagg.get(top, (0, 0.0, 0))
    achieved = (tbytes / 1e9) / (tms / 1e3) if tms > 0 and tbytes else None
    roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                "launches": cnt, "avg_ms": (tms / cnt) if cnt else None,
                "share_of_step": (tms / dev_ms) if dev_ms > 0 else None,
                "algorithmic_bytes_per_launch": (tbytes / cnt) if cnt else None}
    # DRAM traffic per launch of that kernel from the committed `ncu --set full` capture of this
    # same loop (profiles/ncu_r1b_gather.json); only valid for the configuration it was taken on
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "ncu_r1b_gather.json")))
        if cap["kernel"] == top and args.n == 64 and args.workload == "iso":
            roofline["traffic"] = cap["dram_bytes_per_launch"]
            roofline["traffic_source"] = "profiles/ncu_r1b_gather.json (dram__bytes_read.sum + dram__bytes_write.sum, avg of the loop's 16 launches)"
    except Exception:
        pass

    # ---- end to end: host buffers in, host buffers out ---------------------------------------
    e2e = None
    if not args.no_e2e:
        img = host_copy_of(base, lib)
        down = Download(lib)
        for _ in range(2):
            m = upload(img, lib)
            run_loop(m, opts)
            d2h_bytes = down(m)
            del m
        barrier()
        t0 = time.perf_counter()
        e_new = 0
        for _ in range(args.steps):
            m = upload(img, lib)
            run_loop(m, opts)
            d2h_bytes = down(m)
            e_new += m.nelems() - nelems0
            del m
        barrier()
        e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_s = float(t[0])
            t2 = torch.tensor([float(e_new)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t2, op=dist.ReduceOp.SUM)
            e_new = float(t2[0])
        e2e = {"value": e_new / e_s, "unit": UNIT, "h2d_bytes_per_step": int(img["nbytes"]),
               "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": e_s * 1e3 / args.steps,
               "timing": "wall clock around upload + loop + download, stream synchronised, max over ranks"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s = reference_sample(args.n, args.workload)
        if s:
            cpu = {k: s[k] for k in ("value", "unit", "cores", "kind", "sample")}

    gpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gpu_base = gpu_baseline_sample(args.n, args.workload, local_rank)
    also = None
    if rank == 0 and world == 1 and not args.no_also and args.workload == "iso":
        # after the headline regions: the other single-GPU configurations of BASELINE.json, so that they
        # are in the driver's own record (only config[1] was driver-run)
        del base
        also = {}
        try:
            also["aniso_n64"] = also_aniso(lib, 64, 3)
        except Exception as e:  # noqa: BLE001
            also["aniso_n64"] = {"error": str(e)[:300]}
        try:
            also["adj_100M"] = also_adj(lib)
        except Exception as e:  # noqa: BLE001
            also["adj_100M"] = {"error": str(e)[:300]}

    if rank == 0:
        cfg = workload_config(args.n, args.workload)
        cfg["parallelism"] = "1 GPU" if world == 1 else "%d independent replicas (one box per GPU, no ghost exchange yet)" % world
        cfg["passes_per_step"] = npasses
        cfg["tets_per_step"] = "%d -> %d" % (nelems0, nelems1)
        cfg["input_caches"] = ("the input mesh carries its derived R->E, R->V, F->V and the length + quality tags, measured "
                               "before the timed region (the reference arm calls ask_lengths/ask_qualities first too); "
                               "base.copy() shares them, so pass 0 pays no transit / measure_edges (~0.2 ms of the loop)")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "wall_ms_per_step": wall_ms / args.steps, "syncs_per_step": None,
            "peak_device_bytes": lib.peak_bytes(), "gpu_baseline": gpu_base, "also": also,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def box_shape(world):
    """ranks as a block of unit cubes: 2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2 ..."""
    shape = [1, 1, 1]
    i = 0
    w = world
    while w > 1:
        if w % 2:
            shape[i % 3] *= w
            break
        shape[i % 3] *= 2
        w //= 2
        i += 1
    return shape


def main_b200_partitioned(args):
    """N > 1: ONE box of N x n^3 cells, partitioned over the ranks (omega_h_b200/dist.py): every
    rank refines the elements it owns + a halo, the shell's edge data and the global-number scan go
    over NCCL. Weak scaling: the per-rank share is the N = 1 workload."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from omega_h_b200 import AdaptOpts, Lib, VERT, build_box
    from omega_h_b200 import dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    lib = Lib(device=local_rank).init()
    if lib.is_emulation:
        raise SystemExit("bench.py refuses to run on the emulation build")
    D.share_stream(lib, device)   # torch, NCCL and the library on one stream: no host syncs between them
    n = args.n
    shape = box_shape(world)
    parity = None
    if not args.no_parity_check:
        parity = partition_parity_check(lib, device, args.parity_halo or args.halo)
        torch.cuda.empty_cache()
    base = build_box(float(shape[0]), float(shape[1]), float(shape[2]), shape[0] * n, shape[1] * n, shape[2] * n, lib=lib)
    h = 1.0 / n / 2.0
    base.add_tag(VERT, "metric", 1, np.full(base.nverts(), 1.0 / (h * h)))
    base.ask_lengths()
    base.ask_qualities()
    halo = args.halo
    part0 = D.distribute(base, halo, device, parting=args.parting)
    nglobal0 = base.nelems()
    # derive what the first pass asks for once, as the N = 1 arm's input has it cached
    part0.mesh.ask_down(part0.mesh.dim(), 1)
    part0.mesh.ask_verts_of(part0.mesh.dim())
    part0.mesh.ask_verts_of(2)
    del base
    torch.cuda.empty_cache()   # the cut's temporaries: give them back before the library's pool grows
    opts = AdaptOpts(part0.mesh)

    def barrier():
        lib.sync()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def loop(part):
        passes = 0
        while part.refine_by_size(opts):
            passes += 1
        return passes

    for _ in range(max(args.warmup, 3)):
        part = part0.clone()
        npasses = loop(part)
    owned1 = part.owned_nelems()
    t = torch.tensor([float(owned1)], device=device, dtype=torch.float64)
    dist.all_reduce(t)
    nglobal1 = int(t.item())
    local1 = part.mesh.nelems()
    last = dict(part.last)
    del part

    # pick the dominant library kernel (outside the timed region)
    part = part0.clone()
    lib.profile_begin(None)
    loop(part)
    agg = summarize_profile(lib.profile_end())
    top = max(agg.items(), key=lambda kv: kv[1][1])[0]
    del part
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = lib.launch_count()
    lib.profile_begin(top)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        part = part0.clone()
        loop(part)
    lib.sync()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    top_recs = lib.profile_end()
    launches = lib.launch_count() - launches0
    clocks = sampler.stop()
    del part
    # roofline of the dominant library kernel on rank 0 (timed live with CUDA events, as at N = 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    cnt, tms, tbytes = summarize_profile(top_recs).get(top, (0, 0.0, 0))
    achieved = (tbytes / 1e9) / (tms / 1e3) if tms > 0 and tbytes else None
    roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "launches": cnt, "avg_ms": (tms / cnt) if cnt else None,
                "share_of_step": (tms / dev_ms) if dev_ms > 0 else None,
                "algorithmic_bytes_per_launch": (tbytes / cnt) if cnt else None, "rank": 0}
    tm = torch.tensor([dev_ms], device=device, dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tm[0])
    total_new = float(nglobal1 - nglobal0) * args.steps
    value = total_new / (dev_ms_max / 1e3)

    # end to end: every rank uploads its part from pinned host arrays and reads its refined part back
    e2e = None
    if not args.no_e2e:
        img = host_copy_of(part0.mesh, lib)
        down = Download(lib)

        def e2e_step():
            m = upload(img, lib)
            part = D.DistMesh(m, device, halo)
            part.nglobal = list(part0.nglobal)
            loop(part)
            return down(part.mesh)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            d2h_bytes = e2e_step()
        barrier()
        e_s = time.perf_counter() - t0
        te = torch.tensor([e_s], device=device, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_s = float(te[0])
        tb = torch.tensor([float(img["nbytes"]), float(d2h_bytes)], device=device, dtype=torch.float64)
        dist.all_reduce(tb)
        e2e = {"value": total_new / e_s, "unit": UNIT, "h2d_bytes_per_step": int(tb[0]), "d2h_bytes_per_step": int(tb[1]),
               "ms_per_step": e_s * 1e3 / args.steps,
               "timing": "wall clock around upload + partitioned loop + download of every rank's part, max over ranks"}

    def headline(also):
        """rank 0's ONE json line (the headline numbers are complete before the `also` block starts)"""
        cfg = workload_config(n, "iso")
        cfg["workload"] = ("3D tet box build_box %dx%dx%d cells (x6 tets) = %d ranks x %d^3, uniform isotropic metric "
                           "h=1/(2n), while(refine_by_size) loop on the partitioned mesh" %
                           (shape[0] * n, shape[1] * n, shape[2] * n, world, n))
        cfg["parallelism"] = ("%d parts (%s), halo %d layers, per-pass shell exchange + global-number scan over NCCL" % (
            world, "recursive inertial bisection = Mesh::balance" if args.parting == "rib" else
            "contiguous ranges of the Hilbert element order", halo))
        cfg["passes_per_step"] = npasses
        cfg["tets_per_step"] = "%d -> %d" % (nglobal0, nglobal1)
        cfg["local_tets_rank0"] = local1
        cfg["last_pass"] = last
        return {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None,
            "peak_device_bytes": peak_bytes, "parity_check": parity, "also": also,
        }

    peak_bytes = lib.peak_bytes()
    also = None
    if not args.no_also:
        # the `also` block must not be able to lose the headline: if it hangs (a rank that failed leaves the others in
        # a collective) rank 0 prints the line without it after ALSO_TIMEOUT_S and every rank ends
        def give_up():
            if rank == 0:
                emit(headline({"error": "the also block did not finish within %d s" % ALSO_TIMEOUT_S}))
            os._exit(0)

        def run_also():
            torch.cuda.empty_cache()
            lib.trim()
            return {"aniso_n%d" % n: also_aniso_partitioned(lib, device, n, halo, args.parting,
                                                            time_reghost=args.time_reghost)}
        try:
            also = run_guarded(ALSO_TIMEOUT_S, give_up, run_also)
        except Exception as e:  # noqa: BLE001
            also = {"error": str(e)[:300]}
            if rank == 0:
                emit(headline(also))
            os._exit(0)   # the other ranks may be inside a collective of the block: do not wait for them

    if rank == 0:
        emit(headline(also))
    if os.environ.get("OSHB_DIST_CPROFILE") and rank == 0:
        # diagnosis only: where the host spends its time in the partitioned loop
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(3):
            loop(part0.clone())
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(28)
    elif os.environ.get("OSHB_DIST_CPROFILE"):
        for _ in range(3):
            loop(part0.clone())
    if D.TIMING is not None and rank == 0:
        # the library's own kernels during one partitioned loop (CUDA events on its stream)
        part = part0.clone()
        lib.profile_begin(None)
        loop(part)
        agg = summarize_profile(lib.profile_end())
        ksum = sum(v[1] for v in agg.values())
        print("  library kernels in one partitioned loop: %.3f ms, %d launches" % (ksum, sum(v[0] for v in agg.values())),
              file=sys.stderr)
        for nm, (cnt, ms, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
            print("    %-34s n=%4d %8.3f ms" % (nm, cnt, ms), file=sys.stderr)
        del part
    elif D.TIMING is not None:
        loop(part0.clone())
    if D.TIMING is not None:
        nloops = args.steps * (1 if args.no_e2e else 2) + max(args.warmup, 3) + (0 if args.no_e2e else 1)
        for r in range(world):
            if r == rank:
                for k, v in sorted(D.TIMING.items(), key=lambda kv: -kv[1]):
                    print("  rank %d %-28s %9.3f ms/step" % (rank, k, v * 1e3 / nloops), file=sys.stderr)
            dist.barrier()
    dist.barrier()
    dist.destroy_process_group()
    return 0


_JSON_FD = None


def emit(obj):
    """the ONE json line of the contract, written to the process's real stdout (main() points fd 1 at stderr for
    everything else: NCCL prints its version banner to stdout when the library creates its communicator)"""
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse_args()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)   # C-level and Python-level stdout of everything below -> stderr; emit() keeps the real one
    if args.impl in ("reference", "reference-cuda"):
        return main_reference(args)
    if (int(os.environ.get("WORLD_SIZE", "1")) > 1 or os.environ.get("OSHB_FORCE_PARTITIONED")) and not args.replicas:
        return main_b200_partitioned(args)
    return main_b200(args)


if __name__ == "__main__":
    sys.exit(main())
