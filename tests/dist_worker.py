"""Worker of the world_size-N partitioned-refine tests (launched by tests/test_dist.py through
torch.distributed.run). Every rank builds the same box, keeps its part + halo, runs the
partitioned refine loop, and checks the part it owns against the serial loop on the full mesh:
the same elements (by global number) with the same vertices, coordinates and classification,
and the same global numbers / down adjacencies / codes on every entity of its elements."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from omega_h_b200 import _lib, mesh as M  # noqa: E402
from omega_h_b200 import dist as D  # noqa: E402


def add_metric(m, n, aniso):
    coords = m.coords().reshape(-1, m.dim())
    if aniso:
        h = np.full((m.nverts(), 3), 1.0 / n)
        h[:, 2] = (1.0 / n) * (1.0 - 0.75 / np.cosh(20.0 * (coords[:, 2] - 0.5)) ** 2)
        met = np.zeros((m.nverts(), 6))
        met[:, 0:3] = 1.0 / (h * h)
        m.add_tag(0, "metric", 6, met.reshape(-1))
    else:
        h = 1.0 / (2 * n) if m.dim() == 3 else 1.0 / (2 * n)
        m.add_tag(0, "metric", 1, np.full(m.nverts(), 1.0 / (h * h)))


def main():
    lib_path, device, n, halo, aniso, dim = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    parting = sys.argv[7] if len(sys.argv) > 7 else "hilbert"
    backend = "nccl" if device == "cuda" else "gloo"
    dist.init_process_group(backend)
    rank, P = dist.get_rank(), dist.get_world_size()
    if device == "cuda":
        device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(device)
    lib = _lib.Lib(lib_path).init()
    if os.environ.get("OSHB_SHARE_STREAM", "1") == "1":
        D.share_stream(lib, device)
    base = M.build_box(1.0, 1.0, 1.0 if dim == 3 else 0.0, n, n, n if dim == 3 else 0, lib)
    add_metric(base, n, aniso)
    serial = base.copy()
    npass = 0
    while M.refine_by_size(serial):
        npass += 1
    part = D.distribute(base, halo, device, parting=parting)
    dpass = 0
    while part.refine_by_size():
        dpass += 1
    assert dpass == npass, (dpass, npass)
    # ---- compare what this rank owns with the serial result
    m = part.mesh
    gl = [m.globals(d) for d in range(dim + 1)]
    own = [part.owned_mask(d).cpu().numpy() for d in range(dim + 1)]
    part_tag = m.get_array(dim, "own:part")
    erank = part_tag >> 8
    edepth = (part_tag << 24) >> 24
    assert np.array_equal(own[dim], edepth <= 0)
    assert np.all(erank[edepth <= 0] == rank)
    nown = np.array([int(own[dim].sum())], dtype=np.int64)
    t = torch.from_numpy(nown).to(device)
    dist.all_reduce(t)
    assert int(t.item()) == serial.nelems(), (int(t.item()), serial.nelems())
    for d in range(dim + 1):
        g = gl[d][own[d]]
        assert len(np.unique(g)) == len(g)
        assert g.min() >= 0 and g.max() < serial.nents(d)
        for name, ttype, nc in serial.tags(d):
            if name in ("global", "length", "quality"):
                continue
            a = serial.get_array(d, name).reshape(-1, nc)[g]
            b = m.get_array(d, name).reshape(-1, nc)[own[d]]
            assert np.array_equal(a, b), (d, name)
        if d >= 1:
            deg = M.simplex_degree(d, d - 1)
            sd, sc = serial.ask_down(d, d - 1)
            ld, lc = m.ask_down(d, d - 1)
            a = sd.reshape(-1, deg)[g]
            b = gl[d - 1][ld.reshape(-1, deg)[own[d]]]
            assert np.array_equal(a, b), (d, "down")
            if d >= 2:
                assert np.array_equal(sc.reshape(-1, deg)[g], lc.reshape(-1, deg)[own[d]]), (d, "codes")
    sv = serial.ask_verts_of(dim).reshape(-1, dim + 1)[gl[dim][own[dim]]]
    lv = gl[0][m.ask_verts_of(dim).reshape(-1, dim + 1)[own[dim]]]
    assert np.array_equal(sv, lv)
    # ---- the gathered mesh equals the serial one, array by array
    whole = part.gather(0)
    if rank == 0:
        for d in range(dim + 1):
            assert whole.nents(d) == serial.nents(d)
            names = sorted(t[0] for t in serial.tags(d))
            assert names == sorted(t[0] for t in whole.tags(d)), (names, whole.tags(d))
            for name in names:
                assert np.array_equal(serial.get_array(d, name), whole.get_array(d, name)), (d, name)
            if d >= 1:
                a, ac = serial.ask_down(d, d - 1)
                b, bc = whole.ask_down(d, d - 1)
                assert np.array_equal(a, b), (d, "down")
                if d >= 2:
                    assert np.array_equal(ac, bc), (d, "codes")
    else:
        assert whole is None
    if rank == 0:
        print("DIST_OK passes=%d ranks=%d serial_elems=%d local_elems=%d owned=%d reghosts=%d" % (
            npass, P, serial.nelems(), m.nelems(), int(own[dim].sum()), getattr(part, "reghosts", 0)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
