"""compare_meshes on the device against the reference's own compare_meshes (what oshdiff runs,
src/Omega_h_compare.cpp:179-277): the same pairs of meshes, written as .osh files by omega_h_b200.write_osh, are
judged by oracle/_ref/ref_driver `diff` and by oshb_mesh_compare; the verdicts (SAME 0 / MORE 1 / DIFF 2) must agree."""
import subprocess

import numpy as np
import pytest

import parity
from omega_h_b200 import Mesh, build_box, compare_meshes, refine_by_size, write_osh


def _libs():
    return [pytest.param("emu_lib", id="emu"), pytest.param("gpu_lib", id="gpu", marks=pytest.mark.gpu)]


@pytest.fixture
def lib(request):
    return request.getfixturevalue(request.param)


def _clone(m, lib, perturb=None, extra=None, drop=None, permute_elems=False, bad_conn=False):
    """a deep copy through host arrays with an optional defect"""
    dim = m.dim()
    c = Mesh(dim, lib=lib)
    c.set_verts(m.nverts())
    perm = None
    for d in range(1, dim + 1):
        down, codes = m.ask_down(d, d - 1)
        deg = d + 1
        down = down.reshape(-1, deg).copy()
        codes = None if codes is None else codes.reshape(-1, deg).copy()
        if d == dim and permute_elems:
            perm = np.random.default_rng(3).permutation(m.nents(d))
            down = down[perm]
            codes = None if codes is None else codes[perm]
        if d == dim and bad_conn:
            down[0, 0], down[0, 1] = down[0, 1], down[0, 0]
        c.set_ents(d, down.reshape(-1), None if codes is None else codes.reshape(-1))
    for d in range(dim + 1):
        for name, _, nc in m.tags(d):
            if drop == (d, name):
                continue
            a = m.get_array(d, name).reshape(-1, nc).copy()
            if perturb and perturb[0] == (d, name):
                a = a * (1.0 + perturb[1] * np.linspace(-1, 1, a.size).reshape(a.shape))
            if d == dim and perm is not None:
                a = a[perm]
            c.add_tag(d, name, nc, a.reshape(-1), internal=True)
    if extra:
        c.add_tag(extra[0], extra[1], 1, np.zeros(c.nents(extra[0])), internal=True)
    c.class_sets = getattr(m, "class_sets", {})
    return c


CASES = [
    ("identical", {}, 0),
    ("coords_1e-9", {"perturb": ((0, "coordinates"), 1e-9)}, 0),
    ("coords_1e-3", {"perturb": ((0, "coordinates"), 1e-3)}, 2),
    ("metric_1e-4", {"perturb": ((0, "metric"), 1e-4)}, 2),
    ("extra_tag", {"extra": (0, "pressure")}, 1),
    ("missing_tag", {"drop": (0, "metric")}, 2),
    ("permuted_elements", {"permute_elems": True}, 0),
    ("swapped_row", {"bad_conn": True}, 2),
]


@pytest.mark.parametrize("lib", _libs(), indirect=True)
def test_verdicts_agree_with_the_reference(lib, ref_driver, tmp_path):
    base = build_box(1.0, 1.0, 1.0, 3, 3, 3, lib=lib)
    base.add_tag(0, "metric", 1, np.full(base.nverts(), 64.0))
    assert refine_by_size(base)
    a_path = str(tmp_path / "a.osh")
    write_osh(a_path, base)
    for name, kw, expect in CASES:
        other = _clone(base, lib, **kw)
        ours = compare_meshes(base, other, tolerance=1e-6, floor=0.0)
        b_path = str(tmp_path / (name + ".osh"))
        write_osh(b_path, other)
        r = subprocess.run([ref_driver, "diff", a_path, b_path, "1e-6", "0.0"], capture_output=True, text=True, check=True)
        ref = int([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0].split()[1])
        assert ours == ref == expect, (name, ours, ref, expect, r.stdout[-500:])
    # tolerance and floor are honoured like oshdiff's -tolerance / -Floor
    other = _clone(base, lib, perturb=((0, "coordinates"), 1e-3))
    assert compare_meshes(base, other, tolerance=1e-2) == 0
    assert compare_meshes(base, other, tolerance=1e-6, floor=10.0) == 0
    assert compare_meshes(base, other, compare_type="none") == 0
