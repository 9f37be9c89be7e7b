"""Mesh::balance()'s recursive inertial bisection on the device against the reference:
golden element -> part maps written by oracle/_ref/ref_driver `rib` (the reference's own
inertia::mark_bisection applied recursively, reductions by repro_sum), bit for bit; live runs on the GPU box."""
import glob
import os
import subprocess

import numpy as np
import pytest

import parity

GOLD = sorted(glob.glob(os.path.join(parity.HERE, "golden", "rib_*.oshd.gz")))


def _libs():
    return [pytest.param("emu_lib", id="emu"), pytest.param("gpu_lib", id="gpu", marks=pytest.mark.gpu)]


@pytest.fixture
def lib(request):
    return request.getfixturevalue(request.param)


def _check(fx, lib):
    m = parity.mesh_from_fixture(fx, lib)
    nparts = int(fx["rib:nparts"][0])
    parts, axes = m.rib_partition(nparts)
    assert np.array_equal(parts, fx["rib:parts"])
    counts = np.bincount(parts, minlength=nparts)
    assert counts.min() > 0 and counts.max() - counts.min() <= 8      # balanced within the reference's tolerance per cut
    assert axes.shape == (nparts - 1, 3) and np.allclose(np.linalg.norm(axes, axis=1), 1.0)


@pytest.mark.parametrize("lib", _libs(), indirect=True)
@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p).split(".")[0])
def test_rib_matches_reference_golden(lib, path):
    _check(parity.load(path), lib)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,nx,ny,nz,nparts", [(3, 24, 16, 12, 8), (3, 20, 20, 20, 4), (2, 96, 64, 0, 8)])
def test_rib_matches_live_reference(gpu_lib, ref_driver, tmp_path, dim, nx, ny, nz, nparts):
    out = str(tmp_path / "rib.oshd")
    subprocess.run([ref_driver, "rib", str(dim), str(nx), str(ny), str(nz), str(nparts), out], check=True,
                   stdout=subprocess.DEVNULL)
    _check(parity.load(out), gpu_lib)
