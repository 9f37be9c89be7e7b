"""Partitioned refine_by_size (omega_h_b200/dist.py) against the serial loop: world_size 2 and 3
over gloo on CPU (the library is the host emulation build), world_size 2 over NCCL on GPUs."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "dist_worker.py")


def run_worker(nranks, lib_path, device, n, halo, aniso, dim, port, timeout=900, parting="hilbert", py_pass=False,
               py_reghost=False):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, lib_path, device, str(n), str(halo),
           str(aniso), str(dim), parting]
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "1"
    env["OSHB_DIST_CHECK"] = "1"
    env["OSHB_DIST_PY"] = "1" if py_pass else "0"   # 0: the library's C++ pass (csrc/dist.cu); 1: the torch-level pass
    env["OSHB_REGHOST_PY"] = "1" if py_reghost else "0"   # 0: the library's re-ghosting (oshb_dist_reghost); 1: torch ops
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


@pytest.mark.parametrize("nranks,n,halo,aniso,dim", [
    (1, 4, 4, 0, 3),    # one rank: the staged pass + external numbering alone
    (2, 8, 4, 0, 3),    # 4 doubling passes, the halo exactly used up
    (3, 6, 5, 0, 3),    # uneven parts, 5 passes
    (2, 5, 8, 1, 3),    # anisotropic metric (ncomps 6), metric transfer
    (2, 16, 4, 0, 2),   # triangles
])
def test_partitioned_loop_matches_serial_gloo(emu_lib, nranks, n, halo, aniso, dim):
    run_worker(nranks, emu_lib.path, "cpu", n, halo, aniso, dim, 29530 + nranks + n)


@pytest.mark.parametrize("nranks,n,halo,aniso,dim", [
    (2, 8, 2, 0, 3),    # 4 passes on a 2-layer halo: one re-ghosting
    (3, 6, 1, 0, 3),    # a 1-layer halo: a fresh halo before every pass but the first
    (2, 6, 2, 1, 3),    # anisotropic, 8 passes, 3 re-ghostings
    (2, 16, 1, 0, 2),   # triangles
    (4, 8, 1, 0, 3),    # four ranks: neighbours of neighbours take part in the exchange
])
def test_reghosting_gloo(emu_lib, nranks, n, halo, aniso, dim):
    """loops longer than the halo: DistMesh.reghost() must hand every rank a halo on which the
    remaining passes again equal the serial ones"""
    out = run_worker(nranks, emu_lib.path, "cpu", n, halo, aniso, dim, 29560 + nranks + n + halo)
    assert int(out.split("reghosts=")[1].split()[0]) >= 1


@pytest.mark.parametrize("nranks,n,halo,aniso,dim", [(2, 8, 2, 0, 3), (3, 6, 1, 0, 3)])
def test_torch_level_reghost_still_matches_serial_gloo(emu_lib, nranks, n, halo, aniso, dim):
    """the earlier re-ghosting in torch ops (dist.py _reghost_py, OSHB_REGHOST_PY=1): a second, independent
    implementation the library's oshb_dist_reghost is compared with (both must reproduce the serial loop)"""
    out = run_worker(nranks, emu_lib.path, "cpu", n, halo, aniso, dim, 29650 + nranks + n + halo, py_reghost=True)
    assert int(out.split("reghosts=")[1].split()[0]) >= 1


@pytest.mark.parametrize("nranks,n,halo,aniso,dim", [(2, 8, 2, 0, 3), (3, 6, 5, 0, 3)])
def test_torch_level_pass_still_matches_serial_gloo(emu_lib, nranks, n, halo, aniso, dim):
    """the earlier orchestration of the same stages in torch ops (dist.py, OSHB_DIST_PY=1): a second, independent
    implementation of the exchanges the C++ pass is compared with"""
    run_worker(nranks, emu_lib.path, "cpu", n, halo, aniso, dim, 29620 + nranks + n, py_pass=True)


@pytest.mark.parametrize("nranks,n,halo,aniso,dim", [
    (2, 8, 4, 0, 3),    # RIB halves of the cube
    (4, 6, 2, 1, 3),    # four RIB parts, anisotropic, with re-ghosting
    (2, 12, 4, 0, 2),   # triangles
    (8, 8, 2, 0, 3),    # eight RIB octants (the 8-GPU layout), one re-ghosting: seven neighbours per rank
])
def test_partitioned_loop_on_rib_parts_gloo(emu_lib, nranks, n, halo, aniso, dim):
    """parts cut by recursive inertial bisection (Mesh::balance, BASELINE config[3]) instead of ranges of the
    element order: the partitioned loop still equals the serial one array by array"""
    run_worker(nranks, emu_lib.path, "cpu", n, halo, aniso, dim, 29590 + nranks + n, parting="rib")


@pytest.mark.gpu
def test_partitioned_loop_one_gpu_nccl(gpu_lib):
    """the whole partitioned path (staged pass, shared stream, numbering helpers, NCCL collectives
    with itself) on a single GPU"""
    run_worker(1, gpu_lib.path, "cuda", 16, 4, 0, 3, 29540)
    run_worker(1, gpu_lib.path, "cuda", 16, 2, 0, 3, 29539)    # with a re-ghosting (no neighbours: a re-cut)


@pytest.mark.gpu
def test_partitioned_loop_matches_serial_nccl(gpu_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    run_worker(2, gpu_lib.path, "cuda", 16, 4, 0, 3, 29541)
    run_worker(2, gpu_lib.path, "cuda", 24, 5, 0, 3, 29542)
    run_worker(2, gpu_lib.path, "cuda", 10, 10, 1, 3, 29543)
    out = run_worker(2, gpu_lib.path, "cuda", 16, 2, 0, 3, 29544)    # re-ghosting over NCCL point-to-point
    assert int(out.split("reghosts=")[1].split()[0]) >= 1
