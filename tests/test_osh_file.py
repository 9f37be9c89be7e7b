""".osh files (src/Omega_h_file.cpp) both ways against the UNMODIFIED reference: what it writes we
read, what we write it reads -- and writes byte for byte like it."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

import parity
from omega_h_b200 import read_osh, refine_by_size, write_osh


@pytest.mark.parametrize("dim,n,metric,npasses", [(3, 3, 0, 1), (2, 5, 2, 2), (3, 3, 2, 0)])
def test_osh_roundtrip_with_reference(emu_lib, ref_driver, tmp_path, dim, n, metric, npasses):
    ref_osh, ref_dump = str(tmp_path / "ref.osh"), str(tmp_path / "ref.oshd")
    subprocess.run([ref_driver, "writeosh", str(dim), str(n), str(metric), str(npasses), ref_osh, ref_dump], check=True,
                   stdout=subprocess.DEVNULL)
    fx = parity.load(ref_dump)
    # reference -> us
    m = read_osh(ref_osh, emu_lib)
    rep = parity.Report()
    parity.compare_mesh(rep, m, fx, "in:")
    rep.assert_ok()
    assert "body" in m.class_sets
    # us -> reference: the same bytes, and the reference reads the same mesh back
    ours = str(tmp_path / "ours.osh")
    write_osh(ours, m)
    assert filecmp.cmp(os.path.join(ref_osh, "0.osh"), os.path.join(ours, "0.osh"), shallow=False)
    back = str(tmp_path / "back.oshd")
    subprocess.run([ref_driver, "readosh", ours, back], check=True, stdout=subprocess.DEVNULL)
    fb = parity.load(back)
    assert set(fb) == set(fx)
    for k in fx:
        assert np.array_equal(fx[k], fb[k]), k
    # compressed streams (what an OMEGA_H_USE_ZLIB build writes) round-trip too
    zosh = str(tmp_path / "z.osh")
    write_osh(zosh, m, compress=True)
    rep = parity.Report()
    parity.compare_mesh(rep, read_osh(zosh, emu_lib), fx, "in:")
    rep.assert_ok()


def test_osh_after_our_refine_is_readable_by_reference(emu_lib, ref_driver, tmp_path):
    """a mesh refined HERE, written HERE, read by the reference: equal to the reference's own pass"""
    ref_osh, ref_dump = str(tmp_path / "ref.osh"), str(tmp_path / "ref.oshd")
    subprocess.run([ref_driver, "writeosh", "3", "3", "0", "0", ref_osh, str(tmp_path / "in.oshd")], check=True,
                   stdout=subprocess.DEVNULL)
    subprocess.run([ref_driver, "writeosh", "3", "3", "0", "1", str(tmp_path / "ref1.osh"), ref_dump], check=True,
                   stdout=subprocess.DEVNULL)
    m = read_osh(ref_osh, emu_lib)
    assert refine_by_size(m)
    ours = str(tmp_path / "ours.osh")
    write_osh(ours, m)
    back = str(tmp_path / "back.oshd")
    subprocess.run([ref_driver, "readosh", ours, back], check=True, stdout=subprocess.DEVNULL)
    fx, fb = parity.load(ref_dump), parity.load(back)
    for k in fx:
        if fx[k].dtype.kind == "f":
            assert np.allclose(fx[k], fb[k], rtol=1e-12, atol=0), k
        else:
            assert np.array_equal(fx[k], fb[k]), k


def test_partitioned_refine_of_a_file_equals_serial(emu_lib, ref_driver, tmp_path):
    """examples/partitioned_refine.py on 2 ranks (gloo, emulation build): .osh in, .osh out, and the
    output is byte-identical to what the serial loop writes"""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = str(tmp_path / "in.osh")
    subprocess.run([ref_driver, "writeosh", "3", "6", "0", "0", src, str(tmp_path / "in.oshd")], check=True,
                   stdout=subprocess.DEVNULL)
    m = read_osh(src, emu_lib)
    while refine_by_size(m):
        pass
    serial = str(tmp_path / "serial.osh")
    write_osh(serial, m)
    out = str(tmp_path / "out.osh")
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "1"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29581",
                        os.path.join(root, "examples", "partitioned_refine.py"), src, out, "2",
                        "--emulation", emu_lib.path], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert filecmp.cmp(os.path.join(serial, "0.osh"), os.path.join(out, "0.osh"), shallow=False)
