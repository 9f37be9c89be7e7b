"""Partitioned refinement of a mesh file: read `.osh` -> cut into parts -> refine loop on all ranks
-> assemble -> write `.osh`. Launch with torchrun, one process per GPU:

    torchrun --nproc-per-node 8 examples/partitioned_refine.py in.osh out.osh [halo]

The output equals, byte for byte, what the serial loop writes (tests/test_osh_file.py runs it on
CPU over gloo against the host emulation build by passing --emulation <path to the test library>)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omega_h_b200 import _lib, read_osh, write_osh  # noqa: E402
from omega_h_b200 import dist as D  # noqa: E402


def main():
    args = [a for a in sys.argv[1:]]
    emu = None
    if "--emulation" in args:                       # tests only
        i = args.index("--emulation")
        emu = args[i + 1]
        del args[i:i + 2]
    src, dst = args[0], args[1]
    halo = int(args[2]) if len(args) > 2 else 4
    if emu:
        dist.init_process_group("gloo")
        device = "cpu"
        lib = _lib.Lib(emu).init()
    else:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        device = "cuda:%d" % local
        dist.init_process_group("nccl")
        lib = _lib.Lib(device=local).init()
        D.share_stream(lib, device)
    mesh = read_osh(src, lib)                       # every rank reads the (small) input
    part = D.distribute(mesh, halo, device)
    passes = 0
    while part.refine_by_size():
        passes += 1
    whole = part.gather(0)
    if dist.get_rank() == 0:
        whole.class_sets = getattr(mesh, "class_sets", {})
        write_osh(dst, whole)
        print("refined %d -> %d elements in %d passes on %d ranks" % (mesh.nelems(), whole.nelems(), passes,
                                                                     dist.get_world_size()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
