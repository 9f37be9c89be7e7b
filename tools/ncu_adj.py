"""one invert_adj R->F, one invert_adj R->V (the staged-tile row sort) and one reflect_down R->F on a 128^3 box,
bracketed by MARK / END launch counts for ncu --launch-skip / --launch-count"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omega_h_b200 import Lib, build_box
lib = Lib(device=0).init()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = build_box(1., 1., 1., n, n, n, lib=lib)
c = lib.c
nr, nf, nv = m.nents(3), m.nents(2), m.nents(0)
d_ab = lib.empty_device(nr * 4, np.int32); d_co = lib.empty_device(nr * 4, np.int8)
lib.check(c.oshb_mesh_ask_down(m.h, C.c_int(3), C.c_int(2), d_ab.ptr, d_co.ptr, C.c_int(0)))
d_off = lib.empty_device(nf + 1, np.int32); d_up = lib.empty_device(nr * 4, np.int32); d_uc = lib.empty_device(nr * 4, np.int8)
for _ in range(2):
    lib.check(c.oshb_invert_adj(d_ab.ptr, d_co.ptr, C.c_int64(nr), C.c_int(4), C.c_int32(nf), d_off.ptr, d_up.ptr, d_uc.ptr))
lib.sync()
print("MARK", lib.launch_count())
lib.check(c.oshb_invert_adj(d_ab.ptr, d_co.ptr, C.c_int64(nr), C.c_int(4), C.c_int32(nf), d_off.ptr, d_up.ptr, d_uc.ptr))
d_rv = lib.empty_device(nr * 4, np.int32); d_fv = lib.empty_device(nf * 3, np.int32)
lib.check(c.oshb_mesh_ask_down(m.h, C.c_int(3), C.c_int(0), d_rv.ptr, None, C.c_int(0)))
d_voff = lib.empty_device(nv + 1, np.int32)
lib.check(c.oshb_invert_adj(d_rv.ptr, None, C.c_int64(nr), C.c_int(4), C.c_int32(nv), d_voff.ptr, d_up.ptr, d_uc.ptr))
lib.check(c.oshb_mesh_ask_down(m.h, C.c_int(2), C.c_int(0), d_fv.ptr, None, C.c_int(0)))
d_hl = lib.empty_device(nr * 4, np.int32); d_hc = lib.empty_device(nr * 4, np.int8)
lib.check(c.oshb_reflect_down(d_rv.ptr, C.c_int64(nr), C.c_int(3), d_fv.ptr, C.c_int64(nf), C.c_int(2), C.c_int32(nv), d_hl.ptr, d_hc.ptr))
lib.sync()
print("END", lib.launch_count())
