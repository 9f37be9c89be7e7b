"""ctypes binding of the C ABI declared in include/oshb.h (liboshb.so).

The product library is built by nvcc for sm_100a only (omega_h_b200/csrc/Makefile) and has
no CPU path: loading fails loudly when the shared object is missing, and oshb_init fails
loudly when no usable GPU is present. Tests may bind another build of the same ABI
(the host emulation under tests/emu) by passing an explicit path to `Lib`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "lib", "liboshb.so")

I8, I32, I64, F64 = 0, 1, 2, 3
NP_OF = {I8: np.int8, I32: np.int32, I64: np.int64, F64: np.float64}
TYPE_OF = {np.dtype(np.int8): I8, np.dtype(np.int32): I32, np.dtype(np.int64): I64, np.dtype(np.float64): F64}


class OshbError(RuntimeError):
    pass


class AdaptOptsC(C.Structure):
    _fields_ = [
        ("min_length_desired", C.c_double),
        ("max_length_desired", C.c_double),
        ("max_length_allowed", C.c_double),
        ("min_quality_allowed", C.c_double),
        ("min_quality_desired", C.c_double),
        ("verbosity", C.c_int32),
    ]


class PassStatsC(C.Structure):
    _fields_ = [
        ("ncands", C.c_int32),
        ("nkeys", C.c_int32),
        ("indset_rounds", C.c_int32),
        ("nents_before", C.c_int32 * 4),
        ("nents_after", C.c_int32 * 4),
    ]


# every symbol include/oshb.h declares (tests check the library exports all of them)
SYMBOLS = [
    "oshb_init", "oshb_sync", "oshb_set_stream", "oshb_last_error", "oshb_is_emulation", "oshb_launch_count", "oshb_sync_count",
    "oshb_peak_bytes", "oshb_trim", "oshb_set_oom_callback", "oshb_dev_alloc", "oshb_dev_free", "oshb_h2d", "oshb_d2h",
    "oshb_offset_scan_i8", "oshb_offset_scan_i32", "oshb_offset_scan_i32_i64", "oshb_collect_marked",
    "oshb_max_i8", "oshb_minmax_f64", "oshb_sort_by_keys_i32", "oshb_sort_by_keys_i64",
    "oshb_unmap", "oshb_map_into", "oshb_expand_into", "oshb_mark_image", "oshb_invert_injective_map",
    "oshb_compound_maps",
    "oshb_invert_adj", "oshb_transit", "oshb_reflect_down", "oshb_find_unique",
    "oshb_measure_edges_metric", "oshb_measure_qualities", "oshb_libm_eval",
    "oshb_mesh_create", "oshb_mesh_destroy", "oshb_mesh_clone", "oshb_mesh_dim", "oshb_mesh_nents",
    "oshb_mesh_set_verts", "oshb_mesh_set_ents", "oshb_mesh_add_tag", "oshb_mesh_remove_tag", "oshb_mesh_ntags",
    "oshb_mesh_tag_info", "oshb_mesh_get_tag", "oshb_mesh_gather_tag", "oshb_mesh_ask_down", "oshb_mesh_ask_up", "oshb_mesh_ask_star",
    "oshb_mesh_ask_lengths", "oshb_mesh_ask_qualities", "oshb_build_box", "oshb_mesh_rib_partition", "oshb_mesh_compare",
    "oshb_adapt_opts_init", "oshb_mesh_set_transfer", "oshb_set_user_transfer", "oshb_refine_qualities", "oshb_mident_metrics", "oshb_find_indset",
    "oshb_rep_vertex2md_order", "oshb_refine_by_size", "oshb_last_pass_stats",
    "oshb_comm_nccl_unique_id", "oshb_comm_create_nccl", "oshb_comm_create_callbacks", "oshb_comm_destroy",
    "oshb_dist_refine_by_size", "oshb_dist_reghost", "oshb_dist_distribute",
    "oshb_pass_create", "oshb_pass_destroy", "oshb_pass_begin", "oshb_pass_restate", "oshb_pass_indset_round",
    "oshb_pass_select_keys", "oshb_pass_number", "oshb_pass_finish", "oshb_pass_size", "oshb_pass_get",
    "oshb_pass_set", "oshb_pass_gather", "oshb_pass_scatter", "oshb_pass_runs_begin", "oshb_pass_runs_get",
    "oshb_pass_runs_set_bases", "oshb_pass_want_get", "oshb_pass_runs_lookup", "oshb_pass_want_set",
    "oshb_pass_runs_commit",
    "oshb_timer_start", "oshb_timer_stop", "oshb_profile_begin", "oshb_profile_end", "oshb_host_alloc",
    "oshb_host_free", "oshb_host_time_stats",
]


class Lib:
    def __init__(self, path=None, device=None):
        self.path = path or PRODUCT_LIB
        if not os.path.exists(self.path):
            raise OshbError(
                "CUDA extension %s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % self.path)
        self.c = C.CDLL(self.path)
        self.c.oshb_last_error.restype = C.c_char_p
        for n in ("oshb_launch_count", "oshb_sync_count", "oshb_peak_bytes"):
            getattr(self.c, n).restype = C.c_uint64
        self._inited = False
        self._device = device

    def init(self):
        if not self._inited:
            dev = self._device
            if dev is None:
                dev = int(os.environ.get("LOCAL_RANK", "0"))
            self.check(self.c.oshb_init(C.c_int(dev)))
            self._inited = True
        return self

    def check(self, rc):
        if rc != 0:
            raise OshbError(self.c.oshb_last_error().decode(errors="replace"))

    @property
    def is_emulation(self):
        return bool(self.c.oshb_is_emulation())

    def launch_count(self):
        return int(self.c.oshb_launch_count())

    def sync_count(self):
        return int(self.c.oshb_sync_count())

    def trim(self):
        """return the library's wholly free device segments to the driver (oshb_trim)"""
        self.check(self.c.oshb_trim())

    def peak_bytes(self):
        return int(self.c.oshb_peak_bytes())

    def sync(self):
        self.check(self.c.oshb_sync())

    # ---- measurement hooks --------------------------------------------------------------
    def timer_start(self):
        self.check(self.c.oshb_timer_start())

    def timer_stop(self):
        ms = C.c_double()
        self.check(self.c.oshb_timer_stop(C.byref(ms)))
        return ms.value

    def profile_begin(self, kernel_name=None):
        self.check(self.c.oshb_profile_begin(kernel_name.encode() if kernel_name else None))

    def profile_end(self):
        """[(kernel name, milliseconds)] for every profiled launch, in launch order."""
        need = C.c_uint64()
        self.check(self.c.oshb_profile_end(None, C.c_uint64(0), C.byref(need)))
        buf = C.create_string_buffer(int(need.value) + 16)
        self.check(self.c.oshb_profile_end(buf, C.c_uint64(len(buf)), C.byref(need)))
        out = []
        for line in buf.value.decode().splitlines():
            name, ms = line.rsplit("\t", 1)
            out.append((name, float(ms)))
        return out

    def pinned_empty(self, n, dtype):
        """numpy array over cudaHostAlloc'd memory (kept alive by the returned array's base)."""
        dtype = np.dtype(dtype)
        nbytes = max(int(n) * dtype.itemsize, 1)
        p = C.c_void_p()
        self.check(self.c.oshb_host_alloc(C.c_uint64(nbytes), C.byref(p)))
        raw = (C.c_char * nbytes).from_address(p.value)
        holder = _PinnedHolder(self, p, raw)
        arr = np.frombuffer(raw, dtype=dtype, count=int(n))
        _PINNED[id(arr)] = holder
        return arr

    # ---- raw device buffers (used by the primitive-level tests) -------------------------
    def to_device(self, a):
        a = np.ascontiguousarray(a)
        p = C.c_void_p()
        self.check(self.c.oshb_dev_alloc(C.c_uint64(max(a.nbytes, 1)), C.byref(p)))
        if a.nbytes:
            self.check(self.c.oshb_h2d(p, a.ctypes.data_as(C.c_void_p), C.c_uint64(a.nbytes)))
        return DevBuf(self, p, a.nbytes, a.dtype, a.size)

    def empty_device(self, n, dtype):
        dtype = np.dtype(dtype)
        p = C.c_void_p()
        nbytes = int(n) * dtype.itemsize
        self.check(self.c.oshb_dev_alloc(C.c_uint64(max(nbytes, 1)), C.byref(p)))
        return DevBuf(self, p, nbytes, dtype, int(n))


_PINNED = {}


class _PinnedHolder:
    """Keeps a pinned allocation alive for the life of the process (bench buffers are few)."""

    def __init__(self, lib, ptr, raw):
        self.lib, self.ptr, self.raw = lib, ptr, raw


class DevBuf:
    def __init__(self, lib, ptr, nbytes, dtype, n):
        self.lib, self.ptr, self.nbytes, self.dtype, self.n = lib, ptr, nbytes, np.dtype(dtype), n

    def to_host(self, n=None):
        n = self.n if n is None else int(n)
        out = np.empty(n, dtype=self.dtype)
        if n:
            self.lib.check(self.lib.c.oshb_d2h(out.ctypes.data_as(C.c_void_p), self.ptr, C.c_uint64(out.nbytes)))
        return out

    def free(self):
        if self.ptr:
            self.lib.c.oshb_dev_free(self.ptr, C.c_uint64(max(self.nbytes, 1)))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


_default = None


def default_lib():
    """The product library, initialised on cuda:LOCAL_RANK. Raises when it is not built or no GPU is usable."""
    global _default
    if _default is None:
        _default = Lib().init()
    return _default
