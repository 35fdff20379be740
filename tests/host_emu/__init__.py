"""TEST INFRASTRUCTURE: the fast traversal kernels compiled for the host (tests/host_emu/emu.cpp) behind ctypes.

Used only by tests/test_host_emu.py to check the logic of qubatron_b200/csrc's own kernel source on the CPU; the
product never imports this package (tests/test_abi.py greps for it)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "..", "..", "qubatron_b200", "csrc")
# QB_EMU_DEFINES="-DQB_MASK_LATE ...": the host build of an experiment switch of the kernel source (its own library file)
_DEFINES = os.environ.get("QB_EMU_DEFINES", "").split()
_LIB = os.path.join(_HERE, "libqb_host_emu%s.so" % ("_" + "".join(c for c in "".join(_DEFINES) if c.isalnum()) if _DEFINES else ""))
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    srcs += [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)]
    if not force and os.path.exists(_LIB) and all(os.path.getmtime(s) <= os.path.getmtime(_LIB) for s in srcs):
        return _LIB
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-w", "-shared", "-fPIC",
           "-DQB_PTX_HOST_HEADER=\"octree_ptx_host.h\"", "-I" + _HERE, "-I" + cuda_inc, "-I" + _CSRC,
           "-o", _LIB, os.path.join(_HERE, "emu.cpp")] + _DEFINES
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("host emulation build failed:\n" + r.stdout)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.qb_emu_scene_create.restype = C.c_void_p
        _lib.qb_emu_scene_create.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
                                             C.c_long, C.c_void_p, C.c_void_p, C.c_long]
        _lib.qb_emu_scene_destroy.argtypes = [C.c_void_p]
        _lib.qb_emu_render_size.argtypes = [C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        _lib.qb_emu_render.restype = C.c_int
        _lib.qb_emu_render.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                       C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.qb_emu_trace_lines.restype = C.c_int
        _lib.qb_emu_trace_lines.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                            C.c_void_p, C.c_void_p]
        _lib.qb_emu_entry_compare.restype = C.c_long
        _lib.qb_emu_entry_compare.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class EmuScene:
    """The scene's reference-format arrays laid out as the connector lays them out in HBM (host memory here)."""

    def __init__(self, scene):
        self.oct_s = np.ascontiguousarray(scene.oct_s, dtype=np.int32).reshape(-1, 12)
        self.oct_d = np.ascontiguousarray(scene.oct_d, dtype=np.int32).reshape(-1, 12)
        self.col_s = np.ascontiguousarray(scene.col_s, dtype=np.float32).reshape(-1, 3)
        self.nrm_s = np.ascontiguousarray(scene.nrm_s, dtype=np.float32).reshape(-1, 3)
        self.col_d = np.ascontiguousarray(scene.col_d, dtype=np.float32).reshape(-1, 3)
        self.nrm_d = np.ascontiguousarray(scene.nrm_d, dtype=np.float32).reshape(-1, 3)
        self.h = lib().qb_emu_scene_create(_p(self.oct_s), len(self.oct_s), _p(self.oct_d), len(self.oct_d),
                                           _p(self.col_s), _p(self.nrm_s), len(self.col_s),
                                           _p(self.col_d), _p(self.nrm_d), len(self.col_d))

    def __del__(self):
        if getattr(self, "h", None):
            lib().qb_emu_scene_destroy(self.h)
            self.h = None


def render(escene, width, height, position, angle, lighta=0.0, quality=10, maxlevel=12, basesize=1800.0, shoot=0,
           light=None, div=0, rows=None):
    """render_fast_kernel on the host.  Returns dict(rgba [H,W,4] u8, flags [H,W] u8, aux [H,W,6] i32)."""
    W, H = C.c_int(), C.c_int()
    lib().qb_emu_render_size(float(width), float(height), int(quality), C.byref(W), C.byref(H))
    W, H = W.value, H.value
    rgba = np.zeros((H, W, 4), np.uint8)
    flags = np.zeros((H, W), np.uint8)
    aux = np.full((H, W, 6), -1, np.int32)
    pos = np.asarray(position, np.float32)
    ang = np.asarray(angle, np.float32)
    lo = np.asarray(light, np.float32) if light is not None else None
    r0, r1 = (0, H) if rows is None else rows
    rc = lib().qb_emu_render(escene.h, float(width), float(height), int(quality), _p(pos), _p(ang), float(lighta),
                             int(maxlevel), float(basesize), int(shoot), _p(lo) if lo is not None else None, int(div),
                             int(r0), int(r1), _p(rgba), _p(flags), _p(aux))
    if rc != 0:
        raise ValueError("the fast kernel does not take this configuration")
    return {"rgba": rgba, "flags": flags, "aux": aux}


def trace_lines(escene, pos, direction, dynamic=False, maxlevel=12, basesize=1800.0):
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    direction = np.ascontiguousarray(direction, np.float32).reshape(-1, 3)
    n = len(pos)
    idx = np.zeros(n, np.int32)
    tlf = np.zeros((n, 4), np.float32)
    lib().qb_emu_trace_lines(escene.h, n, _p(pos), _p(direction), int(bool(dynamic)), int(maxlevel), float(basesize),
                             _p(idx), _p(tlf))
    return idx, tlf


def entry_compare(pos, direction, basecube, div=0):
    """(rays on which base_cube_entry_compact and base_cube_entry_q differ, index of the first such ray or -1)"""
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    direction = np.ascontiguousarray(direction, np.float32).reshape(-1, 3)
    bc = np.asarray(basecube, np.float32)
    first = C.c_long(-1)
    bad = lib().qb_emu_entry_compare(len(pos), _p(pos), _p(direction), _p(bc), int(div), C.byref(first))
    return int(bad), int(first.value)
