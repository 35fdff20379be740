"""ctypes mirror of the reference connector interface.

The C ABI is include/octree_cuc.h; this module binds it one-to-one so that
Python tests read like calls of the reference's octree_glc.c:

    rc = OctreeGlc("shaders/")                       # octree_glc_init   (octree_glc.c L93)
    rc.upload_texbuffer_data(arr, GL_INT, size, 16, start, end, STATIC_OCTREE)   # L361
    rc.update(width, height, pos, angle, lighta, quality, maxlevel, basesize, shoot)  # L249

There is NO fallback: if liboctree_cuc.so is missing the import of this module
raises, and octree_glc_init aborts the process when CUDA is unusable.
"""
import ctypes as C
import os

import numpy as np

from .build import lib_paths

GL_INT = 0x1404
GL_FLOAT = 0x1406

# octree_glc_buffer_t (octree_glc.c L16-24)
STATIC_COLOR, STATIC_NORMAL, STATIC_OCTREE, DYNAMIC_COLOR, DYNAMIC_NORMAL, DYNAMIC_OCTREE = range(6)

FLAG_DISCARD, FLAG_LEAF, FLAG_SHADED, FLAG_LIT, FLAG_DISC_TEST, FLAG_DISC_ON = 1, 2, 4, 8, 16, 32
AUX_MODEL_S, AUX_MODEL_D, AUX_NODE_S, AUX_NODE_D, AUX_SH_NODE_S, AUX_SH_NODE_D = range(6)
AUX_STRIDE = 6

KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FAST = 0, 1, 2
ERROR_FN = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p)   # octree_cuc_set_error_handler
DIV_GLSL, DIV_IEEE = 0, 1
PARTICLES, DUST = 0, 1          # octree_cuc_particles_*: particle_vsh.c / dust_vsh.c


class v3_t(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class octree_glc_t(C.Structure):
    _fields_ = [("impl", C.c_void_p), ("memsize_bytes", C.c_uint64), ("memsize", C.c_uint)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("rays_primary", "rays_shadow", "rays_disc", "expand_s", "expand_d",
                                         "leaf_s", "leaf_d", "hits", "discards", "descents")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# every symbol include/octree_cuc.h declares (tests check the export list against the header)
_PROTOS = {
    "octree_glc_init": (octree_glc_t, [C.c_char_p]),
    "octree_glc_update": (None, [C.POINTER(octree_glc_t), C.c_float, C.c_float, v3_t, v3_t, C.c_float, C.c_uint8,
                                 C.c_int, C.c_float, C.c_int]),
    "octree_glc_upload_texbuffer_data": (None, [C.POINTER(octree_glc_t), C.c_void_p, C.c_int, C.c_size_t,
                                                C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]),
    "octree_cuc_select_device": (None, [C.c_int]),
    "octree_cuc_destroy": (None, [C.POINTER(octree_glc_t)]),
    "octree_cuc_sync": (None, [C.POINTER(octree_glc_t)]),
    "octree_cuc_frame_size": (None, [C.POINTER(octree_glc_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "octree_cuc_read_frame": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_void_p, C.c_size_t]),
    "octree_cuc_frame_device": (C.c_uint64, [C.POINTER(octree_glc_t)]),
    "octree_cuc_read_frame_async": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_void_p, C.c_size_t]),
    "octree_cuc_wait_reads": (None, [C.POINTER(octree_glc_t)]),
    "octree_cuc_set_frame_target": (None, [C.POINTER(octree_glc_t), C.c_uint64, C.c_size_t]),
    "octree_cuc_enable_aux": (None, [C.POINTER(octree_glc_t), C.c_int]),
    "octree_cuc_read_aux": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_void_p, C.c_void_p]),
    "octree_cuc_enable_counters": (None, [C.POINTER(octree_glc_t), C.c_int]),
    "octree_cuc_read_counters": (None, [C.POINTER(octree_glc_t), C.POINTER(Counters)]),
    "octree_cuc_set_shard": (None, [C.POINTER(octree_glc_t), C.c_int, C.c_int, C.c_int, C.c_int]),
    "octree_cuc_set_light": (None, [C.POINTER(octree_glc_t), C.c_void_p]),
    "octree_cuc_set_kernel": (None, [C.POINTER(octree_glc_t), C.c_int]),
    "octree_cuc_last_kernel": (C.c_int, [C.POINTER(octree_glc_t)]),
    "octree_cuc_set_division": (None, [C.POINTER(octree_glc_t), C.c_int]),
    "octree_cuc_last_frame_ms": (C.c_float, [C.POINTER(octree_glc_t)]),
    "octree_cuc_launch_count": (C.c_uint64, [C.POINTER(octree_glc_t)]),
    "octree_cuc_update_views": (None, [C.POINTER(octree_glc_t), C.c_int, C.c_float, C.c_float, C.c_void_p,
                                       C.c_void_p, C.c_float, C.c_uint8, C.c_int, C.c_float, C.c_int]),
    "octree_cuc_set_stream": (None, [C.POINTER(octree_glc_t), C.c_uint64]),
    "octree_cuc_reserve_frame": (None, [C.POINTER(octree_glc_t), C.c_int, C.c_int, C.c_int]),
    "octree_cuc_ipc_export_frame": (None, [C.POINTER(octree_glc_t), C.c_void_p]),
    "octree_cuc_ipc_open": (C.c_uint64, [C.POINTER(octree_glc_t), C.c_void_p]),
    "octree_cuc_ipc_close": (None, [C.POINTER(octree_glc_t), C.c_uint64]),
    "octree_cuc_build_octree_from_paths": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_void_p, C.c_void_p, C.c_void_p,
                                                        C.c_size_t, C.c_int, C.c_int, C.c_int]),
    "octree_cuc_skeleton_alloc_in": (None, [C.POINTER(octree_glc_t), C.c_void_p, C.c_void_p, C.c_size_t]),
    "octree_cuc_skeleton_set_rotations": (None, [C.POINTER(octree_glc_t), C.c_void_p]),
    "octree_cuc_skeleton_update": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                C.c_float, C.c_int]),
    "octree_cuc_skeleton_read_out": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_void_p]),
    "octree_cuc_read_frame_staged": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_void_p, C.c_size_t]),
    "octree_cuc_set_tile_feedback": (None, [C.POINTER(octree_glc_t), C.c_int]),
    "octree_cuc_set_persisting_window": (None, [C.POINTER(octree_glc_t), C.c_size_t]),
    "octree_cuc_set_occupancy": (None, [C.POINTER(octree_glc_t), C.c_int]),
    "octree_cuc_set_error_handler": (None, [C.c_void_p, C.c_void_p]),
    "octree_cuc_enable_present": (None, [C.POINTER(octree_glc_t), C.c_int]),
    "octree_cuc_read_window": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_void_p, C.c_size_t, C.POINTER(C.c_int),
                                            C.POINTER(C.c_int)]),
    "octree_cuc_window_device": (C.c_uint64, [C.POINTER(octree_glc_t)]),
    "octree_cuc_particles_alloc_in": (None, [C.POINTER(octree_glc_t), C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "octree_cuc_particles_update": (None, [C.POINTER(octree_glc_t), C.c_int, C.c_int, C.c_int, C.c_float, v3_t,
                                           C.c_int]),
    "octree_cuc_particles_read_out": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "octree_cuc_voxelise_and_build": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                   C.c_void_p]),
    "octree_cuc_trace_lines": (None, [C.POINTER(octree_glc_t), C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_float, C.c_void_p, C.c_void_p]),
    "octree_cuc_download_points": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_size_t]),
    "octree_cuc_download_octree": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_int, C.c_void_p, C.c_size_t]),
    "octree_cuc_pin_host_buffer": (None, [C.POINTER(octree_glc_t), C.c_void_p, C.c_size_t]),
    "octree_cuc_unpin_host_buffer": (None, [C.POINTER(octree_glc_t), C.c_void_p]),
    "octree_cuc_set_upload_threads": (None, [C.POINTER(octree_glc_t), C.c_int]),
    "octree_cuc_take_upload_ms": (C.c_double, [C.POINTER(octree_glc_t)]),
    "octree_cuc_selftest_div": (C.c_uint64, [C.POINTER(octree_glc_t), C.c_uint64, C.c_uint64]),
    "octree_cuc_debug_order_lut": (None, [C.c_void_p]),
    "octree_cuc_export_pending": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_void_p, C.c_size_t]),
    "octree_cuc_apply_blob": (None, [C.POINTER(octree_glc_t), C.c_void_p, C.c_size_t]),
    "octree_cuc_version": (C.c_char_p, []),
    "octree_cuc_export_pending_device": (C.c_size_t, [C.POINTER(octree_glc_t), C.c_uint64, C.c_size_t]),
    "octree_cuc_apply_blob_device": (None, [C.POINTER(octree_glc_t), C.c_uint64, C.c_size_t]),
    "octree_cuc_set_gpus": (None, [C.POINTER(octree_glc_t), C.c_int, C.c_void_p]),
    "octree_cuc_gpu_count": (C.c_int, [C.POINTER(octree_glc_t)]),
    "octree_cuc_last_step_ms": (C.c_float, [C.POINTER(octree_glc_t)]),
    "octree_cuc_ipc_export_ptr": (None, [C.POINTER(octree_glc_t), C.c_uint64, C.c_void_p]),
    "octree_cuc_fence_device": (C.c_uint64, [C.POINTER(octree_glc_t)]),
    "octree_cuc_set_fence": (None, [C.POINTER(octree_glc_t), C.c_int, C.c_int, C.c_void_p]),
    "octree_cuc_enable_replication_log": (None, [C.POINTER(octree_glc_t), C.c_int]),
}

_lib = None


def load_library():
    """dlopen liboctree_cuc.so and bind every prototype; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_paths()["cuc"]
    if not os.path.exists(path):
        raise ImportError("liboctree_cuc.so is not built (run __graft_entry__.build()); "
                          "qubatron_b200 has no CPU rendering path")
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _f3(v):
    return v3_t(float(v[0]), float(v[1]), float(v[2]))


class OctreeGlc:
    """Host-side mirror of the reference connector (octree_glc.c), one instance per GPU."""

    def __init__(self, path=b"", device=None):
        self.lib = load_library()
        if device is not None:
            self.lib.octree_cuc_select_device(int(device))
        if isinstance(path, str):
            path = path.encode()
        self.rc = self.lib.octree_glc_init(path)
        self._p = C.byref(self.rc)
        self._keep = None

    # ---- reference API -------------------------------------------------------
    def upload_texbuffer_data(self, data, type_, size, itemsize, start, end, buftype):
        """octree_glc_upload_texbuffer_data: `data` = the WHOLE logical host array."""
        arr = np.ascontiguousarray(data)
        self.lib.octree_glc_upload_texbuffer_data(self._p, arr.ctypes.data_as(C.c_void_p), int(type_), int(size),
                                                  int(itemsize), int(start), int(end), int(buftype))

    def update(self, width, height, position, angle, lighta=0.0, quality=10, maxlevel=12, basesize=1800.0, shoot=0):
        """octree_glc_update: render one frame (asynchronous, like a GL draw)."""
        self.lib.octree_glc_update(self._p, float(width), float(height), _f3(position), _f3(angle), float(lighta),
                                   int(quality), int(maxlevel), float(basesize), int(shoot))

    @property
    def memsize(self):
        return int(self.rc.memsize_bytes)

    # ---- conveniences over the reference API ----------------------------------
    def upload_octree(self, nodes, dynamic=False, start_node=0, end_node=None):
        """nodes: int32 [n,12] (whole array); uploads nodes [start_node,end_node)."""
        nodes = np.ascontiguousarray(nodes, dtype=np.int32).reshape(-1, 12)
        n = nodes.shape[0]
        end_node = n if end_node is None else end_node
        self.upload_texbuffer_data(nodes, GL_INT, n * 48, 16, start_node * 48, end_node * 48,
                                   DYNAMIC_OCTREE if dynamic else STATIC_OCTREE)

    def upload_points(self, arr, buftype, start_point=0, end_point=None):
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1, 3)
        n = arr.shape[0]
        end_point = n if end_point is None else end_point
        self.upload_texbuffer_data(arr, GL_FLOAT, n * 12, 12, start_point * 12, end_point * 12, buftype)

    def upload_scene(self, scene):
        """The six uploads of modelutil_load_flat (modelutil.c L227-321)."""
        self.upload_points(scene.col_s, STATIC_COLOR)
        self.upload_points(scene.nrm_s, STATIC_NORMAL)
        self.upload_octree(scene.oct_s, dynamic=False)
        if scene.col_d is not None and len(scene.col_d):
            self.upload_points(scene.col_d, DYNAMIC_COLOR)
            self.upload_points(scene.nrm_d, DYNAMIC_NORMAL)
        self.upload_octree(scene.oct_d, dynamic=True)

    # ---- extensions ---------------------------------------------------------------
    def sync(self):
        self.lib.octree_cuc_sync(self._p)

    def frame_size(self):
        w, h = C.c_int(0), C.c_int(0)
        self.lib.octree_cuc_frame_size(self._p, C.byref(w), C.byref(h))
        return w.value, h.value

    def read_frame(self, out=None, views=1):
        """RGBA8 frame(s) as uint8 [views*H, W, 4], row 0 = bottom."""
        w, h = self.frame_size()
        if out is None:
            out = np.empty((h * views, w, 4), dtype=np.uint8)
        got = self.lib.octree_cuc_read_frame(self._p, out.ctypes.data_as(C.c_void_p), out.nbytes)
        if got == 0:
            raise RuntimeError("read_frame: no frame or buffer too small")
        return out

    def read_frame_async(self, out):
        """Queue the copy of the last frame into `out` (page-locked uint8 array); see wait_reads()."""
        got = self.lib.octree_cuc_read_frame_async(self._p, out.ctypes.data_as(C.c_void_p), out.nbytes)
        if got == 0:
            raise RuntimeError("read_frame_async: no frame or buffer too small")
        return out

    def read_frame_staged(self, out):
        """Like read_frame_async with the render target left in place (sharded frames); see wait_reads()."""
        got = self.lib.octree_cuc_read_frame_staged(self._p, out.ctypes.data_as(C.c_void_p), out.nbytes)
        if got == 0:
            raise RuntimeError("read_frame_staged: no frame or buffer too small")
        return out

    def wait_reads(self):
        self.lib.octree_cuc_wait_reads(self._p)

    def frame_device(self):
        return int(self.lib.octree_cuc_frame_device(self._p))

    def set_frame_target(self, device_ptr, pitch_pixels=0, keepalive=None):
        self._keep = keepalive
        self.lib.octree_cuc_set_frame_target(self._p, int(device_ptr), int(pitch_pixels))

    def enable_aux(self, on=True):
        self.lib.octree_cuc_enable_aux(self._p, int(bool(on)))

    def read_aux(self, views=1):
        w, h = self.frame_size()
        flags = np.empty((h * views, w), dtype=np.uint8)
        aux = np.empty((h * views, w, AUX_STRIDE), dtype=np.int32)
        got = self.lib.octree_cuc_read_aux(self._p, flags.ctypes.data_as(C.c_void_p), aux.ctypes.data_as(C.c_void_p))
        if got == 0:
            raise RuntimeError("read_aux: aux planes are not enabled")
        return flags, aux

    def enable_counters(self, on=True):
        self.lib.octree_cuc_enable_counters(self._p, int(bool(on)))

    def read_counters(self):
        c = Counters()
        self.lib.octree_cuc_read_counters(self._p, C.byref(c))
        return c.as_dict()

    def set_shard(self, rank, world, tile_w=64, tile_h=64):
        self.lib.octree_cuc_set_shard(self._p, int(rank), int(world), int(tile_w), int(tile_h))

    def set_light(self, light):
        if light is None:
            self.lib.octree_cuc_set_light(self._p, None)
        else:
            arr = (C.c_float * 3)(*[float(v) for v in light])
            self.lib.octree_cuc_set_light(self._p, C.cast(arr, C.c_void_p))

    def set_kernel(self, which):
        self.lib.octree_cuc_set_kernel(self._p, int(which))

    def set_division(self, mode):
        """DIV_GLSL (default): a*(1/b) like the reference shader on llvmpipe; DIV_IEEE: like the CPU twin."""
        self.lib.octree_cuc_set_division(self._p, int(mode))

    def last_kernel(self):
        return int(self.lib.octree_cuc_last_kernel(self._p))

    def last_frame_ms(self):
        return float(self.lib.octree_cuc_last_frame_ms(self._p))

    def launch_count(self):
        return int(self.lib.octree_cuc_launch_count(self._p))

    def update_views(self, width, height, positions, angles, lighta=0.0, quality=10, maxlevel=12, basesize=1800.0,
                     shoot=0):
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        ang = np.ascontiguousarray(angles, dtype=np.float32).reshape(-1, 3)
        assert pos.shape == ang.shape
        self.lib.octree_cuc_update_views(self._p, pos.shape[0], float(width), float(height),
                                         pos.ctypes.data_as(C.c_void_p), ang.ctypes.data_as(C.c_void_p),
                                         float(lighta), int(quality), int(maxlevel), float(basesize), int(shoot))
        return pos.shape[0]

    def set_stream(self, cuda_stream):
        self.lib.octree_cuc_set_stream(self._p, int(cuda_stream))

    def reserve_frame(self, width, height, views=1):
        self.lib.octree_cuc_reserve_frame(self._p, int(width), int(height), int(views))

    def ipc_export_frame(self):
        h = np.zeros(64, dtype=np.uint8)
        self.lib.octree_cuc_ipc_export_frame(self._p, h.ctypes.data_as(C.c_void_p))
        return h

    def ipc_export_ptr(self, device_ptr):
        h = np.zeros(64, dtype=np.uint8)
        self.lib.octree_cuc_ipc_export_ptr(self._p, int(device_ptr), h.ctypes.data_as(C.c_void_p))
        return h

    def set_gpus(self, n, devices=None):
        """octree_cuc_set_gpus: drive n devices from this one connector (call before the first upload)."""
        if devices is None:
            self.lib.octree_cuc_set_gpus(self._p, int(n), None)
        else:
            arr = (C.c_int * int(n))(*[int(d) for d in devices])
            self.lib.octree_cuc_set_gpus(self._p, int(n), C.cast(arr, C.c_void_p))

    def gpu_count(self):
        return int(self.lib.octree_cuc_gpu_count(self._p))

    def last_step_ms(self):
        return float(self.lib.octree_cuc_last_step_ms(self._p))

    def fence_device(self):
        return int(self.lib.octree_cuc_fence_device(self._p))

    def set_fence(self, rank, world, fence_ptrs):
        if fence_ptrs is None:
            self.lib.octree_cuc_set_fence(self._p, 0, 1, None)
        else:
            arr = (C.c_uint64 * int(world))(*[int(v) for v in fence_ptrs])
            self.lib.octree_cuc_set_fence(self._p, int(rank), int(world), C.cast(arr, C.c_void_p))

    def enable_replication_log(self, on=True):
        self.lib.octree_cuc_enable_replication_log(self._p, int(bool(on)))

    def ipc_open(self, handle):
        h = np.ascontiguousarray(handle, dtype=np.uint8)
        assert h.size == 64
        return int(self.lib.octree_cuc_ipc_open(self._p, h.ctypes.data_as(C.c_void_p)))

    def ipc_close(self, ptr):
        self.lib.octree_cuc_ipc_close(self._p, int(ptr))

    def build_octree_from_paths(self, paths, first_modind=0, dynamic=True):
        """paths: int32 [n,12] octant digits (host).  Builds the tree on the GPU with the reference numbering."""
        paths = np.ascontiguousarray(paths, dtype=np.int32).reshape(-1, 12)
        p14 = np.ascontiguousarray(paths[:, 0:4])
        p54 = np.ascontiguousarray(paths[:, 4:8])
        p94 = np.ascontiguousarray(paths[:, 8:12])
        return int(self.lib.octree_cuc_build_octree_from_paths(
            self._p, p14.ctypes.data_as(C.c_void_p), p54.ctypes.data_as(C.c_void_p), p94.ctypes.data_as(C.c_void_p),
            len(paths), int(first_modind), 0, DYNAMIC_OCTREE if dynamic else STATIC_OCTREE))

    def build_octree_from_device_paths(self, p14_ptr, p54_ptr, p94_ptr, n, first_modind=0, dynamic=True):
        return int(self.lib.octree_cuc_build_octree_from_paths(
            self._p, C.c_void_p(int(p14_ptr)), C.c_void_p(int(p54_ptr)), C.c_void_p(int(p94_ptr)), int(n),
            int(first_modind), 1, DYNAMIC_OCTREE if dynamic else STATIC_OCTREE))

    def skeleton_alloc_in(self, pnt, nrm):
        pnt = np.ascontiguousarray(pnt, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        assert pnt.shape == nrm.shape
        self.lib.octree_cuc_skeleton_alloc_in(self._p, pnt.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p),
                                              pnt.nbytes)
        self._skin_n = len(pnt)

    def skeleton_update(self, oldbones, newbones, model_count=None, maxlevel=12, basesize=1800.0, build_tree=True):
        ob = np.ascontiguousarray(oldbones, dtype=np.float32).reshape(20, 4)
        nb = np.ascontiguousarray(newbones, dtype=np.float32).reshape(20, 4)
        n = self._skin_n if model_count is None else int(model_count)
        return int(self.lib.octree_cuc_skeleton_update(self._p, ob.ctypes.data_as(C.c_void_p),
                                                       nb.ctypes.data_as(C.c_void_p), n, int(maxlevel),
                                                       float(basesize), int(bool(build_tree))))

    def skeleton_set_rotations(self, rotations):
        if rotations is None:
            self.lib.octree_cuc_skeleton_set_rotations(self._p, None)
        else:
            r = np.ascontiguousarray(rotations, dtype=np.float32).reshape(10, 9)
            self.lib.octree_cuc_skeleton_set_rotations(self._p, r.ctypes.data_as(C.c_void_p))

    def skeleton_read_out(self, n):
        p14, p54, p94 = (np.zeros((n, 4), np.int32) for _ in range(3))
        nrm = np.zeros((n, 3), np.float32)
        pnt = np.zeros((n, 3), np.float32)
        self.lib.octree_cuc_skeleton_read_out(self._p, p14.ctypes.data_as(C.c_void_p), p54.ctypes.data_as(C.c_void_p),
                                              p94.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p),
                                              pnt.ctypes.data_as(C.c_void_p))
        return np.concatenate([p14, p54, p94], axis=1), nrm, pnt

    def set_tile_feedback(self, on=True):
        self.lib.octree_cuc_set_tile_feedback(self._p, int(bool(on)))

    def set_occupancy(self, ctas_per_sm):
        self.lib.octree_cuc_set_occupancy(self._p, int(ctas_per_sm))

    def set_persisting_window(self, persist_bytes):
        self.lib.octree_cuc_set_persisting_window(self._p, int(persist_bytes))

    def enable_present(self, on=True):
        self.lib.octree_cuc_enable_present(self._p, int(bool(on)))

    def read_window(self):
        """The window image of the last update (uint8 [height,width,4], row 0 = bottom); needs enable_present."""
        w, h = C.c_int(0), C.c_int(0)
        self.lib.octree_cuc_read_window(self._p, None, 0, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value, 4), dtype=np.uint8)
        if out.size:
            got = self.lib.octree_cuc_read_window(self._p, out.ctypes.data_as(C.c_void_p), out.nbytes, None, None)
            assert got == out.nbytes
        return out

    def window_device(self):
        return int(self.lib.octree_cuc_window_device(self._p))

    def particles_alloc_in(self, pos, spd, kind=PARTICLES):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        spd = np.ascontiguousarray(spd, dtype=np.float32).reshape(-1, 3)
        assert pos.shape == spd.shape
        self.lib.octree_cuc_particles_alloc_in(self._p, int(kind), pos.ctypes.data_as(C.c_void_p),
                                               spd.ctypes.data_as(C.c_void_p), pos.nbytes)
        self._part_n = getattr(self, "_part_n", {})
        self._part_n[int(kind)] = len(pos)

    def particles_update(self, kind=PARTICLES, count=None, maxlevel=12, basesize=1800.0, campos=(0.0, 0.0, 0.0),
                         steps=1):
        n = self._part_n[int(kind)] if count is None else int(count)
        self.lib.octree_cuc_particles_update(self._p, int(kind), n, int(maxlevel), float(basesize), v3_t(*campos),
                                             int(steps))

    def particles_read_out(self, kind=PARTICLES, count=None):
        """(pos_out f32 [n,3], spd_out f32 [n,3], parked) -- parked only for PARTICLES."""
        n = self._part_n[int(kind)] if count is None else int(count)
        pos = np.zeros((n, 3), np.float32)
        spd = np.zeros((n, 3), np.float32)
        fin = self.lib.octree_cuc_particles_read_out(self._p, int(kind), n, pos.ctypes.data_as(C.c_void_p),
                                                     spd.ctypes.data_as(C.c_void_p))
        return pos, spd, int(fin)

    def voxelise_and_build(self, pos, col_u8, nrm, size=1800, levels=12, dynamic=False, want_order=True):
        """qmc + bulk tree build on the GPU from raw host arrays; returns (count, order int64[m], pos f32[m,3])."""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        col = np.ascontiguousarray(col_u8, dtype=np.uint8).reshape(-1, 3)
        nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        n = len(pos)
        order = np.zeros(n, dtype=np.int64) if want_order else None
        pout = np.zeros((n, 3), dtype=np.float32) if want_order else None
        m = int(self.lib.octree_cuc_voxelise_and_build(
            self._p, pos.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p),
            n, int(size), int(levels), 0, int(bool(dynamic)),
            order.ctypes.data_as(C.c_void_p) if want_order else None,
            pout.ctypes.data_as(C.c_void_p) if want_order else None))
        return (m, order[:m], pout[:m]) if want_order else (m, None, None)

    def trace_lines(self, pos, direction, dynamic=False, maxlevel=12, basesize=1800.0, want_tlf=True):
        """Batched octree_trace_line on one device tree: (index int32[n], tlf float32[n,4] zeros for misses)."""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        direction = np.ascontiguousarray(direction, dtype=np.float32).reshape(-1, 3)
        n = len(pos)
        idx = np.zeros(n, dtype=np.int32)
        tlf = np.zeros((n, 4), dtype=np.float32) if want_tlf else None
        self.lib.octree_cuc_trace_lines(self._p, n, pos.ctypes.data_as(C.c_void_p),
                                        direction.ctypes.data_as(C.c_void_p), int(bool(dynamic)), int(maxlevel),
                                        float(basesize), idx.ctypes.data_as(C.c_void_p),
                                        tlf.ctypes.data_as(C.c_void_p) if want_tlf else None)
        return idx, tlf

    def download_points(self, dynamic=False):
        m = int(self.lib.octree_cuc_download_points(self._p, int(bool(dynamic)), None, None, 0))
        col = np.zeros((m, 3), dtype=np.float32)
        nrm = np.zeros((m, 3), dtype=np.float32)
        self.lib.octree_cuc_download_points(self._p, int(bool(dynamic)), col.ctypes.data_as(C.c_void_p),
                                            nrm.ctypes.data_as(C.c_void_p), m)
        return col, nrm

    def download_octree(self, dynamic=True):
        bt = DYNAMIC_OCTREE if dynamic else STATIC_OCTREE
        n = int(self.lib.octree_cuc_download_octree(self._p, bt, None, 0))
        out = np.zeros((n, 12), dtype=np.int32)
        self.lib.octree_cuc_download_octree(self._p, bt, out.ctypes.data_as(C.c_void_p), n)
        return out

    def pin_host_buffer(self, arr):
        self.lib.octree_cuc_pin_host_buffer(self._p, arr.ctypes.data_as(C.c_void_p), arr.nbytes)

    def set_upload_threads(self, threads):
        """host threads that fill the page-locked staging of bulk uploads from pageable memory (0 = driver's path)"""
        self.lib.octree_cuc_set_upload_threads(self._p, int(threads))

    def unpin_host_buffer(self, arr):
        self.lib.octree_cuc_unpin_host_buffer(self._p, arr.ctypes.data_as(C.c_void_p))

    def take_upload_ms(self):
        return float(self.lib.octree_cuc_take_upload_ms(self._p))

    def selftest_div(self, seed, count):
        return int(self.lib.octree_cuc_selftest_div(self._p, int(seed), int(count)))

    def export_pending(self):
        need = self.lib.octree_cuc_export_pending(self._p, None, 0)
        buf = np.empty(need, dtype=np.uint8)
        self.lib.octree_cuc_export_pending(self._p, buf.ctypes.data_as(C.c_void_p), need)
        return buf

    def apply_blob(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self.lib.octree_cuc_apply_blob(self._p, blob.ctypes.data_as(C.c_void_p), blob.nbytes)

    def export_pending_into(self, buf_ptr, capacity):
        """export_pending into caller memory (e.g. a page-locked tensor); returns the bytes needed / written"""
        return int(self.lib.octree_cuc_export_pending(self._p, C.c_void_p(int(buf_ptr)) if buf_ptr else None,
                                                      int(capacity)))

    def export_pending_device(self, device_ptr, capacity):
        return int(self.lib.octree_cuc_export_pending_device(self._p, int(device_ptr), int(capacity)))

    def apply_blob_device(self, device_ptr, nbytes):
        self.lib.octree_cuc_apply_blob_device(self._p, int(device_ptr), int(nbytes))

    def destroy(self):
        if self.rc.impl:
            self.lib.octree_cuc_destroy(self._p)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
