"""ctypes bindings of the CPU oracle (oracle/octree_fsh_oracle.c) and of the
reference's own compiled code (oracle/_ref/libqubatron_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by qubatron_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboctree_fsh_oracle.so")            # IEEE division (reference CPU twin)
ORACLE_GLSL_SO = os.path.join(HERE, "liboctree_fsh_oracle_glsl.so")  # a * (1/b) (reference GLSL on llvmpipe)
DIV_GLSL, DIV_IEEE = 0, 1
REF_SO = os.path.join(HERE, "_ref", "libqubatron_ref.so")
REF_QMC = os.path.join(HERE, "_ref", "qmc")
REF_GLSL = os.path.join(HERE, "_ref", "glsl_ref")

AUX_STRIDE = 6
FLAG_DISCARD, FLAG_LEAF, FLAG_SHADED, FLAG_LIT, FLAG_DISC_TEST, FLAG_DISC_ON = 1, 2, 4, 8, 16, 32


def build(ref=True):
    """make the oracle; `ref` also (re)builds oracle/_ref when /root/reference exists."""
    r = subprocess.run(["make", "-C", HERE, "all"] + (["ref"] if ref else []), stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout)
    return r.stdout


class _Scene(C.Structure):
    _fields_ = [("oct_s", C.c_void_p), ("nodes_s", C.c_int64), ("oct_d", C.c_void_p), ("nodes_d", C.c_int64),
                ("col_s", C.c_void_p), ("nrm_s", C.c_void_p), ("points_s", C.c_int64),
                ("col_d", C.c_void_p), ("nrm_d", C.c_void_p), ("points_d", C.c_int64)]


class Uniforms(C.Structure):
    _fields_ = [("camfp", C.c_float * 3), ("angle_in", C.c_float * 3), ("light", C.c_float * 3),
                ("basecube", C.c_float * 4), ("dimensions", C.c_float * 2), ("maxlevel", C.c_int32),
                ("shoot", C.c_int32), ("vp_w", C.c_int32), ("vp_h", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("rays_primary", "rays_shadow", "rays_disc", "expand_s", "expand_d",
                                         "leaf_s", "leaf_d", "hits", "discards", "descents")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def algorithmic_bytes(c, pixels):
    """32(E_s+E_d) + 4(L_s+L_d) + 24 H + 4 W H."""
    return 32 * (c["expand_s"] + c["expand_d"]) + 4 * (c["leaf_s"] + c["leaf_d"]) + 24 * c["hits"] + 4 * pixels


_libs = {}


def lib(div=DIV_GLSL):
    """The oracle library for one division semantics (see oracle/Makefile)."""
    if div not in _libs:
        path = ORACLE_GLSL_SO if div == DIV_GLSL else ORACLE_SO
        if not os.path.exists(path):
            build(ref=False)
        l = C.CDLL(path)
        l.qb_oracle_uniforms.argtypes = [C.POINTER(Uniforms), C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_float,
                                         C.c_uint8, C.c_int, C.c_float, C.c_int]
        l.qb_oracle_render.argtypes = [C.POINTER(_Scene), C.POINTER(Uniforms), C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.POINTER(Counters), C.c_int]
        l.qb_oracle_trace_batch.argtypes = [C.POINTER(_Scene), C.POINTER(Uniforms), C.c_int64, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        l.qb_oracle_pixel_ray.argtypes = [C.POINTER(Uniforms), C.c_int, C.c_int, C.c_void_p]
        l.qb_oracle_skin_rot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.qb_oracle_bone_rotations.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.qb_oracle_particles.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.qb_oracle_present.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
        l.qb_oracle_dust.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _libs[div] = l
    return _libs[div]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class OracleScene:
    """Keeps contiguous copies of a qubatron_b200.scene.Scene-like object alive for the C side."""

    def __init__(self, scene):
        self.oct_s = np.ascontiguousarray(scene.oct_s, dtype=np.int32).reshape(-1, 12)
        self.oct_d = np.ascontiguousarray(scene.oct_d, dtype=np.int32).reshape(-1, 12)
        self.col_s = np.ascontiguousarray(scene.col_s, dtype=np.float32).reshape(-1, 3)
        self.nrm_s = np.ascontiguousarray(scene.nrm_s, dtype=np.float32).reshape(-1, 3)
        self.col_d = np.ascontiguousarray(scene.col_d, dtype=np.float32).reshape(-1, 3)
        self.nrm_d = np.ascontiguousarray(scene.nrm_d, dtype=np.float32).reshape(-1, 3)
        self.c = _Scene(_ptr(self.oct_s), len(self.oct_s), _ptr(self.oct_d), len(self.oct_d),
                        _ptr(self.col_s), _ptr(self.nrm_s), len(self.col_s),
                        _ptr(self.col_d), _ptr(self.nrm_d), len(self.col_d))


def uniforms(width, height, position, angle, lighta=0.0, quality=10, maxlevel=12, basesize=1800.0, shoot=0,
             light=None):
    """octree_glc_update's uniform set-up (octree_glc.c L263-284)."""
    u = Uniforms()
    pos = (C.c_float * 3)(*[float(v) for v in position])
    ang = (C.c_float * 3)(*[float(v) for v in angle])
    lib().qb_oracle_uniforms(C.byref(u), float(width), float(height), C.cast(pos, C.c_void_p),
                             C.cast(ang, C.c_void_p), float(lighta), int(quality), int(maxlevel), float(basesize),
                             int(shoot))
    if light is not None:
        for i in range(3):
            u.light[i] = float(light[i])
    return u


def render(oscene, u, rows=None, threads=0, want_aux=True, div=DIV_GLSL):
    """Returns dict(rgba [H,W,4] u8, flags [H,W] u8, aux [H,W,6] i32, counters dict)."""
    W, H = u.vp_w, u.vp_h
    rgba = np.zeros((H, W, 4), dtype=np.uint8)
    flags = np.zeros((H, W), dtype=np.uint8) if want_aux else None
    aux = np.full((H, W, AUX_STRIDE), -1, dtype=np.int32) if want_aux else None
    cnt = Counters()
    r0, r1 = (0, H) if rows is None else rows
    lib(div).qb_oracle_render(C.byref(oscene.c), C.byref(u), int(r0), int(r1), _ptr(rgba), _ptr(flags), _ptr(aux),
                              C.byref(cnt), int(threads))
    return {"rgba": rgba, "flags": flags, "aux": aux, "counters": cnt.as_dict()}


def trace_batch(oscene, u, pos, direction, threads=0, div=DIV_GLSL):
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
    direction = np.ascontiguousarray(direction, dtype=np.float32).reshape(-1, 3)
    n = len(pos)
    result = np.zeros(n, dtype=np.int32)
    nodes = np.zeros((n, 2), dtype=np.int32)
    models = np.zeros((n, 2), dtype=np.int32)
    isp = np.zeros((n, 4), dtype=np.float32)
    lib(div).qb_oracle_trace_batch(C.byref(oscene.c), C.byref(u), n, _ptr(pos), _ptr(direction), _ptr(result),
                                   _ptr(nodes), _ptr(models), _ptr(isp), int(threads))
    return result, nodes, models, isp


def bone_rotations(oldbones, newbones, div=DIV_GLSL):
    """float32 [10, 9]: rot_quat, axis_quat, has_axis per bone pair, evaluated with libm."""
    ob = np.ascontiguousarray(oldbones, dtype=np.float32).reshape(20, 4)
    nb = np.ascontiguousarray(newbones, dtype=np.float32).reshape(20, 4)
    out = np.zeros((10, 9), dtype=np.float32)
    lib(div).qb_oracle_bone_rotations(_ptr(ob), _ptr(nb), _ptr(out))
    return out


def skin(oldbones, newbones, positions, normals, maxlevel=12, basesize=1800.0, div=DIV_GLSL, rotations=None):
    """skeleton_vsh.c main() for n points: (digits int32[n,12], normal_out f32[n,3], skinned point f32[n,3]).
    rotations: optional float32 [10, 9] per-bone rotations replacing the libm ones (see bone_rotations)."""
    ob = np.ascontiguousarray(oldbones, dtype=np.float32).reshape(20, 4)
    nb = np.ascontiguousarray(newbones, dtype=np.float32).reshape(20, 4)
    pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    rot = None if rotations is None else np.ascontiguousarray(rotations, dtype=np.float32).reshape(10, 9)
    n = len(pos)
    cube = np.array([0.0, basesize, basesize, basesize], dtype=np.float32)
    digits = np.zeros((n, 12), dtype=np.int32)
    nout = np.zeros((n, 3), dtype=np.float32)
    pout = np.zeros((n, 3), dtype=np.float32)
    lib(div).qb_oracle_skin_rot(_ptr(ob), _ptr(nb), None if rot is None else _ptr(rot), _ptr(cube), int(maxlevel), n,
                                _ptr(pos), _ptr(nrm), _ptr(digits), _ptr(nout), _ptr(pout))
    return digits, nout, pout


def particles(oct_s, pos, spd, maxlevel=12, basesize=1800.0, div=DIV_GLSL):
    """particle_vsh.c main() for n particles over the static tree `oct_s` (int32 [nodes,12]):
    (pos_out f32[n,3], spd_out f32[n,3], stuck int32[n])."""
    oct_s = np.ascontiguousarray(oct_s, dtype=np.int32).reshape(-1, 12)
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
    spd = np.ascontiguousarray(spd, dtype=np.float32).reshape(-1, 3)
    n = len(pos)
    cube = np.array([0.0, basesize, basesize, basesize], dtype=np.float32)
    po, so, hit = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.int32)
    lib(div).qb_oracle_particles(_ptr(oct_s), len(oct_s), _ptr(cube), int(maxlevel), n, _ptr(pos), _ptr(spd), _ptr(po),
                                 _ptr(so), _ptr(hit))
    return po, so, hit


def dust(campos, pos, spd, div=DIV_GLSL):
    """dust_vsh.c main(): (pos_out, spd_out)."""
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
    spd = np.ascontiguousarray(spd, dtype=np.float32).reshape(-1, 3)
    cam = np.array(list(campos), dtype=np.float32)
    n = len(pos)
    po, so = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    lib(div).qb_oracle_dust(_ptr(cam), n, _ptr(pos), _ptr(spd), _ptr(po), _ptr(so))
    return po, so


def present(frame, u, width, height):
    """octree_glc.c L308-351: the window image (uint8 [height,width,4]) of a rendered frame (uint8 [vp_h,vp_w,4])."""
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    out = np.zeros((int(height), int(width), 4), dtype=np.uint8)
    lib(DIV_GLSL).qb_oracle_present(_ptr(frame), frame.shape[1], frame.shape[0], float(u.dimensions[0]),
                                    float(u.dimensions[1]), int(width), int(height), _ptr(out))
    return out


def render_with_coords(oscene, u, coord_x, coord_y, div=DIV_GLSL, quats=None):
    """render() with the per-pixel `coord` varying supplied by the caller (float32 [H,W] each) instead of the exact
    pixel centres -- e.g. the values llvmpipe's rasteriser interpolated (glsl_coords) -- and, optionally, the view
    quaternions qz, qx (8 floats) as a GL driver's sin / cos produced them (glsl_quats) instead of libm's."""
    cx = np.ascontiguousarray(coord_x, dtype=np.float32)
    cy = np.ascontiguousarray(coord_y, dtype=np.float32)
    assert cx.shape == cy.shape == (u.vp_h, u.vp_w)
    l = lib(div)
    l.qb_oracle_set_coord_override.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    l.qb_oracle_set_quat_override.argtypes = [C.c_void_p]
    l.qb_oracle_set_coord_override(_ptr(cx), _ptr(cy), u.vp_w)
    q = None if quats is None else np.ascontiguousarray(quats, dtype=np.float32).reshape(8)
    if q is not None:
        l.qb_oracle_set_quat_override(_ptr(q))
    try:
        return render(oscene, u, div=div)
    finally:
        l.qb_oracle_set_coord_override(None, None, 0)
        l.qb_oracle_set_quat_override(None)


def glsl_quats(scene, u):
    """qz, qx of main() (octree_fsh.c L406-408) as llvmpipe evaluates them: float32 [8], None if no pixel survives."""
    out = []
    for e in ("qz.x", "qz.y", "qz.z", "qz.w", "qx.x", "qx.y", "qx.z", "qx.w"):
        os.environ["QB_DEBUG_EXPR"] = e
        a, _ = glsl_render(scene, u, mode=9)
        frame, _ = (a, None)
        v = a.view(np.float32).reshape(-1)
        raw = a.reshape(-1)
        live = raw != 0
        if e == "qz.x":
            rgba, _ = glsl_render(scene, u, mode=0)
            alive = (rgba.reshape(-1, 4) != 0).any(axis=1)
            if not alive.any():
                os.environ.pop("QB_DEBUG_EXPR", None)
                return None
            pick = int(np.nonzero(alive)[0][0])     # a pixel that was not discarded carries the value
        out.append(v[pick])
    os.environ.pop("QB_DEBUG_EXPR", None)
    return np.array(out, dtype=np.float32)


def glsl_coords(scene, u):
    """The `coord` varying of every pixel as llvmpipe interpolated it (glsl_ref mode 9): (x, y) float32 [H,W]."""
    out = []
    for e in ("coord.x", "coord.y"):
        os.environ["QB_DEBUG_EXPR"] = e
        a, _ = glsl_render(scene, u, mode=9)
        out.append(a.view(np.float32).reshape(u.vp_h, u.vp_w).copy())
    os.environ.pop("QB_DEBUG_EXPR", None)
    # a pixel whose primary ray is discarded writes nothing: it reads back as 0 -> use the exact centre there
    sx, sy = u.dimensions[0] / np.float32(u.vp_w), u.dimensions[1] / np.float32(u.vp_h)
    ex = ((np.arange(u.vp_w, dtype=np.float32) + np.float32(0.5)) * np.float32(sx))[None, :].repeat(u.vp_h, 0)
    ey = ((np.arange(u.vp_h, dtype=np.float32) + np.float32(0.5)) * np.float32(sy))[:, None].repeat(u.vp_w, 1)
    blank = (out[0] == 0) & (out[1] == 0)
    return np.where(blank, ex, out[0]).astype(np.float32), np.where(blank, ey, out[1]).astype(np.float32)


def pixel_rays(u):
    """Primary ray direction of every pixel: float32 [H,W,3]."""
    W, H = u.vp_w, u.vp_h
    out = np.zeros((H, W, 3), dtype=np.float32)
    d = (C.c_float * 3)()
    l = lib()
    for y in range(H):
        for x in range(W):
            l.qb_oracle_pixel_ray(C.byref(u), x, y, C.cast(d, C.c_void_p))
            out[y, x] = (d[0], d[1], d[2])
    return out


# ---------------------------------------------------------------------------
# the reference's own compiled code (octree.c unmodified), when oracle/_ref exists
# ---------------------------------------------------------------------------

def have_ref():
    return os.path.exists(REF_SO)


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        l = C.CDLL(REF_SO)
        l.qbref_tree_create.restype = C.c_void_p
        l.qbref_tree_create.argtypes = [C.c_float, C.c_int]
        l.qbref_tree_delete.argtypes = [C.c_void_p]
        l.qbref_tree_reset.argtypes = [C.c_void_p]
        l.qbref_tree_insert_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        l.qbref_tree_insert_point.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.qbref_tree_insert_paths.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        l.qbref_tree_remove_point.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.qbref_tree_len.restype = C.c_int64
        l.qbref_tree_len.argtypes = [C.c_void_p]
        l.qbref_tree_data.restype = C.POINTER(C.c_int32)
        l.qbref_tree_data.argtypes = [C.c_void_p]
        l.qbref_trace_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _ref = l
    return _ref


class RefOctree:
    """The reference's octree_t driven through oracle/ref_wrap.c."""

    def __init__(self, basesize=1800.0, levels=12):
        self.l = ref_lib()
        self.h = self.l.qbref_tree_create(float(basesize), int(levels))

    def insert_points(self, pts, first=0):
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
        self.l.qbref_tree_insert_points(self.h, _ptr(pts), len(pts), int(first))

    def insert_point(self, pnt, modind):
        p = np.ascontiguousarray(pnt, dtype=np.float32)
        t = np.zeros(13, dtype=np.int32)
        self.l.qbref_tree_insert_point(self.h, _ptr(p), int(modind), _ptr(t))
        return t

    def insert_paths(self, paths, first=0):
        paths = np.ascontiguousarray(paths, dtype=np.int32).reshape(-1, 12)
        self.l.qbref_tree_insert_paths(self.h, _ptr(paths), len(paths), int(first))

    def remove_point(self, pnt):
        p = np.ascontiguousarray(pnt, dtype=np.float32)
        m, o = C.c_int32(-1), C.c_int32(-1)
        self.l.qbref_tree_remove_point(self.h, _ptr(p), C.byref(m), C.byref(o))
        return m.value, o.value

    def reset(self):
        self.l.qbref_tree_reset(self.h)

    def __len__(self):
        return int(self.l.qbref_tree_len(self.h))

    def nodes(self):
        return np.ctypeslib.as_array(self.l.qbref_tree_data(self.h), shape=(len(self), 12)).copy()

    def trace(self, pos, direction, threads=0):
        """octree_trace_line over rays: returns (oct[8] of leaf or 0, tlf [n,4])."""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        direction = np.ascontiguousarray(direction, dtype=np.float32).reshape(-1, 3)
        n = len(pos)
        idx = np.zeros(n, dtype=np.int32)
        tlf = np.zeros((n, 4), dtype=np.float32)
        self.l.qbref_trace_batch(self.h, n, _ptr(pos), _ptr(direction), _ptr(idx), _ptr(tlf), int(threads))
        return idx, tlf

    def __del__(self):
        try:
            self.l.qbref_tree_delete(self.h)
        except Exception:
            pass


# ---------------------------------------------------------------------------
# the reference's unmodified GLSL on Mesa llvmpipe (oracle/_ref/glsl_ref), when available
# ---------------------------------------------------------------------------

def have_glsl():
    return os.path.exists(REF_GLSL) and os.path.exists(
        os.environ.get("QB_MESA_LIBGL",
                       "/opt/nvidia/nsight-compute/2025.2.1/host/linux-desktop-glibc_2_11_3-x64/Mesa/libGL.so.1"))


def glsl_render(scene, u, mode=0, repeat=1, threads=None, workdir=None, window=None):
    """Run the reference shader for the uniforms `u`; returns (uint8 [H,W,4], info dict).
    mode 0 = shader as shipped; 1/2/3 = aux dumps (static model index, dynamic model index, shadow bit)
    returned as int32 [H,W] decoded from the RGBA8 bytes."""
    import json
    import tempfile
    W, H = u.vp_w, u.vp_h
    hdr = np.array([len(scene.oct_s), len(scene.oct_d), len(scene.col_s), len(scene.col_d), W, H, u.maxlevel,
                    u.shoot], dtype=np.int64)
    uf = np.zeros(16, dtype=np.float32)
    uf[0:3] = list(u.camfp)
    uf[3:6] = list(u.angle_in)
    uf[6:9] = list(u.light)
    uf[9:13] = list(u.basecube)
    uf[13:15] = list(u.dimensions)
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.rgba")
        with open(fin, "wb") as f:
            f.write(hdr.tobytes())
            f.write(uf.tobytes())
            for a, dt in ((scene.oct_s, np.int32), (scene.oct_d, np.int32), (scene.col_s, np.float32),
                          (scene.nrm_s, np.float32), (scene.col_d, np.float32), (scene.nrm_d, np.float32)):
                f.write(np.ascontiguousarray(a, dtype=dt).tobytes())
        env = dict(os.environ)
        if threads is not None:
            env["LP_NUM_THREADS"] = str(int(threads))
        if window is not None:   # mode 40: present into a window of this size
            env["QB_WINDOW"] = "%dx%d" % tuple(window)
            W, H = window
        r = subprocess.run([REF_GLSL, fin, fout, str(mode), str(repeat)], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError("glsl_ref failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        info = json.loads(r.stdout.strip().splitlines()[-1])
        out = np.fromfile(fout, dtype=np.uint8).reshape(H, W, 4)
    if mode not in (0, 40):
        out = out.view(np.int32).reshape(H, W)
    return out, info


def _glsl_skin_run(mode, ob, nb, pos, nrm, maxlevel, basesize, repeat=1, workdir=None, env=None):
    import json
    import tempfile
    n = len(pos)
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(np.array([n, maxlevel], dtype=np.int64).tobytes())
            f.write(np.array([basesize, 0.0], dtype=np.float32).tobytes())
            for a in (ob, nb, pos, nrm):
                f.write(a.tobytes())
        e = dict(os.environ)
        e.update(env or {})
        r = subprocess.run([REF_GLSL, fin, fout, str(mode), str(repeat)], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, env=e)
        if r.returncode != 0:
            raise RuntimeError("glsl_ref failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        info = json.loads(r.stdout.strip().splitlines()[-1])
        if info["gl_error"]:
            raise RuntimeError("glsl_ref: GL error %d" % info["gl_error"])
        raw = np.fromfile(fout, dtype=np.uint8)
    o = n * 16
    return [raw[k * o:(k + 1) * o] for k in range(3)] + [raw[3 * o:3 * o + n * 12]], info


def _skin_args(oldbones, newbones, positions, normals):
    return (np.ascontiguousarray(oldbones, dtype=np.float32).reshape(20, 4),
            np.ascontiguousarray(newbones, dtype=np.float32).reshape(20, 4),
            np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3),
            np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3))


def glsl_skin(oldbones, newbones, positions, normals, maxlevel=12, basesize=1800.0, points=False, repeat=1,
              workdir=None):
    """Run the reference's skinning vertex program (skeleton_vsh.c) through transform feedback on llvmpipe.
    Returns (digits int32 [n,12], normal_out f32 [n,3], info) with the shader exactly as shipped (glsl_ref mode 20);
    points=True captures main()'s `pnt` in place of normal_out (mode 21, one statement added)."""
    ob, nb, pos, nrm = _skin_args(oldbones, newbones, positions, normals)
    n = len(pos)
    bufs, info = _glsl_skin_run(21 if points else 20, ob, nb, pos, nrm, maxlevel, basesize, repeat, workdir)
    digits = np.concatenate([b.view(np.int32).reshape(n, 4) for b in bufs[:3]], axis=1)
    return digits, bufs[3].view(np.float32).reshape(n, 3).copy(), info


def glsl_bone_rotations(oldbones, newbones, positions, normals, workdir=None):
    """The per-bone rotations as the GL driver evaluates them (sin / cos / acos are implementation-defined):
    glsl_ref mode 22 captures oldbone_rot_quat / bonesangle_rot_quat of ONE chosen bone pair per run.  Returns
    float32 [10, 9] in bone_rotations' layout and a bool [10] mask of the pairs that had any point in range (the
    others never contribute)."""
    ob, nb, pos, nrm = _skin_args(oldbones, newbones, positions, normals)
    n = len(pos)
    out = np.zeros((10, 9), dtype=np.float32)
    seen = np.zeros(10, dtype=bool)
    for k in range(10):
        bufs, _ = _glsl_skin_run(22, ob, nb, pos, nrm, 12, 1800.0, workdir=workdir, env={"QB_SKIN_BONE": str(2 * k)})
        rq = bufs[0].view(np.float32).reshape(n, 4)
        aq = bufs[1].view(np.float32).reshape(n, 4)
        ident = bufs[2].view(np.int32).reshape(n, 4)
        hit = np.nonzero(ident[:, 0] == 2 * k)[0]
        if len(hit):
            j = hit[0]
            assert (rq[hit].view(np.uint32) == rq[j].view(np.uint32)).all()     # per-bone, not per-point
            assert (aq[hit].view(np.uint32) == aq[j].view(np.uint32)).all()
            seen[k] = True
            out[k, 0:4] = rq[j]
            out[k, 4:8] = aq[j]
            out[k, 8] = float(ident[j, 1])
    return out, seen


def glsl_particles(oct_s, pos, spd, maxlevel=12, basesize=1800.0, dust_campos=None, repeat=1, workdir=None):
    """Run the reference's particle step (particle_vsh.c) -- or, with dust_campos, its dust step (dust_vsh.c) --
    unmodified through transform feedback on llvmpipe (glsl_ref modes 30 / 31). Returns (pos_out, spd_out, info)."""
    import json
    import tempfile
    oct_s = np.ascontiguousarray(oct_s, dtype=np.int32).reshape(-1, 12)
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
    spd = np.ascontiguousarray(spd, dtype=np.float32).reshape(-1, 3)
    n = len(pos)
    mode = 30 if dust_campos is None else 31
    cam = [0.0, 0.0, 0.0] if dust_campos is None else list(dust_campos)
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(np.array([n, maxlevel, len(oct_s) if mode == 30 else 0], dtype=np.int64).tobytes())
            f.write(np.array([basesize] + cam, dtype=np.float32).tobytes())
            f.write(pos.tobytes())
            f.write(spd.tobytes())
            if mode == 30:
                f.write(oct_s.tobytes())
        r = subprocess.run([REF_GLSL, fin, fout, str(mode), str(repeat)], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True)
        if r.returncode != 0:
            raise RuntimeError("glsl_ref failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        info = json.loads(r.stdout.strip().splitlines()[-1])
        if info["gl_error"]:
            raise RuntimeError("glsl_ref: GL error %d" % info["gl_error"])
        raw = np.fromfile(fout, dtype=np.float32)
    return raw[:3 * n].reshape(n, 3).copy(), raw[3 * n:].reshape(n, 3).copy(), info
