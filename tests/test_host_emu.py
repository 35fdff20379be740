"""The LOGIC of the product's own kernel source, checked on the CPU.

tests/host_emu compiles qubatron_b200/csrc/octree_trace_fast.cuh (+ body) for the host with g++ -- the only thing
replaced is octree_ptx.cuh, the one header that holds inline PTX -- and runs render_fast_kernel one thread at a time on
host copies of the HBM layout.  What these tests pin without a GPU: candidate ordering and its lookup table, the
packed-pair statements of the arithmetic, pop / backtrack / re-snap, the sequencing of a pixel's three rays, the
shading behind the loop, the compact base-cube entry.  What they cannot pin is the device compiler's output; that is
the job of the `-m gpu` parity tests, which run the same comparisons through the C ABI on the B200.

The oracle stays the checker: every frame is compared with oracle/ (flags, hit indices, node ids, RGBA all identical)
and with the golden frames of the reference shader."""
import os
import sys

import numpy as np
import pytest

import golden_util
import host_emu as E
import parity
from oracle import qb_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def _same(r, ref, what):
    parity.compare(r["rgba"], r["flags"], r["aux"], ref, what=what)
    assert np.array_equal(r["rgba"], ref["rgba"]), what + ": RGBA differs"


@pytest.mark.parametrize("name", golden_util.CASES)
def test_host_build_of_the_fast_kernel_reproduces_the_reference_shader(name):
    sc, args, g = golden_util.load(name)
    if sc is None:
        pytest.skip("scene generator gives different bits on this CPU (hash mismatch)")
    esc, osc = E.EmuScene(sc), O.OracleScene(sc)
    r = E.render(esc, div=0, **args)
    assert np.array_equal(r["rgba"], g["rgba"]), "RGBA differs from the shader's frame"
    leaf = (r["flags"] & O.FLAG_LEAF) > 0
    assert np.array_equal(g["model_s"][leaf], r["aux"][leaf][:, 0])
    assert np.array_equal(g["model_d"][leaf], r["aux"][leaf][:, 1])
    shaded = (r["flags"] & O.FLAG_SHADED) > 0
    assert np.array_equal(g["shadow"][shaded] != 0, (r["flags"][shaded] & O.FLAG_LIT) > 0)
    for ediv, odiv in ((0, O.DIV_GLSL), (1, O.DIV_IEEE)):
        _same(E.render(esc, div=ediv, **args), O.render(osc, O.uniforms(**args), div=odiv), "%s div %d" % (name, ediv))


def test_host_build_fuzz_against_the_oracle():
    """40 random scenes / cameras (inside, outside, on faces and grid planes, axis-parallel) / lights / depths /
    render scales, both divisions (scripts/emu_fuzz.py runs the long version: 300 cases, 5.4 M pixels, identical)."""
    from emu_fuzz import fuzz_case
    for seed in range(5000, 5040):
        sc, W, H, pos, ang, kw = fuzz_case(seed)
        osc, esc = O.OracleScene(sc), E.EmuScene(sc)
        for ediv, odiv in ((0, O.DIV_GLSL), (1, O.DIV_IEEE)):
            _same(E.render(esc, W, H, pos, ang, div=ediv, **kw), O.render(osc, O.uniforms(W, H, pos, ang, **kw), div=odiv),
                  "seed %d div %d" % (seed, ediv))


def test_host_build_rows_of_a_c1_frame(scene_c1):
    """bands of a 1280x720 frame of configs[0] (the start pose and a grazing view)"""
    osc, esc = O.OracleScene(scene_c1), E.EmuScene(scene_c1)
    views = list(scene_c1.cameras[:2]) if getattr(scene_c1, "cameras", None) else []
    views.append(((760.0, 200.0, 420.0), (-0.05, -0.12, 0.0)))
    for pos, ang in views:
        for rows in ((96, 128), (352, 384), (640, 664)):
            ref = O.render(osc, O.uniforms(1280, 720, pos, ang), rows=rows, div=O.DIV_GLSL)
            r = E.render(esc, 1280, 720, pos, ang, div=0, rows=rows)
            band = slice(rows[0] // 8 * 8, min(720, (rows[1] + 7) // 8 * 8))
            inner = slice(rows[0], rows[1])
            for k in ("rgba", "flags", "aux"):
                assert np.array_equal(r[k][inner], ref[k][inner]), (pos, rows, k)
            assert band.start <= rows[0]


def _entry_rays(seed, n, S):
    rng = np.random.default_rng(seed)
    grid = np.array([0, S, S / 2, S / 4, 3 * S / 4, S / 4096, S - S / 4096], np.float32)
    pos = rng.uniform(-0.5 * S, 1.5 * S, (n, 3)).astype(np.float32)
    m = rng.random((n, 3)) < 0.25
    pos[m] = rng.choice(grid, size=int(m.sum()))          # origins on faces, edges, corners, grid planes
    tgt = rng.uniform(0, S, (n, 3)).astype(np.float32)
    m = rng.random((n, 3)) < 0.4
    tgt[m] = rng.choice(grid[:5], size=int(m.sum()))      # aimed at faces, edges, corners
    d = (tgt - pos).astype(np.float32)
    k = rng.random(n) < 0.3
    d[k] = rng.normal(0, 1, (int(k.sum()), 3)).astype(np.float32)
    d[rng.random((n, 3)) < 0.08] = 0.0                    # parallel to one, two or three plane pairs
    d[rng.random((n, 3)) < 0.02] = -0.0
    return pos, d


@pytest.mark.parametrize("S", [1800.0, 2048.0, 1234.567])
def test_compact_base_cube_entry_equals_the_face_by_face_statement(S):
    """base_cube_entry_compact (kernel v15) against base_cube_entry_q on adversarial rays: the discard decision and
    every bit of the entry point, both division rules.  (80 M rays of this generator were run once: no difference.)"""
    for div in (0, 1):
        pos, d = _entry_rays(int(S) + div, 1_000_000, S)
        bad, first = E.entry_compare(pos, d, [0.0, S, S, S], div)
        assert bad == 0, (bad, pos[first], d[first])


def test_host_build_trace_lines_equal_the_reference_cpu_function(scene_random):
    """trace_lines_fast_kernel on the host against the reference's compiled octree_trace_line (oracle/_ref)"""
    rng = np.random.default_rng(52)
    n = 6000
    org = np.stack([rng.uniform(600, 900, n), rng.uniform(100, 250, n), rng.uniform(100, 450, n)], axis=1).astype(
        np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:300, 0] = 0.0
    d[300:600, 1] = 0.0
    d[600:700, :2] = 0.0
    org[700:1000] += np.float32(2500.0)
    pts, nodes = scene_random.pnt_s, scene_random.oct_s
    aim = rng.integers(0, len(pts), n // 2)
    d[n // 2:] = (np.asarray(pts)[aim] - org[n // 2:]).astype(np.float32)
    ref = O.RefOctree()
    ref.insert_points(pts)
    assert np.array_equal(ref.nodes(), nodes)
    want_idx, want_tlf = ref.trace(org, d)
    got_idx, got_tlf = E.trace_lines(E.EmuScene(scene_random), org, d)
    hit = want_idx != 0
    assert hit.sum() > 500
    assert np.array_equal(got_idx, want_idx)
    assert np.array_equal(got_tlf[hit], want_tlf[hit])


def test_product_does_not_touch_the_host_build():
    pkg = os.path.join(ROOT, "qubatron_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h", ".inc", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("host_emu/", "import host_emu", "libqb_host_emu", "cuda_host_shim"):
                    if needle == "host_emu/" and f == "octree_ptx.cuh" or f == "octree_view_host.h":
                        continue  # the two comments that say where the test build lives
                    assert needle not in txt, (dirpath, f, needle)
    assert "QB_PTX_HOST_HEADER" not in open(os.path.join(pkg, "csrc", "Makefile")).read()
