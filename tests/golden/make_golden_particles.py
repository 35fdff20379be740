"""Generate tests/golden/particles_cloud.npz and dust_box.npz: outputs of the reference's UNMODIFIED particle and dust
simulation programs (shaders/particle_vsh.c, shaders/dust_vsh.c from /root/reference) run through transform feedback
on Mesa llvmpipe by oracle/_ref/glsl_ref (modes 30 / 31, see oracle/glsl_ref.c).

Run in the build container only (needs /root/reference for `make -C oracle ref`):

    python tests/golden/make_golden_particles.py

particles_cloud.npz: the static octree (12-int nodes), particle positions and speeds, and the program's outputs
after 1 step and after STEPS steps (each step's outputs fed back as the next step's inputs, as modelutil.c L715-724
does).  The particle set covers: debris spawned around the cloud, axis-parallel and zero speeds (the program's
vec4(0.0) parallel-ray result), parked particles, particles outside the base cube, and particles dropped onto leaves
that hang under child slots 4..7 of nodes whose first texel is the last of a texture row (the program's missing row
wrap, particle_vsh.c L108-119) next to a control group under the following nodes.
particles_rowend.npz: a tree built so that a LEVEL-3 node sits at index 5461 (3 * 5461 mod 8192 == 8191): a sheet of
points inside that node's cube, half of it in child slots 4..7 (invisible to the particle program, which fetches those
children outside the texture), half in slots 0..3; particles dropped onto both halves.
dust_box.npz: camera position, motes, outputs after 1 and STEPS steps.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import qb_oracle as O  # noqa: E402
from qubatron_b200 import scene as S  # noqa: E402

STEPS = 12


def leaves_under(oct, node):
    out, todo = [], [int(node)]
    while todo:
        c = todo.pop()
        kids = [int(k) for k in oct[c, :8] if k]
        if kids:
            todo.extend(kids)
        else:
            out.append(int(oct[c, 8]))
    return out


def row_end_targets(oct, shift):
    """model indices of the leaves under child slots 4..7 of the nodes i (+shift) with (3 i) mod 8192 == 8191"""
    nodes = np.arange(len(oct))
    sel = nodes[((3 * nodes) & 8191) == 8191] + shift
    out = []
    for i in sel[sel < len(oct)]:
        for slot in range(4, 8):
            if oct[i, slot]:
                out.extend(leaves_under(oct, oct[i, slot]))
    return np.array(sorted(set(out)), dtype=np.int64)


def drop_onto(pnt, models):
    pos = (pnt[models] + np.float32([0.0, 3.0, 0.0])).astype(np.float32)
    spd = np.tile(np.float32([0.013, -1.0, 0.017]), (len(models), 1))
    return pos, spd


if __name__ == "__main__":
    O.build(ref=True)
    assert O.have_glsl(), "oracle/_ref/glsl_ref or the Mesa libGL is missing"
    sc = S.make_random(9000, 0, seed=5)
    rng = np.random.default_rng(3)
    n = 6000
    idx = rng.integers(0, len(sc.pnt_s), n)
    pos = (sc.pnt_s[idx] + rng.normal(0, 8, (n, 3))).astype(np.float32)
    spd = rng.normal(0, 2.0, (n, 3)).astype(np.float32)
    spd[:300, 0] = 0
    spd[300:600, 2] = 0
    spd[600:700] = 0
    spd[700:800, 0] = -100000.0
    pos[800:900] = rng.uniform(-50, 1900, (100, 3))
    bug_p, bug_s = drop_onto(sc.pnt_s, row_end_targets(sc.oct_s, 0))
    ctl_p, ctl_s = drop_onto(sc.pnt_s, row_end_targets(sc.oct_s, 1))
    groups = np.array([n, len(bug_p), len(ctl_p)])
    pos = np.concatenate([pos, bug_p, ctl_p])
    spd = np.concatenate([spd, bug_s, ctl_s])
    p, s = pos, spd
    outs = {}
    for k in range(1, STEPS + 1):
        p, s, info = O.glsl_particles(sc.oct_s, p, s)
        if k in (1, STEPS):
            outs["pos_%d" % k], outs["spd_%d" % k] = p, s
    path = os.path.join(HERE, "particles_cloud.npz")
    np.savez_compressed(path, oct_s=sc.oct_s, pos=pos, spd=spd, groups=groups, steps=np.array(STEPS),
                        renderer=np.array(info["renderer"] + " / " + info["version"]), **outs)
    parked = (outs["spd_%d" % STEPS][:, 0] < -900).sum()
    b0, b1 = n, n + len(bug_p)
    print("particles_cloud: %d nodes, %d particles (%d row-end, %d control), %d KB, parked after %d steps: %d; "
          "row-end group stuck at step 1: %d / %d, control: %d / %d" %
          (len(sc.oct_s), len(pos), len(bug_p), len(ctl_p), os.path.getsize(path) // 1024, STEPS, parked,
           (outs["spd_1"][b0:b1, 0] < -900).sum(), len(bug_p), (outs["spd_1"][b1:, 0] < -900).sum(), len(ctl_p)))

    # ---- a level-3 node at a row end -------------------------------------------------------------------------
    host = S.HostOctree()
    a_pts = rng.uniform(100.0, 800.0, (4000, 3)).astype(np.float32)
    pts = []
    target = 5459                                   # nodes 0..5458 exist, the next point creates 5459.. for levels 1..
    for q in a_pts:
        if len(host) >= target - 12:
            break
        host.insert_point(q, len(pts))
        pts.append(q)
    k = 0
    while len(host) < target:                       # a neighbour in the same level-11 cube adds exactly one node
        q = np.floor(pts[k] / np.float32(0.87890625)) * np.float32(0.87890625) + np.float32([0.1, 0.1, 0.1])
        for dq in ([0.0, 0.0, 0.0], [0.5, 0.0, 0.0], [0.0, 0.5, 0.0], [0.0, 0.0, 0.5], [0.5, 0.5, 0.0]):
            if len(host) >= target:
                break
            before = len(host)
            qq = (q + np.float32(dq)).astype(np.float32)
            host.insert_point(qq, len(pts))
            if len(host) - before > 1:
                raise SystemExit("filler added %d nodes" % (len(host) - before))
            pts.append(qq)
        k += 1
    assert len(host) == target
    gx, gz = np.meshgrid(np.arange(910.0, 1116.0, 3.0), np.arange(910.0, 1116.0, 3.0), indexing="ij")
    sheet = np.stack([gx.ravel(), np.full(gx.size, 1000.0), gz.ravel()], axis=1).astype(np.float32)
    first = len(pts)
    for q in sheet:
        host.insert_point(q, len(pts))
        pts.append(q)
    oct_r = host.nodes()
    assert (3 * 5461) % 8192 == 8191 and oct_r[5461, :8].any()
    pos_r = (sheet + np.float32([0.0, 3.0, 0.0])).astype(np.float32)
    spd_r = np.tile(np.float32([0.013, -1.0, 0.017]), (len(sheet), 1))
    p, s = pos_r, spd_r
    outs = {}
    for k in range(1, 4):
        p, s, info = O.glsl_particles(oct_r, p, s)
        if k in (1, 3):
            outs["pos_%d" % k], outs["spd_%d" % k] = p, s
    path = os.path.join(HERE, "particles_rowend.npz")
    np.savez_compressed(path, oct_s=oct_r, pos=pos_r, spd=spd_r, steps=np.array(3),
                        renderer=np.array(info["renderer"] + " / " + info["version"]), **outs)
    low = sheet[:, 2] < 1012.5
    stuck1 = outs["spd_1"][:, 0] < -900
    print("particles_rowend: %d nodes, %d particles, %d KB; stuck at step 1: slots 4..7 half %d / %d, slots 0..3 half "
          "%d / %d" % (len(oct_r), len(sheet), os.path.getsize(path) // 1024, stuck1[low].sum(), low.sum(),
                       stuck1[~low].sum(), (~low).sum()))

    cam = (600.0, 150.0, 200.0)
    m = 8000
    dp = np.stack([rng.uniform(380, 820, m), rng.uniform(-10, 310, m), rng.uniform(-10, 410, m)], axis=1).astype(np.float32)
    ds = rng.normal(0, 1.5, (m, 3)).astype(np.float32)
    p, s = dp, ds
    outs = {}
    for k in range(1, STEPS + 1):
        p, s, info = O.glsl_particles(sc.oct_s, p, s, dust_campos=cam)
        if k in (1, STEPS):
            outs["pos_%d" % k], outs["spd_%d" % k] = p, s
    path = os.path.join(HERE, "dust_box.npz")
    np.savez_compressed(path, campos=np.float32(cam), pos=dp, spd=ds, steps=np.array(STEPS), **outs)
    print("dust_box: %d motes, %d KB" % (m, os.path.getsize(path) // 1024))
