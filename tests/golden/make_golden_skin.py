"""Generate tests/golden/skin_*.npz: outputs of the reference's UNMODIFIED skinning vertex program
(shaders/skeleton_vsh.c from /root/reference) run through transform feedback on Mesa llvmpipe by
oracle/_ref/glsl_ref (modes 20-22, see oracle/glsl_ref.c).

Run in the build container only (needs /root/reference for `make -C oracle ref`):

    python tests/golden/make_golden_skin.py

skin_inputs.npz holds the rest-pose positions and normals of a sample of the capsule figure; each pose fixture
holds its oldbones / newbones (20 x vec4) and what the shader produced:
    digits     int32 [n,12]   oct14 | oct54 | oct94            (mode 20, shader as shipped)
    normal_out f32   [n,3]                                     (mode 20)
    pnt        f32   [n,3]    main()'s skinned position        (mode 21, captured in place of normal_out)
    rotations  f32   [10,9]   oldbone_rot_quat, bonesangle_rot_quat, has_axis per bone pair as llvmpipe evaluated
                              them (mode 22) -- sin / cos / acos are implementation-defined in GLSL ES, so these are
                              the only driver-dependent values; everything else must match bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import qb_oracle as O  # noqa: E402
from qubatron_b200 import scene as S  # noqa: E402

POSES = {"skin_rest": (0.0, (0.0, 0.0, 0.0)), "skin_walk": (1.0, (3.0, 0.0, -2.0)), "skin_bent": (2.5, (7.5, 0.0, -5.0))}

if __name__ == "__main__":
    O.build(ref=True)
    assert O.have_glsl(), "oracle/_ref/glsl_ref or the Mesa libGL is missing"
    pos, col, nrm = S.zombie_raw(spacing=0.6, shells=4)
    pos, nrm = pos[::31].copy(), nrm[::31].copy()
    np.savez_compressed(os.path.join(HERE, "skin_inputs.npz"), positions=pos, normals=nrm)
    for name, (pose, shift) in POSES.items():
        ob, nb = S.zombie_bones(pose=pose, shift=shift)
        digits, normal_out, info = O.glsl_skin(ob, nb, pos, nrm)
        _, pnt, _ = O.glsl_skin(ob, nb, pos, nrm, points=True)
        rot, seen = O.glsl_bone_rotations(ob, nb, pos, nrm)
        assert seen.all()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, oldbones=ob, newbones=nb, digits=digits,
                            normal_out=normal_out, pnt=pnt, rotations=rot,
                            renderer=np.array(info["renderer"] + " / " + info["version"]))
        moved = (np.abs(pnt - pos).max(axis=1) > 1.0).mean()
        print("%-10s n=%d  %d KB  moved %.2f  renderer %s" % (name, len(pos), os.path.getsize(path) // 1024, moved,
                                                               info["renderer"]))
