"""Static checks on the compiled hot kernel (cuobjdump, no GPU needed).

The fast traversal uses Blackwell's packed fp32 instructions (FADD2 / FMUL2, octree_trace_fast_body.inc).  Results stay
identical to the scalar statement only as long as no packed product is contracted with a packed sum: ptxas does that
to mul.rn.f32x2 + add.rn.f32x2 whatever the rounding modifiers say, so the source never hands a packed product to a
packed sum -- and this test makes sure no FFMA2 (and no scalar FFMA outside the reciprocals, square roots and the hoisted
IEEE division) appears."""
import re
import shutil
import subprocess

import pytest

from qubatron_b200.build import lib_paths


def _sass_by_function():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-sass", lib_paths()["cuc"]], stdout=subprocess.PIPE, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[name].append(line)
    return funcs


@pytest.fixture(scope="module")
def sass():
    try:
        return _sass_by_function()
    except (OSError, subprocess.CalledProcessError) as e:
        pytest.skip("cuobjdump unavailable: %r" % (e,))


def test_hot_kernel_is_compiled_for_sm_100a_with_packed_fp32_and_no_contraction(sass):
    hot = {k: v for k, v in sass.items() if "render_fast_kernel" in k}
    assert len(hot) == 16                      # DIV x DYN x AUX x COUNT
    for name, lines in hot.items():
        text = "\n".join(lines)
        assert "FFMA2" not in text, name       # a contracted packed product would change the rounding
        glsl = "render_fast_kernelILi0E" in name
        if glsl:
            assert text.count("FADD2") >= 5 and text.count("FMUL2") >= 5, name
            # GLSL division (a * rcp(b)): the only FFMAs are those of 1.0f / d and sqrt, none in the traversal
            assert len(re.findall(r"\bFFMA\b", text)) < 80, name


def test_single_ray_kernels_share_the_traversal(sass):
    for frag in ("particle_step_kernelILi0ELb1E", "trace_lines_fast_kernel"):
        names = [k for k in sass if frag in k]
        assert names, frag
        for n in names:
            text = "\n".join(sass[n])
            assert "FFMA2" not in text and "LDS" in text and "STS" in text, n   # shared-memory stack, no contraction
