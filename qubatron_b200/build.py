"""In-tree build of liboctree_cuc.so (CUDA, sm_100a) and libqb_host.so (C)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_CC = "/usr/bin/gcc"


def lib_paths():
    return {
        # QB_CUC_LIB: developer override to A/B-test a differently tuned build of the same sources
        "cuc": os.environ.get("QB_CUC_LIB") or os.path.join(_HERE, "liboctree_cuc.so"),
        "host": os.path.join(_HERE, "libqb_host.so"),
    }


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in srcs)


def build_host(force=False):
    out = lib_paths()["host"]
    src = os.path.join(_HERE, "host", "qb_host.c")
    if force or _stale(out, [src]):
        _run([HOST_CC, "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=gnu11", "-Wall", "-shared", "-fPIC",
              "-o", out, src, "-lm"])
    return out


def build_cuc(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (see csrc/Makefile)."""
    out = lib_paths()["cuc"]
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)]
    srcs.append(os.path.join(_HERE, "..", "include", "octree_cuc.h"))
    if os.environ.get("QB_CUC_LIB"):
        return out  # a developer's A/B build (scripts/build_variants.sh): used as it is, never rebuilt here
    if force or _stale(out, srcs):
        # -B: whatever made the library stale (a header the Makefile does not list, the Makefile itself), rebuild
        _run(["make", "-B", "-C", csrc])
    return out


def build_all(force=False):
    return {"host": build_host(force), "cuc": build_cuc(force)}
