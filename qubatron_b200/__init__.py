"""qubatron_b200 -- B200-native renderer for Qubatron's per-pixel octree hot path.

Only what the path needs lives here:

  csrc/          hand-written sm_100a CUDA kernels + the C-ABI connector
                 (include/octree_cuc.h; replaces the reference's octree_glc.c)
  host/          host-side data model in C (12-int octree array, voxeliser)
  connector.py   ctypes mirror of the reference connector interface
  scene.py       synthetic scenes in the reference's data formats
  build.py       in-tree build of the two shared libraries

The CUDA library is mandatory: importing the connector without it raises, there
is no CPU rendering path in this package (the CPU oracle lives in oracle/ and is
test infrastructure only).
"""
from .build import build_all, lib_paths  # noqa: F401

__all__ = ["build_all", "lib_paths"]
