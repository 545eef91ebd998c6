"""Drop-in for the reference's `_gridencoder` extension module (torch-ngp grid encoder).

Bound in the reference by gridencoder/src/bindings.cpp:5-8 and imported by gridencoder/grid.py:9-12
(`import _gridencoder as _backend`).  Same two entry points, same argument order and meaning.
"""
from scenedreamer_b200.ops import grid_encode_backward, grid_encode_forward  # noqa: F401
