"""Drop-in for the reference's top-level `voxlib` extension module (GANcraft voxlib).

The reference imports it as `import voxlib` / `from voxlib import ray_voxel_intersection_perspective`
(imaginaire/model_utils/gancraft/voxlib/{__init__,positional_encoding,sp_trilinear}.py) and the
pybind table it expects is imaginaire/model_utils/gancraft/voxlib/voxlib.cpp:25-31.  Put this
directory (dropin/) on PYTHONPATH ahead of any built reference extension and the reference's
Python runs unchanged on the B200-native kernels of libsdb200 (scenedreamer_b200.ops).
"""
from scenedreamer_b200.ops import (  # noqa: F401
    positional_encoding,
    positional_encoding_backward,
    ray_voxel_intersection_perspective,
    sp_trilinear_worldcoord,
    sp_trilinear_worldcoord_backward,
)
