"""Drop-in for the reference's top-level `voxlib` extension module (GANcraft voxlib).

The reference imports it as `import voxlib` / `from voxlib import ray_voxel_intersection_perspective`
(imaginaire/model_utils/gancraft/voxlib/{__init__,positional_encoding,sp_trilinear}.py) and the
pybind table it expects is imaginaire/model_utils/gancraft/voxlib/voxlib.cpp:25-31.  Put this
directory (dropin/) on PYTHONPATH ahead of any built reference extension and the reference's
Python runs unchanged on the B200-native kernels of libsdb200 (scenedreamer_b200.ops).
"""
from scenedreamer_b200 import ops as _ops
from scenedreamer_b200.ops import (  # noqa: F401
    positional_encoding,
    positional_encoding_backward,
    sp_trilinear_worldcoord,
    sp_trilinear_worldcoord_backward,
)


def ray_voxel_intersection_perspective(in_voxel, cam_ori, cam_dir, cam_up, cam_f, cam_c, img_dims, max_samples):
    """voxlib.cpp:11 -- exactly the reference's eight positional arguments.  The empty-space bound of the volume is
    built on first use and cached per voxel tensor (scenedreamer_b200.ops.height_bound); results are bit-identical."""
    return _ops.ray_voxel_intersection_perspective(in_voxel, cam_ori, cam_dir, cam_up, cam_f, cam_c, img_dims, max_samples)
