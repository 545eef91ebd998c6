"""`gridencoder` package surface of the reference (gridencoder/grid.py) on the libsdb200 backend.

Same constructor arguments, attributes (`embeddings`, `offsets`, `output_dim`, `per_level_scale`, ...),
state-dict names/shapes and autograd behaviour as the reference's GridEncoder / VarGridEncoder
(gridencoder/grid.py:93-156, :158-233) and its `_grid_encode` Function (:19-87).  Only the float32 path is
implemented (the SceneDreamer configs train and run with AMP disabled); under autocast the
embeddings stay float32 instead of being cast to half as the reference does (:38-39).
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

import _gridencoder as _backend

_gridtype_to_id = {'hash': 0, 'tiled': 1}


class _grid_encode(Function):
    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False):
        # inputs [B, D] in [0, 1]; embeddings [sum T_l, C]; offsets [L+1] -> [B, L*C]   (grid.py:22-59)
        inputs = inputs.contiguous().float()
        embeddings = embeddings.contiguous().float()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = np.log2(per_level_scale)
        H = base_resolution
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        if calc_grad_inputs:
            dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype)
        else:
            dy_dx = torch.empty(1, device=inputs.device, dtype=embeddings.dtype)
        _backend.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx,
                                     gridtype, align_corners)
        outputs = outputs.permute(1, 0, 2).reshape(B, L * C)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H, gridtype]
        ctx.calc_grad_inputs = calc_grad_inputs
        ctx.align_corners = align_corners
        return outputs

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype = ctx.dims
        calc_grad_inputs = ctx.calc_grad_inputs
        grad = grad.float().view(B, L, C).permute(1, 0, 2).contiguous()           # grid.py:72
        grad_embeddings = torch.zeros_like(embeddings)
        if calc_grad_inputs:
            grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype)
        else:
            grad_inputs = torch.zeros(1, device=inputs.device, dtype=embeddings.dtype)
        _backend.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H,
                                      calc_grad_inputs, dy_dx, grad_inputs, gridtype, ctx.align_corners)
        if calc_grad_inputs:
            return grad_inputs.to(inputs.dtype), grad_embeddings, None, None, None, None, None, None
        return None, grad_embeddings, None, None, None, None, None, None


grid_encode = _grid_encode.apply


def _level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners):
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        n = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return torch.from_numpy(np.array(offsets, dtype=np.int32)), offset


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.align_corners = align_corners
        self.max_params = 2 ** log2_hashmap_size
        offsets, offset = _level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size,
                                         align_corners)
        self.register_buffer('offsets', offsets)
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        return ('GridEncoder: input_dim=%d num_levels=%d level_dim=%d resolution=%d -> %d per_level_scale=%.4f '
                'params=%s gridtype=%s align_corners=%s' % (
                    self.input_dim, self.num_levels, self.level_dim, self.base_resolution,
                    int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1))),
                    self.per_level_scale, tuple(self.embeddings.shape), self.gridtype, self.align_corners))

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)                                  # grid.py:144
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners)
        return outputs.view(prefix_shape + [self.output_dim])


class VarGridEncoder(nn.Module):
    """grid.py:158-233: the first `hash_entries` rows of the table come from the caller at forward time."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False,
                 hash_entries=None):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.align_corners = align_corners
        self.max_params = 2 ** log2_hashmap_size
        offsets, offset = _level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size,
                                         align_corners)
        self.register_buffer('offsets', offsets)
        self.n_params = offsets[-1] * level_dim
        self.offset = offset
        self.embeddings = nn.Parameter(torch.empty(offset - hash_entries, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def forward(self, inputs, embeddings, bound=1):
        input_embeddings = torch.cat([embeddings, self.embeddings], dim=0)
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, input_embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners)
        return outputs.view(prefix_shape + [self.output_dim])
