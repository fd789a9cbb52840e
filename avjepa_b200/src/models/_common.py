"""Shared constructor / initialisation logic of the encoder and predictor mirrors.

Registration order of sub-modules and the sequence of RNG-consuming initialisers follow the
reference constructors exactly (``src/models/audiovision_transformer.py:27-140``,
``vision_transformer.py:26-131``, ``audiovisionpredictor.py:20-176``, ``predictor.py:26-148``),
so a model built here under ``torch.manual_seed(s)`` has bit-identical parameters to the
reference built under the same seed (checked in tests/test_init_parity.py).
"""
import math

import torch
import torch.nn as nn

from avjepa_b200.src.models.utils.pos_embs import (
    get_2d_sincos_pos_embed, get_2d_sincos_pos_embed_xy, get_3d_sincos_pos_embed)
from avjepa_b200.src.utils.tensors import trunc_normal_

AUDIO_MEL_BINS, AUDIO_FRAMES = 128, 192     # the spectrogram is always 128 x 192 in the reference


def video_sincos(embed_dim, input_size, patch_size, num_frames, tubelet_size, uniform_power):
    grid = input_size // patch_size
    if num_frames > 1:
        table = get_3d_sincos_pos_embed(embed_dim, grid, num_frames // tubelet_size, cls_token=False,
                                        uniform_power=uniform_power)
    else:
        table = get_2d_sincos_pos_embed(embed_dim, grid, cls_token=False)
    return torch.from_numpy(table).float().unsqueeze(0)


def audio_sincos(embed_dim, patch_size):
    table = get_2d_sincos_pos_embed_xy(embed_dim, AUDIO_MEL_BINS // patch_size, AUDIO_FRAMES // patch_size, cls_token=False)
    return torch.from_numpy(table).float().unsqueeze(0)


def init_linear_norm_conv(m, std, convs=True):
    """The reference's per-module ``_init_weights``."""
    if isinstance(m, nn.Linear):
        trunc_normal_(m.weight, std=std)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif convs and isinstance(m, (nn.Conv2d, nn.Conv3d)):
        trunc_normal_(m.weight, std=std)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)


def rescale_blocks(blocks):
    for layer_id, layer in enumerate(blocks, start=1):
        layer.attn.proj.weight.data.div_(math.sqrt(2.0 * layer_id))
        layer.mlp.fc2.weight.data.div_(math.sqrt(2.0 * layer_id))


def interpolate_video_pos(pos_embed, x, is_video, input_size, num_frames, patch_size, tubelet_size):
    """Resize the positional table when the input resolution differs from the construction-time
    one (trilinear for video, bicubic for images); identity otherwise.  Init-/eval-time only."""
    _, N, dim = pos_embed.shape
    if is_video:
        _, _, T, H, W = x.shape
        if H == input_size and W == input_size and T == num_frames:
            return pos_embed
        T, H, W = T // tubelet_size, H // patch_size, W // patch_size
        n_t, n_hw = num_frames // tubelet_size, input_size // patch_size
        assert n_hw * n_hw * n_t == N, 'Positional embedding initialized incorrectly'
        grid = pos_embed.reshape(1, n_t, n_hw, n_hw, dim).permute(0, 4, 1, 2, 3)
        grid = nn.functional.interpolate(grid, scale_factor=(T / n_t, H / n_hw, W / n_hw), mode='trilinear')
        return grid.permute(0, 2, 3, 4, 1).reshape(1, -1, dim)
    _, _, H, W = x.shape
    if H == input_size and W == input_size:
        return pos_embed
    npatch = (H // patch_size) * (W // patch_size)
    side = int(math.sqrt(N))
    grid = pos_embed.reshape(1, side, side, dim).permute(0, 3, 1, 2)
    grid = nn.functional.interpolate(grid, scale_factor=math.sqrt(npatch / N), mode='bicubic')
    return grid.permute(0, 2, 3, 1).reshape(1, -1, dim)
