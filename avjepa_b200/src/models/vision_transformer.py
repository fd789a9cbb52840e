"""Video-only ViT encoder.

Drop-in for the reference's ``src/models/vision_transformer.py`` (``VisionTransformer :24-252``,
factories ``:255-300``, ``VIT_EMBED_DIMS :303-313``): ``forward(x, masks=None)`` with ``masks`` a
[B, K] int64 tensor of KEPT token indices (or a one-element list).  Same kernel schedule as the
audio-video encoder without the audio rows.
"""
from functools import partial

import torch
import torch.nn as nn

from avjepa_b200 import backbone
from avjepa_b200.src.models import _common
from avjepa_b200.src.models.utils.modules import Block
from avjepa_b200.src.models.utils.patch_embed import PatchEmbed, PatchEmbed3D


class VisionTransformer(nn.Module):
    """ Vision Transformer """

    def __init__(
        self,
        img_size=224,
        patch_size=16,
        num_frames=1,
        tubelet_size=2,
        in_chans=3,
        embed_dim=768,
        depth=12,
        num_heads=12,
        mlp_ratio=4.0,
        qkv_bias=True,
        qk_scale=None,
        drop_rate=0.0,
        attn_drop_rate=0.0,
        norm_layer=nn.LayerNorm,
        init_std=0.02,
        out_layers=None,
        uniform_power=False,
        **kwargs
    ):
        super().__init__()
        self.num_features = self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.out_layers = out_layers
        self.input_size = img_size
        self.patch_size = patch_size
        self.num_frames = num_frames
        self.tubelet_size = tubelet_size
        self.is_video = num_frames > 1
        grid_size = img_size // patch_size
        grid_depth = num_frames // tubelet_size

        if self.is_video:
            self.patch_embed = PatchEmbed3D(patch_size=patch_size, tubelet_size=tubelet_size,
                                            in_chans=in_chans, embed_dim=embed_dim)
            self.num_patches = grid_depth * grid_size * grid_size
        else:
            self.patch_embed = PatchEmbed(patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
            self.num_patches = grid_size * grid_size

        self.uniform_power = uniform_power
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, embed_dim), requires_grad=False)

        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, act_layer=nn.GELU, grid_size=grid_size, grid_depth=grid_depth,
                  attn_drop=attn_drop_rate, norm_layer=norm_layer)
            for _ in range(depth)])
        self.norm = norm_layer(embed_dim)

        self.pos_embed.data.copy_(_common.video_sincos(
            embed_dim, img_size, patch_size, num_frames, tubelet_size, uniform_power))
        self.init_std = init_std
        self.apply(self._init_weights)
        self._rescale_blocks()

    def _init_weights(self, m):
        _common.init_linear_norm_conv(m, self.init_std)

    def _rescale_blocks(self):
        _common.rescale_blocks(self.blocks)

    def get_num_layers(self):
        return len(self.blocks)

    def no_weight_decay(self):
        return {}

    def forward(self, x, masks=None):
        """
        :param x: image [B, C, H, W] or video clip [B, C, T, H, W]
        :param masks: indices of the patch tokens to KEEP, [B, K] int64 (or a one-element list)
        """
        return backbone.run_encoder(self, x, None, masks, None)

    def interpolate_pos_encoding(self, x, pos_embed):
        return _common.interpolate_video_pos(pos_embed, x, self.is_video, self.input_size, self.num_frames,
                                             self.patch_size, self.tubelet_size)


def _factory(embed_dim, depth, num_heads, mlp_ratio=4):
    def make(patch_size=16, **kwargs):
        return VisionTransformer(
            patch_size=patch_size, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio,
            qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    return make


vit_tiny = _factory(192, 12, 3)
vit_small = _factory(384, 12, 6)
vit_base = _factory(768, 12, 12)
vit_large = _factory(1024, 24, 16)
vit_huge = _factory(1280, 32, 16)
vit_giant = _factory(1408, 40, 16, mlp_ratio=48 / 11)


def vit_gigantic(patch_size=14, **kwargs):
    # `mpl_ratio` typo kept from the reference (falls into **kwargs -> mlp_ratio stays 4.0)
    return VisionTransformer(
        patch_size=patch_size, embed_dim=1664, depth=48, num_heads=16, mpl_ratio=64 / 13,
        qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


VIT_EMBED_DIMS = {
    'vit_tiny': 192,
    'vit_small': 384,
    'vit_base': 768,
    'vit_large': 1024,
    'vit_huge': 1280,
    'vit_giant': 1408,
    'vit_gigantic': 1664,
}
