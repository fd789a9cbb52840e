"""Video-only JEPA predictor.

Drop-in for the reference's ``src/models/predictor.py`` (``VisionTransformerPredictor :24-239``,
``vit_predictor :242-246``): ``forward(ctxt, tgt, masks_ctxt, masks_tgt, mask_index=1)``.
Sequence layout ``[ctx | tgt]``; same fused token assembly as the audio-video predictor.
"""
from functools import partial

import torch
import torch.nn as nn

from avjepa_b200 import backbone
from avjepa_b200.src.models import _common
from avjepa_b200.src.models.utils.modules import Block
from avjepa_b200.src.utils.tensors import trunc_normal_


class VisionTransformerPredictor(nn.Module):
    """ Vision Transformer predictor """

    def __init__(
        self,
        img_size=224,
        patch_size=16,
        num_frames=1,
        tubelet_size=2,
        embed_dim=768,
        predictor_embed_dim=384,
        depth=6,
        num_heads=12,
        mlp_ratio=4.0,
        qkv_bias=True,
        qk_scale=None,
        drop_rate=0.0,
        attn_drop_rate=0.0,
        norm_layer=nn.LayerNorm,
        init_std=0.02,
        uniform_power=False,
        use_mask_tokens=False,
        num_mask_tokens=2,
        zero_init_mask_tokens=True,
        **kwargs
    ):
        super().__init__()
        self.predictor_embed = nn.Linear(embed_dim, predictor_embed_dim, bias=True)

        self.mask_tokens = None
        self.num_mask_tokens = 0
        if use_mask_tokens:
            self.num_mask_tokens = num_mask_tokens
            self.mask_tokens = nn.ParameterList([
                nn.Parameter(torch.zeros(1, 1, predictor_embed_dim)) for _ in range(num_mask_tokens)])

        self.input_size = img_size
        self.patch_size = patch_size
        self.num_frames = num_frames
        self.tubelet_size = tubelet_size
        self.is_video = num_frames > 1
        grid_size = img_size // patch_size
        grid_depth = num_frames // tubelet_size
        self.num_patches = (grid_depth if self.is_video else 1) * grid_size * grid_size

        self.uniform_power = uniform_power
        self.predictor_pos_embed = nn.Parameter(
            torch.zeros(1, self.num_patches, predictor_embed_dim), requires_grad=False)

        self.predictor_blocks = nn.ModuleList([
            Block(dim=predictor_embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                  qk_scale=qk_scale, drop=drop_rate, act_layer=nn.GELU, attn_drop=attn_drop_rate,
                  grid_size=grid_size, grid_depth=grid_depth, norm_layer=norm_layer)
            for _ in range(depth)])
        self.predictor_norm = norm_layer(predictor_embed_dim)
        self.predictor_proj = nn.Linear(predictor_embed_dim, embed_dim, bias=True)

        self.predictor_pos_embed.data.copy_(_common.video_sincos(
            predictor_embed_dim, img_size, patch_size, num_frames, tubelet_size, uniform_power))
        self.init_std = init_std
        if not zero_init_mask_tokens:
            for mt in self.mask_tokens:
                trunc_normal_(mt, std=init_std)
        self.apply(self._init_weights)
        self._rescale_blocks()

    def _init_weights(self, m):
        _common.init_linear_norm_conv(m, self.init_std, convs=False)

    def _rescale_blocks(self):
        _common.rescale_blocks(self.predictor_blocks)

    def _parts(self):
        return (self.predictor_embed, None, self.mask_tokens, None, self.predictor_pos_embed, None)

    def forward(self, ctxt, tgt, masks_ctxt, masks_tgt, mask_index=1):
        """
        :param ctxt: context tokens from the encoder, [B, Kc, D]
        :param tgt: target tokens -- unused with mask tokens
        :param masks_ctxt: indices of the context tokens in the full grid, [B, Kc]
        :param masks_tgt: indices of the target tokens in the full grid, [B, Kt]
        """
        assert (masks_ctxt is not None) and (masks_tgt is not None), 'Cannot run predictor without mask indices'
        return backbone.run_predictor(self, self._parts(), mask_index, ctxt, None, masks_ctxt, None, masks_tgt, None)


def vit_predictor(**kwargs):
    return VisionTransformerPredictor(
        mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
