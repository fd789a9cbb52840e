"""Audio-video JEPA predictor.

Drop-in for the reference's ``src/models/audiovisionpredictor.py``
(``AudioVisionTransformerPredictor :18-301``, ``vit_avpredictor :304-308``):
``forward(ctxt, tgt, masks_ctxt, masks_tgt, mask_index=1)`` with 2-tuples (video, audio) for
each argument.  The token assembly of the reference (two Linears, four ``.repeat`` of the
positional tables, four gathers, two mask-token repeats, five ``torch.cat``) is replaced by two
GEMMs whose epilogue adds the gathered positional rows and writes straight into the
``[ctx_v | tgt_v | ctx_a | tgt_a]`` layout, plus one mask-token fill kernel per modality.
``tgt`` is unused when mask tokens are enabled, exactly as in the reference.
"""
from functools import partial

import torch
import torch.nn as nn

from avjepa_b200 import backbone
from avjepa_b200.src.models import _common
from avjepa_b200.src.models.utils.modules import Block
from avjepa_b200.src.utils.tensors import trunc_normal_


class AudioVisionTransformerPredictor(nn.Module):
    """ Audio-video predictor """

    def __init__(
        self,
        img_size=224,
        a_size=(128, 192),
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
        self.predictor_embed_v = nn.Linear(embed_dim, predictor_embed_dim, bias=True)
        self.predictor_embed_a = nn.Linear(embed_dim, predictor_embed_dim, bias=True)

        self.mask_tokens_v = None
        self.mask_tokens_a = None
        self.num_mask_tokens = 0
        if use_mask_tokens:
            self.num_mask_tokens = num_mask_tokens
            self.mask_tokens_v = nn.ParameterList([
                nn.Parameter(torch.zeros(1, 1, predictor_embed_dim)) for _ in range(num_mask_tokens)])
            self.mask_tokens_a = nn.ParameterList([
                nn.Parameter(torch.zeros(1, 1, predictor_embed_dim)) for _ in range(num_mask_tokens)])

        self.input_size = img_size
        self.patch_size = patch_size
        self.num_frames = num_frames
        self.tubelet_size = tubelet_size
        self.is_video = num_frames > 1
        grid_size = img_size // patch_size
        grid_depth = num_frames // tubelet_size
        self.a_height = a_size[0] // patch_size
        self.a_width = a_size[1] // patch_size
        self.num_patches = grid_depth * grid_size * grid_size
        self.num_patches_a = self.a_height * self.a_width

        self.uniform_power = uniform_power
        self.predictor_pos_embed_v = nn.Parameter(
            torch.zeros(1, self.num_patches, predictor_embed_dim), requires_grad=False)
        self.predictor_pos_embed_a = nn.Parameter(
            torch.zeros(1, self.num_patches_a, predictor_embed_dim), requires_grad=False)

        self.predictor_blocks = nn.ModuleList([
            Block(dim=predictor_embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                  qk_scale=qk_scale, drop=drop_rate, act_layer=nn.GELU, attn_drop=attn_drop_rate,
                  grid_size=grid_size, grid_depth=grid_depth, norm_layer=norm_layer)
            for _ in range(depth)])
        self.predictor_norm = norm_layer(predictor_embed_dim)
        self.predictor_proj = nn.Linear(predictor_embed_dim, embed_dim, bias=True)

        self.predictor_pos_embed_v.data.copy_(_common.video_sincos(
            predictor_embed_dim, img_size, patch_size, num_frames, tubelet_size, uniform_power))
        self.predictor_pos_embed_a.data.copy_(_common.audio_sincos(predictor_embed_dim, patch_size))
        self.init_std = init_std
        if not zero_init_mask_tokens:
            for mt in self.mask_tokens_v:
                trunc_normal_(mt, std=init_std)
            for mt in self.mask_tokens_a:
                trunc_normal_(mt, std=init_std)
        self.apply(self._init_weights)
        self._rescale_blocks()

    def _init_weights(self, m):
        _common.init_linear_norm_conv(m, self.init_std, convs=False)

    def _rescale_blocks(self):
        _common.rescale_blocks(self.predictor_blocks)

    def _parts(self):
        return (self.predictor_embed_v, self.predictor_embed_a, self.mask_tokens_v, self.mask_tokens_a,
                self.predictor_pos_embed_v, self.predictor_pos_embed_a)

    def forward(self, ctxt, tgt, masks_ctxt, masks_tgt, mask_index=1):
        """
        :param ctxt: (video, audio) context tokens from the encoder, [B, Kc_v, D], [B, Kc_a, D]
        :param tgt: (video, audio) target tokens -- unused with mask tokens
        :param masks_ctxt: (video, audio) indices of the context tokens in the full grid
        :param masks_tgt: (video, audio) indices of the target tokens in the full grid
        """
        assert (masks_ctxt is not None) and (masks_tgt is not None), 'Cannot run predictor without mask indices'
        return backbone.run_predictor(self, self._parts(), mask_index, ctxt[0], ctxt[1],
                                      masks_ctxt[0], masks_ctxt[1], masks_tgt[0], masks_tgt[1])


def vit_avpredictor(**kwargs):
    return AudioVisionTransformerPredictor(
        mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
