"""Transformer building blocks: parameter containers with the reference's names and shapes.

Mirror of ``src/models/utils/modules.py`` (``MLP :13-36``, ``Attention :39-78``,
``Block :81-120``, ``CrossAttention :123-159``, ``CrossAttentionBlock :162-183``).  The sub-modules are ordinary ``nn.Linear`` / ``nn.LayerNorm`` holders so
state-dict keys, parameter order (and therefore same-seed initialisation and the optimizer's
name filters) are identical to the reference; the arithmetic does not go through their
``forward`` -- a backbone runs all of its blocks as one explicit kernel schedule
(:class:`avjepa_b200.engine.StackRun`).  ``Block.forward`` is provided for callers that use a
block on its own (probes, debugging): it is a depth-1 stack.
"""
import torch
import torch.nn as nn


class MLP(nn.Module):

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError('only exact-erf nn.GELU is implemented (the reference never uses another)')
        if drop != 0.:
            raise NotImplementedError('dropout > 0 is not implemented (every shipped config uses 0)')
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class Attention(nn.Module):

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., use_sdpa=True):
        super().__init__()
        if attn_drop != 0. or proj_drop != 0.:
            raise NotImplementedError('dropout > 0 is not implemented (every shipped config uses 0)')
        self.num_heads = num_heads
        head_dim = dim // num_heads
        # NB: like the reference's SDPA branch, the kernel always scales by head_dim**-0.5
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop_prob = proj_drop
        self.proj_drop = nn.Dropout(proj_drop)
        self.use_sdpa = use_sdpa


class Block(nn.Module):

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, grid_size=None, grid_depth=None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                              attn_drop=attn_drop, proj_drop=drop)
        self.norm2 = norm_layer(dim)
        self.mlp = MLP(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, return_attention=False, mask=None):
        """x: [B, N, D] CUDA tensor.  ``mask`` is accepted and ignored, like the reference."""
        if return_attention:
            raise NotImplementedError('return_attention is not available from the fused attention kernel')
        from avjepa_b200 import backbone
        return backbone.run_blocks(self, [self], None, x)


class CrossAttention(nn.Module):
    """Parameter container of the reference's cross-attention (``q``, ``kv``, ``proj`` Linears); used on its own it is the
    ``complete_block=False`` form of the attentive pooler."""

    def __init__(self, dim, num_heads=12, qkv_bias=False, use_sdpa=True):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, int(dim * 2), bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_sdpa = use_sdpa

    def forward(self, q, x):
        from avjepa_b200 import pooler
        return pooler.run_cross_attention(self, None, q, x)


class CrossAttentionBlock(nn.Module):

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.xattn = CrossAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = MLP(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer)

    def forward(self, q, x):
        """q: [B, n, D] queries, x: [B, N, D] encoder tokens -> [B, n, D]:
        ``q + xattn(q, norm1(x))`` then ``+ mlp(norm2(.))`` (``modules.py:179-183``)."""
        from avjepa_b200 import pooler
        return pooler.run_cross_attention(self.xattn, self, q, x)
