"""Multi-clip / multi-view batching around a frozen encoder (drop-in for the two aggregation wrappers of the
reference's ``evals/video_classification_frozen/utils.py``: ``FrameAggregation :22-83``, ``ClipAggregation :86-157``).

Both wrappers push every clip and every spatial view through the encoder in ONE batch and hand the tokens back per
view; with ``attend_across_segments`` the clips of a view are concatenated along time and (optionally) a 1-D temporal
sincos embedding, gathered at the sampled frame indices with ``apply_masks``, is added.  The augmentation helpers of
that reference file (``make_transforms`` etc.) belong to the data pipeline and are out of scope.

Reference quirk kept on purpose: ``ClipAggregation`` sub-samples ``clip_indices`` by the tubelet size INSIDE its loop over
views (``:146``), so a second view would see indices sub-sampled twice; like the reference this only works with one
view per clip when the temporal embedding is enabled.
"""
import torch
import torch.nn as nn

from avjepa_b200.src.masks.utils import apply_masks
from avjepa_b200.src.models.utils.pos_embs import get_1d_sincos_pos_embed


def _temporal_table(embed_dim, length):
    table = nn.Parameter(torch.zeros(1, length, embed_dim), requires_grad=False)
    table.copy_(torch.from_numpy(get_1d_sincos_pos_embed(embed_dim, length)).float().unsqueeze(0))
    return table


def _temporal_embedding(table, batch, indices, n_spatial):
    """[batch, sum_T * n_spatial, D]: rows of `table` at `indices` (list of [batch, T_i]), repeated per spatial token."""
    rows = torch.cat(apply_masks(table.repeat(batch, 1, 1), indices, concat=False), dim=1)      # [B, sum_T, D]
    return rows.unsqueeze(2).repeat(1, 1, n_spatial, 1).flatten(1, 2)


class FrameAggregation(nn.Module):
    """Every frame is an independent encoder input; all tokens of a view are concatenated along time."""

    def __init__(self, model, max_frames=10000, use_pos_embed=False, attend_across_segments=False):
        super().__init__()
        self.model = model
        self.embed_dim = model.embed_dim
        self.num_heads = model.num_heads
        self.attend_across_segments = attend_across_segments
        self.pos_embed = _temporal_table(model.embed_dim, max_frames) if use_pos_embed else None

    def forward(self, x, clip_indices=None):
        n_views = len(x[0])
        frames = torch.cat([torch.cat(views, dim=0) for views in x], dim=2)          # views -> batch, clips -> time
        VB, C, T, H, W = frames.size()
        tokens = self.model(frames.permute(0, 2, 1, 3, 4).reshape(VB * T, C, H, W))
        _, N, D = tokens.size()
        tokens = tokens.reshape(VB, T, N, D).flatten(1, 2)
        B = VB // n_views
        out = []
        for v in range(n_views):
            o = tokens[v * B:(v + 1) * B]
            if self.pos_embed is not None and clip_indices is not None:
                o += _temporal_embedding(self.pos_embed, B, clip_indices, N)
            out.append(o)
        return out


class ClipAggregation(nn.Module):
    """Every clip is an independent encoder input; returns tokens[view][clip], or one tensor per view with the clips
    concatenated along time when ``attend_across_segments`` is set."""

    def __init__(self, model, tubelet_size=2, max_frames=10000, use_pos_embed=False, attend_across_segments=False):
        super().__init__()
        self.model = model
        self.tubelet_size = tubelet_size
        self.embed_dim = model.embed_dim
        self.num_heads = model.num_heads
        self.attend_across_segments = attend_across_segments
        self.pos_embed = _temporal_table(model.embed_dim, max_frames // tubelet_size) if use_pos_embed else None

    def forward(self, x, clip_indices=None):
        n_clips, n_views = len(x), len(x[0])
        B, _, T, _, _ = x[0][0].size()
        tokens = self.model(torch.cat([torch.cat(views, dim=0) for views in x], dim=0))   # [clips * views * B, N, D]
        _, n_tok, D = tokens.size()
        T = T // self.tubelet_size                     # temporal tokens per clip
        n_spatial = n_tok // T
        per_view = [[tokens[(c * n_views + v) * B:(c * n_views + v + 1) * B] for c in range(n_clips)] for v in range(n_views)]
        if not self.attend_across_segments:
            return per_view
        for v, clips in enumerate(per_view):
            merged = torch.cat([o.reshape(B, T, n_spatial, D) for o in clips], dim=1).flatten(1, 2)
            if self.pos_embed is not None and clip_indices is not None:
                clip_indices = [c[:, ::self.tubelet_size] for c in clip_indices]     # (sic) see the module docstring
                merged += _temporal_embedding(self.pos_embed, B, clip_indices, n_spatial)
            per_view[v] = merged
        return per_view
