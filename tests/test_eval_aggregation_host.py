"""Host logic of the frozen-eval aggregation wrappers (evals/video_classification_frozen/utils.py:22-157) against a
literal restatement, with a deterministic stand-in encoder.  The temporal-embedding branch calls apply_masks, which
only exists on CUDA in this package; here it is replaced by torch.gather to check the logic AROUND the kernel."""
import sys

import numpy as np
import pytest
import torch

from avjepa_b200.evals.video_classification_frozen import utils as agg


class StandInEncoder(torch.nn.Module):
    """[B, C, T, H, W] (or [B, C, H, W]) -> [B, N, D] with N = (T/2) * 4 (or 4): deterministic, shape faithful."""
    embed_dim, num_heads = 8, 2

    def forward(self, x):
        if x.dim() == 4:
            x = x.unsqueeze(2).repeat(1, 1, 2, 1, 1)
        B, C, T, H, W = x.shape
        t = x.reshape(B, C, T // 2, 2, 2, H // 2, 2, W // 2).mean(dim=(1, 3, 5, 7))          # [B, T/2, 2, 2]
        return t.reshape(B, -1, 1) * torch.arange(1, 9, dtype=x.dtype).reshape(1, 1, 8)


def _ref_clip_aggregation(model, x, tubelet, attend, table=None, clip_indices=None):
    """Literal restatement of ClipAggregation.forward (reference :115-157)."""
    num_clips, num_views = len(x), len(x[0])
    B, C, T, H, W = x[0][0].size()
    outputs = model(torch.cat([torch.cat(xi, dim=0) for xi in x], dim=0))
    _, N, D = outputs.size()
    T = T // tubelet
    N = N // T
    eff_B = B * num_views
    all_outputs = [[] for _ in range(num_views)]
    for i in range(num_clips):
        o = outputs[i * eff_B:(i + 1) * eff_B]
        for j in range(num_views):
            all_outputs[j].append(o[j * B:(j + 1) * B])
    if not attend:
        return all_outputs
    for i, outs in enumerate(all_outputs):
        outs = torch.cat([o.reshape(B, T, N, D) for o in outs], dim=1).flatten(1, 2)
        if table is not None and clip_indices is not None:
            clip_indices = [c[:, ::tubelet] for c in clip_indices]
            pe = table.repeat(B, 1, 1)
            pe = [torch.gather(pe, 1, m.unsqueeze(-1).repeat(1, 1, pe.size(-1))) for m in clip_indices]
            pe = torch.cat(pe, dim=1).unsqueeze(2).repeat(1, 1, N, 1).flatten(1, 2)
            outs = outs + pe
        all_outputs[i] = outs
    return all_outputs


def _clips(g, n_clips, n_views, B=2, T=4):
    return [[torch.randn(B, 3, T, 4, 4, generator=g) for _ in range(n_views)] for _ in range(n_clips)]


@pytest.mark.parametrize('n_clips,n_views,attend', [(1, 1, False), (3, 2, False), (2, 1, True), (3, 2, True)])
def test_clip_aggregation_matches_restatement(n_clips, n_views, attend):
    g = torch.Generator().manual_seed(5)
    x = _clips(g, n_clips, n_views)
    model = StandInEncoder()
    ours = agg.ClipAggregation(model, tubelet_size=2, attend_across_segments=attend)(x)
    ref = _ref_clip_aggregation(model, x, 2, attend)
    assert len(ours) == len(ref) == n_views
    for a, b in zip(ours, ref):
        if attend:
            assert torch.equal(a, b)
        else:
            assert len(a) == len(b) == n_clips and all(torch.equal(p, q) for p, q in zip(a, b))


def test_clip_aggregation_temporal_embedding(monkeypatch):
    g = torch.Generator().manual_seed(6)
    x = _clips(g, 2, 1, B=2, T=4)
    idx = [torch.tensor([[0, 1, 2, 3], [4, 5, 6, 7]]), torch.tensor([[8, 9, 10, 11], [2, 3, 4, 5]])]
    monkeypatch.setattr(agg, 'apply_masks',
                        lambda t, masks, concat=True: [torch.gather(t, 1, m.unsqueeze(-1).repeat(1, 1, t.size(-1))) for m in masks])
    model = StandInEncoder()
    wrap = agg.ClipAggregation(model, tubelet_size=2, max_frames=32, use_pos_embed=True, attend_across_segments=True)
    assert tuple(wrap.pos_embed.shape) == (1, 16, 8) and not wrap.pos_embed.requires_grad
    ours = wrap(x, clip_indices=idx)
    ref = _ref_clip_aggregation(model, x, 2, True, table=wrap.pos_embed.data, clip_indices=idx)
    assert torch.equal(ours[0], ref[0])
    # the table is the reference's 1-D sincos embedding (pos_embs.py:84-117)
    omega = 1.0 / 10000 ** (np.arange(4, dtype=float) / 4.0)
    want = np.concatenate([np.sin(np.outer(np.arange(16.0), omega)), np.cos(np.outer(np.arange(16.0), omega))], axis=1)
    assert np.array_equal(wrap.pos_embed[0].numpy(), want.astype(np.float32))


def test_frame_aggregation_shapes_and_values():
    g = torch.Generator().manual_seed(7)
    x = _clips(g, 2, 2, B=2, T=2)
    model = StandInEncoder()
    out = agg.FrameAggregation(model)(x)
    # restatement (reference :50-83): views -> batch, clips -> time, one encoder input per frame
    frames = torch.cat([torch.cat(xi, dim=0) for xi in x], dim=2)
    VB, C, T, H, W = frames.shape
    tok = model(frames.permute(0, 2, 1, 3, 4).reshape(VB * T, C, H, W))
    tok = tok.reshape(VB, T, tok.size(1), tok.size(2)).flatten(1, 2)
    assert len(out) == 2 and all(torch.equal(out[v], tok[v * 2:(v + 1) * 2]) for v in range(2))


def _live_reference():
    import os
    if not os.path.isdir('/root/reference'):
        return None
    sys.path.insert(0, '/root/reference')
    try:
        import evals.video_classification_frozen.utils as ref_utils
        return ref_utils
    except Exception:
        return None
    finally:
        sys.path.remove('/root/reference')


@pytest.mark.parametrize('n_clips,n_views,attend', [(2, 2, False), (3, 1, True), (2, 3, True)])
def test_against_live_reference_modules(n_clips, n_views, attend):
    """Only in the build container (the reference is not on the GPU box): the UNMODIFIED reference wrappers around the
    same stand-in encoder."""
    R = _live_reference()
    if R is None:
        pytest.skip('reference not mounted')
    g = torch.Generator().manual_seed(11)
    x = _clips(g, n_clips, n_views)
    model = StandInEncoder()
    ours = agg.ClipAggregation(model, tubelet_size=2, attend_across_segments=attend)(x)
    ref = R.ClipAggregation(model, tubelet_size=2, attend_across_segments=attend)(x)
    for a, b in zip(ours, ref):
        if attend:
            assert torch.equal(a, b)
        else:
            assert all(torch.equal(p, q) for p, q in zip(a, b))
    xf = _clips(g, n_clips, n_views, T=2)
    of, rf = agg.FrameAggregation(model)(xf), R.FrameAggregation(model)(xf)
    assert len(of) == len(rf) and all(torch.equal(p, q) for p, q in zip(of, rf))
