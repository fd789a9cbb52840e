"""Mask collators: bit-exact against fixtures generated from the unmodified reference
(oracle/make_golden.py), for both the product collators and the oracle's restatement."""
import numpy as np
import pytest
import torch

from helpers import MASK_CFG, golden


def fake_batch(b):
    return [([torch.zeros(1)], 0, [0], torch.zeros(1)) for _ in range(b)]


GRID = [(s, b) for s in (0, 234) for b in (1, 2, 8)]


@pytest.mark.parametrize('seed,bsz', GRID)
def test_av_collator_bit_exact(seed, bsz):
    from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator
    g = golden('masks_av.npz')
    torch.manual_seed(seed)
    coll = AVMaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
    for call in range(3):
        if f's{seed}_b{bsz}_c{call}_crash' in g.files:
            with pytest.raises(TypeError):
                coll(fake_batch(bsz))
            continue
        batch, ev, ea, pv, pa = coll(fake_batch(bsz))
        assert len(batch) == 4
        for gi in range(2):
            for nm, t in (('ev', ev), ('ea', ea), ('pv', pv), ('pa', pa)):
                ref = g[f's{seed}_b{bsz}_c{call}_g{gi}_{nm}'].astype(np.int64)
                assert t[gi].dtype == torch.int64
                assert t[gi].shape == ref.shape
                assert np.array_equal(t[gi].numpy(), ref), (seed, bsz, call, gi, nm)


@pytest.mark.parametrize('seed,bsz', GRID)
def test_video_collator_bit_exact(seed, bsz):
    from avjepa_b200.src.masks.multiblock3d import MaskCollator
    g = golden('masks_video.npz')
    torch.manual_seed(seed)
    coll = MaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
    for call in range(3):
        _, e, p = coll(fake_batch(bsz))
        for gi in range(2):
            assert np.array_equal(e[gi].numpy(), g[f's{seed}_b{bsz}_c{call}_g{gi}_e'].astype(np.int64))
            assert np.array_equal(p[gi].numpy(), g[f's{seed}_b{bsz}_c{call}_g{gi}_p'].astype(np.int64))


@pytest.mark.parametrize('seed,bsz', GRID)
def test_oracle_sampler_bit_exact(seed, bsz):
    from oracle import avjepa_oracle as O
    g = golden('masks_av.npz')
    torch.manual_seed(seed)
    samplers = [O.MaskSampler(c) for c in O.VITL16_MASK_CFG]
    for call in range(3):
        if f's{seed}_b{bsz}_c{call}_crash' in g.files:
            with pytest.raises(TypeError):
                O.sample_av_masks(samplers, bsz)
            continue
        ev, ea, pv, pa = O.sample_av_masks(samplers, bsz)
        for gi in range(2):
            for nm, t in (('ev', ev), ('ea', ea), ('pv', pv), ('pa', pa)):
                assert np.array_equal(t[gi].numpy(), g[f's{seed}_b{bsz}_c{call}_g{gi}_{nm}'].astype(np.int64))


def test_known_shapes_from_survey():
    """SURVEY.md section 8c smoke values: seed 0, fresh collator, B=2, generator 0 first call."""
    from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator
    from avjepa_b200.src.masks.multiblock3d import MaskCollator
    torch.manual_seed(0)
    _, ev, ea, pv, pa = AVMaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16,
                                       tubelet_size=2)(fake_batch(2))
    assert (tuple(ev[0].shape), tuple(ea[0].shape), tuple(pv[0].shape), tuple(pa[0].shape)) == \
        ((2, 504), (2, 19), (2, 832), (2, 71))
    torch.manual_seed(0)
    _, e, p = MaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)(fake_batch(2))
    assert (tuple(e[0].shape), tuple(p[0].shape)) == ((2, 416), (2, 1136))


def test_masks_partition_the_grid():
    """Before truncation enc and pred are complements; after it they stay disjoint and ascending."""
    from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator
    torch.manual_seed(7)
    coll = AVMaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
    _, ev, ea, pv, pa = coll(fake_batch(4))
    for e, p, n in ((ev, pv, 1568), (ea, pa, 96)):
        for gi in range(2):
            for b in range(4):
                es, ps = set(e[gi][b].tolist()), set(p[gi][b].tolist())
                assert not (es & ps)
                assert max(es | ps) < n
                assert e[gi][b].tolist() == sorted(es) and p[gi][b].tolist() == sorted(ps)


def test_collator_is_picklable_and_steps():
    import pickle
    from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator
    coll = AVMaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
    coll.step()
    assert coll.mask_generators[0]._itr_counter._v.value == 0
    # multiprocessing.Value only pickles through process inheritance (same as the reference);
    # the collator's own attributes are plain data:
    assert all(isinstance(g.npred, int) for g in coll.mask_generators)
