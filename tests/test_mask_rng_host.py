"""Host half of the device-side mask sampler (row f4): the conversion between torch's CPU generator state and the
626-word layout `avj_mask_collate` works on, and the generator model itself (MT19937, `torch.randint` = next 32-bit
word modulo the range) restated in numpy and checked against torch's own draws.  The CUDA kernel implements exactly
this model; tests/test_mask_collate_gpu.py checks it bit for bit on the device."""
import numpy as np
import torch

from avjepa_b200.src.masks.device_collator import host_rng_to_words, words_to_host_rng


class _MT(object):
    """at::mt19937 as the kernel implements it (mask_collate.cu: mc_next_state / mc_rand32)."""

    def __init__(self, words):
        w = words.numpy().view(np.uint32)
        self.s = [int(x) for x in w[:624]]
        self.left, self.next = int(w[624]), int(w[625])

    def _next_state(self):
        s, n, m = self.s, 624, 397

        def tw(u, v):
            return (((u & 0x80000000) | (v & 0x7fffffff)) >> 1) ^ (0x9908b0df if (v & 1) else 0)
        for j in range(n - m):
            s[j] = s[j + m] ^ tw(s[j], s[j + 1])
        for j in range(n - m, n - 1):
            s[j] = s[j + m - n] ^ tw(s[j], s[j + 1])
        s[n - 1] = s[m - 1] ^ tw(s[n - 1], s[0])
        self.left, self.next = 624, 0

    def rand32(self):
        self.left -= 1
        if self.left == 0:
            self._next_state()
        y = self.s[self.next]
        self.next += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9d2c5680
        y ^= (y << 15) & 0xefc60000
        y ^= y >> 18
        return y & 0xffffffff

    def words(self):
        out = np.array(self.s + [self.left, self.next], dtype=np.uint32)
        return torch.from_numpy(out.view(np.int32).copy())


def test_state_round_trip():
    torch.manual_seed(77)
    for _ in range(5):
        torch.randint(0, 9, (1,))
    st = torch.get_rng_state()
    assert torch.equal(words_to_host_rng(host_rng_to_words(st)), st)


def test_generator_model_reproduces_torch_randint_across_a_state_refill():
    torch.manual_seed(234)
    mt = _MT(host_rng_to_words())
    ranges = [14 - h + 1 for h in (1, 5, 9, 14)] + [8, 3, 7, 1]
    mine, ref = [], []
    for i in range(1400):                              # > 624 draws: crosses two regenerations of the state block
        n = ranges[i % len(ranges)]
        mine.append(mt.rand32() % n)
        ref.append(int(torch.randint(0, n, (1,))))
    assert mine == ref
    # and the model's final state is torch's final state
    assert torch.equal(words_to_host_rng(mt.words()), torch.get_rng_state())


def test_sampler_restatement_matches_the_host_collator():
    """oracle/mask_sampler.py (the algorithm the CUDA kernel implements) against the host AVMaskCollator: same generator state
    in -> same masks out and same generator state afterwards, over several seeds, batch sizes and calls."""
    import pytest
    from oracle import mask_sampler as MS
    from avjepa_b200.src.masks import _blocks
    from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator
    cfgs = {
        'vitl16': [dict(aspect_ratio=(0.75, 1.5), num_blocks=8, spatial_scale=(0.15, 0.15), temporal_scale=(1.0, 1.0), max_temporal_keep=1.0),
                   dict(aspect_ratio=(0.75, 1.5), num_blocks=2, spatial_scale=(0.7, 0.7), temporal_scale=(1.0, 1.0), max_temporal_keep=1.0)],
        'short': [dict(aspect_ratio=(0.3, 3.0), num_blocks=3, spatial_scale=(0.2, 0.6), temporal_scale=(0.4, 0.9), max_temporal_keep=0.5)],
    }
    for name, seed, B in (('vitl16', 234, 6), ('vitl16', 1, 2), ('short', 3, 4), ('short', 11, 1)):
        torch.manual_seed(seed)
        host = AVMaskCollator(cfgs[name], crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
        want = []
        try:
            for _ in range(3):
                want.append(host([torch.zeros(1)] * B)[1:])
        except TypeError:
            continue                                   # the reference's one-element quirk: nothing to compare for this draw
        want_state = torch.get_rng_state()
        torch.manual_seed(seed)
        rng = MS.MT19937(host_rng_to_words().numpy())
        twin = AVMaskCollator(cfgs[name], crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)   # fresh step counters
        for call in range(3):
            gens = []
            for g in twin.mask_generators:
                t, h, w = _blocks.draw_block_size(g.step(), g.duration, g.height, g.width, g.temporal_pred_mask_scale,
                                                  g.spatial_pred_mask_scale, g.aspect_ratio)
                gens.append((t, h, w, g.npred, g.max_context_duration))
            g0 = twin.mask_generators[0]
            got, status = MS.sample(rng, gens, B, g0.duration, g0.height, g0.width, g0.a_height, g0.a_width)
            assert status == 0
            for gi, per_gen in enumerate(got):
                for kind in range(4):                  # enc_v, enc_a, pred_v, pred_a
                    assert torch.equal(torch.from_numpy(per_gen[kind]), want[call][kind][gi]), (name, seed, call, gi, kind)
        assert torch.equal(words_to_host_rng(torch.from_numpy(rng.words().copy())), want_state)
