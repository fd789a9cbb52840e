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
