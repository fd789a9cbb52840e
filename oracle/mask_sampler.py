"""CPU restatement of the device-side mask sampler (avjepa_b200/csrc/mask_collate.cu) -- TEST INFRASTRUCTURE ONLY, like the
rest of oracle/: nothing under avjepa_b200/ imports it.

It follows the reference's per-sample loop (``src/masks/avmultiblock3d.py:172-234``) the way the kernel does: block origins
from an explicit MT19937 replica of torch's CPU generator (``torch.randint(0, n, (1,))`` = next 32-bit word modulo n;
``at::mt19937``: 624 state words, `left`, `next`), draw order video (top, left, start) then audio (top, left) per block, a
sample whose video context comes out empty is drawn again, kept / dropped token ids ascending, and the batch minimum as
the common length.  Pinned against the host collator (itself bit-exact with the reference) by tests/test_mask_rng_host.py;
the CUDA kernel is pinned against the same collator on the device by tests/test_mask_collate_gpu.py.
"""
import numpy as np


class MT19937(object):
    """at::mt19937 on the 626-word layout of `avj_mask_collate` (624 state words, left, next)."""

    def __init__(self, words):
        w = np.asarray(words).view(np.uint32)
        self.s = [int(x) for x in w[:624]]
        self.left, self.next = int(w[624]), int(w[625])

    def _refill(self):
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
            self._refill()
        y = self.s[self.next]
        self.next += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9d2c5680
        y ^= (y << 15) & 0xefc60000
        y ^= y >> 18
        return y & 0xffffffff

    def randint(self, n):
        return self.rand32() % n

    def words(self):
        return np.array(self.s + [self.left, self.next], dtype=np.uint32).view(np.int32)


def sample(rng, gens, B, duration, height, width, a_height, a_width, a_block=(4, 6)):
    """gens: list of (t, h, w, npred, max_ctx).  Returns per generator (enc_v, enc_a, pred_v, pred_a) as int64 arrays [B, K]
    (K = batch minimum) plus a status word (bit 0: some index set had exactly one element)."""
    out, status = [], 0
    for (t, h, w, npred, max_ctx) in gens:
        rows = [[], [], [], []]
        while len(rows[0]) < B:
            keep_v = np.ones((duration, height, width), dtype=bool)
            keep_a = np.ones((a_height, a_width), dtype=bool)
            if max_ctx < duration:
                keep_v[max_ctx:] = False
            for _ in range(npred):
                top, left, start = rng.randint(height - h + 1), rng.randint(width - w + 1), rng.randint(duration - t + 1)
                keep_v[start:start + t, top:top + h, left:left + w] = False
                atop, aleft = rng.randint(a_height - a_block[0] + 1), rng.randint(a_width - a_block[1] + 1)
                keep_a[atop:atop + a_block[0], aleft:aleft + a_block[1]] = False
            kv, dv = np.flatnonzero(keep_v.ravel()), np.flatnonzero(~keep_v.ravel())
            ka, da = np.flatnonzero(keep_a.ravel()), np.flatnonzero(~keep_a.ravel())
            if kv.size == 0:
                continue
            if 1 in (kv.size, ka.size, dv.size, da.size):
                status |= 1
            for lst, v in zip(rows, (kv, ka, dv, da)):
                lst.append(v.astype(np.int64))
        out.append(tuple(np.stack([r[:min(x.size for x in lst)] for r in lst]) for lst in rows))
    return out, status
