"""Audio-video multiblock mask sampling ON THE DEVICE.

``DeviceAVMaskCollator`` has the constructor and call signature of ``AVMaskCollator`` (reference
``src/masks/avmultiblock3d.py:20-67``) but returns the four mask lists as CUDA tensors produced by one kernel launch
(``avj_mask_collate``): block positions are drawn on the GPU from a bit-exact replica of torch's CPU generator (MT19937 --
the reference draws them from the GLOBAL generator with ``torch.randint``), keep / drop index sets are compacted there,
and only ``4 * generators * B`` int32 counts come back to the host to size the ``[B, K]`` views (K = batch minimum, like
``:225-233``).  With the same generator state the masks are bit-identical to the host collator's.

The generator state is taken from ``torch.get_rng_state()`` at the first call and then lives on the device; with
``sync_host_rng=True`` it is written back into torch's global generator after every call, so host and device collators can
be interleaved freely (used by the tests).  The per-call block SIZE (three seeded ``torch.rand`` draws from a fresh generator,
``:105-129``) stays on the host: it is three numbers per mask generator and does not touch the global stream.

One deviation: when an index set has exactly one element the reference (and the host collator in strict mode) raises
``TypeError`` in the middle of the batch; here the whole batch is sampled first and the same ``TypeError`` is raised afterwards.
"""
import numpy as np
import torch

from avjepa_b200 import _cabi, engine
from avjepa_b200.src.masks import _blocks
from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator

_STATE_WORDS = 626            # 624 state words + left + next (at::mt19937)


def host_rng_to_words(state=None):
    """torch's CPU generator state (``torch.get_rng_state()``: legacy POD {seed u64, left i32, seeded i32, next u64,
    state u64[624], ...}) -> int32[626] = 624 state words, left, next."""
    st = (torch.get_rng_state() if state is None else state).numpy()
    left = int(st[8:12].view(np.int32)[0])
    nxt = int(st[16:24].view(np.uint64)[0])
    words = st[24:24 + 624 * 8].view(np.uint64).astype(np.uint32)
    out = np.empty(_STATE_WORDS, dtype=np.uint32)
    out[:624] = words
    out[624] = np.uint32(left)
    out[625] = np.uint32(nxt)
    return torch.from_numpy(out.view(np.int32).copy())


def words_to_host_rng(words):
    """Inverse of host_rng_to_words: writes the 624 words / left / next into a copy of torch's current CPU generator state."""
    w = words.cpu().numpy().view(np.uint32)
    st = torch.get_rng_state().numpy().copy()
    st[8:12] = np.array([int(w[624])], dtype=np.int32).view(np.uint8)
    st[16:24] = np.array([int(w[625])], dtype=np.uint64).view(np.uint8)
    st[24:24 + 624 * 8] = w[:624].astype(np.uint64).view(np.uint8)
    return torch.from_numpy(st)


class DeviceAVMaskCollator(object):

    def __init__(self, cfgs_mask, crop_size=(224, 224), num_frames=16, patch_size=(16, 16), tubelet_size=2, device='cuda',
                 sync_host_rng=False, prefetch=False):
        self._host = AVMaskCollator(cfgs_mask, crop_size=crop_size, num_frames=num_frames, patch_size=patch_size,
                                    tubelet_size=tubelet_size)
        self.mask_generators = self._host.mask_generators      # block-size draws and step counters are shared with the host class
        self.device = torch.device(device)
        self.sync_host_rng = sync_host_rng
        # prefetch: the masks of the NEXT call are sampled on a side stream as soon as this call returns, so the one
        # device->host read (the counts) never waits behind the training step's kernels.  The look-ahead assumes the next
        # call asks for the same batch size; if it does not, the prefetched draw is discarded (its random numbers are
        # consumed), so the mask sequence then differs from the host collator's from that call on.
        self.prefetch = prefetch
        if prefetch and sync_host_rng:
            raise ValueError('prefetch keeps the device generator one call ahead; it cannot be mirrored into the host generator')
        self._stream = None
        self._pending = None
        self._rng = None
        g0 = self.mask_generators[0]
        self._grid = (g0.duration, g0.height, g0.width, g0.a_height, g0.a_width)
        for g in self.mask_generators:
            if (g.duration, g.height, g.width, g.a_height, g.a_width) != self._grid:
                raise ValueError('all mask generators must share one token grid')

    def step(self):
        self._host.step()

    def load_host_rng(self):
        """(Re)load the device generator from torch's global CPU generator."""
        self._rng = host_rng_to_words().to(self.device)

    def store_host_rng(self):
        """Write the device generator back into torch's global CPU generator (synchronises)."""
        torch.set_rng_state(words_to_host_rng(self._rng))

    def sample(self, batch_size):
        """-> (masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a): lists (one entry per mask generator) of int64 CUDA
        tensors [B, K]."""
        if not self.prefetch:
            return self._finish(self._launch(batch_size))
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=self.device)
        pend = self._pending if (self._pending is not None and self._pending[0] == int(batch_size)) else None
        if pend is None:
            with torch.cuda.stream(self._stream):
                pend = self._launch(batch_size)
        out = self._finish(pend)
        torch.cuda.current_stream(self.device).wait_stream(self._stream)
        for lst in out:
            for t in lst:
                t.record_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._stream):
            self._pending = self._launch(batch_size)
        return out

    def _launch(self, batch_size):
        if self._rng is None:
            self.load_host_rng()
        gens = self.mask_generators
        n_gen, B = len(gens), int(batch_size)
        D, H, W, AH, AW = self._grid
        NV, NA = D * H * W, AH * AW
        params = np.empty((n_gen, 5), dtype=np.int32)
        for i, g in enumerate(gens):
            t, h, w = _blocks.draw_block_size(g.step(), g.duration, g.height, g.width, g.temporal_pred_mask_scale,
                                              g.spatial_pred_mask_scale, g.aspect_ratio)
            params[i] = (t, h, w, g.npred, g.max_context_duration)
        dev = self.device
        enc_v = torch.empty((n_gen, B, NV), dtype=torch.int64, device=dev)
        pred_v = torch.empty((n_gen, B, NV), dtype=torch.int64, device=dev)
        enc_a = torch.empty((n_gen, B, NA), dtype=torch.int64, device=dev)
        pred_a = torch.empty((n_gen, B, NA), dtype=torch.int64, device=dev)
        counts = torch.empty((n_gen, B, 4), dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _cabi.call('avj_mask_collate', self._rng.data_ptr(), params.ctypes.data, n_gen, B, D, H, W, AH, AW,
                       gens[0].AUDIO_BLOCK[0], gens[0].AUDIO_BLOCK[1], enc_v.data_ptr(), pred_v.data_ptr(), enc_a.data_ptr(),
                       pred_a.data_ptr(), counts.data_ptr(), status.data_ptr(), engine.stream())
        pinned = torch.empty(n_gen * B * 4 + 1, dtype=torch.int32).pin_memory()
        pinned.copy_(torch.cat([counts.flatten(), status]), non_blocking=True)   # the one device->host read: sizes of the views
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        return int(B), n_gen, pinned, ev, (enc_v, enc_a, pred_v, pred_a)

    def _finish(self, pend):
        B, n_gen, host, ev, (enc_v, enc_a, pred_v, pred_a) = pend
        gens = self.mask_generators
        ev.synchronize()
        st = int(host[-1])
        if self.sync_host_rng:
            self.store_host_rng()
        if st & 2:
            raise RuntimeError('device mask sampling: no non-empty video context found (block size covers the whole grid)')
        strict = all(g.strict for g in gens)
        if (st & 1) and strict:
            raise TypeError('len() of a 0-d tensor')              # the reference's one-element quirk (see _blocks.strict_len)
        k = host[:-1].view(n_gen, B, 4).min(dim=1).values          # batch minimum per generator and index set
        out = ([], [], [], [])
        # compact [B, K] copies, on the stream that produced the rows
        with (torch.cuda.stream(self._stream) if self.prefetch else torch.cuda.device(self.device)):
            for i in range(n_gen):
                for j, full in enumerate((enc_v, enc_a, pred_v, pred_a)):
                    out[j].append(full[i, :, :int(k[i, j])].contiguous())
        return out

    def __call__(self, batch):
        collated_batch = torch.utils.data.default_collate(batch)
        ev, ea, pv, pa = self.sample(len(batch))
        return collated_batch, ev, ea, pv, pa
