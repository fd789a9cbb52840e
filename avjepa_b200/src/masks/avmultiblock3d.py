"""Audio-video multiblock-3D mask collator (host, DataLoader workers).

Drop-in for the reference's ``src/masks/avmultiblock3d.py``: ``AVMaskCollator(...)(batch)``
returns the 5-tuple ``(collated_batch, masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a)``
(``:20-67``) that ``app/avjepa/train.py:389`` unpacks.  One ``_AVMaskGenerator`` per
``cfgs_mask`` entry (``:70-234``): the video mask is the multiblock3d mask, the audio mask
punches one 4x6 block per video block out of the 8x12 mel-patch grid.  RNG draw order per
block is video (top, left, start) then audio (top, left); ``max_keep`` is accepted and
ignored, exactly like the reference.  Outputs are bit-exact with the reference.
"""
import torch

from avjepa_b200.src.masks import _blocks

_GLOBAL_SEED = 0


class AVMaskCollator(object):

    def __init__(self, cfgs_mask, crop_size=(224, 224), num_frames=16, patch_size=(16, 16), tubelet_size=2):
        super(AVMaskCollator, self).__init__()
        self.mask_generators = [
            _AVMaskGenerator(
                crop_size=crop_size,
                num_frames=num_frames,
                spatial_patch_size=patch_size,
                temporal_patch_size=tubelet_size,
                spatial_pred_mask_scale=m.get('spatial_scale'),
                temporal_pred_mask_scale=m.get('temporal_scale'),
                aspect_ratio=m.get('aspect_ratio'),
                npred=m.get('num_blocks'),
                max_context_frames_ratio=m.get('max_temporal_keep', 1.0),
                max_keep=m.get('max_keep', None),
            ) for m in cfgs_mask]

    def step(self):
        for g in self.mask_generators:
            g.step()

    def __call__(self, batch):
        collated_batch = torch.utils.data.default_collate(batch)
        enc_v, enc_a, pred_v, pred_a = [], [], [], []
        for g in self.mask_generators:
            ev, ea, pv, pa = g(len(batch))
            enc_v.append(ev)
            enc_a.append(ea)
            pred_v.append(pv)
            pred_a.append(pa)
        return collated_batch, enc_v, enc_a, pred_v, pred_a


class _AVMaskGenerator(object):

    AUDIO_BLOCK = (4, 6)   # fixed audio block (mel-patch rows, time-patch cols)

    def __init__(
        self,
        crop_size=(224, 224),
        a_size=(128, 192),
        num_frames=16,
        spatial_patch_size=(16, 16),
        temporal_patch_size=2,
        spatial_pred_mask_scale=(0.2, 0.8),
        temporal_pred_mask_scale=(1.0, 1.0),
        aspect_ratio=(0.3, 3.0),
        npred=1,
        max_context_frames_ratio=1.0,
        max_keep=None,
        strict=True,
    ):
        super(_AVMaskGenerator, self).__init__()
        if not isinstance(crop_size, tuple):
            crop_size = (crop_size, ) * 2
        self.crop_size = crop_size
        self.height, self.width = crop_size[0] // spatial_patch_size, crop_size[1] // spatial_patch_size
        self.a_size = a_size
        self.a_height = a_size[0] // spatial_patch_size
        self.a_width = a_size[1] // spatial_patch_size
        self.duration = num_frames // temporal_patch_size
        self.spatial_patch_size = spatial_patch_size
        self.temporal_patch_size = temporal_patch_size
        self.aspect_ratio = aspect_ratio
        self.spatial_pred_mask_scale = spatial_pred_mask_scale
        self.temporal_pred_mask_scale = temporal_pred_mask_scale
        self.npred = npred
        self.max_context_duration = max(1, int(self.duration * max_context_frames_ratio))
        self.max_keep = max_keep   # kept for signature parity; unused by the AV generator
        self.strict = strict
        self._itr_counter = _blocks.StepCounter()

    def step(self):
        return self._itr_counter.next()

    def _punch_video(self, keep, b_size):
        t, h, w = b_size
        top = _blocks.draw_offset(self.height, h)
        left = _blocks.draw_offset(self.width, w)
        start = _blocks.draw_offset(self.duration, t)
        keep[start:start + t, top:top + h, left:left + w] = False
        if self.max_context_duration < self.duration:
            keep[self.max_context_duration:, :, :] = False

    def _punch_audio(self, keep):
        h, w = self.AUDIO_BLOCK
        top = _blocks.draw_offset(self.a_height, h)
        left = _blocks.draw_offset(self.a_width, w)
        keep[top:top + h, left:left + w] = False

    def __call__(self, batch_size):
        b_size = _blocks.draw_block_size(
            self.step(), self.duration, self.height, self.width,
            self.temporal_pred_mask_scale, self.spatial_pred_mask_scale, self.aspect_ratio)
        enc_v, enc_a, pred_v, pred_a = [], [], [], []
        while len(enc_v) < batch_size:
            keep_v = torch.ones((self.duration, self.height, self.width), dtype=torch.bool)
            keep_a = torch.ones((self.a_height, self.a_width), dtype=torch.bool)
            for _ in range(self.npred):
                self._punch_video(keep_v, b_size)
                self._punch_audio(keep_a)
            kept_v, drop_v = _blocks.split_keep_drop(keep_v.flatten())
            kept_a, drop_a = _blocks.split_keep_drop(keep_a.flatten())
            if _blocks.strict_len(kept_v, self.strict) == 0:
                continue    # empty video context: resample this sample
            for v in (drop_v, drop_a, kept_a):
                _blocks.strict_len(v, self.strict)
            enc_v.append(kept_v)
            enc_a.append(kept_a)
            pred_v.append(drop_v)
            pred_a.append(drop_a)
        return (_blocks.stack_truncated(enc_v), _blocks.stack_truncated(enc_a),
                _blocks.stack_truncated(pred_v), _blocks.stack_truncated(pred_a))
