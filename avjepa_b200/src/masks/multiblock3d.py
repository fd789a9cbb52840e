"""Video multiblock-3D mask collator (host, DataLoader workers).

Drop-in for the reference's ``src/masks/multiblock3d.py``: ``MaskCollator(cfgs_mask,
crop_size, num_frames, patch_size, tubelet_size)(batch) -> (collated_batch, masks_enc,
masks_pred)`` (``:20-63``) with one ``_MaskGenerator`` per ``cfgs_mask`` entry (``:66-203``).
Outputs are bit-exact with the reference for the same global ``torch.manual_seed`` and call
index: ascending int64 token indices into the t*h*w grid, every row truncated to the
batch-min length, ``max_keep`` applied to the encoder mask only.
"""
import torch

from avjepa_b200.src.masks import _blocks

_GLOBAL_SEED = 0


class MaskCollator(object):

    def __init__(self, cfgs_mask, crop_size=(224, 224), num_frames=16, patch_size=(16, 16), tubelet_size=2):
        super(MaskCollator, self).__init__()
        self.mask_generators = [
            _MaskGenerator(
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
        masks_enc, masks_pred = [], []
        for g in self.mask_generators:
            e, p = g(len(batch))
            masks_enc.append(e)
            masks_pred.append(p)
        return collated_batch, masks_enc, masks_pred


class _MaskGenerator(object):

    def __init__(
        self,
        crop_size=(224, 224),
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
        super(_MaskGenerator, self).__init__()
        if not isinstance(crop_size, tuple):
            crop_size = (crop_size, ) * 2
        self.crop_size = crop_size
        self.height, self.width = crop_size[0] // spatial_patch_size, crop_size[1] // spatial_patch_size
        self.duration = num_frames // temporal_patch_size
        self.spatial_patch_size = spatial_patch_size
        self.temporal_patch_size = temporal_patch_size
        self.aspect_ratio = aspect_ratio
        self.spatial_pred_mask_scale = spatial_pred_mask_scale
        self.temporal_pred_mask_scale = temporal_pred_mask_scale
        self.npred = npred
        # number of time-steps the context may span, and the cap on kept context patches
        self.max_context_duration = max(1, int(self.duration * max_context_frames_ratio))
        self.max_keep = max_keep
        self.strict = strict
        self._itr_counter = _blocks.StepCounter()

    def step(self):
        return self._itr_counter.next()

    def _punch_block(self, keep, b_size):
        """Zero one (t,h,w) block of the [T,H,W] keep-mask; draws top, left, start in that order."""
        t, h, w = b_size
        top = _blocks.draw_offset(self.height, h)
        left = _blocks.draw_offset(self.width, w)
        start = _blocks.draw_offset(self.duration, t)
        keep[start:start + t, top:top + h, left:left + w] = False
        if self.max_context_duration < self.duration:
            keep[self.max_context_duration:, :, :] = False

    def __call__(self, batch_size):
        b_size = _blocks.draw_block_size(
            self.step(), self.duration, self.height, self.width,
            self.temporal_pred_mask_scale, self.spatial_pred_mask_scale, self.aspect_ratio)
        rows_enc, rows_pred = [], []
        while len(rows_enc) < batch_size:
            keep = torch.ones((self.duration, self.height, self.width), dtype=torch.bool)
            for _ in range(self.npred):
                self._punch_block(keep, b_size)
            kept, dropped = _blocks.split_keep_drop(keep.flatten())
            if _blocks.strict_len(kept, self.strict) == 0:
                continue    # empty context: resample this sample
            _blocks.strict_len(dropped, self.strict)
            rows_enc.append(kept)
            rows_pred.append(dropped)
        return _blocks.stack_truncated(rows_enc, self.max_keep), _blocks.stack_truncated(rows_pred)
