"""Shared test helpers: golden loading, product-model construction, oracle bridging."""
import os

import numpy as np
import torch

from conftest import GOLDEN

MASK_CFG = [
    dict(aspect_ratio=(0.75, 1.5), num_blocks=8, spatial_scale=(0.15, 0.15), temporal_scale=(1.0, 1.0),
         max_temporal_keep=1.0, max_keep=None),
    dict(aspect_ratio=(0.75, 1.5), num_blocks=2, spatial_scale=(0.7, 0.7), temporal_scale=(1.0, 1.0),
         max_temporal_keep=1.0, max_keep=None),
]


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def checksum(t):
    t = t.detach().double().flatten().cpu()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64) % 7 + 1
    return np.array([float(t.sum()), float(t.abs().sum()), float((t * w).sum())])


def build_product(model_name='vit_tiny', seed=0, device='cpu', pred_depth=12):
    """The product's init_audio_video_model under the reference's seeding protocol."""
    import logging
    logging.disable(logging.CRITICAL)
    from avjepa_b200.app.avjepa.utils import init_audio_video_model
    torch.manual_seed(seed)
    np.random.seed(seed)
    return init_audio_video_model(
        device=torch.device(device), patch_size=16, num_frames=16, tubelet_size=2, model_name=model_name, crop_size=224,
        pred_depth=pred_depth, pred_embed_dim=384, uniform_power=True, use_mask_tokens=True, num_mask_tokens=2,
        zero_init_mask_tokens=True, use_sdpa=True)


def backbone_params(wrapper):
    """{reference backbone key: tensor} for the oracle."""
    return {k[len('backbone.'):]: v.detach().cpu().float().clone() for k, v in wrapper.state_dict().items()}


def step_inputs():
    """The seeded synthetic inputs of tests/golden/step_tiny.npz."""
    g = torch.Generator().manual_seed(1234)
    clips = torch.randn(2, 3, 16, 224, 224, generator=g)
    asgram = -80.0 * torch.rand(2, 1, 128, 192, generator=g)
    gz = golden('step_tiny.npz')
    masks = {nm: [torch.from_numpy(gz[f'mask_g{gi}_{nm}'].astype(np.int64)) for gi in range(2)]
             for nm in ('ev', 'ea', 'pv', 'pa')}
    return clips, asgram, masks, gz


def rel_err(a, b):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
