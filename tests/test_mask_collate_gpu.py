"""Device-side mask sampling: ``DeviceAVMaskCollator`` against the host ``AVMaskCollator``
(itself pinned bit-exactly to the reference's collator by tests/test_masks_host.py and the golden mask vectors): same
torch generator state in -> identical masks out, identical generator state afterwards."""
import pytest
import torch

pytestmark = pytest.mark.gpu

MASK_CFG = [   # configs/pretrain/vitl16.yaml:38-62
    dict(aspect_ratio=(0.75, 1.5), num_blocks=8, spatial_scale=(0.15, 0.15), temporal_scale=(1.0, 1.0), max_temporal_keep=1.0, max_keep=None),
    dict(aspect_ratio=(0.75, 1.5), num_blocks=2, spatial_scale=(0.7, 0.7), temporal_scale=(1.0, 1.0), max_temporal_keep=1.0, max_keep=None),
]
SHORT_CFG = [  # temporal blocks shorter than the clip and a context limited to the first frames
    dict(aspect_ratio=(0.3, 3.0), num_blocks=3, spatial_scale=(0.2, 0.6), temporal_scale=(0.4, 0.9), max_temporal_keep=0.5, max_keep=None),
]


def _host_run(cfg, seed, B, calls):
    from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator
    torch.manual_seed(seed)
    c = AVMaskCollator(cfg, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
    out = []
    for _ in range(calls):
        try:
            out.append(c([torch.zeros(1)] * B)[1:])
        except TypeError:
            out.append(None)         # the reference's one-element quirk; the stream position is then undefined
            break
    return out, torch.get_rng_state()


def _device_run(cfg, seed, B, calls):
    from avjepa_b200.src.masks.device_collator import DeviceAVMaskCollator
    torch.manual_seed(seed)
    c = DeviceAVMaskCollator(cfg, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2, device='cuda', sync_host_rng=True)
    out = []
    for _ in range(calls):
        try:
            out.append(c([torch.zeros(1)] * B)[1:])
        except TypeError:
            out.append(None)
            break
    return out, torch.get_rng_state()


@pytest.mark.parametrize('cfg_name,seed,B', [('vitl16', 234, 24), ('vitl16', 0, 2), ('vitl16', 7, 5), ('short', 3, 8), ('short', 11, 1)])
def test_device_collator_is_bit_identical_to_the_host_collator(cfg_name, seed, B):
    cfg = MASK_CFG if cfg_name == 'vitl16' else SHORT_CFG
    host, host_state = _host_run(cfg, seed, B, calls=4)
    dev, dev_state = _device_run(cfg, seed, B, calls=4)
    assert len(host) == len(dev)
    for h, d in zip(host, dev):
        if h is None or d is None:
            assert h is None and d is None
            return
        for hk, dk in zip(h, d):                      # enc_v, enc_a, pred_v, pred_a
            assert len(hk) == len(dk)
            for hm, dm in zip(hk, dk):
                assert dm.is_cuda and dm.dtype == torch.int64
                assert hm.shape == dm.shape, (hm.shape, dm.shape)
                assert torch.equal(hm, dm.cpu())
    assert torch.equal(host_state, dev_state)          # the global generator ends at the same position


def test_device_masks_drive_a_training_step():
    """The device masks go straight into TrainStep and give the loss the host masks give."""
    import step_support as S
    from helpers import build_product
    from avjepa_b200.src.masks.device_collator import DeviceAVMaskCollator
    from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator
    enc, pred = build_product('vit_tiny', seed=0, device='cuda', pred_depth=2)
    step = S.make_train_step(enc, pred, True)
    g = torch.Generator().manual_seed(1)
    clips = torch.randn(2, 3, 16, 224, 224, generator=g).cuda()
    asgram = (-80.0 * torch.rand(2, 1, 128, 192, generator=g)).cuda()
    torch.manual_seed(99)
    _, ev, ea, pv, pa = DeviceAVMaskCollator(MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2, sync_host_rng=True)([0, 0])
    loss_d, _, _ = step.forward_loss(clips, asgram, ev, ea, pv, pa)
    torch.manual_seed(99)
    _, ev, ea, pv, pa = AVMaskCollator(MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)([0, 0])
    loss_h, _, _ = step.forward_loss(clips, asgram, [m.cuda() for m in ev], [m.cuda() for m in ea], [m.cuda() for m in pv], [m.cuda() for m in pa])
    assert float(loss_d) == float(loss_h)
