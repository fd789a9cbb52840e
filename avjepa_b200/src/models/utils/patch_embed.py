"""Patch-embedding parameter containers.

Mirror of ``src/models/utils/patch_embed.py`` (``PatchEmbed :15-32``, ``PatchEmbed3D :35-61``,
``AudioVisionPatchEmbed3D :64-102``): the weights live in ``nn.Conv3d`` / ``nn.Conv2d`` holders
named ``proj`` / ``audio_proj`` so state-dict keys and shapes ([D,3,2,16,16], [D,1,16,16]) match
the reference.  The convolution itself never runs: a non-overlapping conv is a GEMM over
flattened patches, executed by ``avj_patchify`` + ``avj_gemm`` on the KEPT tokens only, with
bias and positional embedding fused in the GEMM epilogue (see avjepa_b200.backbone).
"""
import torch.nn as nn


class PatchEmbed(nn.Module):
    """Image to patch embedding (2-D)."""

    def __init__(self, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class PatchEmbed3D(nn.Module):
    """Video to tubelet embedding (3-D)."""

    def __init__(self, patch_size=16, tubelet_size=2, in_chans=3, embed_dim=768):
        super().__init__()
        self.patch_size = patch_size
        self.tubelet_size = tubelet_size
        self.proj = nn.Conv3d(in_channels=in_chans, out_channels=embed_dim,
                              kernel_size=(tubelet_size, patch_size, patch_size),
                              stride=(tubelet_size, patch_size, patch_size))


class AudioVisionPatchEmbed3D(nn.Module):
    """Video tubelet embedding plus log-mel spectrogram patch embedding."""

    def __init__(self, patch_size=16, tubelet_size=2, in_chans=3, embed_dim=768):
        super().__init__()
        self.patch_size = patch_size
        self.tubelet_size = tubelet_size
        self.proj = nn.Conv3d(in_channels=in_chans, out_channels=embed_dim,
                              kernel_size=(tubelet_size, patch_size, patch_size),
                              stride=(tubelet_size, patch_size, patch_size))
        self.audio_proj = nn.Conv2d(in_channels=1, out_channels=embed_dim,
                                    kernel_size=(patch_size, patch_size), stride=(patch_size, patch_size))
