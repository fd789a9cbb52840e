"""Checkpoint hand-off from pre-training to the frozen evaluations.

Mirror of ``load_pretrained`` in the reference's ``evals/video_classification_frozen/eval.py:419-444``: a
pre-training checkpoint stores ``nn.DataParallel(MultiMaskWrapper(backbone))`` state dicts, i.e. keys of the form
``module.backbone.<param>``, while the evaluations build a bare backbone.  The rest of that file (data loaders,
the attentive-probe training loop) is outside the hot path.
"""
import logging

import torch

logger = logging.getLogger()

_WRAPPER_PREFIXES = ('module.', 'backbone.')


def bare_backbone_keys(state_dict):
    """Drop every occurrence of the wrapper prefixes from the keys (the reference uses str.replace, so a prefix is
    removed wherever it occurs; parameter names never contain these substrings elsewhere)."""
    out = {}
    for key, value in state_dict.items():
        for prefix in _WRAPPER_PREFIXES:
            key = key.replace(prefix, '')
        out[key] = value
    return out


def load_pretrained(encoder, pretrained, checkpoint_key='target_encoder'):
    """Load `checkpoint_key` (falling back to ``'encoder'``) of the checkpoint file `pretrained` into the bare
    `encoder`.  Parameters absent from the file keep their initial values, parameters whose shape differs keep
    the model's own tensor (both are logged), and the load is non-strict -- the reference's behaviour."""
    logger.info(f'Loading pretrained model from {pretrained}')
    checkpoint = torch.load(pretrained, map_location='cpu')
    source = checkpoint[checkpoint_key] if checkpoint_key in checkpoint else checkpoint['encoder']
    incoming = bare_backbone_keys(source)
    for name, own in encoder.state_dict().items():
        theirs = incoming.get(name)
        if theirs is None:
            logger.info(f'key "{name}" could not be found in loaded state dict')
        elif theirs.shape != own.shape:
            logger.info(f'key "{name}" is of different shape in model and loaded state dict')
            incoming[name] = own
    result = encoder.load_state_dict(incoming, strict=False)
    logger.info(f'loaded pretrained model with msg: {result}')
    logger.info(f'loaded pretrained encoder from epoch: {checkpoint["epoch"]}\n path: {pretrained}')
    return encoder
