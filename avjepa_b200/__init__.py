"""avjepa_b200 -- B200-native (sm_100a) AV-JEPA masked encoder/predictor training step.

Layout
  csrc/ , lib/          hand-written CUDA kernels and the C-ABI shared library they build into
  _cabi.py              ctypes binding of include/avjepa_b200.h (no CPU fallback)
  engine.py             raw-pointer launches + explicit fwd/bwd schedule of a transformer stack
  backbone.py           encoder / predictor schedules and their autograd nodes
  loss.py, optim.py     fused latent loss; fused AdamW + EMA over flat buffers
  dist.py               NCCL data-parallel gradient averaging
  src/ , app/           mirrors of the reference's Python interface (same module paths, names,
                        signatures and state-dict keys as johnshizhu/AVJEPA's src/ and app/)
"""
__version__ = '0.1.0'
