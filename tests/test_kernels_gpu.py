"""GPU kernel numerics: every C-ABI kernel against a plain PyTorch fp32 reference of the same op."""
import pytest

pytestmark = pytest.mark.gpu

F32, BF16 = 0, 1
NT, NN, TN = 0, 1, 2

GEMM_CASES = [
    # (code, layout, M, N, K, epilogue)
    (F32, NT, 100, 192, 192, 'bias'), (F32, NN, 77, 64, 136, 'none'), (F32, TN, 192, 64, 333, 'accum'),
    (F32, NT, 130, 768, 192, 'bias_gelu'), (F32, NT, 64, 192, 768, 'bias_res'), (F32, NN, 96, 768, 192, 'dact'),
    (F32, NT, 120, 192, 256, 'pos_map'),
    (BF16, NT, 128, 256, 64, 'none'), (BF16, NT, 1000, 1024, 1024, 'bias'), (BF16, NT, 333, 576, 192, 'bias'),
    (BF16, NT, 700, 768, 192, 'bias_gelu'), (BF16, NT, 515, 192, 768, 'bias_res'), (BF16, NT, 480, 384, 1536, 'pos_map'),
    (BF16, NN, 128, 256, 64, 'none'), (BF16, NN, 900, 1024, 3072, 'none'), (BF16, NN, 650, 768, 192, 'dact'),
    (BF16, TN, 128, 256, 64, 'accum'), (BF16, TN, 3072, 1024, 5000, 'accum'), (BF16, TN, 192, 1536, 777, 'accum'),
    (BF16, TN, 384, 192, 130, 'accum'), (BF16, NT, 96, 1280, 1280, 'bias'),
]


@pytest.mark.parametrize('code,layout,M,N,K,epi', GEMM_CASES)
def test_gemm(code, layout, M, N, K, epi):
    import kernel_checks as kc
    ok, err = kc.check_gemm(code, layout, M, N, K, epi)
    assert ok, f'rel err {err}'


def test_bf16_gemm_the_tensor_core_kernel_refuses_is_an_error_not_a_silent_simt_fallback():
    """N = 72 is not a multiple of 64: the tcgen05 kernel cannot take it.  The call must fail loudly (a production
    call that silently ran the fp32-FMA check kernel would be a 100x cliff)."""
    import kernel_checks as kc
    from avjepa_b200._cabi import AvjError
    with pytest.raises(AvjError, match='not supported by the tcgen05 kernel'):
        kc.check_gemm(BF16, NT, 50, 72, 64, 'bias')
    ok, err = kc.check_gemm(F32, NT, 50, 72, 64, 'bias')          # fp32 check mode takes any shape
    assert ok, err


def test_launch_counter_counts_kernels():
    import torch
    from avjepa_b200 import _cabi, engine
    x = torch.randn(64, 256, device='cuda')
    y = torch.empty(64, 256, device='cuda')
    n0 = _cabi.launch_count()
    engine.layernorm_fwd(x.data_ptr(), None, None, y.data_ptr(), F32, None, None, 64, 256, 1e-5)
    assert _cabi.launch_count() == n0 + 1
    ws = torch.empty(int(_cabi.load().avj_sumsq_ws_floats(0)), device='cuda')
    out = torch.empty(1, device='cuda')
    _cabi.call('avj_sumsq', x.data_ptr(), x.numel(), out.data_ptr(), ws.data_ptr(), engine.stream())
    assert _cabi.launch_count() == n0 + 3                          # partial + final kernels
    assert float(out) == pytest.approx(float((x.double() ** 2).sum()), rel=1e-5)


def test_segment_stats_kernel():
    """avj_segment_stats: per-parameter sum of squares / sum of |x| over a flat buffer with ragged segments."""
    import torch
    from avjepa_b200 import _cabi, engine
    g = torch.Generator(device='cuda').manual_seed(0)
    sizes = [8, 40000, 16, 8, 9000, 1024, 24, 70000, 8]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    x = torch.randn(offs[-1], generator=g, device='cuda')
    seg = torch.tensor(offs, dtype=torch.int64, device='cuda')
    for mode in (0, 1):
        out = torch.zeros(len(sizes), dtype=torch.float64, device='cuda')
        _cabi.call('avj_segment_stats', x.data_ptr(), seg.data_ptr(), len(sizes), x.numel(), mode, out.data_ptr(), engine.stream())
        ref = torch.stack([(x[a:b].double() ** 2).sum() if mode == 0 else x[a:b].double().abs().sum()
                           for a, b in zip(offs[:-1], offs[1:])])
        assert torch.allclose(out, ref, rtol=1e-5), (mode, out, ref)


@pytest.mark.parametrize('code,rows,D,affine', [(F32, 333, 192, True), (BF16, 1000, 1024, True), (BF16, 77, 384, True),
                                               (F32, 50, 1280, True), (F32, 64, 768, False)])
def test_layernorm(code, rows, D, affine):
    import kernel_checks as kc
    ok, err = kc.check_layernorm(code, rows, D, affine)
    assert ok, f'rel err {err}'


@pytest.mark.parametrize('code,B,N,H,hd', [(F32, 2, 100, 3, 64), (F32, 1, 333, 2, 24), (BF16, 2, 257, 4, 64),
                                          (BF16, 1, 130, 2, 24), (BF16, 1, 96, 2, 80), (BF16, 2, 40, 2, 128),
                                          (F32, 1, 70, 2, 32), (BF16, 1, 1216, 2, 24), (BF16, 2, 1664, 2, 64),
                                          (BF16, 2, 375, 3, 64), (BF16, 1, 64, 1, 64), (BF16, 3, 159, 2, 64),
                                          (BF16, 1, 300, 2, 32), (BF16, 1, 200, 2, 16), (BF16, 1, 100, 2, 48),
                                          (BF16, 1, 129, 1, 8), (BF16, 2, 333, 3, 80), (BF16, 1, 260, 2, 96),
                                          (BF16, 1, 200, 2, 112), (BF16, 1, 513, 2, 128),
                                          (BF16, 1, 50, 2, 24), (BF16, 2, 100, 2, 32), (BF16, 1, 64, 3, 16)])
def test_attention(code, B, N, H, hd):
    import kernel_checks as kc
    ok, err = kc.check_attention(code, B, N, H, hd)
    assert ok, f'rel err {err} {kc.LAST_ATTENTION_ERRORS}'


@pytest.mark.parametrize('code', [F32, BF16])
def test_gather_rows_bit_exact(code):
    import kernel_checks as kc
    ok, _ = kc.check_gather(code)
    assert ok


@pytest.mark.parametrize('code,audio', [(F32, False), (BF16, False), (F32, True), (BF16, True)])
def test_patchify(code, audio):
    import kernel_checks as kc
    ok, err = kc.check_patchify(code, audio=audio)
    assert ok, f'rel err {err}'


@pytest.mark.parametrize('audio,masked,D', [(False, True, 192), (False, False, 128), (True, True, 192), (True, False, 64),
                                            (False, True, 1024)])
def test_patch_embed_im2col_free(audio, masked, D):
    import kernel_checks as kc
    ok, err = kc.check_patch_embed(D=D, audio=audio, masked=masked)
    assert ok, f'rel err {err}'


@pytest.mark.parametrize('audio,masked,D', [(False, True, 192), (False, False, 128), (True, True, 384), (True, False, 64),
                                            (False, True, 1024)])
def test_patch_embed_weight_gradient_without_patch_matrix(audio, masked, D):
    import kernel_checks as kc
    ok, err = kc.check_patch_embed_wgrad(D=D, audio=audio, masked=masked)
    assert ok, f'rel err {err}'


def test_row_kernels():
    import kernel_checks as kc
    ok, err = kc.check_rows()
    assert ok, err


def test_loss():
    import kernel_checks as kc
    ok, err = kc.check_loss()
    assert ok, err


def test_adamw_ema():
    import kernel_checks as kc
    ok, err = kc.check_adamw()
    assert ok, err
