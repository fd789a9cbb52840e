"""ctypes binding of ``libavjepa_sm100.so`` (the C ABI declared in ``include/avjepa_b200.h``).

There is no CPU or PyTorch fallback: if the library is missing or a call fails the caller
gets an exception.  Build the library with ``python -c 'import __graft_entry__ as g; g.build()'``
or ``make -C avjepa_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libavjepa_sm100.so')

F32, BF16 = 0, 1
GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2


class RowMap(C.Structure):
    _fields_ = [('rows_per_group', C.c_int32), ('group_stride', C.c_int32), ('row_offset', C.c_int32)]


IDENTITY = RowMap(0, 0, 0)


class Epilogue(C.Structure):
    _fields_ = [
        ('bias', C.c_void_p), ('residual', C.c_void_p), ('pos', C.c_void_p), ('pos_idx', C.c_void_p),
        ('pos_rows', C.c_int32), ('act', C.c_int32),
        ('pre_out', C.c_void_p), ('dact_aux', C.c_void_p),
        ('accumulate', C.c_int32), ('out_dtype', C.c_int32),
        ('out_map', RowMap),
    ]


class AdamWArgs(C.Structure):
    _fields_ = [
        ('p', C.c_void_p), ('g', C.c_void_p), ('m', C.c_void_p), ('v', C.c_void_p),
        ('target', C.c_void_p), ('p_lp', C.c_void_p), ('target_lp', C.c_void_p),
        ('n', C.c_int64),
        ('lr', C.c_float), ('wd', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
        ('step', C.c_int32), ('ema_m', C.c_float), ('skip_update', C.c_int32), ('zero_grad', C.c_int32),
        ('scale_ptr', C.c_void_p),
    ]


class LinearWS(C.Structure):
    _fields_ = [('w', C.c_void_p), ('b', C.c_void_p), ('gw', C.c_void_p), ('gb', C.c_void_p)]


class NormWS(C.Structure):
    _fields_ = [('w', C.c_void_p), ('b', C.c_void_p), ('gw', C.c_void_p), ('gb', C.c_void_p),
                ('eps', C.c_float), ('pad_', C.c_int32)]


class Layer(C.Structure):
    _fields_ = [
        ('n1', NormWS), ('qkv', LinearWS), ('proj', LinearWS), ('n2', NormWS), ('fc1', LinearWS), ('fc2', LinearWS),
        ('x', C.c_void_p), ('mean1', C.c_void_p), ('rstd1', C.c_void_p), ('h1', C.c_void_p), ('qkv_act', C.c_void_p),
        ('o', C.c_void_p), ('lse', C.c_void_p), ('x1', C.c_void_p), ('mean2', C.c_void_p), ('rstd2', C.c_void_p),
        ('h2', C.c_void_p), ('pre', C.c_void_p), ('act', C.c_void_p), ('x_out', C.c_void_p),
    ]


MAX_SEGMENTS = 8


class Stack(C.Structure):
    _fields_ = [('dtype', C.c_int32), ('B', C.c_int32), ('N', C.c_int32), ('D', C.c_int32), ('H', C.c_int32),
                ('hidden', C.c_int32), ('L', C.c_int32), ('n_seg', C.c_int32),
                ('seg_B', C.c_int32 * MAX_SEGMENTS), ('seg_N', C.c_int32 * MAX_SEGMENTS)]


class StackScratch(C.Structure):
    _fields_ = [('dxa', C.c_void_p), ('dxb', C.c_void_p), ('dx_lp', C.c_void_p), ('d_hid', C.c_void_p),
                ('d_qkv', C.c_void_p), ('d_h', C.c_void_p), ('d_o', C.c_void_p), ('ws', C.c_void_p),
                ('layer_done', C.POINTER(C.c_void_p))]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); mirrors include/avjepa_b200.h one to one
PROTOTYPES = {
    'avj_version': (_i, []),
    'avj_last_error_string': (C.c_char_p, []),
    'avj_device_ok': (_i, []),
    'avj_launch_count': (_i64, []),
    'avj_gemm': (_i, [_i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, C.POINTER(Epilogue), _vp]),
    'avj_patchify': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'avj_patch_embed': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(Epilogue), _vp]),
    'avj_patch_embed_supported': (_i, [_i, _i, _i, _i, _i, _i]),
    'avj_patch_embed_wgrad': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'avj_mask_collate': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'avj_gather_rows_fwd': (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'avj_gather_rows_bwd': (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'avj_copy_rows': (_i, [_vp, _i, _i, RowMap, _vp, _i, _i, RowMap, _i, _i, _i, _vp]),
    'avj_fill_mask_tokens': (_i, [_vp, _vp, _vp, _vp, _i, RowMap, _i, _i, _vp]),
    'avj_colsum_ws_floats': (_i64, [_i, _i]),
    'avj_colsum': (_i, [_vp, _i, _i, RowMap, _vp, _i, _i, _vp, _vp]),
    'avj_layernorm_fwd': (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _f, _vp]),
    'avj_layernorm_bwd_ws_floats': (_i64, [_i, _i]),
    'avj_layernorm_bwd': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    'avj_colsum2': (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp]),
    'avj_attention_fwd': (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    'avj_attention_bwd_ws_floats': (_i64, [_i, _i, _i, _i]),
    'avj_attention_bwd': (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    'avj_xattn_fwd': (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    'avj_xattn_bwd': (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    'avj_loss_ws_floats': (_i64, [_i64]),
    'avj_loss_fwd_bwd': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _f, _i, _f, _f, _vp, _vp]),
    'avj_reg_accumulate': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    'avj_reg_finish': (_i, [_vp, _vp, _i, _vp]),
    'avj_adamw_ema_step': (_i, [C.POINTER(AdamWArgs), _vp]),
    'avj_sumsq_ws_floats': (_i64, [_i64]),
    'avj_sumsq': (_i, [_vp, _i64, _vp, _vp, _vp]),
    'avj_clip_coef': (_i, [_vp, _f, _f, _vp, _vp]),
    'avj_segment_stats': (_i, [_vp, _vp, _i, _i64, _i, _vp, _vp]),
    'avj_cast': (_i, [_vp, _vp, _i, _i64, _vp]),
    'avj_memset_zero': (_i, [_vp, _i64, _vp]),
    'avj_prof_enable': (_i, [_i]),
    'avj_prof_collect': (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    'avj_prof_dump': (_i, [C.c_char_p]),
    'avj_stack_forward': (_i, [C.POINTER(Stack), C.POINTER(Layer), _vp]),
    'avj_stack_backward': (_i, [C.POINTER(Stack), C.POINTER(Layer), C.POINTER(StackScratch), _vp]),
}

_lib = None


class AvjError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AvjError(
            f'{LIB_PATH} not found: the CUDA library is not built. Run '
            f'`python -c "import __graft_entry__ as g; g.build()"` (or `make -C avjepa_b200/csrc`). '
            f'There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.avj_version() != 1:
        raise AvjError(f'ABI version mismatch: library reports {lib.avj_version()}, binding expects 1')
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().avj_last_error_string()
        raise AvjError(f'{what} failed (rc={rc}): {msg.decode() if msg else "?"}')


def launch_count():
    """Kernels launched by the library in this process so far -- counted inside the library at every launch site
    (``AVJ_LAUNCH_CHECK``), not estimated; bench.py reports the difference over its timed region."""
    return int(load().avj_launch_count())


def call(name, *args):
    check(getattr(load(), name)(*args), name)


PROF_FAMILIES = ('gemm', 'attention_fwd', 'attention_bwd', 'layernorm_fwd', 'layernorm_bwd', 'colsum', 'optimizer', 'other')


def prof_enable(on):
    check(load().avj_prof_enable(1 if on else 0), 'avj_prof_enable')


def prof_dump(path):
    """CSV of every timed launch since prof_enable(True): family,work,ms,d0..d3 (include/avjepa_b200.h)."""
    check(load().avj_prof_dump(str(path).encode()), 'avj_prof_dump')


def prof_collect():
    """{family: (ms, work, launches)} for the records since prof_enable(True)."""
    out = {}
    for i, name in enumerate(PROF_FAMILIES):
        ms, work, n = C.c_double(0), C.c_double(0), C.c_int(0)
        check(load().avj_prof_collect(i, C.byref(ms), C.byref(work), C.byref(n)), 'avj_prof_collect')
        out[name] = (ms.value, work.value, n.value)
    return out
