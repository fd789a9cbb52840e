"""Fixed sin/cos positional tables (host, init-time, float64 numpy).

Mirror of the reference's ``src/models/utils/pos_embs.py`` public functions
(``get_3d_sincos_pos_embed :11-44``, ``get_2d_sincos_pos_embed :47-63``,
``get_2d_sincos_pos_embed_xy :65-81``, ``get_1d_sincos_pos_embed :84-96``,
``get_1d_sincos_pos_embed_from_grid :99-117``): same names, arguments and token order.
The tables are built once per model and live in HBM as fp32 ``[1, N, D]`` buffers.
"""
import numpy as np


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    """(M,) positions -> (M, embed_dim): first half sines, second half cosines."""
    assert embed_dim % 2 == 0
    n_freq = embed_dim // 2
    freq = 1.0 / 10000 ** (np.arange(n_freq, dtype=float) / (embed_dim / 2.0))
    phase = np.asarray(pos, dtype=float).reshape(-1, 1) * freq.reshape(1, -1)
    return np.concatenate([np.sin(phase), np.cos(phase)], axis=1)


def _with_cls(table, cls_token):
    if cls_token:
        table = np.concatenate([np.zeros([1, table.shape[1]]), table], axis=0)
    return table


def get_1d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    table = get_1d_sincos_pos_embed_from_grid(embed_dim, np.arange(grid_size, dtype=float))
    return _with_cls(table, cls_token)


def get_2d_sincos_pos_embed_xy(embed_dim, grid_h, grid_w, cls_token=False):
    """Row-major (h, w) grid; columns = [height half | width half]."""
    hh, ww = np.meshgrid(np.arange(grid_h, dtype=float), np.arange(grid_w, dtype=float), indexing='ij')
    table = np.concatenate([
        get_1d_sincos_pos_embed_from_grid(embed_dim // 2, hh),
        get_1d_sincos_pos_embed_from_grid(embed_dim // 2, ww)], axis=1)
    return _with_cls(table, cls_token)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    return get_2d_sincos_pos_embed_xy(embed_dim, grid_size, grid_size, cls_token=cls_token)


def get_3d_sincos_pos_embed(embed_dim, grid_size, grid_depth, cls_token=False, uniform_power=False):
    """Row-major (d, h, w) grid; columns = [depth | height | width], truncated to embed_dim.

    ``uniform_power`` gives each axis ceil(D/6)*2 columns; otherwise D/2, D/4, D/4.
    """
    dd, hh, ww = np.meshgrid(np.arange(grid_depth, dtype=float), np.arange(grid_size, dtype=float),
                             np.arange(grid_size, dtype=float), indexing='ij')
    if uniform_power:
        dim_d = dim_h = dim_w = int(np.ceil(embed_dim / 6) * 2)
    else:
        dim_d, dim_h, dim_w = embed_dim // 2, embed_dim // 4, embed_dim // 4
    table = np.concatenate([
        get_1d_sincos_pos_embed_from_grid(dim_d, dd),
        get_1d_sincos_pos_embed_from_grid(dim_h, hh),
        get_1d_sincos_pos_embed_from_grid(dim_w, ww)], axis=1)[:, :embed_dim]
    return _with_cls(table, cls_token)
