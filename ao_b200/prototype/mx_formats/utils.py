"""Scale-layout helpers (reference: torchao/prototype/mx_formats/utils.py:25-134)."""
import torch


def ceil_div(a, b):
    return (a + b - 1) // b


def to_blocked(input_matrix: torch.Tensor) -> torch.Tensor:
    """[H, W] -> flat 32*ceil(H/128) x 16*ceil(W/4): the cuBLAS / tcgen05.cp block-scale layout
    (each 128x4 tile becomes 32 rows of 16 bytes: (r%32)*16 + (r/32)*4 + c)."""
    rows, cols = input_matrix.shape
    rb, cb = ceil_div(rows, 128), ceil_div(cols, 4)
    padded = input_matrix
    if (rows, cols) != (rb * 128, cb * 4):
        padded = torch.zeros((rb * 128, cb * 4), device=input_matrix.device, dtype=input_matrix.dtype)
        padded[:rows, :cols] = input_matrix
    blocks = padded.view(rb, 128, cb, 4).permute(0, 2, 1, 3)
    return blocks.reshape(-1, 4, 32, 4).transpose(1, 2).reshape(-1, 32, 16).flatten()


def from_blocked(blocked: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    rb, cb = ceil_div(rows, 128), ceil_div(cols, 4)
    t = blocked.reshape(rb * cb, 32, 4, 4).transpose(1, 2).reshape(rb, cb, 128, 4).permute(0, 2, 1, 3)
    return t.reshape(rb * 128, cb * 4)[:rows, :cols]


def hp_data_dims_to_swizzled_scale_dims_mx(M, K):
    return ceil_div(M, 128) * 32, ceil_div(K // 32, 4) * 16


def hp_data_dims_to_swizzled_scale_dims_nvfp4(M, K):
    return ceil_div(M, 128) * 32, ceil_div(K // 16, 4) * 16
