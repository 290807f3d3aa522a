"""Drop-in for `src.torch_utils.ops.grid_sample_gradfix` (reference: src/torch_utils/ops/grid_sample_gradfix.py).
Only the ADA augmentation pipeline uses it (augment.py:297) — outside the hot path; current
torch differentiates grid_sample to second order natively, so this is a thin pass-through with the same surface."""
import torch

enabled = False


def grid_sample(input, grid):
    return torch.nn.functional.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros', align_corners=False)
