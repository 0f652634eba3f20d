"""Host-side mirror of the position-to-grid assignment snippet every reference caller copy-pastes
(image_model/inference.py:113-125 `find_permutation`, :294-306 rearrange / mean / pairwise_distances / argsort).

The work runs in the warp-per-puzzle kernels of csrc/assign.cu; results are integer and bit-exact against the
numpy/sklearn snippet on identical score matrices.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import ops
from .models import get_2d_sincos_pos_embed


def find_permutation(distance_matrix, sentinel: float = 1e9):
    """Greedy column-by-column arg-min (inference.py:113-125).  Accepts a [n,n] numpy array / tensor and returns the
    reference's `sort_list` (python list of row indices); a [B,n,n] input returns an int32 tensor [B,n]."""
    if isinstance(distance_matrix, np.ndarray):
        scores = torch.from_numpy(np.ascontiguousarray(distance_matrix, dtype=np.float64)).cuda()
    else:
        scores = distance_matrix.to(device="cuda", dtype=torch.float64)
    single = scores.dim() == 2
    if single:
        scores = scores.unsqueeze(0)
    order, _ = ops.assign_from_scores(scores.contiguous(), sentinel)
    return [int(v) for v in order[0].tolist()] if single else order


def canonical_embeddings(grid_size: int, device) -> torch.Tensor:
    """The assignment targets: get_2d_sincos_pos_embed(8, G) cast to fp32 (inference.py:220)."""
    return torch.tensor(get_2d_sincos_pos_embed(8, grid_size)).float().to(device).contiguous()


def solve_puzzles(latents: torch.Tensor, grid_size: int, sentinel: float = 1e9, return_scores: bool = False):
    """latents [B,T,8] (the p_sample_loop result) -> (order, pred[, scores]); pred[b, i] is the grid cell assigned to
    slot i, i.e. the reference's `np.asarray(order).argsort()`; the puzzle is solved iff pred == the scramble indices."""
    canon = canonical_embeddings(grid_size, latents.device)
    return ops.assign_greedy_l1(latents.float().contiguous(), canon, grid_size, sentinel, return_scores)


def accuracy(pred: torch.Tensor, indices: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(puzzle_correct [B] bool, patch_matches [B] int) as logged by inference.py:309-316."""
    eq = pred.to(torch.int64) == indices.to(device=pred.device, dtype=torch.int64)
    return eq.all(dim=1), eq.sum(dim=1)
