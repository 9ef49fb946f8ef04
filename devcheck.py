"""Device-side checks of what a benchmark or a full-size test just computed (measurement support, not product).

Size-independent parity properties of the LU path (sizes the oracle cannot reach):
the scaled backward error ||P A - L U||_F / (max(m, n) eps ||A||_F) of getrf.rs:12-27's contract, evaluated with
torch in f64 on the GPU, blockwise so that a 16 384^2 (or 65 536-row) problem needs no n x n temporaries beyond
the inputs.  eps = epsilon / 2 as in the reference's `Real::eps()` (src/scalar.rs:393-402).
"""
from __future__ import annotations

import numpy as np


def perm_from_pivots(piv, m: int) -> np.ndarray:
    """Row permutation of the sequential interchanges (laswp.rs:11-40): (P A)[i] = A[perm[i]]."""
    perm = np.arange(m)
    piv = np.asarray(piv)
    for i in np.nonzero(piv != np.arange(len(piv)))[0]:
        p = piv[i]
        perm[i], perm[p] = perm[p], perm[i]
    return perm


def backward_error_dev(a0, lu, piv, block: int = 4096) -> float:
    """a0, lu: torch CUDA tensors (m x n, any float dtype); piv: host or device int vector of length min(m, n)."""
    import torch
    m, n = a0.shape
    k = min(m, n)
    piv = piv.cpu().numpy() if hasattr(piv, "cpu") else np.asarray(piv)
    perm = torch.from_numpy(perm_from_pivots(piv, m)).to(a0.device)
    eps = 2.0 ** -53 if a0.dtype == torch.float64 else 2.0 ** -24
    num2 = torch.zeros((), dtype=torch.float64, device=a0.device)
    den2 = torch.zeros((), dtype=torch.float64, device=a0.device)
    U = torch.triu(lu[:k, :].double())                       # k x n
    for r0 in range(0, m, block):
        r1 = min(m, r0 + block)
        Lb = lu[r0:r1, :k].double()
        rows = torch.arange(r0, r1, device=a0.device)
        cols = torch.arange(k, device=a0.device)
        Lb = torch.where(cols[None, :] < rows[:, None], Lb, torch.zeros((), dtype=torch.float64, device=a0.device))
        diag = rows[rows < k]
        Lb[diag - r0, diag] = 1.0
        PA = a0[perm[r0:r1]].double()
        num2 += ((PA - Lb @ U) ** 2).sum()
        den2 += (PA ** 2).sum()
        del Lb, PA
    return float(torch.sqrt(num2) / (max(m, n) * eps * torch.sqrt(den2)))
