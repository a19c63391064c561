"""Shared helpers for the parity tests."""
import numpy as np
import torch

REL_TOL = 1e-3        # north_star: accent outputs / CTC loss within 1e-3 relative of the reference forward
FLOOR = 1e-6          # elementwise relative error with a 1e-6 floor


def rel_err(got, want, floor=FLOOR):
    got = got.detach().cpu().double().numpy() if isinstance(got, torch.Tensor) else np.asarray(got, dtype=np.float64)
    want = want.detach().cpu().double().numpy() if isinstance(want, torch.Tensor) else np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor))) if got.size else 0.0


def norm_err(got, want):
    """max-abs error relative to the largest reference magnitude (for wide-range tensors)."""
    got = got.detach().cpu().double().numpy() if isinstance(got, torch.Tensor) else np.asarray(got, dtype=np.float64)
    want = want.detach().cpu().double().numpy() if isinstance(want, torch.Tensor) else np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    return float(np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-30))


def t64(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def dev(a, device="cuda"):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
