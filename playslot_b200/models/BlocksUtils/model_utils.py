"""
Model utilities -- mirror of src/models/BlocksUtils/model_utils.py (same public names).
"""
import numpy as np
import torch
import torch.nn as nn

from .._common import print_


def build_grid(resolution, vmin=-1., vmax=1., device=None):
    """
    (1, H, W, 4) grid of coordinates [y, x, 1-y, 1-x] (reference model_utils.py:12-34): float64 linspace,
    cast to float32, then the complement computed in float32.
    """
    axes = [np.linspace(vmin, vmax, num=r) for r in resolution]
    yy, xx = np.meshgrid(axes[0], axes[1], indexing="ij")
    g = np.stack([yy, xx], axis=-1).astype(np.float32)[None]
    return torch.from_numpy(np.concatenate([g, 1.0 - g], axis=-1)).to(device)


def count_model_params(model, verbose=False):
    n = sum(p.numel() for p in model.parameters() if p.requires_grad)
    if verbose:
        print_(f"  --> Number of learnable parameters: {n}")
    return n


def freeze_params(model):
    print_(f"Freezing {model.__class__.__name__}")
    for p in model.parameters():
        p.requires_grad = False


def unfreeze_params(model):
    print_(f"Unfreezing {model.__class__.__name__}")
    for p in model.parameters():
        p.requires_grad = True


@torch.no_grad()
def init_xavier_(model: nn.Module):
    """ biases -> 0, matrices -> Xavier uniform, vectors untouched (reference model_utils.py:72-90) """
    print_("Xavier Initialization of model parameters:")
    for name, t in model.named_parameters():
        if name.endswith(".bias"):
            t.zero_()
        elif t.dim() > 1:
            try:
                nn.init.xavier_uniform_(t)
            except Exception:
                print_(f" --> {name} of shape {t.shape} could not be initialized...")


@torch.no_grad()
def init_kaiming_(model: nn.Module):
    """ as in the reference (model_utils.py:94-110) only the biases end up being touched """
    print_("Kaiming Initialization of model parameters:")
    for name, t in model.named_parameters():
        if name.endswith(".bias"):
            t.zero_()


def build_slot_causal_mask(seq_len, num_slots, device):
    """ (T*S, T*S) staircase mask: token j visible to token i iff frame(j) <= frame(i) (model_utils.py:115-124) """
    f = torch.arange(seq_len * num_slots, device=device) // num_slots
    return (f[None, :] <= f[:, None]).float()
