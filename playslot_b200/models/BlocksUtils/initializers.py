"""
Slot initialisers -- mirror of src/models/BlocksUtils/initializers.py.
They only produce the (B, S, D) starting slots (a broadcast or one randn); no kernel is involved.
"""
from math import sqrt

import torch
import torch.nn as nn

from .._common import print_

INITIALIZERS = ["Learned", "LearnedRandom", "CoM", "BBox"]


def get_initializer(mode, slot_dim, num_slots):
    """ Factory with the reference's mode names (initializers.py:21-50) """
    if mode not in INITIALIZERS:
        raise ValueError(f"Unknown initializer {mode = }. Available {INITIALIZERS = }")
    if mode == "Learned":
        init = Learned(slot_dim=slot_dim, num_slots=num_slots)
    elif mode == "LearnedRandom":
        init = LearnedRandom(slot_dim=slot_dim, num_slots=num_slots)
    else:
        init = CoordInit(slot_dim=slot_dim, num_slots=num_slots, mode=mode)
    print_("Initializer:")
    print_(f"  --> mode={mode}")
    print_(f"  --> slot_dim={slot_dim}")
    print_(f"  --> num_slots={num_slots}")
    return init


def _uniform_limit(slot_dim):
    return sqrt(6.0 / (1 + slot_dim))


class Learned(nn.Module):
    """ Learned slots, repeated over the batch (initializers.py:54-83); parameter `slots` (1, S, D) """

    def __init__(self, slot_dim, num_slots):
        super().__init__()
        self.slot_dim = slot_dim
        self.num_slots = num_slots
        self.slots = nn.Parameter(torch.empty(1, num_slots, slot_dim).uniform_(-_uniform_limit(slot_dim), _uniform_limit(slot_dim)))

    def forward(self, batch_size, **kwargs):
        return self.slots.repeat(batch_size, 1, 1)


class LearnedRandom(nn.Module):
    """ Slots sampled from a learned diagonal Gaussian (initializers.py:87-124); the noise is drawn with
    torch.randn on the parameters' device, exactly like the reference, so RNG streams line up. """

    def __init__(self, slot_dim, num_slots):
        super().__init__()
        self.slot_dim = slot_dim
        self.num_slots = num_slots
        lim = _uniform_limit(slot_dim)
        self.slots_mu = nn.Parameter(torch.empty(1, 1, slot_dim).uniform_(-lim, lim))
        self.slots_sigma = nn.Parameter(torch.empty(1, 1, slot_dim).uniform_(-lim, lim))

    def forward(self, batch_size, **kwargs):
        mu = self.slots_mu.expand(batch_size, self.num_slots, -1)
        sigma = self.slots_sigma.expand(batch_size, self.num_slots, -1)
        return mu + sigma * torch.randn(mu.shape, device=self.slots_mu.device)


class CoordInit(nn.Module):
    """ Slots from centre-of-mass / bounding-box annotations through a small MLP (initializers.py:128-195).
    Not used by any shipped PlaySlot config; kept for surface parity.  Missing objects are padded with -1
    along the object axis. """

    MODE_REP = {"CoM": "com_coords", "BBox": "bbox_coords"}
    IN_FEATS = {"CoM": 2, "BBox": 4}

    def __init__(self, slot_dim, num_slots, mode):
        assert mode in CoordInit.MODE_REP, f"Unknown {mode = }. Use one of {list(CoordInit.MODE_REP)}"
        super().__init__()
        self.slot_dim = slot_dim
        self.num_slots = num_slots
        self.mode = mode
        self.in_feats = CoordInit.IN_FEATS[mode]
        self.coord_encoder = nn.Sequential(
                nn.Linear(self.in_feats, 256),
                nn.ReLU(),
                nn.Linear(256, slot_dim),
            )
        self.dummy_parameter = nn.Parameter(torch.tensor([0.]))

    def forward(self, **kwargs):
        device = self.dummy_parameter.device
        name = CoordInit.MODE_REP[self.mode]
        coords = kwargs.get(name, None)
        if coords is None or coords.sum() == 0:
            raise ValueError(f"{self.mode} Initializer requires '{name}' as input...")
        if coords.dim() == 4:
            coords = coords[:, 0]
        coords = coords.to(device)
        B, n = coords.shape[0], coords.shape[1]
        if n > self.num_slots:
            raise ValueError(f"There shouldnt be more {n = } than {self.num_slots = }! ")
        if n < self.num_slots:
            pad = -torch.ones((B, self.num_slots - n, self.in_feats), device=device)
            coords = torch.cat([coords, pad], dim=1)
        return self.coord_encoder(coords)
