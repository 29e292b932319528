"""
Transition modules -- mirror of src/models/BlocksUtils/transition_models.py.
"""
import torch.nn as nn

from .._common import print_
from .attention import TransformerEncoderBlock

TRANSITION_MODULES = [None, "", "None", "none", "TransformerBlock"]


def get_transition_module(model_name, **kwargs):
    """ None-like names -> identity; "TransformerBlock" -> one pre-LN encoder block over the slots
    (transition_models.py:19-53); extra kwargs are the block's num_heads / head_dim / mlp_size. """
    if model_name not in TRANSITION_MODULES:
        raise ValueError(f"Unknown transition module: {model_name = }... Use one of {TRANSITION_MODULES = }.")
    if model_name == "TransformerBlock":
        kwargs.pop("num_slots", None)
        module = TransformerEncoderBlock(embed_dim=kwargs.pop("slot_dim"), **kwargs)
    else:
        module = nn.Identity()
    print_("Transition Module:")
    print_(f"  --> model-name: {model_name}")
    for k, v in kwargs.items():
        print_(f"  --> {k}: {v}")
    return module
