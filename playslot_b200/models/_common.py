""" Helpers shared by the module mirrors: logging shim, derived-weight cache, CUDA-only guard. """
import torch

from .. import _lib


def print_(message, message_type="info"):
    """ Uses the reference's logger when the mirror runs inside its scripts (src/lib/logger.py:49-58). """
    try:
        from lib.logger import print_ as ref_print   # only importable inside the reference's src/ tree
        ref_print(message, message_type)
    except Exception:
        pass


class DerivedWeights:
    """
    Cache of tensors derived from parameters (packed / transposed / concatenated weights, weight-only
    position-embedding maps).  Rebuilt whenever any source tensor changes storage, device or version
    (optimizer steps and load_state_dict bump `_version`; graph-replayed optimiser steps bump ops.weight_generation()).
    """

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, sources, builder):
        from .. import ops
        key = (ops.weight_generation(),) + tuple((t.data_ptr(), t._version, str(t.device)) for t in sources)
        if key != self._key:
            with torch.no_grad():
                self._val = builder()
            self._key = key
        return self._val


def require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.PlaySlotKernelError(
            f"{what}: input lives on {t.device}. playslot_b200 modules execute hand-written sm_100a kernels "
            "and have no CPU / eager-PyTorch fallback; move the module and its inputs to a CUDA device.")


def no_grad_only(what, *tensors):
    """ Round-1 scope: forward (inference) kernels.  Fail loudly instead of silently dropping gradients. """
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            f"{what}: backward kernels are not implemented yet; call under torch.no_grad() "
            "(see DESIGN.md, 'Out of scope / next').")


def f32c(t):
    """ contiguous fp32 view/copy of a tensor (plumbing) """
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()
