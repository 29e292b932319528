"""
CPU-side checks (run with -m "not gpu"): C-ABI exports, state_dict layout of the module mirror against the
manifests recorded from the real reference, loud failure without CUDA, host-side logic.
"""
import copy
import ctypes
import os
import re

import pytest
import torch

from golden_util import GOLDEN_DIR, PIPELINE_CASES
from build_util import make_savi, make_predictor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    h = open(os.path.join(ROOT, "include", "playslot_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    from playslot_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/playslot_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert _lib.load().ps_version() == 100


@pytest.mark.parametrize("case", PIPELINE_CASES)
def test_state_dict_layout_matches_reference(case):
    fx = torch.load(os.path.join(GOLDEN_DIR, f"{case}.pt"), weights_only=False)
    savi = make_savi(fx["savi_cfg"])
    pred = make_predictor(fx["predictor_name"], fx["predictor_cfg"])
    sm = {k: list(v.shape) for k, v in savi.state_dict().items()}
    pm = {k: list(v.shape) for k, v in pred.state_dict().items()}
    assert list(sm.items()) == list(fx["savi_manifest"].items())          # same keys, order and shapes
    assert list(pm.items()) == list(fx["predictor_manifest"].items())
    for v in list(savi.state_dict().values()) + list(pred.state_dict().values()):
        assert v.dtype == torch.float32


def test_attribute_surface():
    fx = torch.load(os.path.join(GOLDEN_DIR, "blockpush.pt"), weights_only=False)
    savi = make_savi(fx["savi_cfg"])
    pred = make_predictor(fx["predictor_name"], fx["predictor_cfg"])
    for a in ("num_slots", "slot_dim", "slot_attention", "transition_module", "initializer", "encoder", "decoder",
              "encode", "decode", "broadcast", "decoder_resolution"):
        assert hasattr(savi, a), a
    for a in ("latent_action", "dynamics_model", "use_variability", "num_actions", "action_dim",
              "forward_multiple_samples", "autoregressive_inference", "get_actions_emb", "get_random_actions"):
        assert hasattr(pred, a), a
    for a in ("quantizer", "compute_actions", "get_action", "decompose_action_latent", "mean_fc", "variance_fc"):
        assert hasattr(pred.latent_action, a), a
    for a in ("enforce_window", "forward_single", "forward_autoregressive", "context_length", "pos_emb"):
        assert hasattr(pred.dynamics_model, a), a
    assert pred.latent_action.quantizer.get_codewords().shape == (8, 16)
    assert sum(p.numel() for p in pred.parameters()) > 7_000_000


def test_no_cpu_fallback():
    from playslot_b200._lib import PlaySlotKernelError
    fx = torch.load(os.path.join(GOLDEN_DIR, "buttonpress.pt"), weights_only=False)
    savi = make_savi(fx["savi_cfg"])
    with torch.no_grad():
        with pytest.raises(PlaySlotKernelError):
            savi(torch.rand(1, 2, 3, 64, 64), num_imgs=2, decode=False)
        with pytest.raises(PlaySlotKernelError):
            savi.decode(torch.rand(1, 4, 128))
        with pytest.raises(PlaySlotKernelError):
            savi.slot_attention(torch.rand(1, 64, 128), torch.rand(1, 4, 128))


def test_mask_to_slots_per_frame():
    from playslot_b200.models.BlocksUtils.attention import _mask_to_slots_per_frame
    from playslot_b200.models.BlocksUtils.model_utils import build_slot_causal_mask
    m = build_slot_causal_mask(seq_len=8, num_slots=6, device="cpu")
    assert _mask_to_slots_per_frame(m, 48) == 6
    assert _mask_to_slots_per_frame(m, 18) == 6           # cropped mask, 3 frames
    assert _mask_to_slots_per_frame(None, 10) == 0
    with pytest.raises(NotImplementedError):
        _mask_to_slots_per_frame(torch.eye(6), 6) if False else _mask_to_slots_per_frame(torch.rand(6, 6) > 0.5, 6)


def test_enforce_window_and_pe_buffer():
    from oracle import playslot_oracle as O
    fx = torch.load(os.path.join(GOLDEN_DIR, "sketchy_sharp.pt"), weights_only=False)
    pred = make_predictor(fx["predictor_name"], fx["predictor_cfg"])
    dm = pred.dynamics_model
    assert dm.context_length == 10
    s, a = torch.zeros(1, 13, 8, 128), torch.zeros(1, 13, 8, 16)
    s2, a2, v2 = dm.enforce_window(s, a, None)
    assert s2.shape[1] == 10 and a2.shape[1] == 10 and v2 is None
    assert torch.equal(dm.pos_emb.pe[0, :, 0], O.sinusoid_pe(10, 256))


def test_border_class_tap_sums_match_direct_conv():
    """ the collapsed first decoder conv (SAVi._decoder_constants) is exact: checked on the CPU with torch ops """
    import torch.nn.functional as F
    torch.manual_seed(0)
    w = torch.randn(6, 5, 5, 5)
    slot = torch.randn(5)
    H = W = 9
    full = F.conv2d(slot.view(1, 5, 1, 1).expand(1, 5, H, W), w, padding=2)[0]      # (6, H, W)
    k, p = 5, 2

    def taps(c):
        return slice(p - c, k) if c < p else (slice(0, k - (c - p)) if c > p else slice(0, k))

    def cls(v, n):
        return v if v < p else (2 * p - (n - 1 - v) if v > n - 1 - p else p)
    for y in range(H):
        for x in range(W):
            ws = w[:, :, taps(cls(y, H)), taps(cls(x, W))].sum(dim=(2, 3))
            assert torch.allclose(ws @ slot, full[:, y, x], atol=1e-4)


@pytest.mark.skipif(not os.path.isdir("/root/reference/experiments"), reason="reference tree not mounted")
@pytest.mark.parametrize("name", ["BlockPush", "ButtonPress", "GridShapes5Objs", "Sketchy"])
def test_restated_configs_equal_reference_json(name):
    import json
    from playslot_b200 import configs
    j = json.load(open(f"/root/reference/experiments/{name}/predictors/PlaySlot/experiment_params.json"))
    c = configs.experiment(name)
    assert c["model"] == j["model"] and c["predictor"] == j["predictor"]
