"""
CPU-side checks (run with -m "not gpu"): C-ABI exports, state_dict layout of the module mirror against the
manifests recorded from the real reference, loud failure without CUDA, host-side logic.
"""
import copy
import ctypes
import json
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


def test_packed_operands_of_the_tensor_core_kernels():
    """ host-side weight packing (pure tensor code, no kernel): the row-merged compositing operand [3 filter rows][hi 16 | lo 16][64]
    of ps_dec_composite_tc reconstructs the 3x3 weight to 22 bits in the (kx, o) order the kernel's epilogue assumes, and the
    block-diagonal pair weight of a 32-channel ConvBlock convolves the two images of a pair independently """
    import torch.nn.functional as F
    from playslot_b200 import ops, configs
    torch.manual_seed(0)
    w = torch.randn(4, 64, 3, 3) * 0.05
    packed, scale = ops.pack_final_conv_tc(w)
    assert packed.shape == (3, 32, 64) and packed.dtype == torch.int16
    h = packed.view(torch.float16).float()
    rec = (h[:, 0:12] + h[:, 16:28] / 2048.0) / scale                                   # [ky][kx * 4 + o][cin]
    ref = w.permute(2, 3, 0, 1).reshape(3, 12, 64)
    assert (rec - ref).abs().max() <= 2.0 ** -21 * ref.abs().max()
    assert h[:, 12:16].abs().max() == 0 and h[:, 28:32].abs().max() == 0
    # block-diagonal pairing
    cfg = configs.experiment("BlockPush")
    savi, _ = configs.build_models(cfg, "cpu")
    blk = list(savi.encoder.encoder)[1]
    c = blk.conv
    co, ci = c.weight.shape[:2]
    wp = torch.zeros(2 * co, 2 * ci, 5, 5)
    wp[:co, :ci] = c.weight.detach(); wp[co:, ci:] = c.weight.detach()
    x = torch.randn(2, ci, 16, 16)
    y_pair = F.conv2d(x.reshape(1, 2 * ci, 16, 16), wp, torch.cat([c.bias, c.bias]).detach(), padding=2)
    y_ref = F.conv2d(x, c.weight.detach(), c.bias.detach(), padding=2)
    assert torch.allclose(y_pair.reshape(2, co, 16, 16), y_ref, atol=1e-5)


@pytest.mark.skipif(not os.path.isdir("/root/reference/experiments"), reason="reference tree not mounted")
@pytest.mark.parametrize("name", ["BlockPush", "ButtonPress", "GridShapes5Objs", "Sketchy"])
def test_restated_configs_equal_reference_json(name):
    import json
    from playslot_b200 import configs
    j = json.load(open(f"/root/reference/experiments/{name}/predictors/PlaySlot/experiment_params.json"))
    c = configs.experiment(name)
    assert c["model"] == j["model"] and c["predictor"] == j["predictor"]


def test_launcher_aliases_models_package():
    """ `models.*` resolves to the mirror after launch.install(); reference-only sub-modules stay reachable """
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from playslot_b200 import launch\n"
        "names = launch.install(reference_src=%r)\n"
        "from models.SAVi import SAVi\n"
        "from models.Predictors.LatentActionPredictors import SlotSingleActionPredictor\n"
        "from models.Predictors.predictor_wrappers import PredictorWrapper\n"
        "from models.Predictors.transformers import OCVPSeq\n"
        "from models.Predictors.ActionConditional_OCVP import ActionCondtionalOCVP\n"
        "import models\n"
        "assert SAVi.__module__ == 'playslot_b200.models.SAVi', SAVi.__module__\n"
        "assert models.SlotAttention.__module__.startswith('playslot_b200')\n"
        "print('ok', len(names))\n"
    ) % (ROOT, "/root/reference/src" if os.path.isdir("/root/reference/src") else None)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().startswith("ok")


def test_ocvp_and_action_conditional_state_dict_layout():
    """ OCVPSeq / ActionCondtionalOCVP parameter trees (a12, a13): key names as the reference builds them """
    from playslot_b200.models.Predictors.transformers import OCVPSeq
    from playslot_b200.models.Predictors.ActionConditional_OCVP import ActionCondtionalOCVP
    m = OCVPSeq(num_slots=8, slot_dim=128, input_buffer_size=14, token_dim=256, hidden_dim=1024, num_layers=4, n_heads=8, residual=False)
    keys = list(m.state_dict().keys())
    assert "mlp_in.weight" in keys and "mlp_out.bias" in keys and "pe.pe" in keys
    assert "transformer_encoders.3.time_encoder_block.attn.q.weight" in keys
    assert "transformer_encoders.0.object_encoder_block.mlp.mlp.2.bias" in keys
    assert m.state_dict()["pe.pe"].shape == (1, 14, 1, 256)
    a = ActionCondtionalOCVP(slot_dim=128, raw_action_dim=4, action_embed_dim=16, condition_mode="sum", autoregressive_dynamics=True,
                             DynamicsModel={"model_name": "SlotGPTDymamicsModel", "model_params": {
                                 "residual": False, "embed_dim": 256, "head_dim": 64, "num_heads": 8, "mlp_size": 1024,
                                 "num_layers": 4, "pos_enc_dropout": 0}})
    keys = list(a.state_dict().keys())
    assert "action_encoder.0.weight" in keys and "dynamics_model.pos_emb.pe" in keys
    assert "dynamics_model.variability_emb.weight" not in keys


def test_aux_predictor_manifests_match_reference():
    from playslot_b200.models.Predictors.transformers import OCVPSeq
    from playslot_b200.models.Predictors.predictor_wrappers import PredictorWrapper
    from playslot_b200.models.Predictors.ActionConditional_OCVP import ActionCondtionalOCVP
    aux = torch.load(os.path.join(GOLDEN_DIR, "aux_predictors.pt"), weights_only=False)
    fx = aux["ocvp_seq"]
    exp = {"prediction_params": fx["prediction_params"], "predictor": {"predictor_name": "OCVPSeq", "predictor_params": fx["cfg"]}}
    pred = PredictorWrapper(exp_params=exp, predictor=OCVPSeq(num_slots=8, slot_dim=128, input_buffer_size=14, **fx["cfg"]))
    assert list({k: list(v.shape) for k, v in pred.state_dict().items()}.items()) == list(fx["manifest"].items())
    fx = aux["action_cond"]
    m = ActionCondtionalOCVP(**copy.deepcopy(fx["cfg"]))
    assert list({k: list(v.shape) for k, v in m.state_dict().items()}.items()) == list(fx["manifest"].items())


def _reference_root():
    """ a tree holding the reference's src/lib/setup_model.py: the mounted reference or the copy made by oracle/install_ref.py """
    for root in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if os.path.isfile(os.path.join(root, "src", "lib", "setup_model.py")):
            return root
    return None


@pytest.mark.skipif(_reference_root() is None, reason="neither /root/reference nor oracle/_ref is present")
@pytest.mark.parametrize("exp", ["BlockPush", "ButtonPress", "GridShapes5Objs", "Sketchy"])
def test_reference_factory_builds_and_loads_the_mirror(exp, tmp_path):
    """ the drop-in boundary, driven by the REFERENCE's own code: lib/setup_model.py's setup_model / setup_predictor
    (:29-116) instantiate the mirror through launch.install(), save_checkpoint-style files round-trip through its
    load_checkpoint (:245-298, strict load_state_dict + the pos_emb.pe re-injection), and the state_dict layout equals the one
    the real reference modules have (manifests recorded in the golden fixtures) """
    import subprocess
    import sys
    root = _reference_root()
    code = f"""
import copy, json, os, sys, torch
sys.path.insert(0, {ROOT!r})
os.chdir({root!r})
sys.path.insert(0, os.path.join({root!r}, "src"))
from playslot_b200 import launch
launch.install(reference_src=os.path.join({root!r}, "src"))
from lib.setup_model import setup_model, setup_predictor, load_checkpoint
j = json.load(open(os.path.join({root!r}, "experiments", {exp!r}, "predictors", "PlaySlot", "experiment_params.json")))
savi = setup_model(copy.deepcopy(j["model"]))
pred = setup_predictor(copy.deepcopy(j))
assert type(savi).__module__ == "playslot_b200.models.SAVi", type(savi).__module__
assert type(pred).__module__.startswith("playslot_b200.models.Predictors"), type(pred).__module__
for name, m in (("savi", savi), ("pred", pred)):
    path = os.path.join({str(tmp_path)!r}, name + ".pth")
    sd = {{k: torch.randn_like(v) if v.is_floating_point() else v.clone() for k, v in m.state_dict().items()}}
    torch.save({{"model_state_dict": {{"module." + k: v for k, v in sd.items()}}}}, path)     # DataParallel-style prefix
    load_checkpoint(path, model=m, only_model=True)
    for k, v in m.state_dict().items():
        if not k.endswith("pos_emb.pe"):
            assert torch.equal(v, sd[k]), k
print("MANIFEST " + json.dumps({{"savi": {{k: list(v.shape) for k, v in savi.state_dict().items()}},
                                  "pred": {{k: list(v.shape) for k, v in pred.state_dict().items()}}}}))
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("MANIFEST ")][-1]
    got = json.loads(line[len("MANIFEST "):])
    case = {"BlockPush": "blockpush", "ButtonPress": "buttonpress", "GridShapes5Objs": "gridshapes5", "Sketchy": "sketchy_sharp"}[exp]
    fx = torch.load(os.path.join(GOLDEN_DIR, f"{case}.pt"), weights_only=False)
    assert list(got["savi"].items()) == [(k, list(v)) for k, v in fx["savi_manifest"].items()]
    assert list(got["pred"].items()) == [(k, list(v)) for k, v in fx["predictor_manifest"].items()]
