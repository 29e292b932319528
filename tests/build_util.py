""" Builds playslot_b200 modules from a golden fixture / config with the seeded synthetic weights. """
import copy

import torch

from synth import synth_state_dict


def make_savi(scfg, ssd=None, device="cpu"):
    from playslot_b200.models.SAVi import SAVi
    m = SAVi(**copy.deepcopy(scfg)).eval()
    if ssd is not None:
        m.load_state_dict(ssd, strict=True)
    return m.to(device)


def make_predictor(predictor_name, pcfg, psd=None, device="cpu"):
    from playslot_b200.models.Predictors.LatentActionPredictors import SlotSingleActionPredictor, SlotLatentPredictor
    cls = {"SlotSingleAction": SlotSingleActionPredictor, "SlotLatentPredictor": SlotLatentPredictor}[predictor_name]
    m = cls(**copy.deepcopy(pcfg)).eval()
    if psd is not None:
        psd = dict(psd)
        # like the reference's load_checkpoint (lib/setup_model.py:277-279): keep the module's own sinusoid buffer
        psd["dynamics_model.pos_emb.pe"] = m.state_dict()["dynamics_model.pos_emb.pe"]
        m.load_state_dict(psd, strict=True)
    return m.to(device)


def to_dev(sd, device):
    return {k: v.to(device) for k, v in sd.items()}


class FixedRandn:
    """ context manager that makes the next torch.randn calls return the given tensor (InvDyn noise / LearnedRandom) """

    def __init__(self, value):
        self.value = value

    def __enter__(self):
        self._orig = torch.randn
        val = self.value

        def fake(*a, **k):
            dev = k.get("device", None)
            return val.clone().to(dev) if dev is not None else val.clone()
        torch.randn = fake
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig
        return False


class ReplayDraws:
    """ context manager replaying recorded torch.randint / torch.rand results in call order (prior-sample fixtures made by
    oracle/make_golden.py::_RecordDraws): the CUDA path consumes exactly the random numbers the reference drew """

    def __init__(self, randint_draws, rand_draws):
        self.ri, self.r = list(randint_draws), list(rand_draws)

    def __enter__(self):
        self._ri, self._r = torch.randint, torch.rand

        def pop(queue, what, a, k):
            assert queue, f"the CUDA path drew more torch.{what} tensors than the reference"
            t = queue.pop(0)
            size = k.get("size", None)
            if size is not None:
                assert tuple(t.shape) == tuple(size), f"torch.{what}: shape {tuple(size)} vs recorded {tuple(t.shape)}"
            dev = k.get("device", None)
            return t.clone().to(dev) if dev is not None else t.clone()
        torch.randint = lambda *a, **k: pop(self.ri, "randint", a, k)
        torch.rand = lambda *a, **k: pop(self.r, "rand", a, k)
        return self

    def __exit__(self, *exc):
        torch.randint, torch.rand = self._ri, self._r
        return False
