"""
Vector quantisation of latent actions -- mirror of src/models/BlocksUtils/quantization.py
(VectorQuantizer, EmaVectorQuantizer, EmaCodebookMeter; same state_dict: `embeddings.0.weight`,
`ema_count`, `ema_weight`).  The nearest-codeword search runs in csrc/invdyn.cu (fp32, first-index ties).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from .._common import print_, require_cuda, no_grad_only, f32c


def L2_efficient(x, y):
    """ pairwise squared distances |x|^2 - 2 x y + |y|^2 (quantization.py:18-20); kept for API parity """
    return (x.pow(2).sum(1, keepdim=True) - 2 * x @ y + y.pow(2).sum(0, keepdim=True))


class EmaCodebookMeter:
    """ EMA estimate of codeword usage (quantization.py:24-83); python-side statistics, not in the state_dict """

    def __init__(self, codebook_size, ema_alpha=0.05):
        if ema_alpha < 0 or ema_alpha > 1.0:
            raise ValueError(f"{ema_alpha = } must be in [0, 1]...")
        self.codebook_size = codebook_size
        self.ema_alpha = ema_alpha
        self.reset()

    def reset(self):
        self.bins = nn.Parameter(torch.ones(self.codebook_size) / self.codebook_size, requires_grad=False)
        self._iters = torch.zeros((), dtype=torch.float32)      # device-side counter: updates stay sync-free (CUDA-graph safe)

    @property
    def iters(self):
        return int(self._iters.item())

    @iters.setter
    def iters(self, value):
        self._iters = torch.full((), float(value), dtype=torch.float32, device=self._iters.device)

    def bincount(self, val, weights=None):
        norm = val.shape[0]
        flat = val.reshape(-1)
        if weights is None:
            count = F.one_hot(flat, num_classes=self.codebook_size).sum(0).to(torch.float32)
        else:
            count = torch.zeros(self.codebook_size, device=flat.device, dtype=torch.float32)
            count.index_add_(0, flat, weights.reshape(-1).to(torch.float32))
        self._iters = self._iters.to(flat.device) if self._iters.device != flat.device else self._iters
        self._iters += 1
        return count.detach() / norm

    def load(self, bins):
        self.bins = torch.load(bins)

    def update(self, val, weights=None):
        """ same statistics as the reference (which keeps `bins` on the CPU and rebinds it every call); here the histogram
        follows the indices' device and is updated in place, without a host synchronisation """
        if self.bins.device != val.device:
            self.bins = nn.Parameter(self.bins.data.to(val.device), requires_grad=False)
        count = self.bincount(val, weights=weights)
        alpha = torch.clamp(1.0 / (self._iters + 1.0), min=self.ema_alpha)
        self.bins.data.mul_(1. - alpha).add_(alpha * count)

    def get_hist(self):
        return self.bins


class VectorQuantizer(nn.Module):
    """ Nearest-codeword quantiser with straight-through value (quantization.py:177-372); nbooks == 1 only """

    def __init__(self, num_embs, emb_dim, nbooks=1, balance=False):
        if nbooks != 1:
            raise NotImplementedError("So far we only support 'nbooks' = 1")
        if balance is not False:
            raise NotImplementedError("So far we only support 'balance' = False")
        super().__init__()
        print_("Instanciating VectorQuantizer")
        self.num_embs = num_embs
        self.emb_dim = emb_dim
        self.nbooks = nbooks
        self.balance = balance
        self.num_embs_i = num_embs // nbooks
        self.chunk_dims = [emb_dim]
        self.embeddings = torch.nn.ModuleDict({"0": nn.Embedding(self.num_embs_i, emb_dim)})
        self.embeddings["0"].weight.data.uniform_(-1.0 / self.num_embs_i, 1.0 / self.num_embs_i)
        self.trackers = {0: EmaCodebookMeter(self.num_embs_i)}

    def get_codewords(self, i='0'):
        return self.embeddings[i].weight

    def get_hist(self, i='0'):
        return self.trackers[int(i)].get_hist()

    def get_state(self):
        return {i: self.get_hist(i).cpu().data.numpy() for i in self.trackers.keys()}

    def load_state(self, bins):
        for i, b in bins.items():
            self.trackers[i].load(b)

    def reset(self, i):
        """ as upstream (quantization.py:247-250), including its quirk: the tracker DICT is replaced by one meter sized
        by the chunk dimension """
        for i in self.trackers.keys():
            self.trackers = EmaCodebookMeter(self.chunk_dims[int(i)])

    def track_assigment(self, emb_ind, i):
        self.trackers[i].update(emb_ind)

    def lookup(self, z):
        """ kernel path: (proto = z + (z_q - z), exact z_q, idx (..., 1), per-vector loss) """
        require_cuda(z, "VectorQuantizer")
        no_grad_only("VectorQuantizer", z)
        out = ops.vq_lookup(f32c(z), f32c(self.embeddings["0"].weight.detach()))
        if self.training:
            self.track_assigment(out[2].reshape(-1, 1).detach(), 0)
        return out

    def forward_one(self, z, i, weights=None):
        _, zq, idx, _ = self.lookup(z)
        return zq, idx

    def forward(self, z, **kwargs):
        assert z.size(-1) == self.emb_dim, f"{z.shape = } must have {self.emb_dim = }"
        proto, _, idx, loss = self.lookup(z)
        return proto, {"quant_loss": loss, "commit_loss": loss.clone()}, idx

    def get_codebook_entry(self, indices):
        return self.embeddings["0"](indices.squeeze(-1))

    def get_variability(self, z, action_embs=None, action_idxs=None):
        """ variability = z - codeword (re-looked-up unless both action_embs and action_idxs are given) """
        shape, A = z.shape[:-1], z.shape[-1]
        if action_embs is None or action_idxs is None:
            action_embs, action_idxs = self.forward_one(z, i=0, weights=None)
        v = (z.reshape(-1, A) - action_embs.reshape(-1, A)).reshape(*shape, A)
        scores = F.one_hot(action_idxs, num_classes=self.num_embs)
        return v, scores


class EmaVectorQuantizer(nn.Module):
    """ VectorQuantizer whose codebook follows an exponential moving average in train mode
    (quantization.py:87-173); buffers `ema_count` (K,), `ema_weight` (K, A). """

    def __init__(self, num_embs, emb_dim, nbooks=1, balance=False, ema_alpha=0.05):
        assert 0 <= ema_alpha <= 1.0, f"{ema_alpha = } must be in [0, 1]"
        if nbooks != 1:
            raise NotImplementedError(f"Only nbooks=1 is implemented, but {nbooks = }")
        super().__init__()
        self.num_embs = num_embs
        self.emb_dim = emb_dim
        self.ema_alpha = ema_alpha
        self.epsilon = 1e-6
        self.update = True
        self.sync_stats = True        # all-reduce the EMA statistics when torch.distributed is initialised
        self.vq = VectorQuantizer(num_embs=num_embs, emb_dim=emb_dim, nbooks=nbooks, balance=balance).requires_grad_(False)
        self.register_buffer('ema_count', torch.zeros(self.num_embs))
        self.register_buffer('ema_weight', torch.empty((self.num_embs, self.emb_dim)).uniform_(-1 / num_embs, 1 / num_embs))

    @torch.no_grad()
    def ema_update(self, z, idx):
        """ quantization.py:137-150: counts and sums of the assigned vectors, Laplace-smoothed counts """
        B = z.shape[0]
        zf = z.flatten(0, -2)
        enc = F.one_hot(idx.reshape(-1), num_classes=self.num_embs).to(zf.dtype)
        enc_sum, dw = enc.sum(0), enc.t() @ zf
        if self.sync_stats and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            # batch-sharded training: the statistics of the GLOBAL batch make every rank's codebook identical to the
            # single-GPU run on the concatenated batch (one all-reduce of K + K*A + 1 floats)
            packed = torch.cat([enc_sum, dw.flatten(), torch.full((1,), float(B), device=zf.device, dtype=zf.dtype)])
            torch.distributed.all_reduce(packed)
            enc_sum, dw, B = packed[:self.num_embs], packed[self.num_embs:-1].view_as(dw), packed[-1]   # B stays on the device
        # same arithmetic as the reference; written IN PLACE (the reference rebinds the buffers) so that a training step
        # captured in a CUDA graph (train.GraphedTrainStep) keeps advancing the same tensors on every replay
        cnt = self.ema_alpha * self.ema_count + (1 - self.ema_alpha) * enc_sum
        self.ema_count.copy_((cnt + self.epsilon) / (B + self.num_embs * self.epsilon) * B)
        self.ema_weight.copy_(self.ema_alpha * self.ema_weight + (1 - self.ema_alpha) * dw)
        self.vq.embeddings['0'].weight.copy_(self.ema_weight / (self.ema_count.unsqueeze(1) + 1e-8))    # bumps _version

    def forward(self, z):
        z_q, losses, idx = self.vq(z)
        if self.training and self.update:
            self.ema_update(z, idx)
        return z_q, losses, idx

    def get_codebook_entry(self, indices):
        return self.vq.get_codebook_entry(indices)

    def get_hist(self):
        return self.ema_count.cpu().detach()

    def get_codewords(self, i='0'):
        return self.vq.get_codewords(i=i)

    def get_variability(self, z, action_embs=None, action_idxs=None):
        return self.vq.get_variability(z=z, action_embs=action_embs, action_idxs=action_idxs)
