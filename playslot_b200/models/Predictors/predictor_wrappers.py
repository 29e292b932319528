"""
Autoregressive wrapper for the unconditional predictors -- mirror of
src/models/Predictors/predictor_wrappers.py (used for OCVPSeq / VanillaTransformer by lib/setup_model.py:83-101).
"""
import torch
import torch.nn as nn

from .._common import print_

__all__ = ["PredictorWrapper"]


class PredictorWrapper(nn.Module):
    """ Rolls a predictor over a sliding buffer of at most `input_buffer_size` frames (predictor_wrappers.py:13-122) """

    def __init__(self, exp_params, predictor):
        super().__init__()
        self.exp_params = exp_params
        self.predictor = predictor
        pp = exp_params["prediction_params"]
        self.num_context = pp["num_context"]
        self.num_preds = pp["num_preds"]
        self.teacher_force = pp["teacher_force"]
        self.video_length = pp["sample_length"]
        self.input_buffer_size = pp["input_buffer_size"]
        self.predictor_name = exp_params["predictor"]["predictor_name"]
        self.predictor_params = exp_params["predictor"]["predictor_params"]
        self._set_buffer_size()

    def forward(self, slot_history, get_pred_only=False, **kwargs):
        """ slot_history (B, T, S, D) -> (pred_slots (B, n, S, D), None); n = num_preds when get_pred_only else
        num_context + num_preds - 1.  Teacher forcing only in train mode (and only if configured). """
        self._is_teacher_force()
        start = self.num_context if get_pred_only else 1
        n = self.num_preds if get_pred_only else self.num_context + self.num_preds - 1
        buf = slot_history[:, :start].clone()
        preds = []
        for t in range(n):
            cur = self.predictor(buf)[:, -1]
            nxt = slot_history[:, start + t] if self.teacher_force else cur
            buf = self._update_buffer_size(torch.cat([buf, nxt.unsqueeze(1)], dim=1))
            preds.append(cur)
        return torch.stack(preds, dim=1), None

    def _is_teacher_force(self):
        self.teacher_force = self.exp_params["prediction_params"]["teacher_force"] if self._is_in_train_mode() else False

    def _is_in_train_mode(self):
        return self.predictor.training

    def _update_buffer_size(self, inputs):
        extra = inputs.shape[1] - self.input_buffer_size
        return inputs[:, extra:] if extra > 0 else inputs

    def _set_buffer_size(self):
        if self.input_buffer_size is None:
            print_(f" --> {self.predictor_name} buffer size is 'None'. Setting it as {self.num_context}")
            self.input_buffer_size = self.num_context
        if self.input_buffer_size < self.num_context:
            print_(f"  --> {self.predictor_name}'s {self.input_buffer_size = } is too small. Using {self.num_context} instead...")
            self.input_buffer_size = self.num_context
        else:
            print_(f"  --> Using buffer size {self.input_buffer_size}...")
