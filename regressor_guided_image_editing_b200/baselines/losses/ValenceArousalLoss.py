"""Drop-in for src/baselines/losses/ValenceArousalLoss.py: same constructor arguments, methods and attributes
(`model`, `fake_loss_metric`, `real_loss_metric`, `output_ixs`, `get_error`, `is_minimized`, `weight`).

`self.model` is the native regressor (this package's baselines/models/EmotionPredictionModel.py or EmoNet.py).  The loss
itself is a handful of [B, 2]-sized torch expressions; they are written so that autograd sees the same operations, in the
same order, as the reference (target - predicted, squared, weighted, mean), and so that the default targets and the
random-condition draws are bit-identical to it.  Reference lines are cited per method.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from ... import _lib
from ..models.EmotionPredictionModel import DEFAULT_PRECISION, load_model_eval
from .EmotionImageLoss import EmotionImageLoss

# default target of each dimension when no target is given: (minimised, maximised)   -- reference :75-112
_DEFAULT_TARGET = {"valence": (0.5, 1.0), "arousal": (0.0, 1.0)}
# which output columns a loss mode reads and which error it uses                     -- reference :41-57
_MODES = {"valence": ([0], "get_valence_error"), "arousal": ([1], "get_arousal_error"),
          "va": ([0, 1], "get_valence_arousal_error")}


def _head_config(checkpoint_name: str) -> Tuple[int, Optional[torch.nn.Module]]:
    """(number of regressor outputs, output activation) encoded in the checkpoint's file name (reference :35-47)."""
    num_classes, activation = 4, torch.nn.Sigmoid()
    if "no_sigmoid" in checkpoint_name:
        activation = None
    if "mse" in checkpoint_name:
        num_classes, activation = 2, None
    if "arousal_nll" in checkpoint_name:
        num_classes = 2
    return num_classes, activation


class ValenceArousalLoss(EmotionImageLoss):
    def __init__(self, path_to_model, device, weight: float, is_minimized: bool = True, loss: str = "va",
                 is_input_range_0_1=True, input_size=480, crop_size=448, requires_grad=False,
                 precision: str = DEFAULT_PRECISION):
        super().__init__(device, weight, is_minimized)
        name = path_to_model if isinstance(path_to_model, str) else ""
        if "EmoNet" in name:                                   # :31-33  EmoNet ten-crop regressor
            from ..models.EmoNet import load_model_eval as load_emonet
            self.model = load_emonet(path_to_model, normalize=is_input_range_0_1, requires_grad=requires_grad,
                                     precision=precision)
        else:
            num_classes, activation = _head_config(name)
            self.model = load_model_eval(path_to_model, num_classes, normalize=is_input_range_0_1,
                                         activation_function=activation, input_size=input_size, crop_size=crop_size,
                                         is_ten_crop=True, requires_grad=requires_grad, precision=precision)
        columns, error_name = _MODES.get(loss, _MODES["va"])
        self.output_ixs = list(columns)
        self.get_error = getattr(self, error_name)

    # ------------------------------------------------------------------------------------------------ forward (:59-73)
    def forward(self, fake_imgs: Tensor, real_imgs: Tensor = None, target: Tensor = None) -> Tensor:
        self.fake_loss_metric = self._metric(fake_imgs)
        if real_imgs is not None:
            self.real_loss_metric = self._metric(real_imgs)
        return torch.mean(self.weight * self.get_error(self.fake_loss_metric, target))

    def _metric(self, imgs: Tensor) -> Tensor:
        return self.model(imgs)[:, self.output_ixs]

    # ------------------------------------------------------------------------------------------------ errors (:75-129)
    def _squared_error(self, predicted: Tensor, target: Optional[Tensor], dimension: str) -> Tensor:
        if target is None:
            value = _DEFAULT_TARGET[dimension][0 if self.is_minimized else 1]
            target = (value * torch.ones(predicted.size(0))).to(predicted.device)
        error = target - predicted
        return error * error

    def get_valence_error(self, predicted, target):
        return self._squared_error(predicted, target, "valence")

    def get_arousal_error(self, predicted, target):
        return self._squared_error(predicted, target, "arousal")

    def get_valence_arousal_error(self, predicted, target):
        t_val, t_ar = (None, None) if target is None else (target[:, 0], target[:, 1])
        return self.get_valence_error(predicted[:, 0], t_val) + self.get_arousal_error(predicted[:, 1], t_ar)

    # ------------------------------------------------------------------------------------------------ helpers (:131-147)
    def predict_loss_metric(self, imgs: Tensor) -> Tensor:
        with torch.no_grad():
            return self._metric(imgs)

    def get_random_condition_tensor(self, batch_size):
        k = len(self.output_ixs)                  # one draw of batch_size * k values, as the reference, then reshaped
        return torch.randint(0, 2, (batch_size * k,)).view(batch_size, k).to(self.device)
