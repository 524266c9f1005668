"""Drop-in for src/baselines/losses/ValenceArousalLoss.py (same constructor, methods and attributes).

The regressor behind `self.model` is the native one (baselines/models/EmotionPredictionModel.py in this package); the
few [B,2]-sized tensor expressions below are kept literally as in the reference (:59-138) so autograd sees the same
graph down to the logits.
"""
from __future__ import annotations

import torch
from torch import Tensor

from ... import _lib
from ..models.EmotionPredictionModel import DEFAULT_PRECISION, load_model_eval
from .EmotionImageLoss import EmotionImageLoss


class ValenceArousalLoss(EmotionImageLoss):
    def __init__(self, path_to_model, device, weight: float, is_minimized: bool = True, loss: str = "va",
                 is_input_range_0_1=True, input_size=480, crop_size=448, requires_grad=False,
                 precision: str = DEFAULT_PRECISION):
        super().__init__(device, weight, is_minimized)
        name = path_to_model if isinstance(path_to_model, str) else ""
        if "EmoNet" in name:                                                                             # :31-33
            from ..models.EmoNet import load_model_eval as load_model_eval_emo_net
            self.model = load_model_eval_emo_net(path_to_model, normalize=is_input_range_0_1,
                                                 requires_grad=requires_grad, precision=precision)
        else:
            num_classes = 4
            activ_func = torch.nn.Sigmoid()
            if "no_sigmoid" in name:
                activ_func = None
            if "mse" in name:
                num_classes = 2
                activ_func = None
            if "arousal_nll" in name:
                num_classes = 2
            self.model = load_model_eval(path_to_model, num_classes, normalize=is_input_range_0_1,
                                         activation_function=activ_func, input_size=input_size, crop_size=crop_size,
                                         is_ten_crop=True, requires_grad=requires_grad, precision=precision)
        if loss == "valence":
            self.get_error = self.get_valence_error
            self.output_ixs = [0]
        elif loss == "arousal":
            self.get_error = self.get_arousal_error
            self.output_ixs = [1]
        else:
            self.get_error = self.get_valence_arousal_error
            self.output_ixs = [0, 1]

    def forward(self, fake_imgs: Tensor, real_imgs: Tensor = None, target: Tensor = None) -> Tensor:      # :59-73
        self.fake_loss_metric = self.model(fake_imgs)[:, self.output_ixs]
        if real_imgs is not None:
            self.real_loss_metric = self.model(real_imgs)[:, self.output_ixs]
        return torch.mean(self.weight * self.get_error(self.fake_loss_metric, target))

    def get_valence_error(self, predicted, target):                                                       # :75-93
        if target is None:
            if self.is_minimized:
                target = 0.5 * torch.ones(predicted.size(0)).to(predicted.device)
            else:
                target = torch.ones(predicted.size(0)).to(predicted.device)
        error = target - predicted
        return error * error

    def get_arousal_error(self, predicted, target):                                                       # :95-112
        if target is None:
            if self.is_minimized:
                target = torch.zeros(predicted.size(0)).to(predicted.device)
            else:
                target = torch.ones(predicted.size(0)).to(predicted.device)
        error = target - predicted
        return error * error

    def get_valence_arousal_error(self, predicted, target):                                               # :114-129
        if target is not None:
            val_error = self.get_valence_error(predicted[:, 0], target[:, 0])
            ar_error = self.get_arousal_error(predicted[:, 1], target[:, 1])
        else:
            val_error = self.get_valence_error(predicted[:, 0], None)
            ar_error = self.get_arousal_error(predicted[:, 1], None)
        return val_error + ar_error

    def predict_loss_metric(self, imgs: Tensor) -> Tensor:                                                # :131-138
        with torch.no_grad():
            return self.model(imgs)[:, self.output_ixs]

    def get_random_condition_tensor(self, batch_size):                                                    # :140-147
        dim_space = len(self.output_ixs)
        return torch.randint(0, 2, (batch_size * dim_space,)).reshape(batch_size, dim_space).to(self.device)
