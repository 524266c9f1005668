"""Base class of the image-emotion losses -- drop-in for src/baselines/losses/EmotionImageLoss.py:6-51.

It only fixes the attribute set the optimisation scripts rely on (`model`, `device`, `weight`, `is_minimized`, and the last
predictions in `fake_loss_metric` / `real_loss_metric`); the subclasses (ValenceArousalLoss) implement the rest.
"""
import torch.nn as nn


class EmotionImageLoss(nn.Module):
    def __init__(self, device, weight: float, is_minimized: bool = True):
        super().__init__()
        self.device, self.weight, self.is_minimized = device, weight, is_minimized
        self.model = None                    # set by the subclass: images [B,3,H,W] -> predictions [B,k]
        self.fake_loss_metric = None         # predictions of the last forward() for the edited images ...
        self.real_loss_metric = None         # ... and, when given, for the originals

    def forward(self, fake_imgs, real_imgs=None, condition=None):
        """Loss of a batch of edited images (subclass responsibility; the base returns None like the reference's stub)."""
        return None

    def predict_loss_metric(self, imgs):
        """The quantity the loss is computed from, for a batch of images."""
        return self.model(imgs)

    def get_random_condition_tensor(self, batch_size):
        """A random target for `batch_size` images (subclass responsibility)."""
        return None
