"""Drop-in for src/baselines/losses/EmotionImageLoss.py:6-51 (base class: attributes only)."""
import torch.nn as nn


class EmotionImageLoss(nn.Module):
    def __init__(self, device, weight: float, is_minimized: bool = True):
        super().__init__()
        self.model = None
        self.device = device
        self.is_minimized = is_minimized
        self.weight = weight
        self.fake_loss_metric = None
        self.real_loss_metric = None

    def forward(self, fake_imgs, real_imgs=None, condition=None):
        pass

    def predict_loss_metric(self, imgs):
        return self.model(imgs)

    def get_random_condition_tensor(self, batch_size):
        pass
