"""Drop-in for src/guidance_classifier/GuidanceClassifier.py (base class)."""
import torch
import torch.nn as nn


class GuidanceClassifier(nn.Module):
    def __init__(self, device: str):
        super().__init__()
        self.device = torch.device(device)
        self.model = None

    def forward(self, latents, t, prompt_embeds=None):
        pass

    def get_loss(self, latents, label, t, prompts):
        pass

    def predict_score(self, latents, t, prompts):
        pass
