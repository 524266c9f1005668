"""Interface of the diffusion-guidance classifiers -- drop-in for src/guidance_classifier/GuidanceClassifier.py.

A guidance classifier scores noisy latents at a timestep; the pipelines differentiate `forward` with respect to the latents
(pipelines/InversionResamplingStableDiffusionPipeline.py:126-142).  MiduClassifier implements the three entry points.
"""
import torch
import torch.nn as nn


class GuidanceClassifier(nn.Module):
    def __init__(self, device: str):
        super().__init__()
        self.device = torch.device(device)
        self.model = None                    # the trainable head (state_dict-compatible with the reference checkpoints)

    def forward(self, latents, t, prompt_embeds=None):
        """Scalar guidance score of `latents` at timestep `t` (differentiable w.r.t. the latents)."""
        return None

    def get_loss(self, latents, label, t, prompts):
        """Training loss of the head against `label`."""
        return None

    def predict_score(self, latents, t, prompts):
        """The head's prediction, without a graph."""
        return None
