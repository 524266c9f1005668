"""Drop-in for src/baselines/models/utilities/MeanReplicatedCrops.py:18-27 (a view + mean over [B*reps, k] logits)."""
import torch.nn as nn


class MeanReplicatedCrops(nn.Module):
    def __init__(self, num_replications=10):
        super().__init__()
        self.num_replications = num_replications

    def forward(self, x):
        b, w = x.size()
        return x.view(b // self.num_replications, self.num_replications, w).mean(dim=1)
