"""Drop-in for src/guidance_classifier/guidance_scores.py:4-22 (`valence_arousal_score`): a [B,2]-sized expression kept
literally as in the reference so autograd produces d(score)/d(pred) for the native head's backward."""
import torch


def valence_arousal_score(predicted_va, device, is_minimized=True, reference_value=None):
    if reference_value is not None:
        target = reference_value
    else:
        target = torch.ones(predicted_va.size(0), 2).to(device)
        if is_minimized:
            target[:, 0] = 0.5 * target[:, 0]
            target[:, 1] = 0.0 * target[:, 1]
    error = (target - predicted_va).squeeze().squeeze()
    return torch.sum(error * error)
