"""`valence_arousal_score` -- drop-in for src/guidance_classifier/guidance_scores.py:4-22: squared distance of the predicted
(valence, arousal) pairs to a target, summed over the batch.  Without `reference_value` the target is (0.5, 0.0) when the
score is minimised and (1, 1) otherwise.  The difference / squeeze / square / sum chain is the reference's, so autograd hands
the native head's backward the same d(score)/d(prediction)."""
import torch

_MINIMISED_TARGET = (0.5, 0.0)


def valence_arousal_score(predicted_va, device, is_minimized=True, reference_value=None):
    target = reference_value
    if target is None:
        rows = predicted_va.size(0)
        target = torch.ones(rows, 2)
        if is_minimized:
            target = torch.tensor([_MINIMISED_TARGET]).repeat(rows, 1)
        target = target.to(device)
    error = (target - predicted_va).squeeze().squeeze()
    return torch.sum(error * error)
