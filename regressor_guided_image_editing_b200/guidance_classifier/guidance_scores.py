"""Drop-ins for src/guidance_classifier/guidance_scores.py.  `valence_arousal_score` (:4-22): squared distance of the predicted
(valence, arousal) pairs to a target, summed over the batch.  Without `reference_value` the target is (0.5, 0.0) when the
score is minimised and (1, 1) otherwise.  The difference / squeeze / square / sum chain is the reference's, so autograd hands
the native head's backward the same d(score)/d(prediction).  `arousal_score` (:25-46) and `valence_score` (:49-73) are the
single-axis forms: per-sample squared distance (no batch sum) of column 1 / column 0 to 0 / 0.5 (minimised) or 1."""
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


def _axis_score(pred, column, device, neutral, is_minimized, reference_value):
    target = reference_value
    if target is None:
        target = torch.full((pred.size(0),), neutral if is_minimized else 1.0).to(device)
    value = pred[:, column] if pred.size(1) > 1 else pred
    error = (target - value).squeeze().squeeze()
    return error * error


def arousal_score(predicted_arousal, device, is_minimized=True, reference_value=None):
    return _axis_score(predicted_arousal, 1, device, 0.0, is_minimized, reference_value)


def valence_score(predicted_valence, device, is_minimized=True, reference_value=None):
    return _axis_score(predicted_valence, 0, device, 0.5, is_minimized, reference_value)
