"""Drop-in for `compare_emotions` of src/baselines/run_img_trans.py:361-386 -- the evaluation the reference runs on every
edited image right after the optimisation loop.  Both regressor predictions go through the native
`ValenceArousalLoss.predict_loss_metric` (tcgen05 forward); the bookkeeping is host code, as in the reference."""
from __future__ import annotations

import torch


def _table(title, names, rows):
    width = max(len(n) for n in names) + 2
    print(f"{title:<10s}" + "".join(f"{n:>{max(width, 10)}s}" for n in names))
    for label, vals in rows:
        print(f"{label:<10s}" + "".join(f"{v:>{max(width, 10)}.6f}" for v in vals))


def compare_emotions(loss, image, image_adapted, emotion_type_labels=None, stats=None):
    """Prints mean prediction before / after the edit, the mean change and the L1 reconstruction error, and appends the
    FIRST image's adapted prediction, its change and the error to `stats` (keys: lower-cased labels, delta_<label>,
    rec_error) -- the same quantities, in the same order, as the reference (:363-386)."""
    labels = list(emotion_type_labels) if emotion_type_labels is not None else ['Valence', 'Arousal']
    before = loss.predict_loss_metric(image)
    after = loss.predict_loss_metric(image_adapted)
    delta = after - before
    _table("type", labels, [("adapted", after.mean(0).tolist()), ("original", before.mean(0).tolist())])
    _table("delta", labels, [("mean", delta.mean(0).tolist())])
    rec_error = (image_adapted - image).abs().mean().item()            # == F.l1_loss(image_adapted, image)
    print(f"reconstruction error: {rec_error}\n")
    if stats is None:
        return
    stats["rec_error"].append(rec_error)
    for ix, name in enumerate(labels):
        key = name.lower()
        stats[key].append(after[0, ix].item())
        stats[f"delta_{key}"].append(delta[0, ix].item())
