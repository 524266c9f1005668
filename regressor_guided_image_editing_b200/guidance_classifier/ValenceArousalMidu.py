"""Valence/arousal guidance head on the UNet mid-block features -- drop-in for src/guidance_classifier/ValenceArousalMidu.py:10-33:
MiduClassifier with two outputs whose score is `valence_arousal_score`."""
from .MiduClassifier import DEFAULT_PRECISION, MiduClassifier
from .guidance_scores import valence_arousal_score


class ValenceArousalMidu(MiduClassifier):
    NUM_OUTPUTS = 2                          # (valence, arousal)

    def __init__(self, pipe, device: str, is_minimized: bool = True, ckp_path: str = None, is_sdxl: bool = False,
                 precision: str = DEFAULT_PRECISION):
        super().__init__(pipe, device, ckp_path, num_outputs=self.NUM_OUTPUTS, is_minimized=is_minimized, is_sdxl=is_sdxl,
                         precision=precision)

    @staticmethod
    def _calculate_score(x, m, device, is_minimized=True, reference_value=None):
        prediction = m(x)
        return valence_arousal_score(prediction, device, is_minimized, reference_value)
