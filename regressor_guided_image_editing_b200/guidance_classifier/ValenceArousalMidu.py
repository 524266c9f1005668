"""Drop-in for src/guidance_classifier/ValenceArousalMidu.py:10-33."""
from .MiduClassifier import DEFAULT_PRECISION, MiduClassifier
from .guidance_scores import valence_arousal_score


class ValenceArousalMidu(MiduClassifier):
    def __init__(self, pipe, device: str, is_minimized: bool = True, ckp_path: str = None, is_sdxl: bool = False,
                 precision: str = DEFAULT_PRECISION):
        super().__init__(pipe, device, ckp_path, num_outputs=2, is_minimized=is_minimized, is_sdxl=is_sdxl,
                         precision=precision)

    @staticmethod
    def _calculate_score(x, m, device, is_minimized=True, reference_value=None):
        return valence_arousal_score(m(x), device, is_minimized, reference_value)
