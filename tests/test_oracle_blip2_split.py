"""The image/text split of the BLIP-2 oracle (used to score many frames x prompts with one ViT pass per frame) is the
same computation as Blip2ForImageTextRetrieval.forward's ITC branch."""
import numpy as np

from oracle.blip2_oracle import Blip2Oracle
from vlfm_b200.utils.synthetic import make_rgb
from vlfm_b200.vlm.blip2_config import TINY, random_state_dict


def test_split_equals_forward():
    for outl in (False, True):
        orc = Blip2Oracle(TINY, random_state_dict(TINY, 1, outliers=outl))
        rng = np.random.default_rng(0)
        for ids in ([3, 14, 15, 9, 2], [7, 8], [1, 50, 60, 70, 80, 90, 2]):
            img = make_rgb(rng, 120, 160)
            a = orc.cosine(img, ids)
            b = orc.cosine_from(orc.image_features(img), orc.text_feature(ids))
            assert abs(a - b) < 1e-6, (a, b)
