"""DiffCollage long-sequence score composition for the sampling path -- mirror of the reference's diff_collage package
(condind_long.py, condind_circle.py, w_img.py, generic_sampler.SimpleWork).  Only the workers sample_rule.py uses."""
from .condind_circle import CondIndCircle
from .condind_long import CondIndSimple
from .generic_sampler import SimpleWork
from .w_img import avg_merge_wimg, split_wimg

__all__ = ["CondIndSimple", "CondIndCircle", "SimpleWork", "split_wimg", "avg_merge_wimg"]
