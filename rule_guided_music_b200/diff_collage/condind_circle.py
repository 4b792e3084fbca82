"""CondIndCircle: the loop-able variant -- the latent is wrapped around by one overlap before windowing and the two
ends of the merged eps are averaged.  Mirror of diff_collage/condind_circle.py:7-84."""
import torch as th

from .condind_long import _window_eps
from .generic_sampler import SimpleWork
from .w_img import avg_merge_wimg, split_wimg


class CondIndCircle(SimpleWork):
    def __init__(self, shape, eps_scalar_t_fn, num_img, overlap_size=32):
        c, h, w = shape
        assert overlap_size == w // 2
        self.overlap_size = overlap_size
        self.num_img = num_img
        final_img_w = w * num_img - self.overlap_size * num_img
        super().__init__((c, h, final_img_w), self.get_eps_t_fn(eps_scalar_t_fn))

    def _wrap(self, in_x):
        return th.cat([in_x, in_x[:, :, :, : self.overlap_size]], dim=-1)

    def _unwrap(self, long_x, overlap_size):
        return th.cat([(long_x[:, :, :, :overlap_size] + long_x[:, :, :, -overlap_size:]) / 2.0,
                       long_x[:, :, :, overlap_size:-overlap_size]], dim=-1)

    def circle_split(self, in_x):
        return split_wimg(self._wrap(in_x), self.num_img, rtn_overlap=False)

    def circle_merge(self, xs, overlap_size=None):
        if overlap_size is None:
            overlap_size = self.overlap_size
        return self._unwrap(avg_merge_wimg(xs, overlap_size, n=self.num_img, is_avg=True), overlap_size)

    def get_eps_t_fn(self, eps_scalar_t_fn):
        def eps_t_fn(in_x, scalar_t, y=None):
            xs = split_wimg(self._wrap(in_x), self.num_img, rtn_overlap=False)
            whole = _window_eps(eps_scalar_t_fn, xs, scalar_t, y, self.num_img, self.overlap_size)
            long_eps = avg_merge_wimg(whole, self.overlap_size, n=self.num_img, is_avg=False)
            return self._unwrap(long_eps, self.overlap_size)

        return eps_t_fn
