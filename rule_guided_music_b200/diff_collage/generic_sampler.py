"""SimpleWork base of the DiffCollage workers -- mirror of diff_collage/generic_sampler.py:17-20 (the EDM-style
`generic_sampler` itself is not used by sample_rule.py and is not part of the path)."""


class SimpleWork:
    def __init__(self, shape, eps_scalar_t_fn):
        self.shape = shape
        self.eps_scalar_t_fn = eps_scalar_t_fn
