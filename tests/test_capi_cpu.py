"""CPU-side checks of the round-2 boundary additions and host fixes (no GPU): the coefficient tables a C host gets from
rgm_coeff_tables against numpy's (= the reference's, pinned in test_host_cpu.py), the step-graph cache's signatures /
strong references / LRU bound, and the `--image_size 128 16` command line of the reference's scripts."""
import argparse
import ctypes
import gc
from functools import partial
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from rule_guided_music_b200 import _lib
from rule_guided_music_b200.guided_diffusion import gaussian_diffusion as gd
from rule_guided_music_b200.guided_diffusion import script_util
from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion


@pytest.mark.parametrize("resp", ["", "256", "ddim50", "10,15,20"])
def test_coeff_tables_equal_numpy(resp):
    """rgm_coeff_tables_host == float64 numpy tables of GaussianDiffusion.__init__ cast to fp32, bit for bit."""
    d = create_diffusion(timestep_respacing=resp)
    T = d.num_timesteps
    betas = np.ascontiguousarray(d.betas, dtype=np.float64)
    out = np.zeros((len(_lib.COEF_ROWS), T), dtype=np.float32)
    _lib.call("rgm_coeff_tables_host", betas.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), T,
              out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    fl = np.append(d.posterior_variance[1], d.betas[1:])
    want = {n: getattr(d, n) for n in _lib.COEF_ROWS[:11]}
    want.update(fixed_large_variance=fl, fixed_large_log_variance=np.log(fl), log_betas=np.log(d.betas))
    for i, n in enumerate(_lib.COEF_ROWS):
        ref = want[n].astype(np.float32)
        # sqrt / division are correctly rounded everywhere; log may differ in the last float64 bit between libm and
        # numpy's SIMD loops, which survives the cast to fp32 only on a rounding boundary: allow 1 ulp there
        if "log" in n:
            np.testing.assert_allclose(out[i], ref, rtol=1.2e-7, atol=0, err_msg=n)
        else:
            np.testing.assert_array_equal(out[i], ref, err_msg=n)
    bad = np.array([0.1, 1.5])
    assert _lib.lib().rgm_coeff_tables_host(bad.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 2,
                                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) != 0


def test_image_size_takes_two_values():
    """scripts/sample_rule.py / edit.py are invoked with `--image_size 128 16` and index args.image_size[0], [1]
    (reference script_util.py:510-511)."""
    p = argparse.ArgumentParser()
    script_util.add_dict_to_argparser(p, dict(image_size=128, batch_size=4, use_fp16=False, model="DiTRotary_XL_8"))
    a = p.parse_args(["--image_size", "128", "16", "--batch_size", "2", "--use_fp16", "true"])
    assert a.image_size == [128, 16] and a.batch_size == 2 and a.use_fp16 is True
    assert p.parse_args([]).image_size == 128


def _key(diffusion, model, kw):
    refs = []
    return (diffusion._sig(model, refs), tuple((k, diffusion._sig(v, refs)) for k, v in kw.items())), refs


def test_step_graph_signature_is_by_content_and_holds_references():
    d = create_diffusion(timestep_respacing="4")

    class Model:
        pass

    def model_fn(x, t, y=None, model=None, w=0.0):
        return x

    m = Model()
    y = torch.ones(2, dtype=torch.long)
    kw = dict(model_kwargs={"y": y, "rule": {"pitch_hist": torch.zeros(2, 12)}},
              scg_kwargs={"num_samples": 4, "pitch_hist": 1.0}, eta=1.0, embed_model=None)
    k1, refs1 = _key(d, partial(model_fn, model=m, w=0.0), kw)
    k2, _ = _key(d, partial(model_fn, model=m, w=0.0), kw)           # a NEW partial with the same content
    k3, _ = _key(d, partial(model_fn, model=m, w=3.0), kw)           # another cfg weight
    assert k1 == k2 and k1 != k3
    kw2 = dict(kw, scg_kwargs={"num_samples": 8, "pitch_hist": 1.0})
    assert _key(d, partial(model_fn, model=m, w=0.0), kw2)[0] != k1

    class Cfg(dict):  # a mapping that is not a plain dict (an OmegaConf DictConfig behaves like this)
        pass

    c = Cfg(num_samples=4, pitch_hist=1.0)
    ka, _ = _key(d, m, dict(scg_kwargs=c))
    c["pitch_hist"] = 2.0                                             # mutated in place: a different step
    assert _key(d, m, dict(scg_kwargs=c))[0] != ka
    ns = SimpleNamespace(schedule=True, t_start=750)
    kn, _ = _key(d, m, dict(g=ns))
    ns.t_start = 500
    assert _key(d, m, dict(g=ns))[0] != kn
    # objects keyed by id() are kept alive by the entry, so the id cannot be recycled
    assert any(r is m for r in refs1) and any(r is y for r in refs1)
    ident = id(m)
    del m
    gc.collect()
    assert any(id(r) == ident for r in refs1)


def test_step_graph_cache_is_bounded():
    d = create_diffusion(timestep_respacing="4")
    for i in range(200):
        d._graphs[("k", i)] = gd._StepGraph([])
        d._evict_graphs()
    assert len(d._graphs) == 64 and ("k", 199) in d._graphs and ("k", 0) not in d._graphs
    for i in range(20):  # captured graphs are bounded separately: least recently used goes first
        e = gd._StepGraph([])
        e.graph = object()
        d._graphs[("g", i)] = e
        d._evict_graphs()
    assert d.captured_graphs() == d.MAX_STEP_GRAPHS
    assert ("g", 19) in d._graphs and ("g", 0) not in d._graphs


def test_scg_refuses_learned_variance_like_the_reference():
    """The reference's scg_sample asserts on a learn_sigma model's 2C-channel output (:519 -> :360)."""
    d = create_diffusion(timestep_respacing="4", learn_sigma=True)
    x = torch.zeros(1, 4, 8, 16)
    with pytest.raises(AssertionError):
        d.scg_sample(lambda *a, **k: x, torch.zeros(1, dtype=torch.long), x, x, None, 1.0,
                     model_kwargs={"y": torch.zeros(1, dtype=torch.long), "rule": {}}, scg_kwargs={"num_samples": 2})


def _build_c_host(tmp_path):
    import os
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "scg_step_host")
    cuda_lib = "/usr/local/cuda/lib64"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
           os.path.join(root, "examples", "scg_step_host.c"), "-L", os.path.dirname(_lib.LIB_PATH), "-lrgm_b200", "-L", cuda_lib,
           "-lcudart", "-lm", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(_lib.LIB_PATH) + ":" + cuda_lib + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return exe, env


def test_header_is_valid_c_and_a_c_host_links(tmp_path):
    """include/rgm_b200.h compiles as C99 and examples/scg_step_host.c -- a whole DDIM + SCG step with no Python -- links
    against the library.  Without a B200 the program must refuse to compute (exit code 2), not fall back."""
    import subprocess

    exe, env = _build_c_host(tmp_path)
    if torch.cuda.is_available():
        return
    out = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 2 and "no CPU path" in out.stderr
