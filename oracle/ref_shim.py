"""Import the UNMODIFIED reference (/root/reference/code) on CPU under shims.

TEST INFRASTRUCTURE ONLY.  This module exists only in the build container: the
reference tree is not present on the GPU box, so nothing under ``-m gpu`` tests,
``smoke()`` or ``bench.py`` may import it.  It is used by ``oracle/make_golden.py``
(to generate ``tests/golden/*.npz``) and by the ``not gpu`` test that pins the
restatement in ``oracle/mvsdf_oracle.py`` against the real reference.

Shims, all outside the reference tree:
  1. empty ``imageio`` / ``skimage`` modules (imported by utils/rend_util.py:2-3,
     used only by image loaders that the hot path never calls);
  2. ``numpy.lib.function_base`` with a ``diff`` attribute (stray import at
     model/loss.py:1, removed in numpy 2);
  3. ``Tensor.cuda`` / ``Module.cuda`` -> identity (47 hard-coded ``.cuda()`` calls);
  4. a dict-backed stand-in for the pyhocon config object used by
     ``IDRNetwork.__init__`` (implicit_differentiable_renderer.py:170-177).
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MVSDF_REFERENCE_ROOT", "/root/reference")
REFERENCE_CODE = os.path.join(REFERENCE_ROOT, "code")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_CODE, "model"))


class DictConf:
    """Minimal pyhocon.ConfigTree stand-in: get_int/get_float/get_config/get_string/get_list."""

    def __init__(self, d):
        self._d = d

    def _get(self, key):
        node = self._d
        for part in key.split("."):
            node = node[part]
        return node

    def get_int(self, key):
        return int(self._get(key))

    def get_float(self, key):
        return float(self._get(key))

    def get_string(self, key):
        return str(self._get(key))

    def get_list(self, key):
        return list(self._get(key))

    def get_config(self, key):
        return DictConf(self._get(key))

    # ``**conf.get_config(...)`` needs the mapping protocol
    def keys(self):
        return self._d.keys()

    def __getitem__(self, k):
        v = self._d[k]
        return DictConf(v) if isinstance(v, dict) else v


def model_conf(width=512, render_width=None, line_step_iters=3):
    """The `model{}` block of confs/mvsdf_dtu.conf:17-58 with a selectable hidden width."""
    rw = width if render_width is None else render_width
    return {
        "feature_vector_size": 256,
        "implicit_network": {
            "d_in": 3, "d_out": 1, "dims": [width] * 8, "geometric_init": True, "bias": 0.6,
            "skip_in": [4], "weight_norm": True, "multires": 6,
        },
        "rendering_network": {
            "mode": "idr", "d_in": 9, "d_out": 3, "dims": [rw] * 4, "weight_norm": True,
            "multires_view": 4,
        },
        "ray_tracer": {
            "object_bounding_sphere": 1.0, "sdf_threshold": 5.0e-5, "line_search_step": 0.5,
            "line_step_iters": line_step_iters, "sphere_tracing_iters": 10, "n_steps": 100,
            "n_secant_steps": 8,
        },
    }


_loaded = None


def load():
    """Returns a namespace with the reference modules (imported once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_CODE}")
    import numpy as np
    import torch

    for name in ("imageio", "skimage"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if "numpy.lib.function_base" not in sys.modules:
        fb = types.ModuleType("numpy.lib.function_base")
        fb.diff = np.diff
        sys.modules["numpy.lib.function_base"] = fb
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)
    with contextlib.redirect_stdout(io.StringIO()):
        import model.implicit_differentiable_renderer as idr
        import model.ray_tracing as ray_tracing
        import model.sample_network as sample_network
        import model.embedder as embedder
        import model.loss as loss
        import model.conf as conf
        import utils.rend_util as rend_util
        import utils.my_utils as my_utils
    _loaded = types.SimpleNamespace(
        idr=idr, ray_tracing=ray_tracing, sample_network=sample_network, embedder=embedder,
        loss=loss, conf=conf, rend_util=rend_util, my_utils=my_utils)
    return _loaded


@contextlib.contextmanager
def quiet():
    """The reference prints tracer statistics every forward (ray_tracing.py:63-66)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
