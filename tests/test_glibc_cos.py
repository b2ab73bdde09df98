"""The libm-exact cosine: C restatement (oracle/glibc_cos.c) pinned against the host's real libm,
and the device kernel (csrc/glibc_cos.cuh) pinned against both.

Why it exists: the reference builds Chebyshev nodes with numpy.cos (grid.py:75) = glibc's cos, which
is not correctly rounded; a device generator can only be bit-exact by reproducing that algorithm."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import glibc_cos as oc
from pytristan_b200 import _lib

CHEB_SIZES = list(range(2, 300)) + [511, 512, 513, 1024, 1025, 2048, 4096, 8191, 16384, 65537]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_host_libm_is_a_modelled_glibc_build():
    model = _lib.host_libm_model()
    assert model in (_lib.LIBM_GLIBC_FMA, _lib.LIBM_GLIBC_NOFMA), "numpy.cos here is neither glibc build"


def test_c_restatement_matches_numpy_cos_on_cheb_arguments():
    fma = _lib.host_libm_model() == _lib.LIBM_GLIBC_FMA
    bad = 0
    for n in CHEB_SIZES:
        ref = np.cos(np.arange(n) * np.pi / (n - 1))          # the reference's expression, grid.py:75
        bad += int(np.count_nonzero(_bits(oc.cheb_nodes(n, fma)) != _bits(ref)))
    assert bad == 0


def test_c_restatement_matches_libm_on_random_arguments():
    fma = _lib.host_libm_model() == _lib.LIBM_GLIBC_FMA
    rng = np.random.default_rng(2024)
    x = np.concatenate([rng.uniform(-3.3, 3.3, 2_000_000), rng.uniform(-1e-7, 1e-7, 1000),
                        rng.uniform(-1.0e8, 1.0e8, 500_000), [0.0, -0.0, 0.855469, 2.426265, math.pi, math.pi / 2]])
    assert np.array_equal(_bits(oc.glibc_cos(x, fma)), _bits(np.cos(x)))
    # the other build differs from this host's libm somewhere (otherwise the model probe is meaningless)
    assert not np.array_equal(_bits(oc.glibc_cos(x, not fma)), _bits(np.cos(x)))


def test_model_probes_agree_with_the_restatement():
    x = np.array([p[0] for p in _lib._LIBM_PROBES], dtype=np.uint64).view(np.float64)
    assert [int(v) for v in _bits(oc.glibc_cos(x, True))] == [p[1] for p in _lib._LIBM_PROBES]
    assert [int(v) for v in _bits(oc.glibc_cos(x, False))] == [p[2] for p in _lib._LIBM_PROBES]


def test_non_fma_restatement_matches_libm_without_fma():
    """glibc selects its baseline cos when FMA is masked off through GLIBC_TUNABLES."""
    code = ("import numpy as np, sys; sys.path.insert(0, %r)\n"
            "from oracle import glibc_cos as oc\n"
            "rng = np.random.default_rng(7); x = rng.uniform(-3.3, 3.3, 1_000_000)\n"
            "a = np.cos(x).view(np.uint64)\n"
            "print(int((oc.glibc_cos(x, False).view(np.uint64) != a).sum()), int((oc.glibc_cos(x, True).view(np.uint64) != a).sum()))\n"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, GLIBC_TUNABLES="glibc.cpu.hwcaps=-FMA,-AVX2")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    nofma_bad, fma_bad = (int(v) for v in out.stdout.split())
    if _lib.host_libm_model() == _lib.LIBM_GLIBC_FMA and fma_bad == 0:
        pytest.skip("this glibc ignores the hwcaps tunable: the baseline build cannot be reached here")
    assert nofma_bad == 0


@pytest.mark.gpu
def test_device_cheb_nodes_are_bit_identical_to_the_reference_formula():
    import torch

    import pytristan_b200 as pt

    bad = 0
    for n in CHEB_SIZES:
        got = pt.cheb_device(n).cpu().numpy()
        bad += int(np.count_nonzero(_bits(got) != _bits(pt.cheb(n))))
        gm = pt.cheb_device(n, -2.0, 5.0).cpu().numpy()
        bad += int(np.count_nonzero(_bits(gm) != _bits(pt.cheb(n, -2.0, 5.0))))
    assert bad == 0
    assert pt.cheb_device(0).numel() == 0
    with pytest.raises(TypeError):
        pt.cheb_device(3.0)
    with pytest.raises(ValueError):
        pt.cheb_device(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("model", [_lib.LIBM_GLIBC_FMA, _lib.LIBM_GLIBC_NOFMA])
def test_device_cosine_matches_the_c_restatement(model):
    import torch

    lib = _lib.load()
    rng = np.random.default_rng(99)
    x = np.concatenate([rng.uniform(-3.3, 3.3, 4_000_000), rng.uniform(-1e-7, 1e-7, 1000),
                        rng.uniform(-1.0e8, 1.0e8, 1_000_000), [0.0, -0.0, 0.855469, 2.426265, math.pi, math.pi / 2]])
    t = torch.from_numpy(x).cuda()
    y = torch.empty_like(t)
    _lib.check(lib.pt_cos_libm(_lib.dptr(t), _lib.dptr(y), t.numel(), model, _lib.stream_ptr()), "pt_cos_libm")
    assert np.array_equal(_bits(y.cpu().numpy()), _bits(oc.glibc_cos(x, model == _lib.LIBM_GLIBC_FMA)))
