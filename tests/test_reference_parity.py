"""Parity with the reference itself, imported from /root/reference/src.  These tests only run
where the reference is mounted (the authoring container); on the GPU box the same parity is
checked against the committed golden fixtures (tests/test_gpu_golden.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

REF_SRC = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_SRC, "pytristan")), reason="reference not mounted")


def _ref():
    if REF_SRC not in sys.path:
        sys.path.append(REF_SRC)
    import pytristan as ref

    return ref


def test_oracle_grid_layout_equals_reference():
    from oracle import operators as oracle

    ref = _ref()
    rng = np.random.default_rng(0)
    for shape in [(1,), (7,), (2, 3), (5, 1, 4), (6, 5, 4, 3), (33, 17, 9), (64, 65)]:
        axes = [np.sort(rng.standard_normal(n)) for n in shape]
        assert oracle.grid_expand(axes).tobytes() == np.asarray(ref.Grid.from_arrs(*axes)).tobytes()
    a = oracle.grid_expand([np.arange(3), np.linspace(0, 1, 4)])
    b = np.asarray(ref.Grid.from_arrs(np.arange(3), np.linspace(0, 1, 4)))
    assert a.dtype == b.dtype and np.array_equal(a, b)
    ai = oracle.grid_expand([np.arange(3), np.arange(2)])
    assert ai.dtype == np.asarray(ref.Grid.from_arrs(np.arange(3), np.arange(2))).dtype


def test_cheb_equals_reference():
    import pytristan_b200 as pt
    from oracle import operators as oracle

    ref = _ref()
    for n in (0, 2, 3, 10, 64, 257, 512, 1024, 2048, 4097):
        assert pt.cheb(n).tobytes() == ref.cheb(n).tobytes() == oracle.cheb_nodes(n).tobytes()
        assert pt.cheb(n, -3.0, 0.25).tobytes() == ref.cheb(n, -3.0, 0.25).tobytes()


def test_host_facade_equals_reference(cpu_expand, clean_registries):
    """Same constructors, same results (host logic; device expansion doubled on CPU)."""
    import pytristan_b200 as pt

    ref = _ref()
    cases = [
        dict(bounds=[(0.0, 1.0, 8), (-2.0, 2.0, 6)], axes=[1, 0], mappers=["cheb", "cheb"]),
        dict(bounds=[(0.0, 1.0, 5), (-1.0, 1.0, 4), (2.0, 3.0, 3)], axes=[-2], mappers=["cheb"]),
        dict(bounds=[(0.0, 2 * np.pi - 2 * np.pi / 64, 64), (-1.0, 1.0, 33)], axes=[1], mappers=["cheb"]),
    ]
    for c in cases:
        mine = pt.Grid.from_bounds(*c["bounds"], axes=c["axes"], mappers=[pt.cheb] * len(c["axes"]))
        theirs = ref.Grid.from_bounds(*c["bounds"], axes=c["axes"], mappers=[ref.cheb] * len(c["axes"]))
        assert np.asarray(mine).tobytes() == np.asarray(theirs).tobytes()
        assert mine.npts == theirs.npts and mine.num_dim == theirs.num_dim
        for k in range(mine.num_dim):
            assert mine.axpoints(k).tobytes() == theirs.axpoints(k).tobytes()
        assert all(np.array_equal(a, b) for a, b in zip(mine.mgrids(), theirs.mgrids()))
    for kw in (dict(nt=4, nr=6), dict(nt=16, nr=8, fornberg=True), dict(nt=1024, nr=512)):
        mine = pt.get_polar_grid(axes=[1], mappers=[pt.cheb], **kw)
        theirs = ref.get_polar_grid(axes=[1], mappers=[ref.cheb], **kw)
        assert np.asarray(mine).tobytes() == np.asarray(theirs).tobytes() and mine.num == theirs.num
    ref.drop_all_grids()


def test_reference_test_file_runs_unchanged_against_this_package():
    """The reference's own suite (reference src/pytristan/tests/test_grid.py, 32 cases) with the
    import name aliased to pytristan_b200."""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests"), os.environ.get("PYTHONPATH", "")]))
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-p", "_alias_plugin",
           "--rootdir", "/tmp", os.path.join(REF_SRC, "pytristan", "tests", "test_grid.py")]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    tail = (res.stdout + res.stderr)[-2000:]
    assert res.returncode == 0, tail
    assert "32 passed" in res.stdout, tail
