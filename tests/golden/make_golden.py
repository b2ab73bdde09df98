"""Generate the golden fixtures from the REFERENCE itself (imported from /root/reference/src).

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/reference_small.npz (full arrays of small reference grids and node sets)
and tests/golden/reference_hashes.json (sha256 of the bytes of larger reference grids).
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/src")
import pytristan as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def stretch(npts, lo, hi):
    """A custom non-affine mapper (tanh clustering) with the reference's mapper protocol."""
    s = np.linspace(-1.0, 1.0, npts)
    return lo + (hi - lo) * 0.5 * (1.0 + np.tanh(1.5 * s) / np.tanh(1.5))


# name -> (constructor kwargs); rebuilt verbatim by tests/test_reference_parity.py
SMALL = {
    "raw_2x3": dict(bounds=[(0.0, 1.0, 2), (-1.0, 1.0, 3)], axes=[], mappers=[]),
    "cheb_8x3": dict(bounds=[(0.0, 1.0, 8), (-1.0, 1.0, 3)], axes=[0], mappers=["cheb"]),
    "cheb_8x6_both": dict(bounds=[(0.0, 1.0, 8), (-2.0, 2.0, 6)], axes=[1, 0], mappers=["cheb", "cheb"]),
    "cheb_neg_axis": dict(bounds=[(0.0, 1.0, 8), (-2.0, 2.0, 6)], axes=[-1], mappers=["cheb"]),
    "three_d": dict(bounds=[(0.0, 1.0, 5), (-1.0, 1.0, 4), (2.0, 3.0, 3)], axes=[1], mappers=["cheb"]),
    "stretch_2d": dict(bounds=[(0.0, 2.0, 9), (0.0, 1.0, 17)], axes=[1], mappers=["stretch"]),
    "one_d_cheb64": dict(bounds=[(-1.0, 1.0, 64)], axes=[0], mappers=["cheb"]),
}
LARGE = {
    "channel_1024x257": dict(bounds=[(0.0, 2 * np.pi - 2 * np.pi / 1024, 1024), (-1.0, 1.0, 257)], axes=[1], mappers=["cheb"]),
    "cube_48": dict(bounds=[(0.0, 1.0, 48), (0.0, 2.0, 48), (-1.0, 1.0, 48)], axes=[2], mappers=["cheb"]),
}
MAPPERS = {"cheb": ref.cheb, "stretch": stretch}


def build(spec):
    return ref.Grid.from_bounds(*spec["bounds"], axes=spec["axes"], mappers=[MAPPERS[m] for m in spec["mappers"]])


def main():
    out = {}
    for name, spec in SMALL.items():
        out["grid_" + name] = np.asarray(build(spec))
    for n in (2, 3, 10, 64, 257, 512, 1024, 2048):
        out[f"cheb_{n}"] = ref.cheb(n)
        out[f"cheb_{n}_0_10"] = ref.cheb(n, 0.0, 10.0)
        out[f"cheb_{n}_m1_1"] = ref.cheb(n, -1.0, 1.0)
    # polar grids (reference grid.py:498-511), small enough to store whole
    for name, kw in {"polar_4_6": dict(nt=4, nr=6), "polar_4_6_cheb": dict(nt=4, nr=6, axes=[1], mappers=[ref.cheb]),
                     "polar_4_6_fornberg": dict(nt=4, nr=6, axes=[1], mappers=[ref.cheb], fornberg=True)}.items():
        out["grid_" + name] = np.asarray(ref.get_polar_grid(**kw))
    ref.drop_all_grids()
    np.savez_compressed(os.path.join(HERE, "reference_small.npz"), **out)

    hashes = {}
    for name, spec in LARGE.items():
        g = np.asarray(build(spec))
        hashes[name] = {"sha256": hashlib.sha256(g.tobytes()).hexdigest(), "shape": list(g.shape), "dtype": str(g.dtype)}
    for nt, nr, fb in ((1024, 512, False), (1024, 256, True)):
        g = ref.get_polar_grid(nt, nr, axes=[1], mappers=[ref.cheb], fornberg=fb)
        hashes[f"polar_{nt}_{nr}_{'f' if fb else 'n'}"] = {
            "sha256": hashlib.sha256(np.asarray(g).tobytes()).hexdigest(), "shape": list(g.shape), "dtype": str(g.dtype),
            "theta_sha256": hashlib.sha256(g.axpoints(0).tobytes()).hexdigest(),
            "r_sha256": hashlib.sha256(g.axpoints(1).tobytes()).hexdigest()}
    ref.drop_all_grids()
    with open(os.path.join(HERE, "reference_hashes.json"), "w") as fh:
        json.dump(hashes, fh, indent=1, sort_keys=True)
    print("wrote", len(out), "arrays and", len(hashes), "hashes")


if __name__ == "__main__":
    main()
