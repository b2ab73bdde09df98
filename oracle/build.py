"""Build of the oracle's C restatements (TEST INFRASTRUCTURE — never loaded by the product).

``python -m oracle.build`` compiles ``oracle/glibc_cos.c`` into
``oracle/_build/liboracle_glibc_cos.so`` with gcc (``-ffp-contract=off``: every fused multiply-add
in the restatement is an explicit ``fma()``)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_glibc_cos.so")
SRC = os.path.join(HERE, "glibc_cos.c")
TABLE = os.path.join(HERE, "..", "pytristan_b200", "csrc", "glibc_sincos_table.inc")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(TABLE)):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: cannot build the C oracle")
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call([gcc, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
