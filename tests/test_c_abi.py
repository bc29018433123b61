"""The boundary from C: tests/c_abi/abi_smoke.c is compiled with gcc against include/b381.h and linked with libb381.so, i.e.
what the cgo binding of INTEGRATION.md does.  Every function the header declares is referenced from C (a declaration the
library does not export fails the link), struct layouts and enum values are checked by the C compiler, and -- on a GPU --
pairings and CompareTwoPairings-style checks run through the C caller and are compared with the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from bls_b200 import hostgen as hg, layout as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CDIR = os.path.join(ROOT, "tests", "c_abi")


def declared_functions():
    src = open(os.path.join(ROOT, "include", "b381.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    names = re.findall(r"\b(b381_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as g
    g.build()
    names = declared_functions()
    assert len(names) > 50 and "b381_pairing_batch" in names and "b381_test_op" in names
    with open(os.path.join(CDIR, "symbols.inc"), "w") as f:
        f.write("".join("    (anyfn)%s,\n" % n for n in names))
    out = os.path.join(CDIR, "abi_smoke")
    lib = os.path.join(ROOT, "bls_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-I", os.path.join(ROOT, "include"),
                           "-I", CDIR, "-o", out, os.path.join(CDIR, "abi_smoke.c"), "-L", lib, "-l:libb381.so",
                           "-Wl,-rpath," + lib])
    return out


def test_c_caller_layout_and_exports(exe):
    """no GPU needed: sizes, offsets, enum values, every declared function linked; without a device b381_init says so"""
    r = subprocess.run([exe, "--layout"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "layout ok" in r.stdout


@pytest.mark.gpu
def test_c_caller_pairings_vs_oracle(exe, orc, tmp_path):
    n = 5
    P = hg.g1_progression(0xC0FFEE, 3, n); Q = hg.g2_progression(0xBEEF, 5, n)
    P[n - 1] = hg.g1_neg(P[:1])[0]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(np.uint64(n).tobytes()); f.write(P.tobytes()); f.write(Q.tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(fout, "rb").read()
    got = np.frombuffer(raw[:n * 576], dtype=np.uint64)
    assert got.tobytes() == orc.pairing_batch(P, Q, threads=4).tobytes()
    assert list(raw[n * 576:n * 576 + 2]) == [1, 0]
    assert int(np.frombuffer(raw[n * 576 + 2:n * 576 + 10], dtype=np.uint64)[0]) >= 4
