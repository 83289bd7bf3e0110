"""The generated Montgomery multiply (csrc/mont_asm.inc) on CPU: gen_mont.py's self-test runs the very instruction list
it emits through a Python emulation of the PTX carry semantics against big-integer arithmetic (the same integer
(a*b + m*p) / 2^256 the reference computes, bb/ecc/fields/field_impl_generic.hpp:392-499), and the committed .inc file
must be exactly what the generator produces."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "aztec-2.0_b200", "csrc")


def test_generator_selftest():
    out = subprocess.run([sys.executable, os.path.join(CSRC, "gen_mont.py"), "--selftest"], capture_output=True, text=True, check=True).stdout
    for key in ("fq mul: ok", "fq sqr: ok", "fr mul: ok", "fr sqr: ok",
                # the dedicated squaring kept behind --dedicated-sqr (100 wide multiply-adds) must stay exact as well
                "fq sqr (dedicated, --dedicated-sqr): ok", "fr sqr (dedicated, --dedicated-sqr): ok", "100 wide pairs"):
        assert key in out, out


def test_committed_include_is_what_the_generator_emits(tmp_path):
    work = tmp_path / "csrc"
    work.mkdir()
    shutil.copy(os.path.join(CSRC, "gen_mont.py"), work / "gen_mont.py")
    subprocess.run([sys.executable, str(work / "gen_mont.py")], capture_output=True, text=True, check=True)
    with open(work / "mont_asm.inc") as f, open(os.path.join(CSRC, "mont_asm.inc")) as g:
        assert f.read() == g.read()
