"""csrc/glibc_pow.h restates glibc's pow() (the libm the golden fixtures were generated with) for the reference-order
Jacobian kernels; here the HOST instantiation of that header is compared bit for bit with this machine's libm.
The device instantiation is compared with the host's libm in tests/test_parity_gpu.py."""
import os
import subprocess

from conftest import ROOT


def test_glibc_pow_restatement_is_bit_exact_on_the_host(tmp_path):
    exe = str(tmp_path / "glibc_pow_check")
    src = os.path.join(ROOT, "tests", "helpers", "glibc_pow_check.cc")
    inc = os.path.join(ROOT, "pressio-demoapps_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + inc, src, "-o", exe], check=True)
    out = subprocess.run([exe, "2000000"], check=True, stdout=subprocess.PIPE, text=True).stdout.split()
    bad, skipped = int(out[0]), int(out[1])
    assert bad == 0, "%d results differ from libm's pow" % bad
    assert skipped == 0


def test_tables_match_this_libm(tmp_path):
    """the committed tables are the ones tools/gen_glibc_pow_tables.py reads out of this image's libm"""
    import importlib.util
    import shutil
    gen = os.path.join(ROOT, "tools", "gen_glibc_pow_tables.py")
    spec = importlib.util.spec_from_file_location("gen_glibc_pow_tables", gen)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(mod.LIBM):
        import pytest
        pytest.skip("no libm at the expected path")
    committed = open(mod.OUT).read()
    mod.OUT = str(tmp_path / "tables.inc")
    mod.main()
    assert open(mod.OUT).read() == committed
    shutil.rmtree(str(tmp_path), ignore_errors=True)
