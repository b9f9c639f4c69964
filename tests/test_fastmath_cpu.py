"""Host check of the branch-free math helpers used inside the trajectory kernels (csrc/u1_fastmath.cuh):
fast_sin / fast_cos / fast_sincos against 80-bit libm, wrap_fast and winding against the reference's
regularize expression (utils.py:182-187), neg2_log (Box-Muller radius) against logl.  The header is
host/device code, so the same source the kernels compile is exercised here with g++."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_fastmath_header_against_long_double(tmp_path):
    exe = str(tmp_path / "check_fast")
    flags = ["-O2", "-std=c++17"]
    try:
        if " fma " in open("/proc/cpuinfo").read():
            flags.append("-mfma")
    except OSError:
        pass
    subprocess.run(["g++", *flags, "-o", exe, os.path.join(ROOT, "tools", "check_fast_sin.cpp")], check=True)
    r = subprocess.run([exe, "400000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-2000:] + r.stderr[-500:]


def test_log_table_is_reproducible(tmp_path):
    """csrc/u1_logtab.inc is exactly what tools/gen_log_table.py writes (committed data, generating script)."""
    mp = pytest.importorskip("mpmath")  # noqa: F841
    inc = os.path.join(ROOT, "hmc_ft_b200", "csrc", "u1_logtab.inc")
    out = str(tmp_path / "logtab.inc")
    subprocess.run(["python", os.path.join(ROOT, "tools", "gen_log_table.py"), out], check=True, capture_output=True)
    assert open(out).read() == open(inc).read()
