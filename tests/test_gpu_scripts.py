"""The reference's four scripts (BASELINE configs 1-4) as they read on this library: scripts/*.py mirror
scripts/*.jl line for line (only the import and Julia's `!` change).  Run here at reduced particle counts / lengths."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_script(name, *args, cwd):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", name), *args], capture_output=True, text=True,
                         timeout=300, cwd=cwd)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    return out.stdout


def test_vlasov_poisson_script_as_shipped(tmp_path):
    """config 1: 1e4 particles, 200 steps, every step written to the HDF5 file and read back (:30-38)"""
    out = run_script("vlasov_poisson.py", "--h5file", str(tmp_path / "vp.hdf5"), cwd=tmp_path)
    assert "z: (2, 10000, 201)" in out
    assert os.path.getsize(tmp_path / "vp.hdf5") > 201 * 2 * 10000 * 8


def test_bump_on_tail_script(tmp_path):
    """config 2 at 2e6 particles: the instability grows and W + K is conserved"""
    out = run_script("bump_on_tail.py", "--npart", "2e6", "--T", "40", cwd=tmp_path)
    drift = float(re.search(r"relative energy drift: (\S+)", out).group(1))
    assert drift < 5e-3
    m = re.search(r"W\(0\) = (\S+), max W = (\S+) at t = (\S+)", out)
    assert float(m.group(2)) > 5 * float(m.group(1).rstrip(",")) and 10 < float(m.group(3)) < 40


@pytest.mark.parametrize("plain", [False, True])
def test_lenard_bernstein_conservative_script(tmp_path, plain):
    """config 4 at 2e5 particles x 200 steps: Σv, Σv² from the frames read back from the HDF5 file (:46-50,64)"""
    args = ["--npart", "2e5", "--tend", "2", "--save-stride", "50", "--h5file", str(tmp_path / "clb.hdf5")]
    out = run_script("lenard_bernstein_conservative.py", *args, *(["--plain"] if plain else []), cwd=tmp_path)
    rows = re.findall(r"t =\s*(\S+), mom = (\S+), enr = (\S+),", out)
    assert [float(r[0]) for r in rows] == [0.0, 0.5, 1.0, 1.5, 2.0]
    enr = [abs(float(r[2])) for r in rows]
    if plain:
        assert enr[-1] > 1e-3          # plain LB relaxes towards the unit Maxwellian: energy is not conserved
    else:
        assert max(enr) < 1e-9 and max(abs(float(r[1])) for r in rows) < 1e-6
    assert "normalised fourth moment" in out


def test_lenard_bernstein_script(tmp_path):
    """config 3 at 1e5 particles (RK438 instead of the shipped TRBDF2)"""
    out = run_script("lenard_bernstein.py", "--npart", "1e5", cwd=tmp_path)
    m = re.search(r"momentum change / N: (\S+); energy change: (\S+)", out)
    assert float(m.group(1)) < 1e-12 and float(m.group(2)) < 1e-6
