"""CPU test of the any-length tile FFT core (csrc/fft_any.cuh compiles for the host): tools/test_fft_any.cpp runs every
stage function serially, exactly as the device code indexes it, for the reference's production mesh sizes and a sweep of
prime / composite / tiny lengths against a long-double DFT."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_any_length_core_against_naive_dft(tmp_path):
    exe = str(tmp_path / "test_fft_any")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", os.path.join(ROOT, "tools", "test_fft_any.cpp"),
                    "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "n= 274 plan 137 2" in r.stdout and "n= 307 plan 307" in r.stdout and r.stdout.strip().endswith("OK")
