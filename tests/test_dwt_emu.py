"""The streaming DWT kernels (csrc/dwt_stream.cuh), compiled for the CPU with one OS thread per lane, against the oracle:
checks parities, reflections, vector / scalar paths and strip / chunk borders without a GPU (tests/dwt_emu.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_streaming_dwt_kernels_on_cpu_lanes():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libgb_oracle.so"])
    exe = os.path.join(ROOT, "tests", "_dwt_emu")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-DGB_EMU", os.path.join(ROOT, "tests", "dwt_emu.cpp"),
                           "-L" + os.path.join(ROOT, "oracle"), "-lgb_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe])
    r = subprocess.run([exe, "quick"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 failed" in r.stdout
    # a few random geometries on top (origins, sizes, levels, rows per item, halo lanes, register / ring queue); the round's
    # offline sweep was `tests/_dwt_emu fuzz 150 <seed>` for seeds 1-4: 1200 cases, all exact
    r = subprocess.run([exe, "fuzz", "25", "7"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "0 failed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
