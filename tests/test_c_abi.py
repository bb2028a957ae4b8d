"""include/sart.h is plain C: a C99 driver (tests/c/abi_driver.c) compiled with gcc -pedantic -Werror links against
libsart.so and uses the boundary the way a Nim/cgo/JNI binding would — no Python in the data path. On a CPU-only machine
it checks the host-side entry points and that sart_create refuses (no CPU fallback); on the B200 it runs all three
pipelines through the C-ABI."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIBDIR = ROOT / "solaraxionraytracing_b200"


def _build(tmp_path):
    exe = tmp_path / "abi_driver"
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O1", "-I", str(ROOT / "include"),
           str(ROOT / "tests" / "c" / "abi_driver.c"), "-L", str(LIBDIR), "-lsart", f"-Wl,-rpath,{LIBDIR}", "-lm", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_c99_and_library_links_from_c(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_driver:" in r.stdout


@pytest.mark.gpu
def test_c_driver_runs_all_pipelines_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("abi_driver: mode") == 3, r.stdout
