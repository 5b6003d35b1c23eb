"""The host DEFLATE decoder (csrc/host_inflate.cpp) on its own, under AddressSanitizer / UBSan: every zlib level,
stored / fixed / dynamic blocks, sizes from 1 byte to a full BGZF block, and 20 000 corrupted streams. CPU only."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_decoder_harness_under_sanitizers(tmp_path):
    exe = str(tmp_path / "inflate_harness")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           os.path.join(ROOT, "tests", "native", "inflate_harness.cpp"),
           os.path.join(ROOT, "midas_b200", "csrc", "host_inflate.cpp"), "-lz", "-o", exe]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "valid streams: 0 of 400 failed" in out.stdout
    assert "20000 runs, 0 guard violations" in out.stdout
