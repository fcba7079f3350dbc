"""Host build of the resampling guide-table logic (same PSVO_HD functions as the CUDA kernels) against the plain
upper-bound search: the bracketed search must return the identical ancestor for every threshold."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_guide_table_brackets_every_threshold(tmp_path):
    exe = str(tmp_path / "guide_table_check")
    src = os.path.join(ROOT, "tests", "host", "guide_table_check.cu")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "psvo_b200", "csrc"),
                    "-o", exe, src], check=True, capture_output=True, timeout=300)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad=0" in r.stdout, r.stdout
    probes = float(r.stdout.split("probes/child=")[1].split()[0])
    assert probes < 2.0, "the bracket should be about one particle wide, got %.2f probes per child" % probes
