"""CPU: the scoring kernels' exact-arithmetic helpers, compiled for the host FROM THEIR OWN SOURCE
(tailseeker_b200/csrc/tsk_exact.cuh + tsk_device.cuh, intrinsics mapped by tsk_hostcheck.h), against
IEEE division and the host libm's logf -- the two things the reference's scores are made of
(signalproc.c:39-51 shannon_entropy, :246-266 normalize_signals) -- and the ruler's histogram bin
(tsk_scorebin.cuh) against the fp64 expression of polyaruler.c:163-168.  Exhaustive over every normal float
in (0, 1] for the three logf forms; the GPU self-test's own operand pairs for the division."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read().replace("\n", " ")
    except OSError:
        return False


@pytest.mark.skipif(shutil.which("g++") is None or not _cpu_has_fma(), reason="needs g++ and a CPU with FMA")
def test_kernel_helpers_are_exact_on_the_host(tmp_path):
    exe = str(tmp_path / "exact_check")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-DTSK_HOST_CHECK",
                           "-I", os.path.join(ROOT, "tailseeker_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "exact_check.cpp"), "-lm", "-lpthread"])
    r = subprocess.run([exe, "1", str(1 << 26), "17"], capture_output=True, timeout=600)
    out = json.loads(r.stdout.decode())
    assert r.returncode == 0, out
    assert out["floats"] == 0x3f800000 - 0x00800000 + 1          # every normal float in (0, 1]
    assert out["tsk_logf"] == out["logf_core_t"] == out["logf_ki_t"] == out["div_by_rcp"] == 0
    assert out["division_pairs"] == 2 << 26
    # quotients nearest to a rounding midpoint, for every 17th divisor significand
    assert out["hard_quotients"] > 20_000_000 and out["div_by_rcp_hard"] == 0
    # the ruler's integer histogram bin against the reference's fp64 expression, every packed score
    assert out["score_bins"] == 14 * ((1 << 24) - 1) and out["score_bin"] == 0
