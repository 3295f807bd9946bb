"""The one bitwise difference the GPU showed in a long AMR run, pinned on the CPU: the oracle with the DEVICE's
guard-free division formula in the PLM limiter (oracle/tools/fastdiv_experiment.py) differs from the unmodified reference in
exactly one cycle of the 2-D three-level AMR blast, in a handful of x2f faces at the 1e-175 level, U / time / dt equal; with the
guard (numerator below the fast path's range and non-zero -> IEEE division) every cycle is bitwise.  Round 2 moved that guard
from plm() into ddiv / drcp / dsqrt themselves (athenak_b200/csrc/mhd_device.cuh), so EVERY division site has it.
Needs oracle/_ref/athena_blast (live reference run)."""
import os
import re
import subprocess
import sys

import pytest

import oracle_lib as O

TOOL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "tools", "fastdiv_experiment.py")


def run(*args):
    return subprocess.run([sys.executable, TOOL, *args], capture_output=True, text=True, timeout=600, check=True).stdout


def test_guard_free_division_reproduces_the_gpu_difference_and_the_guard_removes_it():
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref/athena_blast not built")
    out = run()
    m = re.search(r"guard=False: (\d+) of (\d+) cycles differ", out)
    assert m and int(m.group(1)) == 1, out
    line = [l for l in out.splitlines() if l.startswith("cycle ")]
    assert len(line) == 1 and "U faces differing 0" in line[0] and "time equal True, dt equal True" in line[0], out
    mx = [float(x) for x in re.search(r"max\|dB\| \[(.*?)\]", line[0]).group(1).split(",")]
    assert mx[0] == 0.0 and 0.0 < mx[1] < 1e-150 and mx[2] == 0.0, out
    out = run("--guard")
    assert re.search(r"guard=True: 0 of \d+ cycles differ", out), out


def test_device_division_has_the_guard_at_every_site():
    """The range guard lives inside ddiv / drcp / dsqrt (one policy for every call site) and is on by default; the repeat paths'
    AKB_COLD mode (scaled exact division first) ends in the same cold operators."""
    src = open(os.path.join(os.path.dirname(TOOL), "..", "..", "athenak_b200", "csrc", "mhd_device.cuh")).read()
    assert re.search(r"#ifndef AKB_DIV_GUARD\s*\n#define AKB_DIV_GUARD 1", src)
    for fn, cold in (("double drcp(", "drcp_cold(x)"), ("double ddiv(", "ddiv_cold(a, b)"), ("double dsqrt(", "dsqrt_cold(x)")):
        body = src[src.index(fn):src.index(fn) + 1800]
        assert "#if AKB_DIV_GUARD" in body and cold in body, fn
    plm = src[src.index("void plm("):src.index("// ppm.hpp:44-77")]
    assert "ddiv(dq2, dql + dqr, &o)" in plm and "ddiv(dq2, dql + dqr, oor)" in plm and " / " not in "\n".join(l.split("//")[0] for l in plm.splitlines())
    # no call site may bypass the guard: every division / reciprocal / square root of the device arithmetic goes through the
    # guarded helpers (the cold helpers themselves, host-uniform constants and comments excepted)
    for name in ("mhd_device.cuh", "hydro_device.cuh", "iso_device.cuh", "c2p_device.cuh"):
        txt = open(os.path.join(os.path.dirname(TOOL), "..", "..", "athenak_b200", "csrc", name)).read()
        code = "\n".join(l.split("//")[0] for l in txt.splitlines())
        assert "sqrt(" not in re.sub(r"\b(dsqrt|dsqrt_cold|rsqrt)\b", "", code).replace("return sqrt(x)", ""), name
