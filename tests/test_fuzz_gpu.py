"""Randomised differential test (tools/fuzz_parity.py): every pipe vs the CPU oracle with random channel counts,
lengths, chunkings (blocking and asynchronous mode), impairments and error rates.  A short budget here; the long
runs of the round are recorded in profiles/r01_fuzz.txt."""
import os
import subprocess
import sys

import pytest

import oracle_lib

pytestmark = pytest.mark.gpu


def test_fuzz_all_pipes_short():
    tool = os.path.join(oracle_lib.ROOT, "tools", "fuzz_parity.py")
    # a fixed number of rounds (deterministic set of cases), generous time budget
    r = subprocess.run([sys.executable, tool, "600", "3", "60"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "fuzz: all equal" in r.stdout


def test_fuzz_decoders_short():
    tool = os.path.join(oracle_lib.ROOT, "tools", "fuzz_decoders.py")
    r = subprocess.run([sys.executable, tool, "600", "4", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "decoder fuzz: all equal" in r.stdout
