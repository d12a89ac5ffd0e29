"""The C++ blocks with SDR++'s dsp::Processor / dsp::stream surface (sdrpp_tetra_demodulator_b200/host), built
against the stand-in core headers, wired and start()ed like src/main.cpp:84-110 does -- one worker thread per
block, double-buffered streams between them -- must deliver the oracle's bit stream."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sdrpp_tetra_demodulator_b200", "host", "test_host_block")


def test_host_block_builds_against_sdrpp_surface():
    """compile check (no GPU): the block sources build against the SDR++-shaped headers and link the C ABI"""
    subprocess.run(["make", "-C", os.path.dirname(BIN)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert os.path.exists(BIN)


@pytest.mark.gpu
@pytest.mark.parametrize("buffer_samples", [32768, 4097])
def test_threaded_block_chain_matches_oracle(O, tmp_path, buffer_samples):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(BIN):
        subprocess.run(["make", "-C", os.path.dirname(BIN)], check=True)
    N = 150_000
    iq = O.generate(1, N)
    ob = O.OracleB(1)
    counts, _, dibits, bits = ob.process(iq, want_bits=True)
    fin, fout = tmp_path / "in.f32", tmp_path / "out.bits"
    iq[0].tofile(fin)
    r = subprocess.run([BIN, str(fin), str(fout), str(buffer_samples)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(fout, dtype=np.uint8)
    n = int(counts[0])
    assert len(got) == 2 * n, (len(got), 2 * n, r.stdout)
    assert np.array_equal(got, bits[0, :2 * n])
    assert "sync 1" in r.stdout
