import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def unpack2(packed: np.ndarray, n: int) -> np.ndarray:
    """inverse of tests/golden/make_golden.py::pack2"""
    p = np.asarray(packed, np.uint8)
    d = np.stack([(p >> 6) & 3, (p >> 4) & 3, (p >> 2) & 3, p & 3], axis=1).reshape(-1)
    return d[:n].astype(np.uint8)


class GoldenCase:
    """A fixture written by tests/golden/make_golden.py: generator parameters + Oracle A's outputs."""

    def __init__(self, name):
        from oracle import oracle as O
        self.name = name
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.n_channels = int(z["n_channels"])
        self.n_samples = int(z["n_samples"])
        self.snr_db = float(z["snr_db"])
        self.first_channel = int(z["first_channel"])
        self.fastamp_re_only = bool(int(z["fastamp_re_only"])) if "fastamp_re_only" in z else False
        self.counts = z["counts"]
        self.dibits = [unpack2(z[f"dibits_packed_{c}"], int(self.counts[c])) for c in range(self.n_channels)]
        self.ill = [np.unpackbits(z[f"ill_packed_{c}"])[:int(self.counts[c])].astype(bool)
                    for c in range(self.n_channels)]
        self.lock_index = z["lock_index"]
        self.ref_standarderr = z["ref_standarderr"]
        self.ref_sync = z["ref_sync"]
        self.ref_syms = z["ref_syms"] if "ref_syms" in z else None
        if "iq" in z:
            self.iq = z["iq"]
        else:
            self.iq = O.generate(self.n_channels, self.n_samples, O.default_sg_params(snr_db=self.snr_db),
                                 first_channel=self.first_channel)
        self.input_matches = hashlib.sha256(self.iq.tobytes()).digest() == z["input_sha256"].tobytes()


    def assert_dibits_match(self, c: int, dibits: np.ndarray, count: int) -> int:
        """The contract against the reference's own output (Oracle A):
          * identical from the reference's lock point to the end;
          * before lock, identical except where the REFERENCE's decision sits within ill_margin of a
            quadrant boundary (a flipped quadrant touches the dibit it ends and the one it starts);
          * symbol counts within one (the tail symbol can land either side of the buffer end).
        Returns the number of differing dibits (0 in all but slow-acquiring channels)."""
        ref = self.dibits[c]
        assert abs(int(count) - len(ref)) <= 1, (self.name, c, count, len(ref))
        n = min(int(count), len(ref))
        diff = np.flatnonzero(np.asarray(dibits[:n]) != ref[:n])
        if len(diff) == 0:
            return 0
        lock = int(self.lock_index[c])
        assert diff.max() < lock, f"{self.name} ch{c}: dibit {diff.max()} differs after the reference locked at {lock}"
        ill = self.ill[c]
        near = ill.copy()
        near[1:] |= ill[:-1]
        assert near[diff].all(), f"{self.name} ch{c}: differs at well-conditioned decisions {diff[~near[diff]][:8]}"
        assert len(diff) <= max(8, lock // 500), f"{self.name} ch{c}: {len(diff)} differences before lock"
        return len(diff)


_cases = {}


def golden_case(name) -> GoldenCase:
    if name not in _cases:
        _cases[name] = GoldenCase(name)
    return _cases[name]


@pytest.fixture(scope="session")
def O():
    from oracle import oracle as mod
    mod.lib_b()
    return mod


@pytest.fixture(scope="session")
def pkg():
    import sdrpp_tetra_demodulator_b200 as p
    p.capi.lib()
    return p


def require_golden_input(case: GoldenCase):
    if not case.input_matches:
        pytest.fail(f"{case.name}: regenerated capture differs from the one the fixture was made from "
                    f"(libm drift?) -- regenerate with tests/golden/make_golden.py")
