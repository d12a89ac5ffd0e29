"""CPU tests: the two oracles against each other, against the golden fixtures made from the
reference itself, and the host-side design code against the reference's own tap tables."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden_case, require_golden_input

GOLDEN_CASES = ["small_c2_n4096", "batch_c8_n60000_snr30", "batch_c4_n60000_snr20", "cfg1_c1_n1e6_snr30",
                "batch_c8_n60000_snr30_reonly", "batch_c4_n60000_snr20_reonly", "cfg1_c1_n1e6_snr30_reonly"]


def test_state_struct_sizes(O, pkg):
    assert O.STATE_DTYPE.itemsize == 720 == pkg.capi.STATE_DTYPE.itemsize
    assert C.sizeof(O.TdmDesign) == C.sizeof(pkg.TdmDesign)
    assert C.sizeof(O.TdmConfig) == C.sizeof(pkg.TdmConfig) == 80


def _design_arrays(d):
    return (np.array(d.rrc[:], np.float32), np.array(d.be_a[:], np.float32), np.array(d.be_b[:], np.float32),
            np.array(d.bank, np.float32).reshape(128, 8))


def test_design_matches_reference_fixture(O, pkg):
    """Product design code and Oracle B's independent restatement both reproduce, bit for bit, the tables
    read out of the reference's own objects (tests/golden/design_default.npz, written by Oracle A)."""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "design_default.npz"))
    for d in (pkg.design_from_config(pkg.default_config()), O.OracleB(1).design):
        rrc, a, b, bank = _design_arrays(d)
        assert np.array_equal(rrc, z["rrc"])
        assert np.array_equal(a, z["hbe"][:, 0]) and np.array_equal(b, z["hbe"][:, 1])
        # lower band-edge taps are the exact conjugate (src/dsp/fll.cpp:89-93)
        assert np.array_equal(a, z["lbe"][:, 0]) and np.array_equal(-b, z["lbe"][:, 1])
        assert np.array_equal(bank, z["bank"])
        co = z["coeffs"]  # fll a,b,min,max | timing a,b,min,max | costas a,b,min,max | agc rate,set,max,init
        got = np.array([0.0, d.fll_beta, d.fll_min_freq, d.fll_max_freq, d.tr_alpha, d.tr_beta, d.tr_min_omega,
                        d.tr_max_omega, d.costas_alpha, d.costas_beta, d.costas_min_freq, d.costas_max_freq,
                        d.agc_rate, d.agc_set_point, d.agc_max_gain, d.agc_init_gain], np.float32)
        assert np.array_equal(got, co)


def test_design_matches_live_reference(O, pkg):
    if not O.have_ref():
        pytest.skip("oracle/_ref not built here")
    a = O.OracleA(1)
    nt, rrc, lbe, hbe, bank, P, T = a.taps()
    d = pkg.design_from_config(pkg.default_config())
    r2, a2, b2, bank2 = _design_arrays(d)
    assert (nt, P, T) == (65, 128, 8)
    assert np.array_equal(r2, rrc) and np.array_equal(a2, hbe[:, 0]) and np.array_equal(b2, hbe[:, 1])
    assert np.array_equal(bank2, bank)
    assert abs(float(rrc.sum()) - 1.0001) < 2e-4 and abs(float(rrc[32]) - 0.547817) < 1e-5   # SURVEY.md A.6 probe


def test_default_config_matches_plugin_constants(O, pkg):
    """src/main.cpp:35-44,78-84"""
    for cfg in (pkg.default_config(), O.OracleB.default_config()):
        assert (cfg.symbolrate, cfg.samplerate, cfg.rrc_tap_count) == (18000.0, 36000.0, 65)
        assert cfg.rrc_beta == float(np.float32(0.35)) and cfg.agc_rate == float(np.float32(0.02))
        assert cfg.costas_bandwidth == float(np.float32(0.01)) and cfg.fll_bandwidth == float(np.float32(0.006))
        assert abs(cfg.mu_gain - 0.017603) < 1e-6 and abs(cfg.omega_gain - 1.5636e-4) < 1e-8
        assert cfg.omega_rel_limit == float(np.float32(0.02))


def test_design_rejects_unsupported_tap_counts(pkg):
    cfg = pkg.default_config()
    cfg.rrc_tap_count = 66
    with pytest.raises(pkg.TdmError) as e:
        pkg.design_from_config(cfg)
    assert e.value.code == pkg.capi.TDM_ERR_UNSUPPORTED


def test_canonical_sincos_accuracy(O):
    """the shared sin/cos must be a faithful stand-in for cosf/sinf: <= 2 ulp-ish absolute error on [-2pi, 2pi]"""
    L = O.lib_b()
    xs = np.linspace(-2 * np.pi, 2 * np.pi, 200001).astype(np.float32)
    s, c = C.c_float(), C.c_float()
    worst = 0.0
    for x in xs[::7]:
        L.ob_sincos(float(x), C.byref(s), C.byref(c))
        worst = max(worst, abs(s.value - np.sin(np.float64(x))), abs(c.value - np.cos(np.float64(x))))
    assert worst < 2.5e-7, worst


def test_canonical_fll_nco_accuracy(O):
    """The FLL's NCO (range reduction prepared one sample ahead, oracle_b.c ob_fll_*) must stay a faithful stand-in
    for math::phasor's cosf/sinf of the wrapped loop phase (src/dsp/fll.cpp:137) for EVERY state the loop can be
    in: any phase, any frequency within the +-pi/2 clamp, any jump of the frequency (fallback path)."""
    L = O.lib_b()
    rng = np.random.default_rng(11)
    s, c = C.c_float(), C.c_float()
    worst_near, worst_far, worst_small_f = 0.0, 0.0, 0.0
    n0 = L.ob_fll_fallback_count()
    pi_f = np.float32(3.1415926535)
    two_pi = np.float32(pi_f - (-pi_f))
    for i in range(60000):
        phi = np.float32(rng.uniform(-pi_f, pi_f))
        f_old = np.float32(rng.uniform(-1.5707, 1.5707) if i % 3 else rng.normal(0, 0.05))
        jump = [1e-4, 1e-2, 0.5][i % 3]
        f_new = np.float32(np.clip(f_old + rng.normal(0, jump), -1.5707963, 1.5707963))
        L.ob_fll_nco(float(phi), float(f_old), float(f_new), C.byref(s), C.byref(c))
        # two equally legitimate targets: the phase as the reference holds it (float sum, wrapped by the float 2 pi:
        # what the classic fall-back path evaluates), and the unrounded sum (what the prepared reduction evaluates:
        # it never rounds phi + f, and a wrap costs it nothing).  They differ by up to 3e-7 rad themselves.
        ph = np.float32(phi + f_new)
        ph = np.float32(ph - two_pi) if ph > pi_f else (np.float32(ph + two_pi) if ph < -pi_f else ph)
        ex = np.float64(phi) + np.float64(f_new)
        e_ref = max(abs(s.value - np.sin(np.float64(ph))), abs(c.value - np.cos(np.float64(ph))))
        e_ex = max(abs(s.value - np.sin(ex)), abs(c.value - np.cos(ex)))
        worst_near = max(worst_near, min(e_ref, e_ex))
        worst_far = max(worst_far, max(e_ref, e_ex))
        if abs(f_old) < 0.3 and abs(f_new) < 0.3:       # +-1.7 kHz at 36 kS/s: where a locked loop lives
            worst_small_f = max(worst_small_f, min(e_ref, e_ex))
    # r = r0 + f cancels two numbers of size |f| + pi/4: half an ulp of that is the price of the short recurrence
    assert worst_small_f < 1.2e-7, worst_small_f
    assert worst_near < 2.5e-7, worst_near
    assert worst_far < 6.5e-7, worst_far
    assert L.ob_fll_fallback_count() - n0 > 3000       # the big jumps went through the classic reduction


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_b_bits_equal_reference_golden(O, name):
    """Decoded dibits of the canonical-order restatement == the reference's own chain (fixture from Oracle A)."""
    g = golden_case(name)
    require_golden_input(g)
    b = O.OracleB(g.n_channels, fastamp_re_only=g.fastamp_re_only)
    counts, syms, dibits, _ = b.process(g.iq, nthreads=4)
    for c in range(g.n_channels):
        g.assert_dibits_match(c, dibits[c], counts[c])
        assert abs(b.states[c]["standarderr"] - g.ref_standarderr[c]) < 5e-3
    if g.ref_syms is not None:
        n = g.ref_syms.shape[1]
        err = np.abs(syms[:, :n] - g.ref_syms).max()
        assert err < 0.15   # float trajectories are NOT expected to match the reference's (chaotic at 1 ulp)


@pytest.mark.parametrize("name", GOLDEN_CASES[:3] + GOLDEN_CASES[4:6])
def test_live_reference_reproduces_golden(O, name):
    """Where oracle/_ref exists, it must reproduce the committed fixture exactly (the fixture is its output)."""
    g = golden_case(name)
    if not O.have_ref(g.fastamp_re_only):
        pytest.skip("oracle/_ref not built here")
    require_golden_input(g)
    a = O.OracleA(g.n_channels, fastamp_re_only=g.fastamp_re_only)
    counts, _, dibits, bits = a.process(g.iq, want_syms=False, want_bits=True)
    assert np.array_equal(counts, g.counts)
    for c in range(g.n_channels):
        n = int(counts[c])
        assert np.array_equal(dibits[c, :n], g.dibits[c])
        # BitUnpacker, src/dsp/bit_unpacker.cpp:6-7
        assert np.array_equal(bits[c, 0:2 * n:2], dibits[c, :n] >> 1) and np.array_equal(bits[c, 1:2 * n:2], dibits[c, :n] & 1)
    a.close()


@pytest.mark.parametrize("chunk", [32768, 4097, 7])
def test_oracle_b_chunk_invariance(O, chunk):
    """SURVEY.md [PROBE]: the reference is bit-identical for chunk sizes 1e6/32768/4097/7; so is the restatement."""
    iq = O.generate(2, 50001)
    one = O.OracleB(2)
    c1, s1, d1, _ = one.process(iq)
    many = O.OracleB(2)
    outs, outd, tot = [], [], np.zeros(2, int)
    for n0 in range(0, iq.shape[1], chunk):
        c, s, d, _ = many.process(np.ascontiguousarray(iq[:, n0:n0 + chunk]))
        outs.append([s[k, :c[k]] for k in range(2)])
        outd.append([d[k, :c[k]] for k in range(2)])
        tot += c
    assert np.array_equal(tot, c1)
    for k in range(2):
        assert np.array_equal(np.concatenate([o[k] for o in outd]), d1[k, :c1[k]])
        assert np.array_equal(np.concatenate([o[k] for o in outs]).view(np.uint32), s1[k, :c1[k]].view(np.uint32))
    for f in O.EXACT_STATE_FIELDS + O.METRIC_STATE_FIELDS:
        assert np.array_equal(one.states[f], many.states[f]), f


def test_reference_chunk_invariance(O):
    if not O.have_ref():
        pytest.skip("oracle/_ref not built here")
    iq = O.generate(1, 30011)
    a1 = O.OracleA(1)
    c1, _, d1, _ = a1.process(iq, want_syms=False)
    a2 = O.OracleA(1)
    ds = []
    for n0 in range(0, iq.shape[1], 4097):
        c, _, d, _ = a2.process(np.ascontiguousarray(iq[:, n0:n0 + 4097]), want_syms=False)
        ds.append(d[0, :c[0]])
    assert np.array_equal(np.concatenate(ds), d1[0, :c1[0]])


def test_tx_rx_roundtrip(O):
    """Self-consistency: decoded dibits == transmitted dibits after a fixed lag once locked."""
    C_, N = 6, 80000
    iq = O.generate(C_, N)
    b = O.OracleB(C_)
    counts, _, dibits, bits = b.process(iq, want_syms=False, want_bits=True, nthreads=4)
    for c in range(C_):
        n = int(counts[c])
        tx = O.tx_dibits(c, n + 64)
        errs = [np.count_nonzero(dibits[c, lag + 20000:n] != tx[20000:n - lag]) for lag in range(10, 30)]
        assert min(errs) == 0, (c, min(errs))
        assert np.array_equal(bits[c, 0:2 * n:2], dibits[c, :n] >> 1)
    assert b.states["sync"].all()


def test_generator_mapping_is_bits2phase(O):
    """src/decoder/src/phy/tetra_burst.c:99-104: 00->+pi/4, 01->+3pi/4, 11->-3pi/4, 10->-pi/4, i.e. phase
    increments of 1,3,5,7 eighth-turns; the generator's dibits must be uniform over the four values."""
    d = O.tx_dibits(3, 40000)
    hist = np.bincount(d, minlength=4) / len(d)
    assert set(np.unique(d)) == {0, 1, 2, 3} and np.all(np.abs(hist - 0.25) < 0.02)


def test_setter_contract_matches_the_reference(O):
    """tdm_set_params' contract (restated in OracleB.set_params) against the reference driven through ITS OWN setters
    (oracle/_ref, ref_driver.cpp tref_set): after the same sequence every loop coefficient, the timing restart and the
    decoded dibits agree."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built here")
    C_, N1, N2 = 3, 30000, 60000
    iq = O.generate(C_, N1 + N2)
    a, b = np.ascontiguousarray(iq[:, :N1]), np.ascontiguousarray(iq[:, N1:])
    ob, oa, ocfg = O.OracleB(C_), O.OracleA(C_), O.OracleB.default_config()
    ob.process(a)
    oa.process(a, want_syms=False)
    for field, value, what, code in [("agc_rate", 0.01, 4, 4), ("costas_bandwidth", 0.015, 8, 5), ("fll_bandwidth", 0.004, 16, 6),
                                     ("samplerate", 36000.0, 1, 2)]:
        setattr(ocfg, field, value)
        ob.set_params(ocfg, what)
        oa.set(code, value)
    ocfg.omega_gain, ocfg.mu_gain, ocfg.omega_rel_limit = ocfg.omega_gain * 1.5, ocfg.mu_gain * 1.5, 0.03
    ob.set_params(ocfg, 32)
    oa.set(7, ocfg.omega_gain, ocfg.mu_gain, 0.03)
    d = ob.design
    got = np.array([0, d.fll_beta, d.fll_min_freq, d.fll_max_freq, d.tr_alpha, d.tr_beta, d.tr_min_omega, d.tr_max_omega, d.costas_alpha,
                    d.costas_beta, d.costas_min_freq, d.costas_max_freq, d.agc_rate, d.agc_set_point, d.agc_max_gain, d.agc_init_gain], np.float32)
    assert np.array_equal(got, oa.coeffs())
    st = oa.loop_state(0)
    assert (st.tr_mu, st.tr_omega, st.tr_offset) == (0.0, 2.0, 0) == (float(ob.states["tr_mu"][0]), float(ob.states["tr_omega"][0]), int(ob.states["tr_offset"][0]))
    # an RRC redesign through setRRCParams leaves the band-edge filters alone in the reference; so does the contract
    ocfg.rrc_beta = 0.5
    ob.set_params(ocfg, 2)
    oa.set(3, 65, 0.5)
    nt, rrc, lbe, hbe, _, _, _ = oa.taps()
    assert np.array_equal(np.array(ob.design.rrc[:], np.float32), rrc)
    assert np.array_equal(np.array(ob.design.be_a[:], np.float32), hbe[:, 0]) and np.array_equal(np.array(ob.design.be_b[:], np.float32), hbe[:, 1])
    cb, _, db, _ = ob.process(b)
    ca, _, da, _ = oa.process(b, want_syms=False)
    for c in range(C_):
        n = min(int(ca[c]), int(cb[c]))
        diff = np.flatnonzero(db[c, :n] != da[c, :n])
        assert len(diff) == 0 or diff.max() < n // 2
    oa.close()


def test_reset_contract_matches_the_reference(O):
    """tdm_reset's contract (restated in OracleB.reset) against the reference's own PI4DQPSK::reset()
    (src/dsp/pi4dqpsk.cpp:120-130): loops restart from their initial values, both acquire again and the decoded dibits
    agree from the reference's lock point on.  (The reference keeps the band-edge filters' delay line across a reset
    and clears the matched filter's; the product has ONE delay line for both and clears it -- a 64-sample transient,
    before either has locked.)"""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built here")
    C_, N = 8, 60000
    iq = O.generate(C_, N)
    oa, ob = O.OracleA(C_), O.OracleB(C_)
    oa.process(iq[:, :20000], want_syms=False)
    ob.process(iq[:, :20000])
    oa.reset()
    ob.reset()
    st = oa.loop_state(0)
    assert (st.tr_mu, st.tr_omega, st.tr_offset) == (0.0, 2.0, 0)
    assert float(ob.states["agc_gain"][0]) == 1.0 and float(ob.states["fll_phase"][0]) == 0.0
    ca, _, da, _ = oa.process(iq, want_syms=False)
    cb, _, db, _ = ob.process(iq)
    for c in range(C_):
        assert ca[c] == cb[c]
        diff = np.flatnonzero(da[c, :ca[c]] != db[c, :ca[c]])
        assert len(diff) == 0 or diff.max() < ca[c] // 2, (c, diff[-5:])      # slowest channel of this capture locks near 9100
    oa.close()


def test_simd_timing_build_of_the_reference_decodes_the_same_dibits(O):
    """oracle/_ref/libtetra_ref_simd.so -- the reference with 256-bit FMA dot products in the VOLK stand-in, the build
    bench.py TIMES as the CPU baseline -- against the generic-order build that pins parity: same symbol counts, same
    dibits from the lock point on (the floats differ in the last places, as real VOLK's do between machines)."""
    if not O.have_ref() or not O.have_ref_simd():
        pytest.skip("oracle/_ref SIMD build not present or host without AVX2+FMA")
    g = golden_case("batch_c8_n60000_snr30")
    a = O.OracleA(g.n_channels, simd=True)
    counts, _, dibits, _ = a.process(g.iq, want_syms=False)
    for c in range(g.n_channels):
        g.assert_dibits_match(c, dibits[c], counts[c])
    a.close()
