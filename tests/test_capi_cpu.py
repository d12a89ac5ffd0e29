"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/tdm_b200.h declares,
its host-only entry points work, and it refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions(name="tdm_b200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tdm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.capi.lib()
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/tdm_b200.h but not exported"
    assert set(declared) == set(pkg.capi.EXPORTED_SYMBOLS)
    assert L.tdm_abi_version() == 2
    burst = _header_functions("tdm_burst_b200.h")
    assert set(burst) == set(pkg.capi.EXPORTED_BURST_SYMBOLS), set(burst) ^ set(pkg.capi.EXPORTED_BURST_SYMBOLS)
    for name in burst:
        assert hasattr(L, name), f"{name} is declared in include/tdm_burst_b200.h but not exported"
    chan = _header_functions("tdm_chan_b200.h")
    assert len(chan) == 9
    for name in chan:
        assert hasattr(L, name), f"{name} is declared in include/tdm_chan_b200.h but not exported"
    assert sorted(os.listdir(os.path.join(ROOT, "include"))) == ["tdm_b200.h", "tdm_burst_b200.h", "tdm_chan_b200.h"]


def test_library_has_sm100a_code_only(pkg):
    """the product is built for sm_100a (cuobjdump lists the embedded ELF architectures)"""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", pkg.capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_library_does_not_link_the_oracle(pkg):
    """the product path must not depend on anything under oracle/"""
    import subprocess
    out = subprocess.run(["ldd", pkg.capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "tetra_ref" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "sdrpp_tetra_demodulator_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(root, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle_b.h" not in text, f


def test_create_without_gpu_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.TdmError) as e:
        pkg.Demodulator(4, 1024)
    assert e.value.code == pkg.capi.TDM_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_argument_errors(pkg):
    L = pkg.capi.lib()
    h = C.c_void_p()
    assert L.tdm_create(None, 0, 1024, 0, C.byref(h)) == pkg.capi.TDM_ERR_ARG
    assert L.tdm_create(None, 4, 0, 0, C.byref(h)) == pkg.capi.TDM_ERR_ARG
    assert L.tdm_create(None, 4, 1024, 0, None) == pkg.capi.TDM_ERR_ARG
    cfg = pkg.default_config()
    cfg.rrc_tap_count = 0
    assert L.tdm_create(C.byref(cfg), 4, 1024, 0, C.byref(h)) == pkg.capi.TDM_ERR_UNSUPPORTED
    assert L.tdm_process(None, None, 0, 0, None, None, None, 0, None, 0, 0) == pkg.capi.TDM_ERR_ARG
    assert L.tdm_destroy(None) == pkg.capi.TDM_OK
    assert b"null handle" in L.tdm_last_error()


def test_short_filter_is_zero_padded_at_the_old_end(pkg):
    cfg = pkg.default_config()
    cfg.rrc_tap_count = 33
    d = pkg.design_from_config(cfg)
    rrc = np.array(d.rrc[:])
    assert d.ntaps == 33 and np.all(rrc[:32] == 0) and np.all(rrc[32:] != 0)
    assert np.all(np.array(d.be_a[:32]) == 0) and np.all(np.array(d.be_b[:32]) == 0)


def test_channel_partition():
    from sdrpp_tetra_demodulator_b200.sharding import channel_range
    for C_, W in [(4096, 8), (4096, 1), (10, 4), (3, 8)]:
        ranges = [channel_range(r, W, C_) for r in range(W)]
        assert ranges[0][0] == 0 and ranges[-1][1] == C_
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(W - 1))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_burst_sync_refuses_without_gpu_and_checks_arguments(pkg):
    import torch
    L = pkg.capi.lib()
    h = C.c_void_p()
    assert L.tdm_bsync_create(0, 100, 0, C.byref(h)) == pkg.capi.TDM_ERR_ARG
    assert L.tdm_bsync_create(4, 0, 0, C.byref(h)) == pkg.capi.TDM_ERR_ARG
    assert L.tdm_bsync_in(None, None, 0, None, 0, 0, 432, None, 0, None, 0, 0) == pkg.capi.TDM_ERR_ARG
    assert L.tdm_bsync_destroy(None) == pkg.capi.TDM_OK
    if not torch.cuda.is_available():
        assert L.tdm_bsync_create(4, 100, 0, C.byref(h)) == pkg.capi.TDM_ERR_NO_DEVICE
        assert b"no CPU fallback" in L.tdm_last_error()


def test_burst_demux_is_host_only_and_matches_the_restatement(pkg):
    """tdm_burst_demux (no GPU involved) against oracle_bsync.c's split for the three burst types"""
    from oracle import oracle_bsync as B
    B.build()
    rng = np.random.default_rng(4)
    P = B.PortBsync(1)
    for typ, bits in [(B.TRAIN_SYNC, B.sync_burst(rng)), (B.TRAIN_NORM_1, B.norm_burst(rng, False)), (B.TRAIN_NORM_2, B.norm_burst(rng, True))]:
        rec = np.zeros(1, dtype=pkg.capi.BURST_UNPACKED_DTYPE)       # same layout as the checker's record
        rec["train_seq"] = typ
        rec["bits"][0, :510] = bits
        mine, ref = pkg.burst_demux(rec[0]), P.demux(rec[0])         # burst_demux packs it for the ABI
        packed = np.zeros(1, dtype=pkg.capi.BURST_DTYPE)
        packed["bits"][0] = np.packbits(rec["bits"][0]).view(">u4").astype(np.uint32)
        back = np.zeros(510, dtype=np.uint8)
        assert pkg.capi.lib().tdm_burst_unpack(packed.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(back, bits)
        assert len(mine) == len(ref) and len(mine) in (2, 3)
        for f in ["type", "blk_num", "n_bits", "bits"]:
            assert np.array_equal(mine[f], ref[f]), (typ, f)


def test_headers_are_plain_c_and_the_example_fails_loudly_without_a_gpu(pkg, tmp_path):
    """the boundary is a C ABI: both headers must compile as C11 with -pedantic, and a plain-C host linking the
    library stops at tdm_create with TDM_ERR_NO_DEVICE when there is no B200 (no CPU fallback behind the ABI)"""
    import shutil
    import subprocess
    import torch
    if not shutil.which("gcc"):
        pytest.skip("gcc not on PATH")
    exe = tmp_path / "demod_then_bursts"
    libdir = os.path.dirname(pkg.capi.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "demod_then_bursts.c"), "-L", libdir, "-ltdm_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        return
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


def _build_sharded_gather(pkg, tmp_path):
    import subprocess
    exe = tmp_path / "sharded_gather"
    libdir = os.path.dirname(pkg.capi.LIB_PATH)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                        os.path.join(ROOT, "examples", "sharded_gather.c"), "-L", libdir, "-ltdm_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                        f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_multi_gpu_example_builds_and_fails_loudly_without_a_gpu(pkg, tmp_path):
    """examples/sharded_gather.c: the multi-GPU epilogue (tdm_comm_*, tdm_gather_packed, tdm_unpack_dibits) from plain C"""
    import shutil
    import subprocess
    import torch
    if not shutil.which("gcc") or not os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h"):
        pytest.skip("gcc or the CUDA headers are missing")
    exe = _build_sharded_gather(pkg, tmp_path)
    if torch.cuda.is_available():
        return
    r = subprocess.run([str(exe), "0", "1", str(tmp_path / "id")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_multi_gpu_example_runs(pkg, tmp_path):
    """one rank per visible GPU (two at most): every rank demodulates its shard, rank 0 receives all rows over NCCL and its
    own channels equal the transmitted dibits after lock"""
    import subprocess
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = _build_sharded_gather(pkg, tmp_path)
    world = min(2, torch.cuda.device_count())
    # single node, one or two ranks: keep NCCL's bootstrap on the loopback interface and skip the InfiniBand / NVLS /
    # multi-node-NVLink probing (on some boxes the system libnccl this plain-C process loads spent two minutes in its
    # initialisation; the container's hostname may not resolve)
    env = dict(os.environ, NCCL_SOCKET_IFNAME="lo", NCCL_IB_DISABLE="1", NCCL_NVLS_ENABLE="0", NCCL_MNNVL_ENABLE="0")
    procs = [subprocess.Popen([str(exe), str(r), str(world), str(tmp_path / "id"), "64", "60000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
             for r in range(world)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, (o, e)
    assert f"rank 0 holds {64 * world} channels" in outs[0][0] and " 0 of its own 64 channels differ" in outs[0][0], outs[0][0]


def _build_example(pkg, tmp_path, name):
    import subprocess
    exe = tmp_path / name
    libdir = os.path.dirname(pkg.capi.LIB_PATH)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                        os.path.join(ROOT, "examples", name + ".c"), "-L", libdir, "-ltdm_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                        f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_wideband_example_builds_and_fails_loudly_without_a_gpu(pkg, tmp_path):
    """examples/wideband_chain.c: int16 capture -> channeliser -> demodulator from plain C"""
    import shutil
    import subprocess
    import torch
    if not shutil.which("gcc") or not os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h"):
        pytest.skip("gcc or the CUDA headers are missing")
    exe = _build_example(pkg, tmp_path, "wideband_chain")
    if torch.cuda.is_available():
        return
    (tmp_path / "empty.cs16").write_bytes(b"")
    r = subprocess.run([str(exe), str(tmp_path / "empty.cs16"), "4"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_wideband_example_finds_the_carriers(O, pkg, tmp_path):
    """three TETRA carriers on a 144-channel raster, written as an int16 capture file: the C host reports exactly those
    channels (and no empty one) as locked"""
    import subprocess
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import oracle_chan as OC
    exe = _build_example(pkg, tmp_path, "wideband_chain")
    M, D, n = 144, 100, 50000
    carriers = [3, 40, 100]
    sp = O.default_sg_params(snr_db=60.0, max_freq_off_hz=200.0, min_amp=0.5, max_amp=1.0)
    nb = O.generate(len(carriers), n, sp)
    wide = OC.place_on_raster(nb[..., 0] + 1j * nb[..., 1], carriers, M, D)
    rng = np.random.default_rng(2)
    wide += 1e-3 * (rng.standard_normal(len(wide)) + 1j * rng.standard_normal(len(wide)))
    scale = 20000.0 / np.abs(np.concatenate([wide.real, wide.imag])).max()
    cs16 = np.stack([wide.real, wide.imag], axis=1) * scale
    np.round(cs16).astype(np.int16).tofile(tmp_path / "capture.cs16")
    r = subprocess.run([str(exe), str(tmp_path / "capture.cs16"), "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    locked = sorted(int(line.split()[1]) for line in r.stdout.splitlines() if line.startswith("channel "))
    assert set(carriers) <= set(locked), r.stdout
    assert not (set(locked) & {20, 50, 70, 120}), r.stdout
    assert f"{n * D} wideband samples, {M} channels" in r.stdout
