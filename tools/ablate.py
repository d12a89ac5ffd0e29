"""Development aid: time the ws4 pipeline with roles switched off (library built with -DTDM_ABLATE; results are wrong,
only the tick time means something).  python tools/ablate.py 4096 262144 4 "0,1,2,..." (masks: bit = Role enum)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdrpp_tetra_demodulator_b200 as pkg
C_, N, v = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
names = ["LOOP", "MID", "TIMING", "COSTAS", "SLICER", "AGC", "PFAR", "QFAR", "RRCA", "RRCB"]
iq, _ = pkg.synth_capture(C_, N)
dm = pkg.Demodulator(C_, N); dm.set_kernel_variant(v); dm.use_torch_stream()
out = None
for m in sys.argv[4].split(","):
    mask = int(m, 0)
    os.environ["TDM_DEBUG_MASK"] = str(mask)
    best = 1e9
    for _ in range(2):
        dm.reset_all()
        out = dm.process(iq, dibits=True, out=out)
        torch.cuda.synchronize()
        best = min(best, dm.last_kernel_ms())
    off = [n for i, n in enumerate(names) if mask >> i & 1]
    # the far and RRC roles are switched by the bit of their first warp (PFAR switches both far warps, RRCA both RRC warps)
    print(f"mask {mask:#06x} off={off}: {best:.2f} ms  {best * 1e-3 * 1.965e9 / (N / 8):.0f} cycles/tick", flush=True)
