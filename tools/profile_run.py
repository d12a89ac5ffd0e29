"""Run one (channels, samples, variant) configuration a few times -- target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdrpp_tetra_demodulator_b200 as pkg
C_, N, v = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
iq, _ = pkg.synth_capture(C_, N)
dm = pkg.Demodulator(C_, N); dm.set_kernel_variant(v); dm.use_torch_stream()
out = None
for _ in range(reps):
    out = dm.process(iq, dibits=True, out=out)
torch.cuda.synchronize()
print("ms", dm.last_kernel_ms(), "Gsps", C_ * N / dm.last_kernel_ms() / 1e6)
