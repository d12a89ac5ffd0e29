import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import sdrpp_tetra_demodulator_b200.capi as capi
capi.LIB_PATH = os.path.join(root, "gpurun_dbg_libtdm.so")
import torch
import sdrpp_tetra_demodulator_b200 as pkg
C_, N = int(sys.argv[1]), int(sys.argv[2])
iq, _ = pkg.synth_capture(C_, N)
dm = pkg.Demodulator(C_, N); dm.set_kernel_variant(int(sys.argv[3]) if len(sys.argv) > 3 else 5)
dm.process(iq, dibits=True); torch.cuda.synchronize()
print("ms", dm.last_kernel_ms())
