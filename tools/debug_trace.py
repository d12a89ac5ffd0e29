import os, sys, shutil
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import sdrpp_tetra_demodulator_b200.capi as capi
capi.LIB_PATH = os.path.join(root, "gpurun_dbg_libtdm.so")
import numpy as np, torch
import sdrpp_tetra_demodulator_b200 as pkg
from oracle import oracle as O
iq = np.ascontiguousarray(O.generate(1, 4096)[:, :63])
dm = pkg.Demodulator(1, 63)
r = dm.process(torch.from_numpy(iq).cuda(), symbols=True, dibits=True)
torch.cuda.synchronize()
