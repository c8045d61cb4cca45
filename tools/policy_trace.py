"""Dev probe: timeline of one policy_step_kernel CTA (library built with -DPOSG_POLICY_TRACE, loaded through POSG_B200_LIB)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posggym_baselines_b200 import _lib
from posggym_baselines_b200.envs.lbf import LBFModel
from posggym_baselines_b200.ppo.collector import DeviceCollector
from tests.test_policy_gpu import RefNet
dev = torch.device("cuda")
col = DeviceCollector(LBFModel(device=dev), RefNet(300, 6).to(dev).state_dict(), 8192, rollout_length=8, seed=1)
col.capture()
h = C.CDLL(os.environ["POSG_B200_LIB"])
names = ["start", "inputs requested", "A operands in smem", "weights arrived", "gemm1 done", "trunk epilogue", "gemm2 done", "lstm epilogue", "heads epilogue (gemm3 inside)", "gemm4 done", "categorical done"]
acc = np.zeros(11)
for rep in range(20):
    col.rollout(); torch.cuda.synchronize()
    t = np.zeros(16, np.uint64); h.posg_policy_trace(t.ctypes.data_as(C.c_void_p))
    acc += (t[:11] - t[0]).astype(np.float64)
acc /= 20
for n, a, d in zip(names, acc, np.diff(np.concatenate([[0], acc]))):
    print(f"{n:34s} at {a / 1e3:6.2f} us  (+{d / 1e3:5.2f})")
