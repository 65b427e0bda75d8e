"""In-library multi-GPU multi-start (tspcuda_mgpu, what TSPCUDA_GPUS selects in the drop-in CLI): NN + 2-opt over
S starts sharded over every visible GPU vs the same starts on one GPU; NCCL incumbent reduction time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance
G = T.device_count()
for name, per_gpu in [("usa13509", 592), ("pla85900", 148)]:
    xy, _ = load_instance(name)
    S = per_gpu * 2          # strong scaling: the SAME S starts on 1 GPU and on all of them (S >> SMs per GPU needed to fill them)
    with T.MultiGpuInstance(xy, ndev=G) as M, T.MultiGpuInstance(xy, ndev=1) as M1:
        M.multistart(True, 0, G * 2, 1); M1.multistart(True, 0, 2, 1)          # warm: kernels loaded, NCCL channels up
        t0 = time.perf_counter(); bs, bp, bc, tc, tr = M.multistart(True, 0, S, 1); dt = time.perf_counter() - t0
        t0 = time.perf_counter(); b1, p1, c1, tc1, tr1 = M1.multistart(True, 0, S, 1); dt1 = time.perf_counter() - t0
        assert (bs, bc) == (b1, c1) and np.array_equal(bp, p1)
        print("%-9s %d starts: %d GPUs %.3f s (slowest shard %.3f s, NCCL reduction %.3f ms) | 1 GPU %.3f s | speed-up %.2fx | best %d @ start %d" % (
            name, S, G, dt, tc * 1e-3, tr, dt1, dt1 / dt, bc, bs), flush=True)
