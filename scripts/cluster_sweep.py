"""K8 with thread-block clusters: S starts (fewer than SMs) of NN + 2-opt, one descent per cluster of CL CTAs, against
CL = 1 (one CTA per descent) and against K7 (whole GPU, one start after the other).  Same tours in every mode."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance
cases = {"usa13509": (9, 18, 37, 74), "d15112": (18, 74), "pla85900": (18, 74)}
for name in sys.argv[1:] or list(cases):
    xy, _ = load_instance(name)
    for S in cases[name]:
        ref = None
        line = "%-9s %3d starts:" % (name, S)
        for cl in (0, 1, 2, 4, 8):
            if cl * S > 148:
                continue
            os.environ["TSPCUDA_K8_CLUSTER"] = str(cl)
            with T.Instance(xy) as I:
                I.multistart_two_opt(0, 2, 1)
                t0 = time.perf_counter(); r = I.multistart_two_opt(0, S, 1); dt = time.perf_counter() - t0
            key = (r[0], r[2], r[1].tobytes(), r[4].tobytes(), r[5].tobytes())
            if ref is None:
                ref = key
            assert key == ref, (name, S, cl)
            line += "  %s %.3f s" % ("K7 serial" if cl == 0 else "CL=%d" % cl, dt)
        print(line, flush=True)
os.environ.pop("TSPCUDA_K8_CLUSTER", None)
