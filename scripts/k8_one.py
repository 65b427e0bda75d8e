"""One K8 batch for an ncu capture: python scripts/k8_one.py INSTANCE STARTS [CL] (CL: TSPCUDA_K8_CLUSTER)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 3:
    os.environ["TSPCUDA_K8_CLUSTER"] = sys.argv[3]
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance
xy, _ = load_instance(sys.argv[1])
with T.Instance(xy) as I:
    t0 = time.perf_counter(); r = I.multistart_two_opt(0, int(sys.argv[2]), 1); dt = time.perf_counter() - t0
    print(sys.argv[1:], "%.3f s, kernel %.3f s, best %d @ %d" % (dt, I.stats()["pruned_kernel_ms"] * 1e-3, r[2], r[0]))
