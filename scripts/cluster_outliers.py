"""Repeatability of the K8 cluster timings (profiles/r01_cluster.md): (starts, CL) pairs on usa13509, three runs each,
wall clock and the kernel's own event time; then the library's default choice for a range of start counts."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance
os.environ["TSPCUDA_VERBOSE"] = "1"
xy, _ = load_instance(sys.argv[1] if len(sys.argv) > 1 else "usa13509")
for S, cl in ((9, 1), (9, 2), (9, 4), (37, 2), (37, 4), (36, 4), (32, 4), (18, 8), (16, 8)):
    os.environ["TSPCUDA_K8_CLUSTER"] = str(cl)
    line = "%3d starts CL=%d:" % (S, cl)
    with T.Instance(xy) as I:
        I.multistart_two_opt(0, 2, 1)
        for rep in range(3):
            I.reset_stats()
            t0 = time.perf_counter(); r = I.multistart_two_opt(0, S, 1); dt = time.perf_counter() - t0
            line += "  %.3f s (kernel %.3f)" % (dt, I.stats()["pruned_kernel_ms"] * 1e-3)
    print(line, flush=True)
os.environ.pop("TSPCUDA_K8_CLUSTER", None)
with T.Instance(xy) as I:
    I.multistart_two_opt(0, 2, 1)
    for S in (12, 16, 18, 24, 32, 36, 37, 48, 64, 74, 100, 148):
        I.reset_stats()
        t0 = time.perf_counter(); r = I.multistart_two_opt(0, S, 1); dt = time.perf_counter() - t0
        print("default %3d starts: %.3f s (kernel %.3f)" % (S, dt, I.stats()["pruned_kernel_ms"] * 1e-3), flush=True)
