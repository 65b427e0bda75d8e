import sys; sys.path.insert(0,'.')
import numpy as np, travellingsalesmanoptimization_b200 as T
I=T.Instance(np.random.rand(100,2)*100)
for k,n in enumerate(["DADD","DMUL","DFMA","F2F pair","MUFU.RSQ","FFMA","K general","K int","F2F.F32.F64","K2 int packed"]):
    print(n, "%.3e"%I.pipe_rate(k))
