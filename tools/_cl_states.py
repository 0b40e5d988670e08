import sys
sys.path.insert(0, "/root/repo")
from ethicaltrajectoryplanning_b200 import workloads
prec = sys.argv[1]
_, st = workloads.closed_loop(12, "weights_ethical.json", "cfg1", materialize=False, precision=prec)
for k, s in enumerate(st):
    print(k, "%.9f %.9f %.9f %d" % s)
