import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ethicaltrajectoryplanning_b200 import workloads
workloads.closed_loop(20, "weights_ethical.json", "cfg1")
pr = cProfile.Profile()
pr.enable()
workloads.closed_loop(300, "weights_ethical.json", "cfg1")
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
