"""Per-statement wall time of the closed loop's Python methods (development aid, GPU box: python tools/closed_loop_lineprof.py [cfg1|cfg2])."""
import inspect, os, sys, textwrap, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ethicaltrajectoryplanning_b200 import scoring, workloads
from ethicaltrajectoryplanning_b200.planner.Frenet import frenet_planner as fp

ACC = {}
LAST = [0.0]


def _T(tag):
    now = time.perf_counter()
    ACC.setdefault(tag, []).append(now - LAST[0])
    LAST[0] = time.perf_counter()


def instrument(cls, name, module):
    fn = getattr(cls, name)
    src = textwrap.dedent(inspect.getsource(fn))
    lines = src.split("\n")
    body_indent = None
    out = []
    depth_open = 0
    for i, l in enumerate(lines):
        out.append(l)
        if i == 0 or not l.strip():
            continue
        if l.startswith("def ") or l.lstrip().startswith('"""') and body_indent is None:
            continue
        ind = len(l) - len(l.lstrip())
        if body_indent is None and ind > 0 and not l.lstrip().startswith(('"""', "#")):
            body_indent = ind
        nxt = lines[i + 1] if i + 1 < len(lines) else ""
        nind = len(nxt) - len(nxt.lstrip()) if nxt.strip() else body_indent
        s = l.rstrip()
        if body_indent is not None and nind == body_indent and not s.endswith((":", ",", "(", "\\", "[", "{")) and \
                not nxt.lstrip().startswith(("else", "elif", "except", "finally", ")", "]", "}")) and \
                s.count("(") - s.count(")") <= 0 and '"""' not in s:
            # only if parentheses are balanced over the statement so far (approximation: line-level)
            out.append(" " * body_indent + f"_T({(name, i, l.strip()[:70])!r})")
    ns = dict(vars(module))
    ns["_T"] = _T
    try:
        exec("\n".join(out), ns)
    except SyntaxError as e:
        print("could not instrument", name, e)
        return
    setattr(cls, name, ns[name])


instrument(fp.FrenetPlanner, "_step_planner_timed", fp)
instrument(fp.FrenetPlanner, "_predictions", fp)
instrument(fp.FrenetPlanner, "_step_fused", fp)
instrument(scoring.ScoringContext, "plan_cycle", scoring)
instrument(scoring.ScoringContext, "set_raw_predictions", scoring)
instrument(scoring.ScoringContext, "make_step_struct", scoring)
lat, _ = workloads.closed_loop(300, "weights_ethical.json", sys.argv[1] if len(sys.argv) > 1 else "cfg1")
print("step p50 %.1f us (instrumented)" % (np.percentile(lat[5:], 50) * 1e6))
rows = [(np.median(v[5:]) * 1e6, k) for k, v in ACC.items() if len(v) > 10]
for us, k in sorted(rows, reverse=True)[:40]:
    print("%7.2f us  %-22s %4d  %s" % (us, k[0], k[1], k[2]))
