"""One post-observation trial followed by a few post-intervention trials on a config-5-shaped set (development aid for
profiling the appended-row path: prior_rows_kernel<4>, posterior_fit_kernel, sweep_kernel with one cached set).
python tools/refresh_probe.py [--n-obs 10000] [--p 32 32 32] [--sets 2]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
ap = argparse.ArgumentParser()
ap.add_argument("--n-obs", type=int, default=10000)
ap.add_argument("--p", type=int, nargs="+", default=[32, 32, 32])
ap.add_argument("--trials", type=int, default=4)
ap.add_argument("--sets", type=int, default=2)
args = ap.parse_args()
import numpy as np, torch
from cbo_with_oop_b200.engine import SweepEngine
from cbo_with_oop_b200.synthetic import best_of, scaled_set
probs = []
for i in range(args.sets):
    pr = scaled_set(i, n_obs=args.n_obs, p=max(args.p), d=len(args.p), c=3, n_int=32, device_fit=True)
    pr.grid = [np.linspace(-2.0, 2.0, pk) for pk in args.p]
    probs.append(pr)
eng = SweepEngine(probs)
best = best_of(probs)
eng.sweep(best, "min")
rng = np.random.default_rng(0)
x, y = probs[0].x_int.copy(), probs[0].y_int.copy()
ms, host = [], []
import time
fin = eng._finish
marks = {}
def timed_finish(ev):
    marks["launched"] = time.perf_counter()
    return fin(ev)
eng._finish = timed_finish
for t in range(args.trials):
    x = np.vstack([x, rng.uniform(-2, 2, (1, len(args.p)))]); y = np.append(y, 0.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); eng.set_interventional(0, x, y)
    t1 = time.perf_counter()
    out = eng.refresh(best, "min", refit=[0])
    t2 = time.perf_counter()
    e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
    # host timeline of the trial in microseconds: data upload, Python + library call up to the last launch, wait for the result
    host.append({"set_interventional_us": (t1 - t0) * 1e6, "refresh_until_launched_us": (marks["launched"] - t1) * 1e6,
                 "finish_wait_us": (t2 - marks["launched"]) * 1e6})
print(json.dumps({"n_obs": args.n_obs, "p": args.p, "sets": args.sets, "ms_per_post_intervention_trial": ms, "host_timeline": host, "selected": [out.set, out.index]}))
