"""Dump the in-kernel event timeline (CTA 0) of one tensor-core chain.  The tracer exists only in the diagnostics build:

    python tools/ablate.py --build-only                      # here, in the build container -> build_abl/libgwb200.so
    GW_ABLATE=0 python tools/trace_chain.py --tag proc_edge   # on the GPU box (GW_B200_LIB defaults to build_abl/)
"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

ap = argparse.ArgumentParser(); ap.add_argument("--tag", default="proc_edge"); ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--out", default="gpurun_out/trace.npy"); a = ap.parse_args()
os.environ.setdefault("GW_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build_abl", "libgwb200.so"))
ge.build()
from graph_weather_b200 import GraphWeatherForecaster
ll = [(float(x), float(y)) for x in range(-90, 90) for y in range(0, 360)]
torch.manual_seed(0)
torch.set_grad_enabled(False)
m = GraphWeatherForecaster(ll, precision="fp32").cuda().eval()
x = torch.randn(a.batch, len(ll), 102, device="cuda")
for _ in range(2): m(x)
torch.cuda.synchronize()
buf = m._engine.plan.trace_next(a.tag)
m(x); torch.cuda.synchronize()
t = buf.cpu().numpy()
np.save(a.out, t)
roles = ["producer", "mma", "-", "-", "-", "worker_q0_h0", "worker_q0_h1", "-"]
t0 = min(int(t[r, 0, 0]) for r in range(8) if t[r, 0, 0] > 0)
for r in range(8):
    ev = [(int(c) - t0, int(k)) for c, k in t[r] if c > 0]
    print(roles[r], len(ev), "events; first 260:")
    print("  " + " ".join(f"{k}@{c}" for c, k in ev[:260]))
