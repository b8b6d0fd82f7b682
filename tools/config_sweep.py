"""BASELINE.json configs beyond the headline one: bf16 at 1 deg, the 0.25 deg ERA5 grid (fp32-faithful and bf16), the assimilator."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
from graph_weather_b200 import GraphWeatherAssimilator, GraphWeatherForecaster

def timeit(fn, iters=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

res = {}
# ---- 1 deg, batch 8: bf16 vs fp32-faithful
ll = [(float(a), float(b)) for a in range(-90, 90) for b in range(0, 360)]
torch.manual_seed(0)
m32 = GraphWeatherForecaster(ll, precision="fp32").cuda()
mbf = GraphWeatherForecaster(ll, precision="bf16").cuda()
mbf.load_state_dict(m32.state_dict())
x = torch.randn(8, len(ll), 102, device="cuda")
y32, ybf = m32(x), mbf(x)
mbf._engine.plan.status()
res["1deg_b8_bf16"] = {"ms": timeit(lambda: mbf(x)), "max_abs_vs_fp32": float((ybf - y32).abs().max()), "rms_vs_fp32": float((ybf - y32).pow(2).mean().sqrt())}
res["1deg_b8_fp32"] = {"ms": timeit(lambda: m32(x))}
print(json.dumps(res), flush=True)
del m32, mbf, x, y32, ybf
torch.cuda.empty_cache()
# ---- 0.25 deg ERA5 grid (721 x 1440), configs[2]: batch 4 bf16; fp32-faithful at batch 2
t0 = time.time()
ll = [(-90.0 + 0.25 * i, 0.25 * j) for i in range(721) for j in range(1440)]
for prec, B in (("bf16", 4), ("fp32", 2)):
    torch.manual_seed(0)
    t0 = time.time()
    m = GraphWeatherForecaster(ll, precision=prec).cuda()
    tc = time.time() - t0
    x = torch.randn(B, len(ll), 102, device="cuda")
    y = m(x)
    m._engine.plan.status()
    res[f"0.25deg_b{B}_{prec}"] = {"construct_s": tc, "n_points": len(ll), "n_dec_edges": int(m.decoder._g_dec.src.size), "ms": timeit(lambda: m(x), 2),
                                   "finite": bool(torch.isfinite(y).all()), "plan_GiB": m._engine.plan.device_bytes() / 2**30}
    print(json.dumps(res[f"0.25deg_b{B}_{prec}"]), flush=True)
    del m, x, y
    torch.cuda.empty_cache()
# ---- assimilator README config (analysis_dim 24, 5 deg output grid, 3380 observations)
rng = np.random.default_rng(42)
obs = [(float(lat), float(lon), float(rng.random())) for lat in range(-90, 90, 7) for lon in rng.uniform(0, 360, 100)]
obs += [(float(lat), float(lon), float(rng.random())) for lat in range(-90, 90, 45) for lon in range(0, 360, 24)]
out_ll = [(float(a), float(b)) for a in range(-90, 90, 5) for b in range(0, 360, 5)]
for prec in ("fp32", "fp32_simt"):
    a = GraphWeatherAssimilator(output_lat_lons=out_ll, analysis_dim=24, precision=prec).cuda()
    o = torch.tensor(obs, device="cuda"); f = torch.randn(1, len(obs), 2, device="cuda")
    y = a(f, o)
    res[f"assimilator_{prec}"] = {"ms": timeit(lambda: a(f, o)), "shape": list(y.shape), "finite": bool(torch.isfinite(y).all())}
print(json.dumps(res))
json.dump(res, open("gpurun_out/config_sweep.json", "w"), indent=1)
# ---- loss boundary: NormalizedMSELoss at 1 deg, batch 8 (HBM-bound reduction; algorithmic bytes = 2 x 4 x B x N x F)
from graph_weather_b200 import NormalizedMSELoss
ll = [(float(a), float(b)) for a in range(-90, 90) for b in range(0, 360)]
crit = NormalizedMSELoss([1.0 + 0.01 * i for i in range(78)], ll, normalize=True)
p, t = torch.randn(8, len(ll), 78, device="cuda"), torch.randn(8, len(ll), 78, device="cuda")
ms = timeit(lambda: crit.local_sum(p, t), iters=20)
res["loss_1deg_b8"] = {"ms": ms, "GBps": 2 * p.numel() * 4 / ms / 1e6, "value": float(crit(p, t))}
print(json.dumps(res), flush=True)
