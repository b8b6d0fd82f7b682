"""Times one training step (forward with tape + NormalizedMSELoss + backward + SGD update) of GraphWeatherForecaster on the
exact-fp32 training path.    python tools/train_step_bench.py [--grid 1deg|10deg] [--batch B] [--steps K]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="1deg", choices=["1deg", "10deg"])
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import __graft_entry__ as ge

    ge.build()
    from graph_weather_b200 import GraphWeatherForecaster, NormalizedMSELoss

    step = 1 if a.grid == "1deg" else 10
    ll = [(float(lat), float(lon)) for lat in range(-90, 90, step) for lon in range(0, 360, step)]
    torch.manual_seed(0)
    model = GraphWeatherForecaster(ll).cuda().train()
    crit = NormalizedMSELoss([1.0] * 78, ll, normalize=True)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    x = torch.randn(a.batch, len(ll), 102, device="cuda")
    y = torch.randn(a.batch, len(ll), 78, device="cuda")
    losses = []

    def one():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats()
    e0.record()
    for _ in range(a.steps):
        losses.append(one())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    free, total = torch.cuda.mem_get_info()
    print(json.dumps({"what": "training step (fwd + loss + bwd + SGD), exact fp32 CUDA cores", "grid": a.grid, "batch": a.batch, "ms_per_step": ms,
                      "samples_per_s": a.batch / (ms * 1e-3), "losses": [float(v) for v in losses], "device_mem_used_gib": round((total - free) / 2**30, 1)}))


if __name__ == "__main__":
    main()
