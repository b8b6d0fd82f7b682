"""Quick device probe: time the forward at a named grid/batch and print per-kernel-class device time."""
import argparse
import json
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--step", type=float, default=1.0)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--precision", default="fp32_simt")
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    ge.build()
    from graph_weather_b200 import GraphWeatherForecaster

    n_lat, n_lon = int(round(180 / a.step)), int(round(360 / a.step))
    ll = [(-90.0 + a.step * i, a.step * j) for i in range(n_lat) for j in range(n_lon)]
    t0 = time.time()
    torch.manual_seed(42)
    model = GraphWeatherForecaster(ll, precision=a.precision).cuda().eval()
    print(f"construct {time.time() - t0:.1f}s  N={len(ll)} Ed={model.decoder._g_dec.src.size}", flush=True)
    x = torch.randn(a.batch, len(ll), 102, device="cuda")
    t0 = time.time()
    y = model(x)
    torch.cuda.synchronize()
    print(f"first call {time.time() - t0:.2f}s  plan bytes {model._engine.plan.device_bytes() / 2**30:.2f} GiB  finite={bool(torch.isfinite(y).all())}", flush=True)
    plan = model._engine.plan
    for _ in range(2):
        model(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        model(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    plan.timing_enable(True)
    for _ in range(a.iters):
        model(x)
    tags = plan.timing_read()
    plan.timing_enable(False)
    print(json.dumps({"ms_per_step": ms, "steps_per_s": 1000 / ms, "tags": {k: (c // a.iters, round(v / a.iters, 3)) for k, (c, v) in tags.items() if c}}))


if __name__ == "__main__":
    try:
        main()
    except Exception as e:
        import gc

        from graph_weather_b200 import _capi

        print("FAILED:", str(e).splitlines()[0])
        for o in gc.get_objects():
            if isinstance(o, _capi.Plan) and o.handle.value:
                w = o.debug_words()
                if w[0]:
                    names = [f"full_a{i}" for i in range(4)] + [f"empty_a{i}" for i in range(4)] + ["full_b0", "full_b1", "empty_b0", "empty_b1", "full_d0", "full_d1", "empty_d0", "empty_d1"]
                    print("status", w[0])
                    for k in range(20):
                        off, par, blk = w[4 + 3 * k : 7 + 3 * k]
                        if off:
                            print(f"  warp {k:2d}: waiting {names[(off - 196608) // 8] if 0 <= (off - 196608) // 8 < len(names) else off} parity {par} block {blk}")
        raise SystemExit(1)
