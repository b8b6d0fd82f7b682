"""Stage-by-stage comparison of the tensor-core path against the exact-fp32 CUDA-core path on the same weights."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--step", type=float, default=10.0)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--precision", default="fp32")
    a = ap.parse_args()
    ge.build()
    from graph_weather_b200 import Decoder, Encoder, GraphWeatherForecaster, Processor
    from oracle import weights

    n_lat, n_lon = int(round(180 / a.step)), int(round(360 / a.step))
    ll = [(-90.0 + a.step * i, a.step * j) for i in range(n_lat) for j in range(n_lon)]
    sd = weights.make_state_dict(weights.forecaster_shapes(), 9)
    x = weights.make_features(a.batch, len(ll), 102, 9).cuda()
    mods = {}
    for prec in ("fp32_simt", a.precision):
        enc = Encoder(ll, input_dim=102, precision=prec).cuda()
        proc = Processor(precision=prec).cuda()
        dec = Decoder(ll, precision=prec).cuda()
        enc.load_state_dict({k[8:]: v for k, v in sd.items() if k.startswith("encoder.")})
        proc.load_state_dict({k[10:]: v for k, v in sd.items() if k.startswith("processor.")})
        dec.load_state_dict({k[8:]: v for k, v in sd.items() if k.startswith("decoder.")})
        mods[prec] = (enc, proc, dec)
    e0, p0, d0 = mods["fp32_simt"]
    e1, p1, d1 = mods[a.precision]
    t = time.time()
    ex0, ei, ea = e0(x)
    torch.cuda.synchronize()
    print("simt encoder", time.time() - t, flush=True)
    ex1, ei1, ea1 = e1(x)
    torch.cuda.synchronize()
    e1._engine.plan.status()
    print(f"encoder   max|tc-simt| = {float((ex1 - ex0).abs().max()):.3e}   (|x| max {float(ex0.abs().max()):.2f})", flush=True)
    px0 = p0(ex0, ei, ea)
    px1 = p1(ex0, ei, ea)
    torch.cuda.synchronize()
    p1._engine.plan.status()
    print(f"processor max|tc-simt| = {float((px1 - px0).abs().max()):.3e}   (|x| max {float(px0.abs().max()):.2f})", flush=True)
    o0 = d0(px0, x[..., :78])
    o1 = d1(px0, x[..., :78])
    torch.cuda.synchronize()
    d1._engine.plan.status()
    print(f"decoder   max|tc-simt| = {float((o1 - o0).abs().max()):.3e}", flush=True)
    m0 = GraphWeatherForecaster(ll, precision="fp32_simt").cuda()
    m1 = GraphWeatherForecaster(ll, precision=a.precision).cuda()
    m0.load_state_dict(sd), m1.load_state_dict(sd)
    y0, y1 = m0(x), m1(x)
    torch.cuda.synchronize()
    m1._engine.plan.status()
    print(f"end2end   max|tc-simt| = {float((y1 - y0).abs().max()):.3e}   finite={bool(torch.isfinite(y1).all())}", flush=True)


if __name__ == "__main__":
    try:
        main()
    except Exception as e:  # after a device trap the host-mapped status block still tells which barrier timed out
        import gc

        from graph_weather_b200 import _capi

        print("FAILED:", str(e).splitlines()[0])
        for o in gc.get_objects():
            if isinstance(o, _capi.Plan) and o.handle.value:
                w = o.debug_words()
                if w[0]:
                    names = ["full_a0", "full_a1", "empty_a0", "empty_a1", "full_b0", "full_b1", "empty_b0", "empty_b1", "full_d0", "full_d1", "empty_d0", "empty_d1"] + [f"st_ready{i}" for i in range(5)] + [f"st_done{i}" for i in range(5)]
                    print("status", w[0])
                    for k in range(16):
                        off, par, blk = w[4 + 3 * k : 7 + 3 * k]
                        if off:
                            print(f"  warp {k:2d}: waiting {names[(off - 212992) // 8] if 0 <= (off - 212992) // 8 < len(names) else off} parity {par} block {blk}")
        raise SystemExit(1)
