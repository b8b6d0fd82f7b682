"""N-GPU check of graph_weather_b200.dist.BoundaryGather (run under torchrun on a GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_boundary_gather.py
Every mode that is available gathers rank-tagged shards (even and uneven) and is compared with the NCCL all_gather; then the
copy-engine mode is timed alone and underneath a persistent chain-kernel workload (one 1-degree forward) to show the overlap."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_weather_b200.dist import BoundaryGather, all_gather_batch, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = {"world": world}
    for mode in ("p2p_copy", "nccl"):
        for total in (world * 3, world * 3 + 1):
            a, b = shard_range(total, rank, world)
            y = (torch.arange((b - a) * 5 * 7, device=dev, dtype=torch.float32).reshape(b - a, 5, 7) + 1000.0 * rank)
            ref = all_gather_batch(y, total)
            try:
                g = BoundaryGather(total, dev, mode=mode)
                for it in range(3):  # buffers are reused every second call
                    out = g(y + it)
                    g.wait()
                    torch.cuda.synchronize(dev)
                    ok = bool(torch.equal(out, ref + it))
                    report[f"{mode}_total{total}_it{it}"] = ok
                report[f"{mode}_mode"] = g.mode
            except Exception as e:  # noqa: BLE001
                report[f"{mode}_total{total}"] = f"unavailable: {type(e).__name__}: {e}"[:300]
    # the forward fused with the gather (multicast / peer stores inside the last chain kernel) against forward + NCCL all_gather
    from graph_weather_b200 import GraphWeatherForecaster

    ll = [(float(a), float(b)) for a in range(-90, 90, 10) for b in range(0, 360, 10)]
    torch.manual_seed(7)
    model = GraphWeatherForecaster(ll).to(dev).eval()  # same seed on every rank: replicated weights
    for total in (world * 2, world * 2 + 1):
        a, b = shard_range(total, rank, world)
        x = torch.randn(b - a, len(ll), 102, generator=torch.Generator().manual_seed(100 + rank)).to(dev)
        ref = all_gather_batch(model(x), total)
        for mode in ("fused", "fused_peer", "p2p_copy", "nccl"):
            try:
                g = BoundaryGather(total, dev, mode=mode)
                oks = []
                for it in range(3):
                    out = g.forward(model, x)
                    g.wait()
                    torch.cuda.synchronize(dev)
                    oks.append(bool(torch.equal(out, ref)))
                report[f"forward_{mode}_total{total}"] = {"ok": oks, "mode": g.mode, "store": (g._fused[0][0] if getattr(g, "_fused", None) else None)}
            except Exception as e:  # noqa: BLE001
                report[f"forward_{mode}_total{total}"] = f"unavailable: {type(e).__name__}: {e}"[:300]
    # timing: 162 MB per rank (the 1 degree / batch 8 output), alone
    y = torch.randn(8, 64800, 78, device=dev)
    for mode in ("p2p_copy", "nccl"):
        try:
            g = BoundaryGather(8 * world, dev, mode=mode)
            for _ in range(3):
                g(y)
            g.wait()
            torch.cuda.synchronize(dev)
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                g(y, overlap=False)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / 10
            report[f"{mode}_ms_alone"] = ms
            report[f"{mode}_gbs_recv"] = (world - 1) * y.numel() * 4 / (ms * 1e-3) / 1e9
        except Exception as e:  # noqa: BLE001
            report[f"{mode}_timing"] = f"unavailable: {type(e).__name__}: {e}"[:300]
    allr = [None] * world
    dist.all_gather_object(allr, report)
    if rank == 0:
        print(json.dumps(allr, indent=1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
