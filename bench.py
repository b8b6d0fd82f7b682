"""bench.py -- forward steps/s of GraphWeatherForecaster(lat_lons)(features) at the 1-degree / 102->78 configuration.

    python bench.py --gpus 1 --steps 20 --warmup 3                 # this repo's CUDA path (tcgen05 chains)
    python bench.py --impl reference --steps 2 --warmup 1           # the reference's CPU path (oracle port) on the host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU, batch-sharded

One step = one model(features) call at batch 8 per GPU (BASELINE.json configs[1]); `value` is the whole-job aggregate
(steps of batch 8 per second summed over ranks; weak scaling).  With N > 1 every step ends with the single NCCL
all-gather of the outputs at the loss boundary; nothing else is communicated.

The JSON line carries the contract keys plus
  roofline      the dominant kernel class by device time, timed live with CUDA events on the launching stream
                (libgwb200's gw_timing_*), algorithmic FLOPs / time vs the measured dense bf16 peak
  cpu_baseline  the oracle port of the reference forward timed on this box's host cores (rank 0, N = 1 only)
  e2e           the same metric through the public module call with pinned-host inputs copied in and the forecast copied out
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_MESH, EL = 5882, 41162
FIN, FOUT = 102, 78
# per-row Linear FLOPs (2*MACs), SURVEY.md section 8
F_EDGE_MLP = 2 * (768 * 256 + 256 * 256 + 256 * 256)
F_NODE_MLP = 2 * (512 * 256 + 256 * 256 + 256 * 256)
F_NODE_ENC = 2 * (FIN * 256 + 256 * 256 + 256 * 256)
F_NODE_DEC = 2 * (256 * 128 + 128 * 128 + 128 * FOUT)


def grid_1deg():
    return [(float(lat), float(lon)) for lat in range(-90, 90) for lon in range(0, 360)]  # README.md:48-51


def algorithmic_flops(n, ed):
    """F_alg per sample and per kernel class (live outputs, unfactored Linear FLOPs)."""
    per = {
        "enc_grid": n * (F_NODE_ENC + F_EDGE_MLP),
        "enc_mesh": H_MESH * F_NODE_MLP,
        "proc_p": 0.0,  # its products are layer 1 of the edge MLP, counted under proc_edge
        "proc_edge": 9 * EL * F_EDGE_MLP,
        "proc_node": 9 * H_MESH * F_NODE_MLP,
        "dec_p": 0.0,
        "dec_edge": ed * F_EDGE_MLP,
        "dec_node": n * (F_NODE_MLP + F_NODE_DEC),
    }
    return sum(per.values()), per


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_inputs(lat_lons, batch, seed=42):
    """Reference-shaped inputs for the CPU leg: default-initialised weights under the seed the reference tests use."""
    from graph_weather_b200 import GraphWeatherForecaster, graphs

    torch.manual_seed(seed)
    model = GraphWeatherForecaster(lat_lons)  # same init as the reference under the same seed (tests/test_capi.py)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    e, m, d = model.encoder._g_enc, model.encoder._g_lat, model.decoder._g_dec
    g = dict(enc_edge_index=torch.from_numpy(e.edge_index), enc_edge_attr=torch.from_numpy(e.edge_attr),
             lat_edge_index=torch.from_numpy(m.edge_index), lat_edge_attr=torch.from_numpy(m.edge_attr),
             dec_edge_index=torch.from_numpy(d.edge_index), dec_edge_attr=torch.from_numpy(d.edge_attr),
             num_latlons=len(lat_lons), num_h3=m.num_h3)  # fmt: skip
    x = torch.randn(batch, len(lat_lons), FIN)
    return sd, g, x


def time_cpu_reference(lat_lons, step_batch, sample_batch, steps, warmup):
    """The reference's forward (oracle/restate.py: same ops, same replicated-graph batching) on the host cores.
    Each timed forward runs `sample_batch` of the step's `step_batch` samples; steps/s = samples/s / step_batch."""
    from oracle import restate

    cores = os.cpu_count()
    sd, g, x = oracle_inputs(lat_lons, sample_batch)
    # give the CPU path its better thread count: all logical cores or one thread per physical core (probe: one Processor
    # pass on one sample; the untimed warm-up forwards then run at the chosen setting)
    best_n, best_t = cores, None
    with torch.no_grad():
        x1, ei1, ea1 = restate.encoder_forward(sd, g, x[:1])
        for n in sorted({cores, max(1, cores // 2)}, reverse=True):
            torch.set_num_threads(n)
            t0 = time.perf_counter()
            restate.processor_forward(sd, x1, ei1, ea1, 9)
            dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best_n, best_t = n, dt
    torch.set_num_threads(best_n)
    cores = best_n
    for _ in range(warmup):
        restate.forecaster_forward(sd, g, x)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        restate.forecaster_forward(sd, g, x)
        ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    steps_per_s = (sample_batch / t) / step_batch
    return steps_per_s, t, cores


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU per step (BASELINE configs[1]: 8)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32_simt", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()

    # stdout carries exactly one JSON line: anything libraries print in between (NCCL's version banner comes from C code)
    # is sent to stderr by pointing file descriptor 1 there until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    lat_lons = grid_1deg()
    workload = f"1deg_grid_64800pts_102to78_batch{a.batch}_per_gpu_fp32"
    cfg = {"workload": workload, "grid": "1deg lat -90..89 x lon 0..359 (README.md:48-51)", "batch_per_gpu": a.batch,
           "global_batch": a.batch * world, "hidden": 256, "processor_blocks": 9, "parallelism": f"dp{world} (batch shards, one all-gather at the loss boundary)",
           "cache": "inputs per step 211 MB + weight-constant edge tables 0.9 GB stream through HBM each step (> 126 MB L2); no explicit flush"}  # fmt: skip

    if a.impl == "reference":
        if rank != 0:
            return
        sample_b = min(2, a.batch)
        sps, t, cores = time_cpu_reference(lat_lons, a.batch, sample_b, max(1, a.steps), max(0, a.warmup))
        sample = f"{sample_b} of the {a.batch} samples of a step per timed forward (1deg grid); steps/s = samples/s / {a.batch}"
        emit(({
            "impl": "reference", "metric": "forward steps/sec", "value": sps, "unit": "steps/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 / sps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))  # fmt: skip
        return

    import __graft_entry__ as ge

    if rank == 0 or not os.path.exists(ge.LIB):
        ge.build()
    from graph_weather_b200 import GraphWeatherForecaster, _capi
    from graph_weather_b200.dist import all_gather_batch

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)
    model = GraphWeatherForecaster(lat_lons, precision=a.precision).to(dev).eval()
    n, ed = len(lat_lons), int(model.decoder._g_dec.src.size)
    torch.manual_seed(1234 + rank)
    x_host = torch.randn(a.batch, n, FIN).pin_memory()
    x = x_host.to(dev)
    out_host = torch.empty(a.batch, n, FOUT).pin_memory()

    def step_resident():
        y = model(x)
        if world > 1:
            y_all = all_gather_batch(y, world * a.batch)  # the one collective: outputs at the loss boundary
            return y_all
        return y

    # End-to-end step through the public module call: every step copies its inputs in from pinned host memory and its
    # forecast back out.  The copies run on their own streams (double-buffered), so step i+1's input upload and step
    # i-1's download overlap step i's compute -- the steady state of a real rollout / evaluation loop.
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    x_bufs = [torch.empty_like(x), torch.empty_like(x)]
    out_bufs = [out_host, torch.empty_like(out_host).pin_memory()]
    e2e_state = {"i": 0, "used": [None, None]}

    def step_e2e():
        i = e2e_state["i"]
        e2e_state["i"] = i + 1
        cur = torch.cuda.current_stream(dev)
        b = i & 1
        if e2e_state["used"][b] is not None:
            s_in.wait_event(e2e_state["used"][b])  # step i-2 has finished reading this input buffer
        with torch.cuda.stream(s_in):
            x_bufs[b].copy_(x_host, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(s_in)
        cur.wait_event(ev_in)
        y = model(x_bufs[b])
        if world > 1:
            all_gather_batch(y, world * a.batch)
        ev_c = torch.cuda.Event()
        ev_c.record(cur)
        e2e_state["used"][b] = ev_c
        s_out.wait_event(ev_c)
        with torch.cuda.stream(s_out):
            out_bufs[b].copy_(y, non_blocking=True)
        y.record_stream(s_out)
        return y

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        cur = torch.cuda.current_stream(dev)
        cur.wait_stream(s_in), cur.wait_stream(s_out)  # the timed region ends when the last download has landed
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # max over ranks
        sync_all()
        return float(ms.item())

    for _ in range(max(3, a.warmup)):
        step_resident()
    plan = model._engine.plan
    plan.timing_enable(True)
    _capi.launch_count_reset()
    with ClockSampler(local) as clk:
        ms_total = timed(step_resident, a.steps)
    launches = _capi.launch_count()
    tags = plan.timing_read()
    plan.timing_enable(False)
    plan.status()  # raises if any kernel flagged fp16-range overflow or a pipeline fault
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    ms_step = ms_total / a.steps
    value = world * a.steps / (ms_total / 1000.0)
    e2e_value = world * a.steps / (ms_e2e / 1000.0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a seconds-long step loop)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    f_alg, per = algorithmic_flops(n, ed)
    dom = max((k for k in tags if tags[k][0] and k != "const"), key=lambda k: tags[k][1])
    cnt, ms_dom = tags[dom]
    per_launch_flops = per[dom] * a.batch * a.steps / cnt
    achieved = per_launch_flops / ((ms_dom / cnt) * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "peak_source": peak_src,
                "note": "achieved = algorithmic (unfactored, SURVEY 8(d)) FLOPs per launch / mean launch time; the fp32-faithful path "
                        "issues 3 fp16 MMAs per product, so frac <= 1/3 x (algorithmic/executed FLOP ratio) of the bf16 peak",
                "whole_step": {"achieved": f_alg * a.batch / (ms_step * 1e-3) / 1e12, "unit": "TFLOP/s",
                               "frac": f_alg * a.batch / (ms_step * 1e-3) / 1e12 / peak_tf},
                "per_kernel_ms_per_step": {k: round(v[1] / a.steps, 4) for k, v in tags.items() if v[0]},
                "kernel_time_share_of_step": round(sum(v[1] for v in tags.values()) / ms_total, 4)}  # fmt: skip
    line = {
        "metric": "forward steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32 (fp16x2-split tcgen05, fp32 accumulate)", "fp32_simt": "f32", "bf16": "bf16"}[a.precision],
        "data": "synthetic", "config": cfg, "samples_per_s": value * a.batch,
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": int(out_host.numel() * 4),
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline,
    }  # fmt: skip
    if world == 1 and not a.no_cpu_baseline:
        sample_b = min(2, a.batch)
        sps, t, cores = time_cpu_reference(lat_lons, a.batch, sample_b, 2, 1)
        line["cpu_baseline"] = {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port", "seconds_per_forward": t,
                                "sample": f"oracle port of the reference forward, {sample_b} of the {a.batch} samples per forward, 1 warm-up + 2 timed; steps/s = samples/s / {a.batch}"}  # fmt: skip
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
