#!/usr/bin/env python
"""bench.py — throughput of the Predict()/Perceive() hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--step-bytes B]
    python bench.py --impl reference ...           # the reference's own CPU implementation

A "step" is one pass of the hot path over one batch of synthetic input: every stream (one
stream = one reference Predictor = one file) advances by --step-bytes bytes (8x as many coded
bits). Workload = BASELINE.json configs[1] ("synthetic ... English-like text (enwik8 shape) on
1xB200, full mixer + LSTM"): synthetic enwik-shaped text from tools/gen_synth.py; the model
groups that are not device resident yet (PAQ8, FXCM: SURVEY §8 a13-a14) enter as synthetic
replay streams of the same shape (2022 12-bit codes per bit). The PPMD byte model runs on the device
(--ppmd resident, the default); --ppmd replay feeds a synthetic 256-float distribution per byte instead.

`value`  : input MB/s with every input already resident in HBM when the timed region starts.
`e2e`    : the same metric through the C-ABI call with HOST (pinned) buffers, copies included.
`roofline`: the mix kernel (dominant): algorithmic bytes (450 000 B per coded bit)
            per launch / its CUDA-event duration on its own stream, against MEASURED_PEAKS.json.
Multi-GPU (torchrun): streams are independent files, sharded across ranks with no data-path
collective (weak scaling); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before the CUDA context: one work queue per library stream

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

NCU_DRAM_BYTES_PER_BIT = 34786      # (63.16 MB read + 8.09 MB written) / 2048 bits of one launch, ncu --set full, profiles/r01_ncu_full_metrics.csv
ALGO_BYTES_PER_BIT = 450_000          # 55 172 fp32 weights read + written, + input vectors
N_EXT = 2022


def measured_hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows = []
        self.stop_flag = False
        self.index = index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons}


def make_inputs(torch, dev, n_bytes, seed):
    """Synthetic enwik-shaped text + replay streams of the right shape, generated on the device."""
    from gen_synth import synth_text
    text = np.frombuffer(synth_text(n_bytes, 0xE9E80001 + seed), dtype=np.uint8).copy()
    vocab = np.zeros(256, dtype=np.uint8)
    vocab[np.unique(text)] = 1
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + seed)
    d_bytes = torch.from_numpy(text).to(dev)
    bits = ((d_bytes[:, None] >> torch.arange(7, -1, -1, device=dev, dtype=torch.uint8)) & 1).reshape(-1).float()
    skill = torch.rand(N_EXT, device=dev, generator=g) * 1.5
    d_ext = torch.empty((n_bytes * 8, N_EXT), dtype=torch.int16, device=dev)
    for lo in range(0, n_bytes * 8, 8192):           # chunked to bound temporaries
        hi = min(lo + 8192, n_bytes * 8)
        logit = torch.randn((hi - lo, N_EXT), device=dev, generator=g) * 1.2 + skill * (2 * bits[lo:hi, None] - 1)
        d_ext[lo:hi] = torch.clamp(torch.round(4095.0 / (1.0 + torch.exp(-logit))), 0, 4095).to(torch.int16)
    d_ext[:, 429:431] = -1                            # 0xFFFF: slots the reference never writes (0.5)
    pp = torch.empty((n_bytes, 256), device=dev).exponential_(3.0, generator=g) + 1e-6
    nxt = torch.roll(d_bytes.long(), -1)
    pp[torch.arange(n_bytes, device=dev), nxt] += torch.rand(n_bytes, device=dev, generator=g) * 8
    pp *= torch.from_numpy(vocab).to(dev)[None, :].float()
    pp = (pp / pp.sum(dim=1, keepdim=True)).contiguous()
    return text, vocab, d_bytes, d_ext, pp


def cpu_baseline(sample_bytes):
    """The reference's own CPU implementation on the box's host cores (single thread: it has no other).
    Uses oracle/_ref/oracle_dump ("reference") when it travelled, else the oracle port ("port")."""
    from gen_synth import synth_text
    import tempfile
    ref = os.path.join(ROOT, "oracle", "_ref", "oracle_dump")
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "in.txt")
        open(src, "wb").write(synth_text(sample_bytes, 0xE9E80001))
        if os.path.exists(ref):
            out = subprocess.run([ref, "time", "n", src, str(sample_bytes + 16)], capture_output=True, text=True, timeout=1800)
            line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
            r = json.loads(line)
            path_s = r["code_s"] - r["big_models_s"]
            return {"value": r["bytes"] / path_s / 1e6, "unit": "MB/s", "cores": 1, "kind": "reference",
                    "sample": "first %d bytes of the synthetic text, cmix -n equivalent, g++ -O2 strict-FP build; "
                              "time of the rows this repo has on the device (predictor total %.2f s minus PAQ8+FXCM+PPMD %.2f s; PPMD, now "
                              "resident here too, is <1 %% of that and stays subtracted, which only favours the CPU); "
                              "whole predictor: %.6f MB/s; constructor %.1f s excluded"
                              % (r["bytes"], r["code_s"], r["big_models_s"], r["bytes"] / r["code_s"] / 1e6, r["ctor_s"]),
                    "full_predictor_value": r["bytes"] / r["code_s"] / 1e6}
    # port fallback: replay synthetic streams through the CPU restatement
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import port_replay, synthetic_streams
    from oracle_io import load_port
    stream, vocab, codes, ppmd = synthetic_streams(sample_bytes, seed=1)
    lib = load_port()
    t0 = time.time()
    port_replay(lib, vocab, stream, codes, ppmd)
    dt = time.time() - t0
    return {"value": sample_bytes / dt / 1e6, "unit": "MB/s", "cores": 1, "kind": "port",
            "sample": "%d bytes of synthetic replay streams through oracle/port (scalar C++)" % sample_bytes}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=int(os.environ.get("CMIXB200_BENCH_STREAMS", "24")), help="independent files per GPU")
    ap.add_argument("--step-bytes", type=int, default=2048)
    ap.add_argument("--ppmd", default=os.environ.get("CMIXB200_BENCH_PPMD", "resident"), choices=["resident", "replay"],
                    help="PPMD byte model: resident on the device (ppmd.cuh) or a synthetic replayed distribution")
    ap.add_argument("--cpu-sample-bytes", type=int, default=4096)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "configs[1]: synthetic enwik8-shaped ASCII text, full mixer + LSTM + small models; "
                          "PAQ8/FXCM outputs as synthetic replay streams; PPMD %s" % ("resident on the device" if args.ppmd == "resident" else "replayed"),
              "streams_per_gpu": args.streams, "step_bytes_per_stream": args.step_bytes,
              "parallelism": "independent streams sharded over %d rank(s), no data-path collective" % max(world, args.gpus),
              "l2": "inputs larger than L2: every step streams %.0f MB of fresh replay codes per stream and walks ~6 GB of "
                    "per-stream HBM tables" % (args.step_bytes * 8 * N_EXT * 2 / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        # each "step" is a bounded sample; K+W steps of the same size, the first W discarded
        sample = max(256, args.cpu_sample_bytes // 4)
        vals = []
        cb = None
        for i in range(args.warmup + args.steps):
            cb = cpu_baseline(sample)
            if i >= args.warmup:
                vals.append(cb["value"])
            if i == 0 and args.warmup + args.steps > 2:
                pass
        v = float(np.mean(vals))
        cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": "input_MB_per_s", "value": v, "unit": "MB/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": sample / v / 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import cmix_b200
    from cmix_b200.capi import code_batch_device
    from cmix_b200.sharding import stream_block, reduce_timing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    S, B, K, W = args.streams, args.step_bytes, args.steps, args.warmup
    total_steps = K + W
    # ---- inputs resident in HBM before the timed region ----
    streams = []
    free0 = torch.cuda.mem_get_info(dev)[0]
    for s in stream_block(S * world, world, rank):      # global stream ids owned by this rank (weak scaling: S per GPU)
        if len(streams) >= S:
            break
        text, vocab, d_bytes, d_ext, d_ppmd = make_inputs(torch, dev, B * total_steps, seed=s)
        P = cmix_b200.Predictor(vocab, device=local_rank)
        d_out = torch.empty(B * total_steps * 8, dtype=torch.float32, device=dev)
        streams.append(dict(P=P, text=text, d_bytes=d_bytes, d_ext=d_ext, d_ppmd=d_ppmd, d_out=d_out))
        if len(streams) == 1:
            # every stream owns ~7 GB of model tables: run as many streams as fit (the same number on every rank)
            # instead of driving the box out of memory
            per_stream = free0 - torch.cuda.mem_get_info(dev)[0] + 2 * 1024 * (8 * N_EXT * 2 + 256 * 4) + (64 << 20)
            s_fit = max(1, int(0.94 * free0 // per_stream))
            if dist:
                t_fit = torch.tensor([s_fit], device=dev, dtype=torch.int64)
                dist.all_reduce(t_fit, op=dist.ReduceOp.MIN)
                s_fit = int(t_fit.item())
            if s_fit < S:
                config["streams_requested"] = S
                S = s_fit
                config["streams_per_gpu"] = S
            config["hbm_per_stream_gb"] = round(per_stream / 1e9, 2)
    torch.cuda.synchronize()

    def run_step(i):
        lo, hi = i * B, (i + 1) * B
        code_batch_device([st["P"] for st in streams], [st["d_bytes"][lo:hi] for st in streams], B,
                          [st["d_ext"][lo * 8:hi * 8] for st in streams],
                          [st["d_ppmd"][lo:hi] for st in streams] if args.ppmd == "replay" else None,
                          [st["d_out"][lo * 8:hi * 8] for st in streams])

    for i in range(W):
        run_step(i)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    for st in streams:
        st["P"].time_mix_kernel(True)
    launches0 = sum(st["P"].kernel_launches for st in streams)
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    for i in range(W, W + K):
        run_step(i)                                   # synchronises its three library streams before returning
    ev1.record()
    torch.cuda.synchronize()
    dt = max(time.perf_counter() - t0, ev0.elapsed_time(ev1) / 1e3)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    bytes_per_rank = S * B * K
    total_bytes = bytes_per_rank * world
    if dist:
        dt, total_bytes = reduce_timing(dist, dev, dt, bytes_per_rank)
        dist.barrier()
    launches = sum(st["P"].kernel_launches for st in streams) - launches0
    mix_ms, mix_n = 0.0, 0                              # summed over the launch-group leaders (others report 0)
    for st in streams:
        ms_, n_ = st["P"].mix_kernel_ms()
        mix_ms += ms_
        mix_n += n_
    value = total_bytes / dt / 1e6

    # ---- end to end through the C-ABI with HOST (pinned) buffers: every stream of this rank, one more step ----
    # The same predictors continue from where the device-resident run stopped; the step's inputs start in
    # pinned host memory and its probabilities end there (cmixb200_code_batch stages them inside the call).
    from cmix_b200.capi import code_batch
    n_e2e = min(2, max(1, K + W - 1))                  # timed steps, after one untimed step that sizes the staging buffers
    lo = (W + K - n_e2e - 1) * B
    nb = (n_e2e + 1) * B
    h_bytes = [torch.from_numpy(st["text"][lo:lo + nb].copy()).pin_memory() for st in streams]
    h_ext = [st["d_ext"][lo * 8:(lo + nb) * 8].cpu().pin_memory() for st in streams]
    h_ppmd = [st["d_ppmd"][lo:lo + nb].cpu().pin_memory() for st in streams]
    h_out = [torch.empty(nb * 8, dtype=torch.float32).pin_memory() for _ in streams]
    preds = [st["P"] for st in streams]

    def e2e_step(j):
        code_batch(preds, [t[j * B:] for t in h_bytes], B, [t[j * B * 8 * N_EXT:] for t in h_ext],
                   [t[j * B * 256:] for t in h_ppmd] if args.ppmd == "replay" else None, [t[j * B * 8:] for t in h_out])

    h_ext = [t.view(-1) for t in h_ext]
    h_ppmd = [t.view(-1) for t in h_ppmd]
    e2e_step(0)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for j in range(1, n_e2e + 1):
        e2e_step(j)
    torch.cuda.synchronize()
    dt_e2e = time.perf_counter() - t0
    if dist:
        tt = torch.tensor([dt_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt_e2e = float(tt.item())
    e2e_value = S * n_e2e * B * world / dt_e2e / 1e6
    h2d = S * (B + B * 8 * N_EXT * 2 + (B * 256 * 4 if args.ppmd == "replay" else 0) + B * 8 * 4)   # per rank: bytes + codes (+ PPMD) + decay table
    d2h = S * B * 8 * 4
    if not all(bool(torch.isfinite(o).all()) and float(o.min()) >= 0.0 and float(o.max()) <= 1.0 for o in h_out):
        raise SystemExit("bench.py: end-to-end probabilities out of range")

    # ---- the decoder's order: Predict()/Perceive(bit) one bit at a time on one stream (not part of `value`) ----
    lock = None
    if rank == 0:
        P0 = streams[0]["P"]
        n_lock = 128
        bits = np.unpackbits(streams[0]["text"][:n_lock // 8])
        for b in bits[:16]:
            P0.Predict(); P0.Perceive(int(b))
        t0 = time.perf_counter()
        for b in bits[16:]:
            P0.Predict(); P0.Perceive(int(b))
        P0.Predict()                                     # drains the last queued Perceive()
        lock = {"us_per_bit": (time.perf_counter() - t0) / (n_lock - 16) * 1e6, "bits": n_lock - 16,
                "note": "cmixb200_predict/perceive through ctypes, replay inputs at 0.5, host clock"}

    if rank == 0:
        peak, peak_kind = measured_hbm_peak()
        # Launch groups run the same kernel concurrently (engine.cu RunPipelined), so the GPU-level figure is the
        # per-launch bandwidth times the MEASURED mean number of launches in flight = sum of launch durations / wall
        # time of the timed region (never more than the number of groups).
        group = int(os.environ.get("CMIXB200_GROUP", "8")) or S
        n_groups = (S + group - 1) // group
        bits_per_launch = (S * B * K * 8) / max(mix_n, 1)
        per_launch = ALGO_BYTES_PER_BIT * bits_per_launch / (mix_ms / max(mix_n, 1) / 1e3) / 1e9 if mix_ms > 0 else None
        in_flight = min(float(n_groups), mix_ms / (dt * 1e3)) if mix_ms > 0 else None
        achieved = per_launch * in_flight if per_launch else None
        out = {
            "metric": "input_MB_per_s", "value": value, "unit": "MB/s", "n_gpus": world if world > 1 else args.gpus,
            "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "all streams through cmixb200_code_batch with pinned host buffers; H2D of the step's inputs and D2H of its "
                            "probabilities inside the timed region (bytes are per rank per step)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": NCU_DRAM_BYTES_PER_BIT * bits_per_launch, "traffic_unit": "B per launch",
                         "traffic_source": "profiles/r01_ncu_full_metrics.csv: dram read+write of one mix_kernel_v3 launch / its 2048 bits",
                         "kernel": "mix_kernel_v3", "peak_source": "MEASURED_PEAKS.json (%s)" % peak_kind,
                         "algorithmic_bytes_per_bit": ALGO_BYTES_PER_BIT, "mix_kernel_ms_total": mix_ms, "mix_launches": mix_n,
                         "launch_groups": n_groups, "mean_launches_in_flight": in_flight, "achieved_per_launch": per_launch,
                         "note": "serial-dependency bound: each dot product is one fp32 FADD chain (bit-exact parity)"},
            "bits_per_s": total_bytes * 8 / dt,
            "lockstep": lock,
        }
        if (world == 1):
            try:
                out["cpu_baseline"] = cpu_baseline(args.cpu_sample_bytes)
            except Exception as e:  # the baseline is reported, never required for the GPU number
                out["cpu_baseline"] = {"value": None, "unit": "MB/s", "cores": 1, "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(out))
    for st in streams:
        st["P"].close()
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
