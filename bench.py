#!/usr/bin/env python
"""Benchmark of the WindGNN GCN-GRU forward hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one forward of the workload (BASELINE.json configs[1]: the shipped 34-station
checkpoint, batch 4096 windows of 168 hours) — per GPU; with N > 1 (launched through
``torch.distributed.run``) every rank runs the same per-GPU batch on its own device (weak scaling:
the sequence batch is sharded, no collective on the data path).

Prints ONE JSON line (rank 0).  ``value`` = sequences/s with the inputs resident in HBM; ``e2e`` =
the same through the host-buffer entry point (pinned host input, H2D + compute + D2H inside the
timed region); ``roofline`` = the dominant kernel (the input-projection FFMA GEMM) against the
FP32 FFMA peak measured live on the same GPU; ``cpu_baseline`` = the oracle's torch-CPU port of
the reference forward on this box's host cores.

``--impl reference`` times that CPU port alone (the reference is pure Python/PyTorch and cannot
travel to the GPU box; the port calls the same torch CPU kernels batch-generalised).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
METRIC = "gcn_gru_forward_sequences_per_s"
UNIT = "sequences/s"

# workload = BASELINE.json configs[1]
S, T, F, B_PER_GPU = 34, 168, 13, 4096
H, I = 3 * S, 13 * S
FLOP_GCN = 2 * (2 * S * S * F + 2 * S * F * F)          # both layers, per (sequence, step)
FLOP_IH = 2 * I * 3 * H
FLOP_HH = 2 * H * 3 * H
FLOP_PER_SEQ = T * (FLOP_GCN + FLOP_IH + FLOP_HH)        # 69,892,032
BYTES_PER_SEQ = T * S * F * 4 + T * H * 4                # 365,568: input read once + output written once


def load_workload():
    sd = torch.load(os.path.join(GOLDEN, "wind_gnn_34.pth"), map_location="cpu", weights_only=True)
    with open(os.path.join(GOLDEN, "coords.json")) as f:
        c = json.load(f)
    idx = [i for i, n in enumerate(c["names"]) if n != "Enchant 2 AGCM"]
    latlon = np.array([[c["lat"][i], c["lon"][i]] for i in idx], dtype=np.float64)
    return sd, latlon


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: one streaming
    `nvidia-smi -lms 20` process started ahead of time; `mark_start()` / `mark_stop()` bracket the
    region and only samples stamped inside it are summarised."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.t0 = self.t1 = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.5)  # let it start streaming before the timed region begins
        except OSError:
            self.proc = None
        return self

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def __exit__(self, *a):
        self.rows = []
        if self.proc is None:
            return
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=10)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        import datetime

        for line in out.splitlines():
            parts = [v.strip() for v in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                continue
            self.rows.append((ts, parts[1:]))

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lo = (self.t0 or 0) - 0.02
        hi = (self.t1 or 1e18) + 0.02
        inside = [r for ts, r in self.rows if lo <= ts <= hi]
        where = "timed region"
        if not inside:  # region shorter than the sampling period: fall back to the nearest samples
            inside = [r for _, r in sorted(self.rows, key=lambda x: abs(x[0] - (self.t0 or 0)))[:3]]
            where = "nearest to timed region"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "window": where}


def ncu_traffic_bytes(kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu summary
    (profiles/, same workload); None if no capture is committed."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    try:
        names = sorted(f for f in os.listdir(pdir) if f.endswith("_ncu_summary.json"))
    except OSError:
        return None
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for name in names:  # the last (latest round) wins
        try:
            with open(os.path.join(pdir, name)) as f:
                summ = json.load(f)
            for k in summ["kernels"]:
                if kernel_substr in k["kernel"]:
                    best = (k["dram__bytes_read.sum"] * scale[k["dram__bytes_read.sum.unit"]]
                            + k["dram__bytes_write.sum"] * scale[k["dram__bytes_write.sum.unit"]])
        except Exception:
            continue
    return best


def cpu_port_seq_per_s(sd, adj32, n_seq: int, min_seconds: float, max_reps: int, seed: int = 0):
    """The oracle's torch-CPU port (same library kernels as the reference, batch-generalised) on all
    host threads.  Returns (sequences/s, threads, reps, seconds)."""
    from oracle import gcn_gru_forward_torch

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((n_seq, T, S, F), generator=g)
    gcn_gru_forward_torch(adj32, x[: max(1, n_seq // 8)], sd)  # warm-up (thread pool, allocator)
    reps, t0 = 0, time.perf_counter()
    while True:
        gcn_gru_forward_torch(adj32, x, sd)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or reps >= max_reps:
            break
    return n_seq * reps / dt, threads, reps, dt


def run_reference(args, rank: int):
    """Reference arm: the CPU implementation of the path on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    sd, latlon = load_workload()
    from oracle import dense_graph_f64

    adj32 = torch.from_numpy(dense_graph_f64(latlon).astype(np.float32))
    from oracle import gcn_gru_forward_torch

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_seq = args.ref_seqs  # bounded sample of the 4096-sequence step
    x = torch.rand((n_seq, T, S, F), generator=torch.Generator().manual_seed(0))
    for _ in range(max(1, min(args.warmup, 2))):
        gcn_gru_forward_torch(adj32, x[: min(64, n_seq)], sd)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gcn_gru_forward_torch(adj32, x, sd)
    dt = time.perf_counter() - t0
    value = n_seq * args.steps / dt
    sample = f"{n_seq} of the {B_PER_GPU} sequences of each step, {args.steps} steps, torch CPU fp32, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "34-station GCN-GRU forward, wind_gnn_34.pth, T=168, CPU sample of B=4096",
                   "S": S, "T": T, "batch_per_step": n_seq},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def run_scaled(args, rank: int, world: int, local_rank: int):
    """The two scaled BASELINE configs (device-resident throughput only; parity for both shapes is in
    tests/test_parity_gpu.py).  fwd34_1m: 2^20 windows in total, rank r owns a contiguous 1/world of them and
    streams them through the library in resident pieces of <= 131072 windows (38.9 GB of input each; one
    synthetic piece is generated on the device and reused for every piece of the shard).  fwd4096: 256 windows
    per GPU of the 4096-station kNN graph (weak scaling)."""
    import windgnn_b200
    from windgnn_b200.shard import max_over_ranks, shard_range

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    warmup, steps = max(args.warmup, 3), max(args.steps, 1)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    if args.workload == "fwd34_1m":
        sd, latlon = load_workload()
        model = windgnn_b200.GCN_GRU(F, F, F, I, H)
        model.load_state_dict(sd, strict=True)
        adj = windgnn_b200.build_graph_from_latlon(latlon, device=dev, dtype=torch.float32)
        total = 1 << 20
        lo, hi = shard_range(total, rank, world)
        piece = min(hi - lo, 131072)
        n_pieces = (hi - lo + piece - 1) // piece
        x = torch.rand((piece, T, S, F), generator=gen, device=dev)
        Sx, Tx, flop_seq, bytes_seq, scaling = S, T, FLOP_PER_SEQ, BYTES_PER_SEQ, "strong"
        units = total
        name = ("34-station GCN-GRU forward (wind_gnn_34.pth), T=168, 2^20 windows sharded contiguously over the "
                f"GPUs, {n_pieces} resident piece(s) of {piece} per GPU (BASELINE.json configs[2])")

        def step():
            for _ in range(n_pieces):
                model(adj, x)
    else:
        S4, Fh4, H4, T4, B4 = 4096, 128, 128, 24, args.batch if args.batch != B_PER_GPU else 256
        torch.manual_seed(4)
        model = windgnn_b200.GCN_GRU(F, Fh4, F, F * S4, H4)
        with torch.no_grad():
            model.conv1.weight.mul_(0.05)
            model.conv2.weight.mul_(0.05)
        ll = windgnn_b200.synthetic_coordinates(S4, seed=0, device=dev).cpu().numpy()
        adj = windgnn_b200.knn_graph_from_latlon(ll, k=8, device=dev)
        nnz = int((adj != 0).sum().item())
        x = torch.rand((B4, T4, S4, F), generator=gen, device=dev)
        Sx, Tx, scaling = S4, T4, "weak"
        flop_seq = T4 * (2 * (nnz * F + S4 * F * Fh4) + 2 * (nnz * Fh4 + S4 * Fh4 * F) + 2 * (F * S4) * 3 * H4
                         + 2 * H4 * 3 * H4)
        bytes_seq = T4 * S4 * F * 4 + T4 * H4 * 4
        units = world * B4
        name = (f"synthetic 4096-station kNN(k=8) graph (nnz {nnz}), GCN_GRU(13,128,13,53248,128), T=24, "
                f"{B4} windows per GPU (BASELINE.json configs[3], SURVEY 8(d) variant C)")

        def step():
            model(adj, x)
    model = model.to(dev).eval()
    with torch.no_grad():
        for _ in range(warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clocks:
            barrier()
            clocks.mark_start()
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            barrier()
            clocks.mark_stop()
        ms = max_over_ranks(e0.elapsed_time(e1), dev)
    if rank == 0:
        value = units * steps / (ms * 1e-3)
        lib_peak = windgnn_b200._lib.load().wg_measure_ffma_tflops(local_rank, 10)
        tfl = flop_seq * value / world / 1e12
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "S": Sx, "T": Tx, "l2_policy": "inputs larger than L2",
                       "station_sequence_predictions_per_s": value * Sx},
            "clocks": clocks.summary(),
            "roofline": {"bound": "fp32", "achieved": tfl, "peak": lib_peak, "unit": "TFLOP/s", "frac": tfl / lib_peak,
                         "scope": "whole path, per GPU, algorithmic FLOPs", "flop_per_seq": flop_seq,
                         "hbm_gbs_algorithmic": bytes_seq * value / world / 1e9},
        }), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="sequences per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tensor-path", action="store_true")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32x3"],
                    help="fp32: every contraction as FP32 FMA; tf32x3: GRU input projection on tcgen05 (3xTF32)")
    ap.add_argument("--ref-seqs", type=int, default=512, help="sequences per step of the reference arm's sample")
    ap.add_argument("--workload", default="fwd34", choices=["fwd34", "fwd34_1m", "fwd4096"],
                    help="fwd34: BASELINE configs[1] (default, the contract line); fwd34_1m: configs[2], 2^20 sequences "
                         "of T=168 sharded over the GPUs (strong scaling); fwd4096: configs[3], 4096-station kNN(8) "
                         "graph, GCN hidden 128, GRU hidden 128, T=24")
    ap.add_argument("--no-train-step", action="store_true")
    ap.add_argument("--train-batch", type=int, default=512,
                    help="windows per GPU of the training-step measurement (BASELINE.json configs[4]: 4096 over 8 GPUs)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — windgnn_b200 has no CPU fallback")
    if args.workload != "fwd34":
        run_scaled(args, rank, world, local_rank)
        return
    import windgnn_b200
    from windgnn_b200 import _lib

    lib = _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: F811

        # keep stdout to the one JSON line: NCCL prints its version banner there at VERSION level
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    Bg = args.batch

    sd, latlon = load_workload()
    model = windgnn_b200.GCN_GRU(F, F, F, I, H)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    model.precision = args.precision
    flags = model._flags()
    adj = windgnn_b200.build_graph_from_latlon(latlon, device=dev, dtype=torch.float32)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.rand((Bg, T, S, F), generator=gen, device=dev)  # 1.22 GB at B=4096: larger than the 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from windgnn_b200.shard import max_over_ranks as _max_over_ranks

    def max_over_ranks(v: float) -> float:
        return _max_over_ranks(v, dev)

    # ---------------- device-resident throughput (`value`) ----------------
    with torch.no_grad():
        for _ in range(warmup):
            y = model(adj, x)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clocks:
            barrier()
            clocks.mark_start()
            e0.record()
            for _ in range(steps):
                y = model(adj, x)
            e1.record()
            barrier()
            clocks.mark_stop()
        ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * Bg * steps / (ms * 1e-3)
    n_chunks = (Bg + 148 * 32 - 1) // (148 * 32)
    launches_per_step = 1 + 3 * n_chunks  # pack + (gcn, inproj, recur) per internal chunk

    # ---------------- the opt-in tensor-core path, same workload, same timing rules ----------------
    tensor_path = None
    if args.precision == "fp32" and not args.no_tensor_path:
        model.precision = "tf32x3"
        try:
            with torch.no_grad():
                for _ in range(warmup):
                    yt = model(adj, x)
                barrier()
                t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0e.record()
                for _ in range(steps):
                    yt = model(adj, x)
                t1e.record()
                barrier()
                ms_t = max_over_ranks(t0e.elapsed_time(t1e))
            err = float((yt - y).abs().max() / y.abs().max())
            tensor_path = {"precision": "tf32x3 (tcgen05 input projection, 3 TF32 products per term, fp32 accumulate)",
                           "value": world * Bg * steps / (ms_t * 1e-3), "unit": UNIT, "ms_per_step": ms_t / steps,
                           "max_abs_diff_vs_fp32_path_normalised": err, "parity_bar": 1e-5}
        finally:
            model.precision = args.precision
    barrier()

    # ---------------- the reference's own call pattern: one window per call (src/main.py:101-103) ----------
    batch1 = None
    if rank == 0:
        with torch.no_grad():
            x1 = x[:1]
            for _ in range(5):
                model(adj, x1)
            torch.cuda.synchronize(dev)
            b0e, b1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0e.record()
            for _ in range(50):
                y1 = model(adj, x1)
            b1e.record()
            torch.cuda.synchronize(dev)
        batch1 = {"gpu_ms_per_window": b0e.elapsed_time(b1e) / 50,
                  "what": "model(adj_matrix, batch_x) with batch_x [1,168,34,13] resident on the GPU, 50 calls back to back"}
    barrier()

    # ---------------- training step (BASELINE.json configs[4]): fwd + MSE + bwd + all-reduce + Adam ----------
    train_step = None
    if not args.no_train_step:
        from windgnn_b200 import train as wtrain

        tmodel = windgnn_b200.GCN_GRU(F, F, F, I, H)
        tmodel.load_state_dict(sd, strict=True)
        tmodel = tmodel.to(dev).train()
        trainer = wtrain.Trainer(tmodel, adj, lr=1e-3)   # main.py:45,52
        Bt = args.train_batch
        xt = x[:Bt] if Bt <= Bg else torch.rand((Bt, T, S, F), generator=gen, device=dev)
        yt_ = torch.rand((Bt, T, H), generator=gen, device=dev)
        for _ in range(warmup):
            trainer.step(xt, yt_)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(steps):
            tl = trainer.step(xt, yt_)
        a1.record()
        barrier()
        ms_tr = max_over_ranks(a0.elapsed_time(a1))
        train_step = {"metric": "gcn_gru_train_step_sequences_per_s", "value": world * Bt * steps / (ms_tr * 1e-3),
                      "unit": UNIT, "ms_per_step": ms_tr / steps, "batch_per_gpu": Bt, "global_batch": Bt * world,
                      "loss_after": float(tl.item()),
                      "what": "forward (gates saved) + MSE + BPTT/GEMM/GCN backward + "
                              + ("NCCL sum all-reduce of one flat 167,440-float gradient bucket + " if world > 1 else "")
                              + "fused Adam, all inside the timed region; FP32",
                      "gpu_launches_per_step": 22 + (1 if world > 1 else 0)}
        del trainer, tmodel, xt, yt_
    barrier()

    # ---------------- per-kernel timing for the roofline (rank 0's GPU, same stream) -------------
    dims = (T, S, F, F, F, H)
    stage_ms = {}
    if rank == 0:
        Bc = min(Bg, 148 * 32)
        nbytes = lib.wg_gcn_gru_workspace_bytes(Bc, *dims, Bc, flags)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty((Bc, T, H), device=dev)
        p = [t.detach().contiguous() for t in (
            model.conv1.weight, model.conv1.bias, model.conv2.weight, model.conv2.bias,
            model.gru.weight_ih_l0, model.gru.weight_hh_l0, model.gru.bias_ih_l0, model.gru.bias_hh_l0)]
        st = torch.cuda.current_stream(dev).cuda_stream
        xs = x[:Bc]
        calls = {
            "pack": lambda: lib.wg_stage_pack_f32(*(t.data_ptr() for t in p[4:]), *dims, Bc, flags, ws.data_ptr(), nbytes, local_rank, st),
            "gcn": lambda: lib.wg_stage_gcn_f32(adj.data_ptr(), xs.data_ptr(), *(t.data_ptr() for t in p[:4]), Bc, *dims, Bc, flags, ws.data_ptr(), nbytes, local_rank, st),
            "inproj": lambda: lib.wg_stage_inproj_f32(Bc, *dims, Bc, flags, ws.data_ptr(), nbytes, local_rank, st),
            "recur": lambda: lib.wg_stage_recur_f32(out.data_ptr(), Bc, *dims, Bc, flags, ws.data_ptr(), nbytes, local_rank, st),
        }
        for name in ("pack", "gcn", "inproj", "recur"):
            for _ in range(3):
                _lib.check(calls[name]())
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(5, min(steps, 20))
            a.record()
            for _ in range(reps):
                _lib.check(calls[name]())
            b.record()
            torch.cuda.synchronize(dev)
            stage_ms[name] = a.elapsed_time(b) / reps
        ffma_peak = lib.wg_measure_ffma_tflops(local_rank, 10)
        del ws, out
    barrier()

    # ---------------- end-to-end through the host-buffer entry point ----------------
    e2e = None
    if not args.no_e2e:
        xh = torch.empty((Bg, T, S, F), dtype=torch.float32, pin_memory=True)
        xh.copy_(x)
        oh = torch.empty((Bg, T, H), dtype=torch.float32, pin_memory=True)
        model.chunk = 256  # pipeline granularity: H2D / compute (two lanes) / D2H of consecutive chunks overlap
        e2e_steps = max(3, min(steps, 10))
        with torch.no_grad():
            for _ in range(2):
                model.forward_host(adj, xh, oh)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                model.forward_host(adj, xh, oh)  # blocks until `oh` is complete
            torch.cuda.synchronize(dev)
            dt = max_over_ranks(time.perf_counter() - t0)
        model.chunk = 0
        e2e = {"value": world * Bg * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": Bg * T * S * F * 4, "d2h_bytes_per_step": Bg * T * H * 4,
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps, "pipeline_chunk": 256, "compute_lanes": 2}
        del xh, oh
    barrier()

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
        Bc = min(Bg, 148 * 32)
        rows = Bc * T
        inproj_flops = FLOP_IH * rows
        inproj_tflops = inproj_flops / (stage_ms["inproj"] * 1e-3) / 1e12
        nominal_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
        path_tflops = FLOP_PER_SEQ * Bg * steps / (ms * 1e-3) / 1e12  # per GPU
        hbm_gbs = BYTES_PER_SEQ * Bg * steps / (ms * 1e-3) / 1e9     # per GPU, algorithmic
        roofline = {
            "bound": "fp32", "kernel": ("inproj_kernel (GRU input projection, FFMA2 GEMM)" if flags == 0 else
                                        "inproj_tc_kernel (GRU input projection, tcgen05 3xTF32; fraction is of the "
                                        "FP32 FMA peak the FP32 path is bound by)"),
            "achieved": inproj_tflops, "peak": ffma_peak, "unit": "TFLOP/s", "frac": inproj_tflops / ffma_peak,
            "peak_source": "FFMA microbenchmark measured live on this GPU (wg_measure_ffma_tflops)",
            "peak_nominal": nominal_peak, "traffic": ncu_traffic_bytes("inproj"), "traffic_unit": "bytes per launch (ncu dram read+write)",
            "algorithmic_flops_per_launch": inproj_flops,
            "kernel_ms": stage_ms,
            "path": {"achieved": path_tflops, "frac_fp32": path_tflops / ffma_peak,
                     "hbm_gbs": hbm_gbs, "hbm_peak": hbm_peak, "hbm_peak_source": hbm_src,
                     "frac_hbm": hbm_gbs / hbm_peak, "flop_per_seq": FLOP_PER_SEQ, "bytes_per_seq": BYTES_PER_SEQ},
        }
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
            adj_cpu = adj.cpu()
            v, threads, reps, secs = cpu_port_seq_per_s(sd, adj_cpu, n_seq=2048, min_seconds=12.0, max_reps=40)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{reps} x 2048 sequences of the same workload in {secs:.1f} s, torch CPU fp32"}
            if batch1 is not None:   # the same port called the way the reference calls its model: one window at a time
                from oracle import gcn_gru_forward_torch

                x1c = torch.rand((1, T, S, F))
                for _ in range(3):
                    gcn_gru_forward_torch(adj_cpu, x1c, sd)
                t0 = time.perf_counter()
                for _ in range(30):
                    gcn_gru_forward_torch(adj_cpu, x1c, sd)
                batch1["cpu_port_ms_per_window"] = (time.perf_counter() - t0) / 30 * 1e3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if flags == 0 else "f32 (input projection: 3xTF32 tensor cores, fp32 accumulate)",
            "data": "synthetic",
            "config": {"precision": args.precision,
                       "workload": "34-station GCN-GRU forward (wind_gnn_34.pth), T=168, batch 4096 per GPU "
                                   "(BASELINE.json configs[1])", "S": S, "T": T, "batch_per_gpu": Bg,
                       "global_batch": Bg * world, "parallelism": f"sequence-sharded x{world}, no collective",
                       "l2_policy": "inputs (1.22 GB per step) larger than L2",
                       "station_sequence_predictions_per_s": value * S},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches_per_step * steps,
            "roofline": roofline, "cpu_baseline": cpu, "tensor_path": tensor_path, "train_step": train_step,
            "batch1": batch1,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
