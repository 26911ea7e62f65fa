#!/usr/bin/env python3
"""Headline benchmark: hybrid window scoring throughput (windows/s), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W [--impl shmfast|reference] [--workload ...]

A "step" is one pass of the hot path over one batch of synthetic 4DOF windows per GPU:
fused gather+normalise from the raw series -> LSTM-VAE score -> stored-threshold compare + ascending
compaction -> second VAE pass on the flagged windows -> residual stack -> CNN -> labels
(06_test_full_pipeline.py:327-383), with the threshold calibrated as 04_vae_thresholding.py does (P99 of
a 2,010-window calibration set => ~1 % flagged).  This is BASELINE.json configs[1] (4DOF scoring +
thresholding on 1xB200) extended by the CNN attribution the metric names (configs[2] per GPU).

`value`      device-resident throughput: inputs (series, eps) already in HBM, CUDA-event timed per step.
`e2e`        the same through the public API with HOST buffers: pinned series H2D every step, eps drawn on
             the device like the reference's forward() does, scores/flags/labels D2H every step.
`roofline`   the dominant kernel (fused VAE scorer): algorithmic FLOPs / CUDA-event launch time vs the
             measured dense bf16 peak (MEASURED_PEAKS.json).
`cpu_baseline` / `--impl reference`: oracle/torch_port.py (the reference's wiring on torch.nn CPU kernels)
             timed on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "hybrid-vae-cnn-for-shm_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "windows/sec (VAE score + CNN attribution)"
# algorithmic work per window, SURVEY.md section 8(d) / BASELINE.md section 3
FLOP_PER_WINDOW = {"4dof": 80_322_560, "openlab": 13_527_040, "1dof": 4_248_512}
CNN_FLOP_PER_FLAGGED = {"4dof": 4_070_912, "openlab": 133_851_648}
# ncu --set full capture of vae_score_tc_kernel<128> (profiles/r01_vae_tc_raw.csv): 15.577 GB read + 15.581 GB written
# for a 151,552-window launch
NCU_DRAM_BYTES_PER_WINDOW_4DOF = (16.249489e9 + 15.895659e9) / 151552         # profiles/r01_vae_tc_v5_raw.csv (final kernel of round 1)
NCU_DRAM_BYTES_PER_WINDOW_OPENLAB = (41.604608e6 + 0.883456e6) / 151552      # profiles/r01_vae_tc_dual_raw.csv (algorithmic: 240 + 32 + 4)


def parse_args():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["shmfast", "reference"], default="shmfast")
    ap.add_argument("--workload", choices=["4dof_hybrid", "4dof_score", "openlab_hybrid", "4dof_train", "1dof_score"], default="4dof_hybrid")
    ap.add_argument("--flag-pct", type=float, default=None, help="percentile of the calibration scores used as gate threshold (default: 4dof 99 = ~1 %% flagged, the 04_vae_thresholding.py rule; openlab 95). 53 / 62 reproduce the repo's real test-set flag rates (47 %% / 38 %%)")
    ap.add_argument("--windows", type=int, default=1 << 20, help="windows per GPU per step")
    ap.add_argument("--engine", choices=["auto", "fp32", "tc"], default="auto")
    ap.add_argument("--cpu-sample", type=int, default=16384, help="windows per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--batch", type=int, default=256, help="4dof_train: windows per GPU per optimisation step (03_train_vae.py:53)")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]), tf_sust=float(d["bf16_tflops_sustained"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")      # B200_PROFILING.md fallback


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(dev_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.strip().lower() == "active"})
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def synth_problem(n_windows: int, seed: int):
    from shmfast import synth
    from shmfast.pipeline import guard_std_4dof
    s = synth.STAGES["4dof"]
    series = synth.series(n_windows + s["T"] - 1, s["D"], seed=seed)
    mean, std = synth.stats(s["D"], seed=0)
    return s, series, mean, guard_std_4dof(std)


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle/torch_port.py on the host cores
# ----------------------------------------------------------------------------------------------
PCT4 = 99.0


def cpu_port_run(sample: int, steps: int, warmup: int, thr: float | None, workload: str):
    from oracle import torch_port as TP     # the one place bench.py executes oracle/: as the timed CPU baseline
    from shmfast import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    s, series, mean, std = synth_problem(sample, seed=123)
    vae = TP.VaePort(synth.stage_vae_weights("4dof", seed=0))
    cnn = TP.Cnn4dofPort(synth.cnn4dof_weights(seed=0))
    torch.manual_seed(42)
    if thr is None:
        # P99 of 2,010 windows spread over the stream, as the GPU arm calibrates it
        starts = np.linspace(0, sample - 1, 2010).astype(np.int64)
        Wc = np.stack([series[i:i + s["T"]] for i in starts]).astype(np.float32)
        Zc = np.nan_to_num((Wc - mean[None, None]) / std[None, None], nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
        thr = float(np.percentile(TP.vae_scores_batched(vae, Zc, None, 512), PCT4))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        r = TP.hybrid_4dof(vae, cnn, series, mean, std, thr if workload == "4dof_hybrid" else float("inf"))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return dict(value=sample / (ms / 1e3), ms_per_step=ms, cores=cores, threads=torch.get_num_threads(),
                sample=f"{sample} windows of the {workload} step per timed pass (series gather + normalise + VAE score, batch 512"
                       f" + threshold + second pass/CNN on the flagged), torch.nn CPU kernels, {steps} passes after {warmup} warm-up",
                flagged=int(r["idx"].size), thr=thr)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = min(a.cpu_sample, a.windows)
    r = cpu_port_run(sample, a.steps, a.warmup, None, a.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "windows/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": a.workload, "windows_per_step": sample, "T": 100, "D": 12, "H": 128, "Z": 16, "L": 2,
                   "note": "reference CPU path (torch.nn on host cores) on a bounded sample of the GPU arm's workload"},
        "cpu_baseline": {"value": r["value"], "unit": "windows/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# the product arm
# ----------------------------------------------------------------------------------------------
def run_shmfast(a):
    import torch.distributed as dist

    from shmfast import ops, synth
    from shmfast.pipeline import Hybrid4dof
    from shmfast.shard import max_over_ranks, sum_over_ranks

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl shmfast needs a CUDA device: libshmfast has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)
    if world != a.gpus and rank == 0:
        print(f"[bench] note: --gpus {a.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    engine = {"auto": ops.ENGINE_AUTO, "fp32": ops.ENGINE_FP32, "tc": ops.ENGINE_TC_BF16X3}[a.engine]
    N = a.windows
    s, series_h, mean, std = synth_problem(N, seed=100 + rank)            # each rank scores its own window range
    T, D, Z = s["T"], s["D"], s["Z"]
    vae = ops.VaeScorer(synth.stage_vae_weights("4dof", seed=0), dev, engine=engine)
    cnn = ops.Cnn4dof(synth.cnn4dof_weights(seed=0), dev)
    pk = peaks()

    # threshold calibration, as 04_vae_thresholding.py: P99 of a 2,010-window normal set (device percentile)
    series_pinned = torch.from_numpy(series_h).pin_memory()
    series_d = series_pinned.to(dev, non_blocking=True)
    src = ops.WindowSource(series_d, T, stride=1, mean=mean, std=std, nan_to_zero=True)
    torch.manual_seed(42 + rank)
    cal_idx = torch.linspace(0, N - 1, 2010, device=dev).to(torch.int32)          # 2,010 windows spread over the stream
    cal_scores = vae.score(src, torch.randn((2010, Z), device=dev), idx=cal_idx)["score"]
    pct4 = 99.0 if a.flag_pct is None else a.flag_pct
    thr = float(ops.percentile(cal_scores, pct4).item()) if a.workload == "4dof_hybrid" else float("inf")
    hyb = Hybrid4dof(vae, cnn, thr)

    eps1 = torch.randn((N, Z), device=dev)
    max_flag = min(N, max(4096, int(min(1.0, 5.0 * (100.0 - pct4) / 100.0) * N)))
    eps2 = torch.randn((max_flag, Z), device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)         # > 126 MB L2

    kern_ev = []

    def step(timed: bool):
        """Device-resident step; the first (dominant) kernel is bracketed by its own CUDA events."""
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = vae.score(src, eps1, n=N, want_latent=hybrid, out=step.first)
            e1.record()
            kern_ev.append((e0, e1))
        else:
            out = vae.score(src, eps1, n=N, want_latent=hybrid, out=step.first)
        score = out["score"]
        flag, idx, count = ops.compact(score, thr)
        launches = 4                                                       # scorer + the three compaction kernels
        if a.workload == "4dof_hybrid":
            # second pass with fresh noise; the (deterministic) encoder's mu / logvar of the first pass are reused
            second = vae.rescore(src, out["mu"], out["logvar"], eps2, idx=idx, n=max_flag, n_dev=count, want_cnn_in=True, out=step.buf)
            if second is None:
                second = vae.score(src, eps2, n=max_flag, idx=idx, n_dev=count, want_score=False, want_cnn_in=True, out=step.buf)
            cnn.forward(second["cnn_in"], n=max_flag, n_dev=count, want_labels=True)
            launches += 3
        return launches, count

    step.buf = {}
    step.first = {}
    hybrid = a.workload == "4dof_hybrid"
    launches_per_step = 0
    for _ in range(a.warmup):
        launches_per_step, count = step(False)
    torch.cuda.synchronize()
    n_flag = int(count.item())
    if n_flag > max_flag:                       # size the flagged-subset buffers to the workload and warm up again
        max_flag = min(N, int(1.25 * n_flag))
        eps2 = torch.randn((max_flag, Z), device=dev)
        step.buf = {}
        for _ in range(max(1, a.warmup)):
            launches_per_step, count = step(False)
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = []
    for _ in range(a.steps):
        flush.zero_()                                                      # L2 flush between timed iterations
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        step(True)
        s1.record()
        ev.append((s0, s1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    step_ms = [x.elapsed_time(y) for x, y in ev]
    total_ms = max_over_ranks(sum(step_ms), dev)
    kern_ms = sum(x.elapsed_time(y) for x, y in kern_ev) / len(kern_ev)
    windows_total = sum_over_ranks(float(N * a.steps), dev)
    value = windows_total / (total_ms / 1e3)

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not a.no_e2e:
        from shmfast.stream import HostStream, scatter_flagged
        outs = {"score": ((N,), torch.float32)}
        if a.workload == "4dof_hybrid":
            outs.update(y_pred=((N,), torch.int64), p_struct=((N,), torch.float32), count=((1,), torch.int32))
        pipe = HostStream(dev, series_pinned.shape, outs)

        def e2e_chunk(sd_, i):                                             # device work of one chunk, current stream, no host sync
            src_ = ops.WindowSource(sd_, T, stride=1, mean=mean, std=std, nan_to_zero=True)
            e1_ = torch.randn((N, Z), device=dev)                          # the reference draws eps on the device too
            if a.workload != "4dof_hybrid":
                return dict(score=vae.score(src_, e1_, n=N)["score"])
            e2_ = torch.randn((max_flag, Z), device=dev)
            res = hyb.run(src_, e1_, e2_, n=N, sync_count=False, max_flagged=max_flag)
            y_pred, p_full = scatter_flagged(res["idx"], res["count"], N, [res["label"], res["p_struct"]], [torch.int64, torch.float32])
            return dict(score=res["score"], y_pred=y_pred, p_struct=p_full, count=res["count"].reshape(1))

        def e2e_run(k):                                                    # k chunks: H2D(i+1) and D2H(i-1) run under chunk i's kernels
            got = 0
            for _, host in pipe.run((series_pinned for _ in range(k)), e2e_chunk):
                got += int(host["score"].shape[0])
            return got

        e2e_run(2)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        assert e2e_run(a.steps) == N * a.steps
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
        d2h = N * 4 + (N * 8 + N * 4 + 4 if a.workload == "4dof_hybrid" else 0)
        e2e = {"value": windows_total / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": int(series_pinned.numel() * 4),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / a.steps,
               "api": "shmfast.stream.HostStream around shmfast.pipeline.Hybrid4dof.run + scatter_flagged (pinned host series in, scores/labels/p_struct out per chunk; copies on side streams under the next chunk's kernels)"}

    cpu_baseline = None
    torch_cuda = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_port_run(min(a.cpu_sample, N), 1, 1, thr, a.workload)
        cpu_baseline = {"value": r["value"], "unit": "windows/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        # the incumbent on the same GPU: the reference's wiring on stock PyTorch CUDA kernels
        # (cuDNN LSTM / conv), batch 512 with the reference's per-batch host<->device copies.  Baseline leg only.
        try:
            from oracle import torch_port as TP
            ns = min(65536, N)
            vae_c = TP.VaePort(synth.stage_vae_weights("4dof", seed=0)).to(dev)
            cnn_c = TP.Cnn4dofPort(synth.cnn4dof_weights(seed=0)).to(dev)
            ser = series_h[: ns + T - 1]
            TP.hybrid_4dof_device(vae_c, cnn_c, ser[: 4096 + T - 1], mean, std, thr, dev)      # warm-up (cuDNN plans)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc_ = TP.hybrid_4dof_device(vae_c, cnn_c, ser, mean, std, thr, dev)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            torch_cuda = {"value": ns / dt, "unit": "windows/s", "kind": "port on torch CUDA kernels (cuDNN LSTM/conv), fp32",
                          "sample": f"{ns} windows, batch 512, host windows + per-batch H2D/D2H as the reference does",
                          "flagged": int(rc_["idx"].size)}
            del vae_c, cnn_c
        except Exception as e:                                   # a baseline must never break the bench line
            torch_cuda = {"unavailable": repr(e)[:200]}

    if rank == 0:
        eng_name = {ops.ENGINE_FP32: "fp32", ops.ENGINE_TC_BF16X3: "tc_bf16x3"}[vae.engine]
        flops = FLOP_PER_WINDOW["4dof"] * N
        achieved = flops / (kern_ms / 1e3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if vae.engine == ops.ENGINE_FP32 else "bf16x3->f32",
            "data": "synthetic",
            "config": {"workload": a.workload, "windows_per_gpu": N, "T": T, "D": D, "H": s["H"], "Z": Z, "L": s["L"],
                       "engine": eng_name, "threshold": f"P{pct4:g} of 2010 calibration windows" if a.workload == "4dof_hybrid" else None,
                       "flagged_per_gpu": n_flag, "input": "raw series, stride 1, gather+normalise fused into the scorer",
                       "l2": "flushed between timed steps (512 MiB memset)", "parallelism": f"window-range shards x{world}, no collective"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * a.steps,
            "roofline": {"bound": "tensor", "kernel": "vae_score (fused LSTM-VAE scorer, first pass)", "achieved": achieved,
                         "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"],
                         "traffic": NCU_DRAM_BYTES_PER_WINDOW_4DOF * N if vae.engine == ops.ENGINE_TC_BF16X3 else None,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of profiles/r01_vae_tc_raw.csv (151,552-window "
                                         "launch: 31.16 GB) scaled per window; it is the layer-0 h_t stream (T x 64 KB per tile, "
                                         "written once + read once), not input re-reads: the input is 48 B/window",
                         "peak_source": pk["source"] + " bf16 dense, sustained", "kernel_ms": kern_ms,
                         "algorithmic_flop_per_window": FLOP_PER_WINDOW["4dof"], "engine": eng_name},
            "cpu_baseline": cpu_baseline, "torch_cuda_baseline": torch_cuda,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# openLAB hybrid (BASELINE.json configs[3] per GPU): gate on 3 clean channels (T=200, stride 20), CNN on the
# flagged raw 4-channel windows; thresholds P95 (default) of a 2,000-window calibration sample and 0.5.
# ----------------------------------------------------------------------------------------------
def openlab_problem(n_windows: int, seed: int):
    from shmfast import synth
    rows = (n_windows - 1) * 20 + 200
    series = synth.series(rows, 4, seed=seed, nan_frac=0.0007)          # ~1.35 % of the raw windows carry a NaN run
    vmu, vsd = synth.stats(3, seed=1)
    cmu, csd = synth.stats(4, seed=2)
    return series, vmu, vsd, cmu, csd


def run_openlab(a):
    import torch.distributed as dist

    from shmfast import ops, synth
    from shmfast.pipeline import HybridOpenLab
    from shmfast.shard import max_over_ranks, sum_over_ranks

    pct = 95.0 if a.flag_pct is None else a.flag_pct
    def cpu_port(warmup, steps):
        """The reference's openLAB wiring (10_test_hybrid_pipeline.py:233-302,351-367) on torch.nn CPU kernels, all host threads."""
        from oracle import torch_port as TP
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n = min(a.cpu_sample, a.windows)
        series, vmu, vsd, cmu, csd = openlab_problem(n, seed=123)
        vae = TP.VaePort(synth.stage_vae_weights("openlab", seed=0))
        cnn = TP.CnnOpenLabPort(synth.cnnol_weights(seed=0))
        torch.manual_seed(42)
        cal = TP.hybrid_openlab(vae, cnn, series[: (min(2000, n) - 1) * 20 + 200], [1, 2, 3], vmu, vsd, cmu, csd, float("inf"), 0.5)
        thr = float(np.percentile(cal["score"], pct))
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            r = TP.hybrid_openlab(vae, cnn, series, [1, 2, 3], vmu, vsd, cmu, csd, thr, 0.5)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        ms = 1e3 * sum(times) / len(times)
        sample = f"{n} windows of the openlab_hybrid step per timed pass (torch.nn CPU kernels, batch 256), flagged {int(r['mask'].sum())}"
        return n / (ms / 1e3), ms, cores, sample, n

    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        val, ms, cores, sample, n = cpu_port(a.warmup, a.steps)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s", "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": {"workload": "openlab_hybrid", "windows_per_step": n},
                          "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)
    N = a.windows
    series_h, vmu, vsd, cmu, csd = openlab_problem(N, seed=100 + rank)
    vae = ops.VaeScorer(synth.stage_vae_weights("openlab", seed=0), dev,
                        engine={"auto": ops.ENGINE_AUTO, "fp32": ops.ENGINE_FP32, "tc": ops.ENGINE_TC_BF16X3}[a.engine])
    cnn = ops.CnnOpenLab(synth.cnnol_weights(seed=0), dev)
    pinned = torch.from_numpy(series_h).pin_memory()
    series_d = pinned.to(dev, non_blocking=True)

    def sources(sd_):
        g = ops.WindowSource(sd_, 200, stride=20, chan=[1, 2, 3], mean=vmu, std=vsd, clip=10.0, nan_to_zero=True)
        r = ops.WindowSource(sd_, 200, stride=20, mean=cmu, std=csd, clip=10.0, nan_to_zero=True)
        return g, r

    src_g, src_r = sources(series_d)
    torch.manual_seed(42 + rank)
    cal_idx = torch.linspace(0, N - 1, 2000, device=dev).to(torch.int32)
    thr = float(ops.percentile(vae.score(src_g, torch.randn((2000, 8), device=dev), idx=cal_idx)["score"], pct).item())
    hyb = HybridOpenLab(vae, cnn, thr, 0.5)
    eps = torch.randn((N, 8), device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    max_flag = min(N, max(1024, int((100.0 - pct) / 100.0 * 1.5 * N)))
    kern_ev = []

    def step(timed):
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); score = vae.score(src_g, eps, n=N)["score"]; e1.record()
            kern_ev.append((e0, e1))
        else:
            score = vae.score(src_g, eps, n=N)["score"]
        flag, idx, count = ops.compact(score, thr)
        cnn.forward(src_r, n=max_flag, idx=idx, n_dev=count, want_prob=True)
        return count

    for _ in range(a.warmup):
        count = step(False)
    torch.cuda.synchronize()
    n_flag = int(count.item())
    if n_flag > max_flag:
        max_flag = min(N, int(1.25 * n_flag))
        count = step(False); torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = []
    for _ in range(a.steps):
        flush.zero_()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); step(True); s1.record()
        ev.append((s0, s1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(sum(x.elapsed_time(y) for x, y in ev), dev)
    kern_ms = sum(x.elapsed_time(y) for x, y in kern_ev) / len(kern_ev)
    windows_total = sum_over_ranks(float(N * a.steps), dev)
    value = windows_total / (total_ms / 1e3)

    from shmfast.stream import HostStream, scatter_flagged
    pipe = HostStream(dev, pinned.shape, {"score": ((N,), torch.float32), "flag": ((N,), torch.uint8), "pred": ((N,), torch.int64),
                                          "prob": ((N,), torch.float64), "count": ((1,), torch.int32)})

    def e2e_chunk(sd_, i):                                                 # device work of one chunk, current stream, no host sync
        g, r = sources(sd_)
        res = hyb.run(g, r, torch.randn((N, 8), device=dev), n=N, sync_count=False, max_flagged=max_flag)
        pred, prob = scatter_flagged(res["idx"], res["count"], N, [res["pred"], res["prob"]], [torch.int64, torch.float64])
        return dict(score=res["score"], flag=res["flag"], pred=pred, prob=prob, count=res["count"].reshape(1))

    def e2e_run(k):                                                        # H2D(i+1) and D2H(i-1) run under chunk i's kernels
        got = 0
        for _, host in pipe.run((pinned for _ in range(k)), e2e_chunk):
            got += int(host["score"].shape[0])
        return got

    e2e_run(2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    assert e2e_run(a.steps) == N * a.steps
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
    if rank == 0:
        cpu_baseline = None
        if world == 1 and not a.no_cpu_baseline:
            val, _, cores, sample, _ = cpu_port(1, 1)
            cpu_baseline = {"value": val, "unit": "windows/s", "cores": cores, "kind": "port", "sample": sample}
        pk = peaks()
        eng = {ops.ENGINE_FP32: "fp32", ops.ENGINE_TC_BF16X3: "tc_bf16x3"}[vae.engine]
        achieved = FLOP_PER_WINDOW["openlab"] * N / (kern_ms / 1e3) / 1e12
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if vae.engine == ops.ENGINE_FP32 else "bf16x3->f32", "data": "synthetic",
            "config": {"workload": "openlab_hybrid", "windows_per_gpu": N, "T": 200, "stride": 20, "D_gate": 3, "D_raw": 4, "H": 64, "Z": 8, "L": 1,
                       "engine": eng, "cnn_engine": {ops.ENGINE_FP32: "fp32", ops.ENGINE_TC_BF16X3: "tc_f16x3 (tcgen05 implicit GEMM)"}[cnn.engine],
                       "gate_threshold": f"P{pct:g} of 2000 calibration windows", "flagged_per_gpu": n_flag, "cnn_threshold": 0.5,
                       "input": "raw 4-channel series with NaN runs, stride 20; gather+standardise fused into scorer and CNN",
                       "l2": "flushed between timed steps (512 MiB memset)", "parallelism": f"window-range shards x{world}, no collective"},
            "clocks": clocks,
            "e2e": {"value": windows_total / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": int(pinned.numel() * 4),
                    "d2h_bytes_per_step": int(N * (4 + 1 + 8 + 8) + 4), "ms_per_step": 1e3 * e2e_s / a.steps,
                    "api": "shmfast.stream.HostStream around shmfast.pipeline.HybridOpenLab.run + scatter_flagged (copies on side streams under the next chunk's kernels)"},
            "gpu_launches": (4 + 12 * ((n_flag + 8191) // 8192)) * a.steps,      # scorer, 3 compaction kernels, 12 CNN kernels per 8192-window chunk
            "roofline": {"bound": "tensor", "kernel": "vae_score (fused LSTM-VAE scorer)", "achieved": achieved, "peak": pk["tf_sust"],
                         "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"],
                         "traffic": NCU_DRAM_BYTES_PER_WINDOW_OPENLAB * N if vae.engine == ops.ENGINE_TC_BF16X3 else None,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of profiles/r01_vae_tc_dual_raw.csv (151,552-window launch) scaled to this launch",
                         "kernel_ms": kern_ms,
                         "peak_source": pk["source"] + " bf16 dense, sustained", "algorithmic_flop_per_window": FLOP_PER_WINDOW["openlab"],
                         "engine": eng},
            "cpu_baseline": cpu_baseline}), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ----------------------------------------------------------------------------------------------
# 4DOF training step (BASELINE.json configs[4]): forward (train mode, dropout 0.3) -> ELBO -> BPTT -> one gradient
# all-reduce -> clip 2.0 + Adam, batch 256 per GPU (03_train_vae.py:53,260-271).  Weak scaling: global batch 256*N.
# ----------------------------------------------------------------------------------------------
TRAIN_FLOP_PER_WINDOW = 3 * 80_322_560       # forward + ~2x for the backward contractions


def run_train(a):
    import torch.distributed as dist

    from shmfast import synth, train
    from shmfast.shard import max_over_ranks, sum_over_ranks

    B, T, D, Z = a.batch, 100, 12, 16
    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        from oracle import torch_port as TP
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        port = TP.VaeTrainPort(synth.stage_vae_weights("4dof", seed=0)).train()
        opt = torch.optim.Adam(port.ordered_parameters(), lr=1e-3, weight_decay=1e-5)
        X = torch.from_numpy(synth.windows(B, T, D, seed=123))
        rng = np.random.Generator(np.random.PCG64(9))
        times = []
        for i in range(a.warmup + a.steps):
            masks = [torch.from_numpy((rng.random((1, B, T, 128)) >= 0.3).astype(np.uint8)) for _ in range(2)]
            t0 = time.perf_counter()
            TP.train_step_port(port, opt, X, torch.randn(B, Z), 0.5, masks[0], masks[1], 0.3)
            if i >= a.warmup:
                times.append(time.perf_counter() - t0)
        ms = 1e3 * sum(times) / len(times)
        val = B / (ms / 1e3)
        print(json.dumps({"impl": "reference", "metric": "training windows/sec (4DOF LSTM-VAE step)", "value": val, "unit": "windows/s",
                          "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "4dof_train", "batch_per_step": B, "T": T, "D": D, "H": 128, "Z": Z, "L": 2},
                          "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": "port",
                                           "sample": f"{a.steps} optimisation steps of batch {B} (torch autograd, nn.LSTM CPU kernels)"},
                          "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)
    from shmfast.models import fourdof
    vae = fourdof.TemporalVAE(12, 16, 128, 2, dropout=0.3)
    vae.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.stage_vae_weights("4dof", seed=0).items()})
    vae = vae.to(dev).train()
    tr = train.VaeTrainer(vae, T, B)
    n_batches = 8
    host = [torch.from_numpy(synth.windows(B, T, D, seed=1000 * rank + i)).pin_memory() for i in range(n_batches)]
    devb = [h.to(dev) for h in host]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    torch.manual_seed(42 + rank)
    for i in range(a.warmup):
        tr.step(devb[i % n_batches], 0.5)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = []
    for i in range(a.steps):
        flush.zero_()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); loss3 = tr.step(devb[i % n_batches], 0.5); s1.record()
        ev.append((s0, s1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(sum(x.elapsed_time(y) for x, y in ev), dev)
    windows_total = sum_over_ranks(float(B * a.steps), dev)
    value = windows_total / (total_ms / 1e3)
    # phase breakdown on rank 0 (forward / ELBO / backward / all-reduce + optimiser), CUDA events on the launch stream
    phases = None
    if rank == 0:
        xb = devb[0]
        eps = torch.randn((B, Z), device=dev)
        masks = train.draw_dropout_masks(2, B, T, 128, 0.3, dev)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        flush.zero_()
        marks[0].record()
        xhat, mu, lv = tr.handle.forward(tr.flat, xb, eps, masks[0], masks[1], 0.3)
        marks[1].record()
        l3, dx, dm, dl = train.elbo_grad(xb, xhat, mu, lv, 0.5)
        marks[2].record()
        tr.handle.backward(tr.flat, dx, dm, dl, tr.grads)
        marks[3].record()
        train.adam_clip_step(tr.flat, tr.grads, tr.exp_avg, tr.exp_avg_sq, tr.steps + 1, 1e-3, weight_decay=1e-5, max_norm=2.0)
        marks[4].record()
        torch.cuda.synchronize()
        phases = {n: marks[i].elapsed_time(marks[i + 1]) for i, n in enumerate(("forward_ms", "elbo_ms", "backward_ms", "clip_adam_ms"))}
    # end to end: pinned host batch in, loss out, every step
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(a.steps):
        xb = host[i % n_batches].to(dev, non_blocking=True)
        loss_host = tr.step(xb, 0.5).cpu()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import torch_port as TP
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        port = TP.VaeTrainPort(synth.stage_vae_weights("4dof", seed=0)).train()
        opt = torch.optim.Adam(port.ordered_parameters(), lr=1e-3, weight_decay=1e-5)
        X = host[0]
        ts = []
        for i in range(4):
            t1 = time.perf_counter()
            TP.train_step_port(port, opt, X, torch.randn(B, Z), 0.5, masks[0].cpu(), masks[1].cpu(), 0.3)
            ts.append(time.perf_counter() - t1)
        cpu_ms = 1e3 * sum(ts[1:]) / 3
        cpu_baseline = {"value": B / (cpu_ms / 1e3), "unit": "windows/s", "cores": cores, "kind": "port",
                        "sample": f"3 optimisation steps of batch {B} after 1 warm-up (torch autograd over nn.LSTM CPU kernels), {cpu_ms:.0f} ms/step"}
    if rank == 0:
        ms = total_ms / a.steps
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        achieved = TRAIN_FLOP_PER_WINDOW * B / (ms / 1e3) / 1e12
        print(json.dumps({
            "metric": "training windows/sec (4DOF LSTM-VAE step)", "value": value, "unit": "windows/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "4dof_train", "batch_per_gpu": B, "global_batch": B * world, "T": T, "D": D, "H": 128, "Z": Z, "L": 2,
                       "dropout": 0.3, "optimizer": "Adam lr 1e-3 wd 1e-5, clip 2.0", "l2": "flushed between timed steps (512 MiB memset)",
                       "parallelism": f"dp{world}, one NCCL all-reduce of the 477,100-float gradient per step"},
            "clocks": clocks, "phases": phases,
            "e2e": {"value": windows_total / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": B * T * D * 4, "d2h_bytes_per_step": 12,
                    "ms_per_step": 1e3 * e2e_s / a.steps, "api": "shmfast.train.VaeTrainer.step (pinned host batch in, loss out)"},
            "gpu_launches": 47 * a.steps,        # profiles/r01_train_launches_v2.csv: 141 libshmfast launches in 3 steps (21 contractions, 8 recurrence, ...)
            "roofline": {"bound": "fp32", "kernel": "whole step (fp32 FMA contractions + resident-weight recurrence)", "achieved": achieved,
                         "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak, "traffic": None,
                         "peak_source": "derived: 148 SM x 128 FMA lanes x 2 x 1.965 GHz", "algorithmic_flop_per_window": TRAIN_FLOP_PER_WINDOW},
            "cpu_baseline": cpu_baseline}), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ----------------------------------------------------------------------------------------------
# 1_DOF (BASELINE.json configs[0], the reference's CPU-runnable case): standardise + window (T=80, stride 1) -> LSTM-VAE
# (H=32, L=2, no LayerNorm; fp32 engine) -> reconstruction -> overlap-average stitch, de-standardise, RMSE per 100-sample
# segment (1_DOF/Scripts/04_test_seen_variants.py:281-311, datasets.py:17-71).
# ----------------------------------------------------------------------------------------------
def run_onedof(a):
    from shmfast import ops, synth
    from shmfast.shard import max_over_ranks, sum_over_ranks
    import torch.distributed as dist

    T, D, Z = 80, 12, 5
    N = min(a.windows, 1 << 18)                               # the reconstruction [N,80,12] is materialised for the stitch
    rows = N + T - 1
    rng = np.random.Generator(np.random.PCG64(7))
    series_h = synth.series(rows, D, seed=5)
    mean = series_h.mean(axis=0).astype(np.float64)
    std = series_h.std(axis=0).astype(np.float64) + 1e-8
    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        from oracle import np_oracle as O, torch_port as TP
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n = min(a.cpu_sample, N)
        vae = TP.VaePort(synth.stage_vae_weights("1dof", seed=0))
        ser = series_h[: n + T - 1]
        times = []
        for i in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            xn = ((ser - mean) / std).astype(np.float32)
            W = np.stack([xn[j:j + T] for j in range(n)])
            with torch.no_grad():
                rec = vae(torch.from_numpy(W), torch.randn(n, Z))[0].numpy()
            y = O.destandardize(O.stitch_windows(rec, ser.shape[0], 1), mean, std)
            O.segment_rmse(ser, y, 100)
            if i >= a.warmup:
                times.append(time.perf_counter() - t0)
        ms = 1e3 * sum(times) / len(times)
        val = n / (ms / 1e3)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s", "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": {"workload": "1dof_score", "windows_per_step": n},
                          "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": "port",
                                           "sample": f"{n} windows per pass: standardise + window + VAE (torch.nn) + stitch + segment RMSE (NumPy)"},
                          "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)
    vae = ops.VaeScorer(synth.stage_vae_weights("1dof", seed=0), dev)
    pinned = torch.from_numpy(series_h).pin_memory()
    series_d = pinned.to(dev)
    eps = torch.randn((N, Z), device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    buf = {}

    def step(sd_):
        src = ops.WindowSource(sd_, T, stride=1, mean=mean.astype(np.float32), std=std.astype(np.float32))
        out = vae.score(src, eps, n=N, want_recon=True, out=buf)
        return ops.stitch_segment_rmse(out["recon"], rows, 1, mean, std, sd_, 100, want_series=False)[1], out["score"]

    for _ in range(a.warmup):
        step(series_d)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = []
    for _ in range(a.steps):
        flush.zero_()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); step(series_d); s1.record()
        ev.append((s0, s1))
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(sum(x.elapsed_time(y) for x, y in ev), dev)
    windows_total = sum_over_ranks(float(N * a.steps), dev)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        rm, sc = step(pinned.to(dev, non_blocking=True))
        rm_h, sc_h = rm.cpu(), sc.cpu()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
    if rank == 0:
        ms = total_ms / a.steps
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        achieved = FLOP_PER_WINDOW["1dof"] * N / (ms / 1e3) / 1e12
        print(json.dumps({
            "metric": METRIC, "value": windows_total / (total_ms / 1e3), "unit": "windows/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "1dof_score", "windows_per_gpu": N, "T": T, "D": D, "H": 32, "Z": Z, "L": 2, "engine": "fp32",
                       "post": "overlap-average stitch + de-standardise + RMSE per 100-sample segment (fp64)",
                       "l2": "flushed between timed steps (512 MiB memset)", "parallelism": f"window-range shards x{world}, no collective"},
            "clocks": clocks,
            "e2e": {"value": windows_total / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": int(pinned.numel() * 4),
                    "d2h_bytes_per_step": int(N * 4 + rm_h.numel() * 8), "ms_per_step": 1e3 * e2e_s / a.steps,
                    "api": "ops.VaeScorer.score(want_recon) + ops.stitch_segment_rmse"},
            "gpu_launches": 2 * a.steps,
            "roofline": {"bound": "fp32", "kernel": "vae_score_fp32_kernel<32> (whole step)", "achieved": achieved, "peak": fp32_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp32_peak, "traffic": None,
                         "peak_source": "derived: 148 SM x 128 FMA lanes x 2 x 1.965 GHz", "algorithmic_flop_per_window": FLOP_PER_WINDOW["1dof"]},
            "cpu_baseline": None}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    global PCT4
    a = parse_args()
    if a.flag_pct is not None:
        PCT4 = a.flag_pct
    if a.workload == "openlab_hybrid":
        run_openlab(a)
    elif a.workload == "4dof_train":
        run_train(a)
    elif a.workload == "1dof_score":
        run_onedof(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_shmfast(a)


if __name__ == "__main__":
    main()
