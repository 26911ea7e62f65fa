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
`cpu_baseline` / `--impl reference`: the reference's own, unmodified 06_test_full_pipeline.main() (files staged into
             oracle/_ref at build time, `kind: "reference"`) on the host cores on a bounded sample of the same workload;
             oracle/torch_port.py (`kind: "port"`) when the staged files are absent.
`torch_cuda_baseline`: the incumbent on the same GPU -- the reference's model classes on stock PyTorch CUDA kernels
             (cuDNN LSTM / conv): with the reference's per-batch host copies, and device resident at batch 512 / 8192.
`secondary`  (default workload only) the other BASELINE configs measured in the same run: openLAB hybrid at the P95 and
             38 % flag rates, 4DOF hybrid at the repo's 47 % flag rate, the DDP training step (NCCL all-reduce at N ranks),
             the HBM-bound kernels against the copy bandwidth, and a one-series N-way shard check (bit equality).
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
# dram__bytes_read.sum + dram__bytes_write.sum per window of the scorer kernels, from the ncu --set full captures named below
NCU_TRAFFIC = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text()) if (ROOT / "profiles" / "ncu_traffic.json").exists() else {}
NCU_DRAM_BYTES_PER_WINDOW_4DOF = NCU_TRAFFIC.get("4dof", {}).get("bytes_per_window", (16.249489e9 + 15.895659e9) / 151552)
NCU_SOURCE_4DOF = NCU_TRAFFIC.get("4dof", {}).get("source", "profiles/r01_vae_tc_v5_raw.csv")
NCU_DRAM_BYTES_PER_WINDOW_OPENLAB = NCU_TRAFFIC.get("openlab", {}).get("bytes_per_window", (41.604608e6 + 0.883456e6) / 151552)
NCU_SOURCE_OPENLAB = NCU_TRAFFIC.get("openlab", {}).get("source", "profiles/r01_vae_tc_dual_raw.csv")
TRAIN_FLOP_PER_WINDOW = 3 * 80_322_560       # forward + ~2x for the backward contractions
FP32_PEAK_TF = 148 * 128 * 2 * 1.965e9 / 1e12


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the `secondary` object of the default workload")
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


class Ctx:
    """One process per GPU: rank / device / (optional) NCCL group, shared by every measurement of the run."""

    def __init__(self):
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py --impl shmfast needs a CUDA device: libshmfast has no CPU fallback")
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = dist
        if self.world > 1:
            dist.init_process_group("nccl", init_method="env://", device_id=self.dev)
        self.pk = peaks()
        self._flush = None

    @property
    def flush(self):
        if self._flush is None:
            self._flush = torch.empty(512 << 20, dtype=torch.uint8, device=self.dev)         # > 126 MB L2
        return self._flush

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max(self, v: float) -> float:
        from shmfast.shard import max_over_ranks
        return max_over_ranks(v, self.dev)

    def sum(self, v: float) -> float:
        from shmfast.shard import sum_over_ranks
        return sum_over_ranks(v, self.dev)

    def timed(self, step, steps: int, sample_clocks: bool = False):
        """EXACTLY `steps` steps, barrier + synchronize on both sides, CUDA events per step on the launch stream, L2 flushed
        between steps; returns (total ms = max over ranks, clocks)."""
        sampler = ClockSampler(self.local) if (sample_clocks and self.rank == 0) else None
        self.barrier()
        ev = []
        for i in range(steps):
            self.flush.zero_()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(); step(i); s1.record()
            ev.append((s0, s1))
        self.barrier()
        clocks = sampler.stop() if sampler else None
        return self.max(sum(x.elapsed_time(y) for x, y in ev)), clocks

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def synth_problem(n_windows: int, seed: int):
    from shmfast import synth
    from shmfast.pipeline import guard_std_4dof
    s = synth.STAGES["4dof"]
    series = synth.series(n_windows + s["T"] - 1, s["D"], seed=seed)
    mean, std = synth.stats(s["D"], seed=0)
    return s, series, mean, guard_std_4dof(std)


ENGINES = {"auto": 0, "fp32": 1, "tc": 2}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own files (oracle/_ref) or the port, on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference_4dof(sample: int, steps: int, warmup: int, pct: float, workload: str):
    """The reference's CPU implementation of the 4DOF step on `sample` windows per pass.  Preferred: its unmodified
    06_test_full_pipeline.main() (eval_group x 3 groups) from oracle/_ref -- kind "reference"; fallback: oracle/torch_port."""
    from oracle import ref_driver as R
    from oracle import torch_port as TP     # bench.py executes oracle/ only here: as the timed CPU baseline
    from shmfast import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vae_sd, cnn_sd = synth.stage_vae_weights("4dof", seed=0), synth.cnn4dof_weights(seed=0)
    s, series, mean, std = synth_problem(sample, seed=123)
    torch.manual_seed(42)
    port = TP.VaePort(vae_sd)

    def calibrate(windows_of):
        """P<pct> of 2,010 windows spread over the timed windows, as the GPU arm calibrates it (04_vae_thresholding.py:283)."""
        if workload != "4dof_hybrid":
            return float("inf")
        Wc = np.stack([windows_of(i) for i in np.linspace(0, sample - 1, 2010).astype(np.int64)]).astype(np.float32)
        Zc = np.nan_to_num((Wc - mean[None, None]) / std[None, None], nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
        return float(np.percentile(TP.vae_scores_batched(port, Zc, None, 512), pct))

    if R.ref_root() is not None:
        per = sample // 3
        rows = R.rows_for_windows(per)
        full = synth.series(3 * rows, s["D"], seed=123)
        series3 = [full[g * rows:(g + 1) * rows] for g in range(3)]
        tails = [x[int(rows * 0.7):int(rows * 1.0)] for x in series3]              # slice_frac(0.7, 1.0), 06:98-103
        per_g = tails[0].shape[0] - s["T"] + 1
        sample = 3 * per_g
        thr = calibrate(lambda i: tails[i // per_g][i % per_g:i % per_g + s["T"]])
        with tempfile.TemporaryDirectory() as td:
            r = R.bench_4dof_reference(Path(td), vae_sd, cnn_sd, mean, std, thr if np.isfinite(thr) else 3.0e38, series3, steps, warmup)
        ms = 1e3 * r["seconds"]
        n = r["windows"]
        return dict(value=n / r["seconds"], ms_per_step=ms, cores=cores, threads=torch.get_num_threads(), kind="reference", windows=n,
                    sample=f"{n} windows per timed pass through the reference's unmodified 06_test_full_pipeline.main() (eval_group x 3 groups: "
                           f"make_windows + normalize_windows + VAE score loop batch 512 + threshold + second pass/CNN on the flagged), reference "
                           f"Models on torch CPU kernels, CSV parsing and plotting stubbed out, {steps} passes after {warmup} warm-up",
                    flagged=r["flagged"], thr=thr)
    cnn = TP.Cnn4dofPort(cnn_sd)
    thr = calibrate(lambda i: series[i:i + s["T"]])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        r = TP.hybrid_4dof(port, cnn, series, mean, std, thr)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return dict(value=sample / (ms / 1e3), ms_per_step=ms, cores=cores, threads=torch.get_num_threads(), kind="port", windows=sample,
                sample=f"{sample} windows of the {workload} step per timed pass (series gather + normalise + VAE score, batch 512"
                       f" + threshold + second pass/CNN on the flagged), torch.nn CPU kernels, {steps} passes after {warmup} warm-up",
                flagged=int(r["idx"].size), thr=thr)


def reference_line(a, r, metric, config):
    return {"impl": "reference", "metric": metric, "value": r["value"], "unit": "windows/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["value"], "unit": "windows/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def run_reference_4dof(a):
    pct = 99.0 if a.flag_pct is None else a.flag_pct
    r = cpu_reference_4dof(min(a.cpu_sample, a.windows), a.steps, a.warmup, pct, a.workload)
    cfg = {"workload": a.workload, "windows_per_step": r["windows"], "T": 100, "D": 12, "H": 128, "Z": 16, "L": 2,
           "flagged_per_step": r["flagged"],
           "note": "reference CPU path on a bounded sample of the GPU arm's workload (same shapes, weights, threshold rule)"}
    print(json.dumps(reference_line(a, r, METRIC, cfg)), flush=True)


# ----------------------------------------------------------------------------------------------
# incumbents on the same GPU: the reference's model classes on stock PyTorch CUDA kernels
# ----------------------------------------------------------------------------------------------
def _ref_models_4dof(dev):
    """(vae(x) -> recon, cnn(x) -> logits, kind): the reference's own classes when staged, else the port."""
    from oracle import ref_driver as R
    from shmfast import synth
    vae_sd, cnn_sd = synth.stage_vae_weights("4dof", seed=0), synth.cnn4dof_weights(seed=0)
    if R.ref_root() is not None:
        V, Cn = R.reference_models("4dof")
        vae = V(input_dim=12, latent_dim=16, hidden_dim=128, num_layers=2, dropout=0.3)
        cnn = Cn(input_channels=2, num_classes=2, dropout_rate=0.5)
        vae.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in vae_sd.items()})
        cnn.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in cnn_sd.items()})
        vae, cnn = vae.to(dev).eval(), cnn.to(dev).eval()
        return (lambda x: vae(x)[0]), cnn, "reference Models/ classes"
    from oracle import torch_port as TP
    vae, cnn = TP.VaePort(vae_sd).to(dev), TP.Cnn4dofPort(cnn_sd).to(dev)
    return (lambda x: vae(x, torch.randn((x.shape[0], 16), device=x.device))[0]), cnn, "port of the reference Models/"


@torch.no_grad()
def incumbents_4dof(ctx, series_h, mean, std, thr, T=100):
    """Stock PyTorch on the same B200 (SURVEY.md section 8d "report all three side by side"), fp32, cuDNN LSTM / conv:
    host_windows  -- 06's data movement: windows pre-built on the HOST (outside the timed region), every batch of 512 goes
                     host -> device and its scores / labels come back (06_test_full_pipeline.py:338-344,358-372);
    resident_b512 / resident_b8192 -- normalised windows already on the device, no host loop work in the timed region."""
    dev = ctx.dev
    vae, cnn, kind = _ref_models_4dof(dev)
    ns = min(65536, series_h.shape[0] - T + 1)
    ser = torch.from_numpy(series_h[: ns + T - 1]).to(dev)
    Zd = ser.unfold(0, T, 1).permute(0, 2, 1)                                   # [ns, T, D] view
    Zd = torch.nan_to_num((Zd - torch.from_numpy(mean).to(dev)) / torch.from_numpy(std).to(dev), nan=0.0, posinf=0.0, neginf=0.0).contiguous()
    Zh = Zd.cpu().numpy()
    out = {"kind": kind + " on torch CUDA kernels (cuDNN LSTM / conv), fp32", "sample_windows": ns}

    def run(batch, host):
        score = torch.empty((ns,), device=dev)
        for i in range(0, ns, batch):
            xb = torch.tensor(Zh[i:i + batch], dtype=torch.float32, device=dev) if host else Zd[i:i + batch]
            s = ((xb - vae(xb)) ** 2).mean(dim=(1, 2))
            if host:
                score[i:i + batch] = torch.from_numpy(s.cpu().numpy()).to(dev)
            else:
                score[i:i + batch] = s
        idx = torch.nonzero(score > thr).flatten()
        y_pred = torch.zeros((ns,), dtype=torch.int64, device=dev)
        idx_h = idx.cpu().numpy() if host else None
        for j in range(0, idx.numel(), batch):
            sel = idx[j:j + batch]
            zb = torch.tensor(Zh[idx_h[j:j + batch]], dtype=torch.float32, device=dev) if host else Zd[sel]
            logits = cnn(torch.stack([zb, (zb - vae(zb)) ** 2], dim=1))
            cls = torch.argmax(logits, dim=1)
            y_pred[sel] = (torch.from_numpy(cls.cpu().numpy()).to(dev) if host else cls) + 1
        return int(idx.numel())

    for name, batch, host in (("host_windows_b512", 512, True), ("resident_b512", 512, False), ("resident_b8192", 8192, False)):
        try:
            run(batch, host)                                                    # warm-up (cuDNN plans, allocator)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            k = run(batch, host)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out[name] = {"value": ns / dt, "unit": "windows/s", "batch": batch, "flagged": k,
                         "data": "host windows, per-batch H2D/D2H as 06 does" if host else "normalised windows resident on the device"}
        except Exception as e:                                                  # a baseline must never break the bench line
            out[name] = {"unavailable": repr(e)[:200]}
    out["value"] = out.get("host_windows_b512", {}).get("value")
    out["resident"] = out.get("resident_b512")
    return out


def incumbent_train(ctx, B=256, T=100, steps=10):
    """cuDNN LSTM forward + backward + torch.optim.Adam at batch 256: the reference's training step (03_train_vae.py:260-271)
    with its own TemporalVAE on the GPU, batch resident on the device."""
    import torch.nn.functional as F
    from oracle import ref_driver as R
    from shmfast import synth
    dev = ctx.dev
    if R.ref_root() is None:
        return {"unavailable": "reference files not staged"}
    try:
        V, _ = R.reference_models("4dof")
        vae = V(input_dim=12, latent_dim=16, hidden_dim=128, num_layers=2, dropout=0.3)
        vae.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.stage_vae_weights("4dof", seed=0).items()})
        vae = vae.to(dev).train()
        opt = torch.optim.Adam(vae.parameters(), lr=1e-3, weight_decay=1e-5)
        xb = torch.from_numpy(synth.windows(B, T, 12, seed=5)).to(dev)

        def step():
            opt.zero_grad()
            xhat, mu, logvar = vae(xb)
            loss = F.mse_loss(xhat, xb, reduction="mean") + 0.5 * (-0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp()))
            loss.backward()
            torch.nn.utils.clip_grad_norm_(vae.parameters(), 2.0)
            opt.step()

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": B / (ms / 1e3), "unit": "windows/s", "ms_per_step": ms, "batch": B,
                "kind": "reference TemporalVAE on torch CUDA kernels (cuDNN LSTM fwd+bwd, torch.optim.Adam), fp32, batch resident"}
    except Exception as e:
        return {"unavailable": repr(e)[:200]}


# ----------------------------------------------------------------------------------------------
# 4DOF hybrid / score (the headline)
# ----------------------------------------------------------------------------------------------
def measure_4dof(ctx, a, N, pct, steps, warmup, workload="4dof_hybrid", e2e=True, baselines=True, sample_clocks=True, full=True):
    from shmfast import ops, synth
    from shmfast.pipeline import Hybrid4dof
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    hybrid = workload == "4dof_hybrid"
    s, series_h, mean, std = synth_problem(N, seed=100 + rank)            # each rank scores its own window range
    T, D, Z = s["T"], s["D"], s["Z"]
    vae = ops.VaeScorer(synth.stage_vae_weights("4dof", seed=0), dev, engine=ENGINES[a.engine])
    cnn = ops.Cnn4dof(synth.cnn4dof_weights(seed=0), dev)
    # threshold calibration, as 04_vae_thresholding.py: P99 of a 2,010-window normal set (device percentile)
    series_pinned = torch.from_numpy(series_h).pin_memory()
    series_d = series_pinned.to(dev, non_blocking=True)
    src = ops.WindowSource(series_d, T, stride=1, mean=mean, std=std, nan_to_zero=True)
    torch.manual_seed(42 + rank)
    cal_idx = torch.linspace(0, N - 1, 2010, device=dev).to(torch.int32)          # 2,010 windows spread over the stream
    cal_scores = vae.score(src, torch.randn((2010, Z), device=dev), idx=cal_idx)["score"]
    thr = float(ops.percentile(cal_scores, pct).item()) if hybrid else float("inf")
    hyb = Hybrid4dof(vae, cnn, thr)
    eps1 = torch.randn((N, Z), device=dev)
    max_flag = min(N, max(4096, int(min(1.0, 5.0 * (100.0 - pct) / 100.0) * N)))
    eps2 = torch.randn((max_flag, Z), device=dev)
    kern_ev = []
    bufs = {"first": {}, "second": {}}

    def step(i, timed=True):
        """Device-resident step; the first (dominant) kernel is bracketed by its own CUDA events."""
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        out = vae.score(src, eps1, n=N, want_latent=hybrid, out=bufs["first"])
        if timed:
            e1.record()
            kern_ev.append((e0, e1))
        flag, idx, count = ops.compact(out["score"], thr)
        launches = 4                                                       # scorer + the three compaction kernels
        if hybrid:
            # second pass with fresh noise; the (deterministic) encoder's mu / logvar of the first pass are reused
            second = vae.rescore(src, out["mu"], out["logvar"], eps2, idx=idx, n=max_flag, n_dev=count, want_cnn_in=True, out=bufs["second"])
            if second is None:
                second = vae.score(src, eps2, n=max_flag, idx=idx, n_dev=count, want_score=False, want_cnn_in=True, out=bufs["second"])
            logits, label, p_struct = cnn.forward(second["cnn_in"], n=max_flag, n_dev=count, want_labels=True)
            ops.scatter_flagged_4dof(idx, count, max_flag, label, p_struct, N)
            launches += 4                                                  # rescore, conv, fc, scatter
        step.launches, step.count = launches, count

    for i in range(max(1, warmup)):
        step(i, False)
    torch.cuda.synchronize()
    n_flag = int(step.count.item())
    if n_flag > max_flag:                       # size the flagged-subset buffers to the workload and warm up again
        max_flag = min(N, int(1.25 * n_flag))
        eps2 = torch.randn((max_flag, Z), device=dev)
        bufs["second"] = {}
        step(0, False)
        torch.cuda.synchronize()
    total_ms, clocks = ctx.timed(step, steps, sample_clocks)
    kern_ms = sum(x.elapsed_time(y) for x, y in kern_ev) / len(kern_ev)
    windows_total = ctx.sum(float(N * steps))
    value = windows_total / (total_ms / 1e3)
    launches_per_step = step.launches

    # ---- end to end through the public API with host buffers: ONE C call per chunk (shm_hybrid4dof_score) ----
    e2e_d = None
    if e2e:
        from shmfast.stream import HostStream
        outs = {"score": ((N,), torch.float32)}
        if hybrid:
            outs.update(y_pred=((N,), torch.int64), p_full=((N,), torch.float32), status=((2,), torch.int32))
        pipe = HostStream(dev, series_pinned.shape, outs)
        res_bufs = [{}, {}, {}]

        def e2e_chunk(sd_, i):                                             # device work of one chunk, current stream, no host sync
            src_ = ops.WindowSource(sd_, T, stride=1, mean=mean, std=std, nan_to_zero=True)
            e1_ = torch.randn((N, Z), device=dev)                          # the reference draws eps on the device too
            if not hybrid:
                return dict(score=vae.score(src_, e1_, n=N)["score"])
            e2_ = torch.randn((max_flag, Z), device=dev)
            res = hyb.run_dense(src_, e1_, e2_, n=N, max_flagged=max_flag, out=res_bufs[i % 3])
            return dict(score=res["score"], y_pred=res["y_pred"], p_full=res["p_full"], status=res["status"])

        def e2e_run(k):                                                    # k chunks: H2D(i+1) and D2H(i-1) run under chunk i's kernels
            got, overflow = 0, 0
            for _, host in pipe.run((series_pinned for _ in range(k)), e2e_chunk):
                got += int(host["score"].shape[0])
                if hybrid:
                    overflow += int(host["status"][1])
            return got, overflow

        e2e_run(2)
        ctx.barrier()
        t0 = time.perf_counter()
        got, overflow = e2e_run(steps)
        torch.cuda.synchronize()
        assert got == N * steps and overflow == 0
        e2e_s = ctx.max(time.perf_counter() - t0)
        d2h = N * 4 + (N * 8 + N * 4 + 8 if hybrid else 0)
        e2e_d = {"value": windows_total / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": int(series_pinned.numel() * 4),
                 "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / steps,
                 "api": "shmfast.stream.HostStream around shmfast.pipeline.Hybrid4dof.run_dense = ONE C call shm_hybrid4dof_score per chunk "
                        "(pinned host series in; scores, dense y_pred / p_struct and the flagged count out; copies on side streams under the "
                        "next chunk's kernels; no ATen kernel on the chunk path except the eps draw)"}
        del pipe, res_bufs

    cpu_baseline = None
    torch_cuda = None
    if baselines and rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_4dof(min(a.cpu_sample, N), 1, 1, pct, workload)
        cpu_baseline = {"value": r["value"], "unit": "windows/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        try:
            torch_cuda = incumbents_4dof(ctx, series_h, mean, std, thr)
        except Exception as e:                                   # a baseline must never break the bench line
            torch_cuda = {"unavailable": repr(e)[:200]}

    eng_name = {ops.ENGINE_FP32: "fp32", ops.ENGINE_TC_BF16X3: "tc_bf16x3"}[vae.engine]
    achieved = FLOP_PER_WINDOW["4dof"] * N / (kern_ms / 1e3) / 1e12
    pk = ctx.pk
    line = {
        "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if vae.engine == ops.ENGINE_FP32 else "bf16x3->f32",
        "data": "synthetic",
        "config": {"workload": workload, "windows_per_gpu": N, "T": T, "D": D, "H": s["H"], "Z": Z, "L": s["L"],
                   "engine": eng_name, "threshold": f"P{pct:g} of 2010 calibration windows" if hybrid else None,
                   "flagged_per_gpu": n_flag, "input": "raw series, stride 1, gather+normalise fused into the scorer",
                   "l2": "flushed between timed steps (512 MiB memset)", "parallelism": f"window-range shards x{world}, no collective"},
        "clocks": clocks, "e2e": e2e_d, "gpu_launches": launches_per_step * steps,
        "roofline": {"bound": "tensor", "kernel": "vae_score (fused LSTM-VAE scorer, first pass)", "achieved": achieved,
                     "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"],
                     "traffic": NCU_DRAM_BYTES_PER_WINDOW_4DOF * N if vae.engine == ops.ENGINE_TC_BF16X3 else None,
                     "traffic_note": f"dram__bytes_read.sum + dram__bytes_write.sum of {NCU_SOURCE_4DOF} scaled per window "
                                     "(the layer-0 h_t stream of the 2-layer stacks, not input re-reads: the input is 48 B/window)",
                     "peak_source": pk["source"] + " bf16 dense, sustained", "kernel_ms": kern_ms,
                     "algorithmic_flop_per_window": FLOP_PER_WINDOW["4dof"], "engine": eng_name},
        "cpu_baseline": cpu_baseline, "torch_cuda_baseline": torch_cuda,
    }
    if not full:
        line = {k: line[k] for k in ("value", "unit", "ms_per_step", "steps", "config", "e2e", "roofline")}
    vae.close(); cnn.close()
    return line


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


def cpu_reference_openlab(a, pct, warmup, steps):
    """The reference's openLAB wiring (10_test_hybrid_pipeline.py:233-302,351-367) on torch.nn CPU kernels, all host threads."""
    from oracle import torch_port as TP
    from shmfast import synth
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
    return dict(value=n / (ms / 1e3), ms_per_step=ms, cores=cores, kind="port", sample=sample, windows=n)


def measure_openlab(ctx, a, N, pct, steps, warmup, e2e=True, cpu=True, sample_clocks=True, full=True):
    from shmfast import ops, synth
    from shmfast.pipeline import HybridOpenLab
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    series_h, vmu, vsd, cmu, csd = openlab_problem(N, seed=100 + rank)
    vae = ops.VaeScorer(synth.stage_vae_weights("openlab", seed=0), dev, engine=ENGINES[a.engine])
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
    max_flag = min(N, max(1024, int((100.0 - pct) / 100.0 * 1.5 * N)))
    kern_ev = []
    sbuf = {}

    def step(i, timed=True):
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        score = vae.score(src_g, eps, n=N, out=sbuf)["score"]
        if timed:
            e1.record()
            kern_ev.append((e0, e1))
        flag, idx, count = ops.compact(score, thr)
        logits, prob = cnn.forward(src_r, n=step.max_flag, idx=idx, n_dev=count, want_prob=True)
        ops.scatter_flagged_openlab(idx, count, step.max_flag, prob, 0.5, N)
        step.count = count

    step.max_flag = max_flag
    for i in range(max(1, warmup)):
        step(i, False)
    torch.cuda.synchronize()
    n_flag = int(step.count.item())
    if n_flag > max_flag:
        max_flag = step.max_flag = min(N, int(1.25 * n_flag))
        step(0, False); torch.cuda.synchronize()
    total_ms, clocks = ctx.timed(step, steps, sample_clocks)
    kern_ms = sum(x.elapsed_time(y) for x, y in kern_ev) / len(kern_ev)
    windows_total = ctx.sum(float(N * steps))
    value = windows_total / (total_ms / 1e3)

    e2e_d = None
    if e2e:
        from shmfast.stream import HostStream
        pipe = HostStream(dev, pinned.shape, {"score": ((N,), torch.float32), "flag": ((N,), torch.uint8), "y_pred": ((N,), torch.int64),
                                              "prob_full": ((N,), torch.float64), "status": ((2,), torch.int32)})
        res_bufs = [{}, {}, {}]

        def e2e_chunk(sd_, i):                                                 # device work of one chunk, current stream, no host sync
            g, r = sources(sd_)
            res = hyb.run_dense(g, r, torch.randn((N, 8), device=dev), n=N, max_flagged=max_flag, out=res_bufs[i % 3])
            return {k: res[k] for k in ("score", "flag", "y_pred", "prob_full", "status")}

        def e2e_run(k):                                                        # H2D(i+1) and D2H(i-1) run under chunk i's kernels
            got, overflow = 0, 0
            for _, host in pipe.run((pinned for _ in range(k)), e2e_chunk):
                got += int(host["score"].shape[0])
                overflow += int(host["status"][1])
            return got, overflow

        e2e_run(2)
        ctx.barrier()
        t0 = time.perf_counter()
        got, overflow = e2e_run(steps)
        torch.cuda.synchronize()
        assert got == N * steps and overflow == 0
        e2e_s = ctx.max(time.perf_counter() - t0)
        e2e_d = {"value": windows_total / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": int(pinned.numel() * 4),
                 "d2h_bytes_per_step": int(N * (4 + 1 + 8 + 8) + 8), "ms_per_step": 1e3 * e2e_s / steps,
                 "api": "shmfast.stream.HostStream around shmfast.pipeline.HybridOpenLab.run_dense = ONE C call shm_hybridol_score per chunk "
                        "(copies on side streams under the next chunk's kernels)"}
        del pipe, res_bufs
    cpu_baseline = None
    if cpu and rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_openlab(a, pct, 1, 1)
        cpu_baseline = {"value": r["value"], "unit": "windows/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    pk = ctx.pk
    eng = {ops.ENGINE_FP32: "fp32", ops.ENGINE_TC_BF16X3: "tc_bf16x3"}[vae.engine]
    achieved = FLOP_PER_WINDOW["openlab"] * N / (kern_ms / 1e3) / 1e12
    n_cnn_chunks = (n_flag + 8191) // 8192
    line = {
        "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if vae.engine == ops.ENGINE_FP32 else "bf16x3->f32", "data": "synthetic",
        "config": {"workload": "openlab_hybrid", "windows_per_gpu": N, "T": 200, "stride": 20, "D_gate": 3, "D_raw": 4, "H": 64, "Z": 8, "L": 1,
                   "engine": eng, "cnn_engine": {ops.ENGINE_FP32: "fp32", ops.ENGINE_TC_BF16X3: "tc_f16x3 (tcgen05 implicit GEMM)"}[cnn.engine],
                   "gate_threshold": f"P{pct:g} of 2000 calibration windows", "flagged_per_gpu": n_flag, "cnn_threshold": 0.5,
                   "input": "raw 4-channel series with NaN runs, stride 20; gather+standardise fused into scorer and CNN",
                   "l2": "flushed between timed steps (512 MiB memset)", "parallelism": f"window-range shards x{world}, no collective"},
        "clocks": clocks, "e2e": e2e_d,
        "gpu_launches": (5 + CNNOL_LAUNCHES_PER_CHUNK * n_cnn_chunks) * steps,   # scorer, 3 compaction kernels, scatter, CNN kernels per 8192-window chunk
        "roofline": {"bound": "tensor", "kernel": "vae_score (fused LSTM-VAE scorer)", "achieved": achieved, "peak": pk["tf_sust"],
                     "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"],
                     "traffic": NCU_DRAM_BYTES_PER_WINDOW_OPENLAB * N if vae.engine == ops.ENGINE_TC_BF16X3 else None,
                     "traffic_note": f"dram__bytes_read.sum + dram__bytes_write.sum of {NCU_SOURCE_OPENLAB} scaled to this launch",
                     "kernel_ms": kern_ms,
                     "peak_source": pk["source"] + " bf16 dense, sustained", "algorithmic_flop_per_window": FLOP_PER_WINDOW["openlab"],
                     "engine": eng},
        "cpu_baseline": cpu_baseline}
    if not full:
        line = {k: line[k] for k in ("value", "unit", "ms_per_step", "steps", "config", "e2e", "roofline")}
        line["scorer_ms"] = kern_ms
        line["rest_ms"] = total_ms / steps - kern_ms
    vae.close(); cnn.close()
    return line


CNNOL_LAUNCHES_PER_CHUNK = 12


# ----------------------------------------------------------------------------------------------
# 4DOF training step (BASELINE.json configs[4]): forward (train mode, dropout 0.3) -> ELBO -> BPTT -> one gradient
# all-reduce -> clip 2.0 + Adam, batch 256 per GPU (03_train_vae.py:53,260-271).  Weak scaling: global batch 256*N.
# ----------------------------------------------------------------------------------------------
def cpu_reference_train(B, steps, warmup):
    from oracle import torch_port as TP
    from shmfast import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    port = TP.VaeTrainPort(synth.stage_vae_weights("4dof", seed=0)).train()
    opt = torch.optim.Adam(port.ordered_parameters(), lr=1e-3, weight_decay=1e-5)
    X = torch.from_numpy(synth.windows(B, 100, 12, seed=123))
    rng = np.random.Generator(np.random.PCG64(9))
    times = []
    for i in range(warmup + steps):
        masks = [torch.from_numpy((rng.random((1, B, 100, 128)) >= 0.3).astype(np.uint8)) for _ in range(2)]
        t0 = time.perf_counter()
        TP.train_step_port(port, opt, X, torch.randn(B, 16), 0.5, masks[0], masks[1], 0.3)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return dict(value=B / (ms / 1e3), ms_per_step=ms, cores=cores, kind="port", windows=B,
                sample=f"{steps} optimisation steps of batch {B} after {warmup} warm-up (torch autograd over nn.LSTM CPU kernels), {ms:.0f} ms/step")


def measure_train(ctx, a, B, steps, warmup, e2e=True, cpu=True, sample_clocks=True, full=True):
    from shmfast import synth, train
    from shmfast.models import fourdof
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    T, D, Z = 100, 12, 16
    vae = fourdof.TemporalVAE(12, 16, 128, 2, dropout=0.3)
    vae.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.stage_vae_weights("4dof", seed=0).items()})
    vae = vae.to(dev).train()
    tr = train.VaeTrainer(vae, T, B)
    n_batches = 8
    host = [torch.from_numpy(synth.windows(B, T, D, seed=1000 * rank + i)).pin_memory() for i in range(n_batches)]
    devb = [h.to(dev) for h in host]
    torch.manual_seed(42 + rank)
    for i in range(max(1, warmup)):
        tr.step(devb[i % n_batches], 0.5)
    torch.cuda.synchronize()
    total_ms, clocks = ctx.timed(lambda i: tr.step(devb[i % n_batches], 0.5), steps, sample_clocks)
    windows_total = ctx.sum(float(B * steps))
    value = windows_total / (total_ms / 1e3)
    # phase breakdown on every rank (the all-reduce is a collective), reported by rank 0
    xb = devb[0]
    eps = torch.randn((B, Z), device=dev)
    masks = train.draw_dropout_masks(2, B, T, 128, 0.3, dev)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ctx.barrier()
    ctx.flush.zero_()
    marks[0].record()
    xhat, mu, lv = tr.handle.forward(tr.flat, xb, eps, masks[0], masks[1], 0.3)
    marks[1].record()
    l3, dx, dm, dl = train.elbo_grad(xb, xhat, mu, lv, 0.5)
    marks[2].record()
    tr.handle.backward(tr.flat, dx, dm, dl, tr.grads)
    marks[3].record()
    scale = train.reduce_gradients(tr.grads, tr.group)
    marks[4].record()
    train.adam_clip_step(tr.flat, tr.grads, tr.exp_avg, tr.exp_avg_sq, tr.steps + 1, 1e-3, weight_decay=1e-5, max_norm=2.0, grad_scale=scale)
    marks[5].record()
    torch.cuda.synchronize()
    phases = {n: marks[i].elapsed_time(marks[i + 1]) for i, n in enumerate(("forward_ms", "elbo_ms", "backward_ms", "allreduce_ms", "clip_adam_ms"))}
    # the collective alone, back to back (NCCL over NVLink; 1.9 MB fp32)
    ar_ms = None
    if world > 1:
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            train.reduce_gradients(tr.grads, tr.group)
        e1.record()
        torch.cuda.synchronize()
        ar_ms = ctx.max(e0.elapsed_time(e1) / 20)
    e2e_d = None
    if e2e:
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            xbh = host[i % n_batches].to(dev, non_blocking=True)
            loss_host = tr.step(xbh, 0.5).cpu()
        torch.cuda.synchronize()
        e2e_s = ctx.max(time.perf_counter() - t0)
        e2e_d = {"value": windows_total / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": B * T * D * 4, "d2h_bytes_per_step": 12,
                 "ms_per_step": 1e3 * e2e_s / steps, "api": "shmfast.train.VaeTrainer.step (pinned host batch in, loss out)"}
    cpu_baseline = None
    if cpu and rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_train(B, 3, 1)
        cpu_baseline = {"value": r["value"], "unit": "windows/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    ms = total_ms / steps
    achieved = TRAIN_FLOP_PER_WINDOW * B / (ms / 1e3) / 1e12
    line = {
        "metric": "training windows/sec (4DOF LSTM-VAE step)", "value": value, "unit": "windows/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "4dof_train", "batch_per_gpu": B, "global_batch": B * world, "T": T, "D": D, "H": 128, "Z": Z, "L": 2,
                   "dropout": 0.3, "optimizer": "Adam lr 1e-3 wd 1e-5, clip 2.0", "l2": "flushed between timed steps (512 MiB memset)",
                   "parallelism": f"dp{world}, one NCCL all-reduce of the 477,100-float gradient per step"},
        "clocks": clocks, "phases": phases, "allreduce_ms": ar_ms, "e2e": e2e_d,
        "gpu_launches": TRAIN_LAUNCHES_PER_STEP * steps,
        "roofline": {"bound": "fp32", "kernel": "whole step (fp32 FMA contractions + resident-weight recurrence)", "achieved": achieved,
                     "peak": FP32_PEAK_TF, "unit": "TFLOP/s", "frac": achieved / FP32_PEAK_TF, "traffic": None,
                     "peak_source": "derived: 148 SM x 128 FMA lanes x 2 x 1.965 GHz", "algorithmic_flop_per_window": TRAIN_FLOP_PER_WINDOW},
        "cpu_baseline": cpu_baseline}
    if not full:
        line = {k: line[k] for k in ("value", "unit", "ms_per_step", "steps", "config", "phases", "allreduce_ms", "e2e")}
    tr.close()
    return line


# ----------------------------------------------------------------------------------------------
# CNN training steps: 4DOF/Scripts/05_train_cnn.py:266-281 (batch 100, CE, Adam) and
# openLAB Codes/06_train_cnn.py:410-421 (batch 128, weighted focal loss, clip 2.0, AdamW); fp32 FMA implicit-GEMM convolutions.
# ----------------------------------------------------------------------------------------------
CNN_TRAIN_FLOP = {"4dof": 3 * 4_070_912, "openlab": 3 * 133_851_648}       # forward + ~2x backward per window


def measure_cnn_train(ctx, arch: str, steps: int, warmup: int, incumbent: bool = True):
    from shmfast import cnn_train as CT, synth
    from shmfast.models import fourdof, openlab
    dev = ctx.dev
    B = 100 if arch == "4dof" else 128
    if arch == "4dof":
        model = fourdof.CNN(2, 2, 0.5)
        sd = synth.cnn4dof_weights(seed=0)
        xs = [np.stack([synth.windows(B, 100, 12, seed=10 + i), synth.windows(B, 100, 12, seed=20 + i) ** 2], axis=1).astype(np.float32) for i in range(4)]
        alpha = None
    else:
        model = openlab.CNN(dropout_rate=0.4)
        sd = synth.cnnol_weights(seed=0)
        xs = [np.clip(2.0 * synth.windows(B, 200, 4, seed=10 + i), -10, 10).astype(np.float32)[:, None] for i in range(4)]
        alpha = torch.tensor([0.8, 1.2])
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    model = model.to(dev).train()
    tr = CT.CnnTrainer(model, B, alpha=alpha)
    xd = [torch.from_numpy(x).to(dev) for x in xs]
    yd = [torch.randint(0, 2, (B,), device=dev) for _ in xs]
    for i in range(max(1, warmup)):
        tr.step(xd[i % 4], yd[i % 4])
    torch.cuda.synchronize()
    total_ms, _ = ctx.timed(lambda i: tr.step(xd[i % 4], yd[i % 4]), steps, False)
    ms = total_ms / steps
    achieved = CNN_TRAIN_FLOP[arch] * B / (ms / 1e3) / 1e12
    out = {"value": ctx.sum(float(B * steps)) / (total_ms / 1e3), "unit": "windows/s", "ms_per_step": ms, "steps": steps, "batch_per_gpu": B,
           "config": {"workload": f"{arch}_cnn_train", "loss": "CrossEntropy" if arch == "4dof" else "weighted focal (gamma 2)",
                      "optimizer": "Adam lr 1e-4 wd 5e-5" if arch == "4dof" else "clip 2.0 + AdamW lr 3e-4 wd 1e-4",
                      "parallelism": f"dp{ctx.world}, one NCCL all-reduce of the flat gradient per step"},
           "roofline": {"bound": "fp32", "achieved": achieved, "peak": FP32_PEAK_TF, "unit": "TFLOP/s", "frac": achieved / FP32_PEAK_TF,
                        "algorithmic_flop_per_window": CNN_TRAIN_FLOP[arch]}}
    tr.close()
    if incumbent and ctx.rank == 0 and ctx.world == 1:
        try:
            from oracle import ref_driver as R
            from oracle import torch_port as TP
            if R.ref_root() is not None:
                _, Cn = R.reference_models("4dof" if arch == "4dof" else "openlab")
                ref = Cn(input_channels=2, num_classes=2, dropout_rate=0.5) if arch == "4dof" else Cn(dropout_rate=0.4)
                ref.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
                ref = ref.to(dev).train()
                if arch == "4dof":
                    opt = torch.optim.Adam(ref.parameters(), lr=1e-4, weight_decay=5e-5)
                    lossf = torch.nn.CrossEntropyLoss()
                else:
                    opt = torch.optim.AdamW(ref.parameters(), lr=3e-4, weight_decay=1e-4)
                    al = alpha.to(dev)
                    lossf = lambda lg, y: TP.focal_loss(lg, y, al, 2.0)

                def ref_step(i):
                    opt.zero_grad()
                    loss = lossf(ref(xd[i % 4]), yd[i % 4])
                    loss.backward()
                    if arch != "4dof":
                        torch.nn.utils.clip_grad_norm_(ref.parameters(), 2.0)
                    opt.step()

                for i in range(3):
                    ref_step(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(10):
                    ref_step(i)
                e1.record()
                torch.cuda.synchronize()
                rms = e0.elapsed_time(e1) / 10
                out["torch_cuda_baseline"] = {"value": B / (rms / 1e3), "unit": "windows/s", "ms_per_step": rms,
                                              "kind": "reference CNN class on torch CUDA kernels (cuDNN conv fwd+bwd, torch optimiser), batch resident"}
        except Exception as e:
            out["torch_cuda_baseline"] = {"unavailable": repr(e)[:200]}
    return out


TRAIN_LAUNCHES_PER_STEP = 47        # profiles/r01_train_launches_v2.csv: 141 libshmfast launches in 3 steps


# ----------------------------------------------------------------------------------------------
# 1_DOF (BASELINE.json configs[0], the reference's CPU-runnable case): standardise + window (T=80, stride 1) -> LSTM-VAE
# (H=32, L=2, no LayerNorm; fp32 engine) -> reconstruction -> overlap-average stitch, de-standardise, RMSE per 100-sample
# segment (1_DOF/Scripts/04_test_seen_variants.py:281-311, datasets.py:17-71).
# ----------------------------------------------------------------------------------------------
def onedof_problem(N):
    from shmfast import synth
    T, D = 80, 12
    series_h = synth.series(N + T - 1, D, seed=5)
    mean = series_h.mean(axis=0).astype(np.float64)
    std = series_h.std(axis=0).astype(np.float64) + 1e-8
    return series_h, mean, std


def cpu_reference_onedof(a, N, steps, warmup):
    from oracle import np_oracle as O, torch_port as TP
    from shmfast import synth
    T, Z = 80, 5
    series_h, mean, std = onedof_problem(N)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = min(a.cpu_sample, N)
    vae = TP.VaePort(synth.stage_vae_weights("1dof", seed=0))
    ser = series_h[: n + T - 1]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        xn = ((ser - mean) / std).astype(np.float32)
        W = np.stack([xn[j:j + T] for j in range(n)])
        with torch.no_grad():
            rec = vae(torch.from_numpy(W), torch.randn(n, Z))[0].numpy()
        y = O.destandardize(O.stitch_windows(rec, ser.shape[0], 1), mean, std)
        O.segment_rmse(ser, y, 100)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return dict(value=n / (ms / 1e3), ms_per_step=ms, cores=cores, kind="port", windows=n,
                sample=f"{n} windows per pass: standardise + window + VAE (torch.nn) + stitch + segment RMSE (NumPy)")


def measure_onedof(ctx, a, steps, warmup):
    from shmfast import ops, synth
    dev, world = ctx.dev, ctx.world
    T, D, Z = 80, 12, 5
    N = min(a.windows, 1 << 18)                               # the reconstruction [N,80,12] is materialised for the stitch
    rows = N + T - 1
    series_h, mean, std = onedof_problem(N)
    vae = ops.VaeScorer(synth.stage_vae_weights("1dof", seed=0), dev)
    pinned = torch.from_numpy(series_h).pin_memory()
    series_d = pinned.to(dev)
    eps = torch.randn((N, Z), device=dev)
    buf = {}

    def step(sd_):
        src = ops.WindowSource(sd_, T, stride=1, mean=mean.astype(np.float32), std=std.astype(np.float32))
        out = vae.score(src, eps, n=N, want_recon=True, out=buf)
        return ops.stitch_segment_rmse(out["recon"], rows, 1, mean, std, sd_, 100, want_series=False)[1], out["score"]

    for _ in range(max(1, warmup)):
        step(series_d)
    torch.cuda.synchronize()
    total_ms, clocks = ctx.timed(lambda i: step(series_d), steps, True)
    windows_total = ctx.sum(float(N * steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        rm, sc = step(pinned.to(dev, non_blocking=True))
        rm_h, sc_h = rm.cpu(), sc.cpu()
    torch.cuda.synchronize()
    e2e_s = ctx.max(time.perf_counter() - t0)
    ms = total_ms / steps
    achieved = FLOP_PER_WINDOW["1dof"] * N / (ms / 1e3) / 1e12
    return {
        "metric": METRIC, "value": windows_total / (total_ms / 1e3), "unit": "windows/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "1dof_score", "windows_per_gpu": N, "T": T, "D": D, "H": 32, "Z": Z, "L": 2, "engine": "fp32",
                   "post": "overlap-average stitch + de-standardise + RMSE per 100-sample segment (fp64)",
                   "l2": "flushed between timed steps (512 MiB memset)", "parallelism": f"window-range shards x{world}, no collective"},
        "clocks": clocks,
        "e2e": {"value": windows_total / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": int(pinned.numel() * 4),
                "d2h_bytes_per_step": int(N * 4 + rm_h.numel() * 8), "ms_per_step": 1e3 * e2e_s / steps,
                "api": "ops.VaeScorer.score(want_recon) + ops.stitch_segment_rmse"},
        "gpu_launches": 2 * steps,
        "roofline": {"bound": "fp32", "kernel": "vae_score_fp32_kernel<32> (whole step)", "achieved": achieved, "peak": FP32_PEAK_TF,
                     "unit": "TFLOP/s", "frac": achieved / FP32_PEAK_TF, "traffic": None,
                     "peak_source": "derived: 148 SM x 128 FMA lanes x 2 x 1.965 GHz", "algorithmic_flop_per_window": FLOP_PER_WINDOW["1dof"]},
        "cpu_baseline": None}


# ----------------------------------------------------------------------------------------------
# secondary: HBM-bound kernels vs the measured copy bandwidth, and the one-series N-way shard check
# ----------------------------------------------------------------------------------------------
def measure_membound(ctx):
    """Algorithmic bytes / CUDA-event time against MEASURED_PEAKS.json hbm_gbs; L2 flushed between repetitions; inputs > L2."""
    from shmfast import ops, synth
    from shmfast.pipeline import guard_std_4dof
    dev, peak = ctx.dev, ctx.pk["hbm_gbs"]

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            ctx.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    out = {}
    N, T, D = 1 << 19, 100, 12
    series = torch.from_numpy(synth.series(N + T - 1, D, seed=1)).to(dev)
    mean, std = synth.stats(D, seed=0)
    src = ops.WindowSource(series, T, stride=1, mean=mean, std=guard_std_4dof(std), nan_to_zero=True)
    buf = torch.empty((N, T, D), dtype=torch.float32, device=dev)
    ms = timed(lambda: ops.window_normalize(src, out=buf))
    b = N * T * D * 4 + series.numel() * 4
    out["gather_4dof_stride1"] = {"gbs": b / ms / 1e6, "frac": b / ms / 1e6 / peak, "ms": ms, "bytes": b}
    del buf, series
    N2 = 1 << 20
    ser2 = torch.from_numpy(synth.series((N2 - 1) * 20 + 200, 4, seed=2, nan_frac=0.0007)).to(dev)
    mu2, sd2 = synth.stats(4, seed=2)
    src2 = ops.WindowSource(ser2, 200, stride=20, mean=mu2, std=sd2, clip=10.0, nan_to_zero=True)
    buf2 = torch.empty((N2, 200, 4), dtype=torch.float32, device=dev)
    ms = timed(lambda: ops.window_normalize(src2, out=buf2))
    b = N2 * 200 * 4 * 4 + ser2.numel() * 4
    out["gather_openlab_stride20"] = {"gbs": b / ms / 1e6, "frac": b / ms / 1e6 / peak, "ms": ms, "bytes": b}
    del buf2, ser2
    M = 1 << 27
    score = torch.rand(M, device=dev)
    for frac, name in ((0.01, "compact_1pct"), (0.47, "compact_47pct")):
        ms = timed(lambda: ops.compact(score, 1.0 - frac))
        b = M * 4 + M * 1 + int(frac * M) * 4
        out[name] = {"gbs": b / ms / 1e6, "frac": b / ms / 1e6 / peak, "ms": ms, "bytes": b}
    ms = timed(lambda: ops.percentile(score, 99.0))
    out["percentile_p99"] = {"gbs": M * 4 / ms / 1e6, "frac": M * 4 / ms / 1e6 / peak, "ms": ms, "bytes": M * 4}
    out["peak_gbs"] = peak
    out["note"] = "algorithmic bytes (unique bytes read + bytes written) / CUDA-event time; 2^27 scores, 2^19 / 2^20 windows"
    return out


def shard_check(ctx, per_rank: int = 1 << 16):
    """ONE 4DOF series of world x 2^16 windows (2 virtual shards when world == 1), split with shard.shard_range /
    series_rows_for (T-1 row halo), each shard scored on its own rank from its own row slice, results gathered with
    gather_by_rank / gather_flagged, and compared BIT FOR BIT on rank 0 with a single-GPU run over the whole series.
    eps follows the reference's order: eps1 by window, eps2 by global flagged ordinal (the flagged-count prefix of the lower
    ranks is the only cross-rank dependency: one 4-byte all_gather)."""
    from shmfast import ops, shard, synth
    from shmfast.pipeline import Hybrid4dof, guard_std_4dof
    dev, world, rank, dist = ctx.dev, ctx.world, ctx.rank, ctx.dist
    T, D, Z = 100, 12, 16
    parts = world if world > 1 else 2
    Ntot = parts * per_rank + 37                                           # ragged: the shards are not all equal
    series = synth.series(Ntot + T - 1, D, seed=77)                        # the same series on every rank (seeded)
    mean, std = synth.stats(D, seed=0)
    std = guard_std_4dof(std)
    g = torch.Generator().manual_seed(1234)
    eps1 = torch.randn((Ntot, Z), generator=g)
    eps2 = torch.randn((Ntot, Z), generator=g)
    vae = ops.VaeScorer(synth.stage_vae_weights("4dof", seed=0), dev)
    cnn = ops.Cnn4dof(synth.cnn4dof_weights(seed=0), dev)
    cal = vae.score(ops.WindowSource(torch.from_numpy(series[: 4096 + T - 1]).to(dev), T, stride=1, mean=mean, std=std, nan_to_zero=True),
                    eps1[:4096].to(dev))["score"]
    thr = float(ops.percentile(cal, 90.0).item())
    if world > 1:                                                          # every rank must use rank 0's threshold
        t = torch.tensor([thr], dtype=torch.float64, device=dev)
        dist.broadcast(t, src=0)
        thr = float(t.item())
    hyb = Hybrid4dof(vae, cnn, thr)

    def run_shard(lo, hi, flag_off=None):
        r0, r1 = shard.series_rows_for(lo, hi, T, 1)
        src = ops.WindowSource(torch.from_numpy(series[r0:r1]).to(dev), T, stride=1, mean=mean, std=std, nan_to_zero=True)
        assert src.n_windows == hi - lo
        first = vae.score(src, eps1[lo:hi].to(dev), want_latent=True)
        flag, idx, count = ops.compact(first["score"], thr)
        k = int(count.item())
        return src, first, flag, idx, k

    def finish(src, first, idx, k, off):
        if k == 0:
            return torch.zeros((0,), dtype=torch.int64, device=dev), torch.zeros((0,), device=dev)
        second = vae.rescore(src, first["mu"], first["logvar"], eps2[off:off + k].to(dev), idx=idx, n=k, want_cnn_in=True)
        logits, label, p = cnn.forward(second["cnn_in"], n=k, want_labels=True)
        return label, p

    shards = [shard.shard_range(Ntot, r, parts) for r in range(parts)]
    if world > 1:
        lo, hi = shards[rank]
        src, first, flag, idx, k = run_shard(lo, hi)
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([k], dtype=torch.int64, device=dev))
        off = int(sum(int(c.item()) for c in counts[:rank]))
        label, p = finish(src, first, idx, k, off)
        score_all = shard.gather_by_rank(first["score"])
        flag_all = shard.gather_by_rank(flag)
        idx_all = shard.gather_flagged(idx, k, lo)
        label_all = shard.gather_by_rank(label)
        p_all = shard.gather_by_rank(p)
    else:
        sc, fl, ix, lb, pp, off = [], [], [], [], [], 0
        for lo, hi in shards:
            src, first, flag, idx, k = run_shard(lo, hi)
            label, p = finish(src, first, idx, k, off)
            off += k
            sc.append(first["score"]); fl.append(flag); ix.append(idx[:k].to(torch.int64) + lo); lb.append(label); pp.append(p)
        score_all, flag_all, idx_all, label_all, p_all = (torch.cat(x) for x in (sc, fl, ix, lb, pp))
    res = None
    if rank == 0:
        src, first, flag, idx, k = run_shard(0, Ntot)
        label, p = finish(src, first, idx, k, 0)
        same = dict(score=torch.equal(score_all, first["score"]), flag=torch.equal(flag_all, flag),
                    idx=torch.equal(idx_all, idx[:k].to(torch.int64)), label=torch.equal(label_all, label), p_struct=torch.equal(p_all, p))
        res = {"shard_check": "bit-identical" if all(same.values()) else "MISMATCH", "fields": same, "windows": Ntot, "shards": parts,
               "flagged": k, "ranks": world,
               "how": "one series split with shard.shard_range + series_rows_for (T-1 row halo), scored per rank, gathered with "
                      "gather_by_rank / gather_flagged, compared with torch.equal against one GPU scoring the whole series"}
    vae.close(); cnn.close()
    return res


def secondary(ctx, a):
    """The other BASELINE configs in the same run.  Bounded: ~60 s at N=1."""
    t_start = time.perf_counter()
    sec = {}

    def leg(name, fn):
        t0 = time.perf_counter()
        try:
            sec[name] = fn()
        except Exception as e:                                   # never lose the headline to a secondary leg
            sec[name] = {"error": repr(e)[:300]}
            if ctx.world > 1:
                raise
        torch.cuda.empty_cache()
        if isinstance(sec[name], dict):
            sec[name]["leg_seconds"] = round(time.perf_counter() - t0, 2)

    N = 1 << 20
    leg("openlab_hybrid_p95", lambda: measure_openlab(ctx, a, N, 95.0, 40, 2, e2e=True, cpu=False, sample_clocks=False, full=False))
    leg("openlab_hybrid_38pct", lambda: measure_openlab(ctx, a, N, 62.0, 8, 1, e2e=False, cpu=False, sample_clocks=False, full=False))
    leg("4dof_hybrid_47pct", lambda: measure_4dof(ctx, a, 1 << 19, 53.0, 4, 1, e2e=False, baselines=False, sample_clocks=False, full=False))
    leg("4dof_train_b256", lambda: measure_train(ctx, a, 256, 40, 3, e2e=False, cpu=False, sample_clocks=False, full=False))
    if ctx.rank == 0 and ctx.world == 1:
        leg("4dof_train_torch_cuda_baseline", lambda: incumbent_train(ctx))
    leg("4dof_cnn_train_b100", lambda: measure_cnn_train(ctx, "4dof", 30, 3))
    leg("openlab_cnn_train_b128", lambda: measure_cnn_train(ctx, "openlab", 20, 3))
    leg("shard_check", lambda: shard_check(ctx))
    if ctx.rank == 0 and ctx.world == 1:
        leg("membound", lambda: measure_membound(ctx))
    sec["seconds"] = round(time.perf_counter() - t_start, 1)
    return sec


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if a.impl == "reference":
        if rank != 0:
            return
        if a.workload in ("4dof_hybrid", "4dof_score"):
            run_reference_4dof(a)
        elif a.workload == "openlab_hybrid":
            r = cpu_reference_openlab(a, 95.0 if a.flag_pct is None else a.flag_pct, a.warmup, a.steps)
            print(json.dumps(reference_line(a, r, METRIC, {"workload": "openlab_hybrid", "windows_per_step": r["windows"]})), flush=True)
        elif a.workload == "4dof_train":
            r = cpu_reference_train(a.batch, a.steps, a.warmup)
            print(json.dumps(reference_line(a, r, "training windows/sec (4DOF LSTM-VAE step)",
                                            {"workload": "4dof_train", "batch_per_step": a.batch, "T": 100, "D": 12, "H": 128, "Z": 16, "L": 2})), flush=True)
        else:
            r = cpu_reference_onedof(a, min(a.windows, 1 << 18), a.steps, a.warmup)
            print(json.dumps(reference_line(a, r, METRIC, {"workload": "1dof_score", "windows_per_step": r["windows"]})), flush=True)
        return
    ctx = Ctx()
    if ctx.world != a.gpus and ctx.rank == 0:
        print(f"[bench] note: --gpus {a.gpus} but WORLD_SIZE={ctx.world}; using {ctx.world}", file=sys.stderr)
    if a.workload in ("4dof_hybrid", "4dof_score"):
        pct = 99.0 if a.flag_pct is None else a.flag_pct
        line = measure_4dof(ctx, a, a.windows, pct, a.steps, a.warmup, workload=a.workload, e2e=not a.no_e2e)
        default = a.workload == "4dof_hybrid" and a.flag_pct is None and a.windows == (1 << 20) and a.engine == "auto"
        if default and not a.no_secondary:
            torch.cuda.empty_cache()
            line["secondary"] = secondary(ctx, a)
    elif a.workload == "openlab_hybrid":
        line = measure_openlab(ctx, a, a.windows, 95.0 if a.flag_pct is None else a.flag_pct, a.steps, a.warmup, e2e=not a.no_e2e)
    elif a.workload == "4dof_train":
        line = measure_train(ctx, a, a.batch, a.steps, a.warmup, e2e=not a.no_e2e)
    else:
        line = measure_onedof(ctx, a, a.steps, a.warmup)
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
