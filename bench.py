#!/usr/bin/env python
"""Benchmark of the MAFED distillation hot path (fused forward + backward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4|C2|C3|C1] [--impl ours|reference]

Metric (BASELINE.json): distill fwd+bwd tokens*layers/sec, and the fraction of HBM peak.
A "step" is one pass of the hot path (forward over all selected layers, epilogue, backward) over one
synthetic batch of hidden states.  Default workload: the VLPythia-1B shape the metric is quoted on
(C4: 16 layers -> 15 distilled, D=2048, T=256+32, bf16) at 64 samples per GPU, i.e. weak scaling to
C4's global batch 512 on 8 GPUs.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "distill fwd+bwd tokens*layers/sec"
UNIT = "tokens*layers/s"

WORKLOADS = {
    # name: (description, hidden-state tuple length, distilled layers (reference rule L-1), per-GPU B, txt, D, dtype)
    "C4": ("VLPythia-1B (16 layers, d=2048) all-layer MAFED distillation, bf16, 64 samples/GPU (C4 shard; 512 on 8 GPUs)",
           17, 15, 64, 32, 2048, "bf16"),
    "C2": ("VLPythia-base (12 layers, d=768) MAFED distillation, batch 128, bf16", 13, 11, 128, 32, 768, "bf16"),
    "C3": ("VLPythia-410M (24 layers, d=1024) MAFED distillation, batch 256, bf16", 25, 23, 256, 32, 1024, "bf16"),
    "C1": ("VLPythia-base MAFED distillation, batch 8, fp32", 13, 11, 8, 32, 768, "fp32"),
}
N_VIS = 256
RECIPE = dict(modality="balanced", layer_strategy="discounted", loss="mse", gamma=0.5)  # scripts/run_seed42.sh:74-93


def measured_traffic(wl):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(wl, {}).get("fused_kernel_dram_bytes")
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (pynvml, else nvidia-smi)."""

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None

    def _loop_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def _loop_smi(self):
        import subprocess
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True).stdout.strip().split(",")
            if len(out) >= 6:
                self.samples.append(int(float(out[0])))
                self.max_mhz = int(float(out[1]))
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)

    def __enter__(self):
        def run():
            try:
                self._loop_nvml()
            except Exception:
                try:
                    self._loop_smi()
                except Exception:
                    pass
        self._thread = threading.Thread(target=run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=5)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- workload
def torch_dtype(name):
    return {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[name]


def make_device_inputs(wl, rank, device):
    _, n_tuple, n_sel, B, txt, D, dt = WORKLOADS[wl]
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    T = N_VIS + txt
    st, te = [], []
    for _ in range(n_sel):
        s = torch.randn(B, T, D, generator=g, device=device, dtype=torch.float32)
        t = s + 0.1 * torch.randn(B, T, D, generator=g, device=device, dtype=torch.float32)
        st.append(s.to(torch_dtype(dt)))
        te.append(t.to(torch_dtype(dt)))
        del s, t
    # headline runs use an all-ones mask so that byte accounting is unambiguous
    am = torch.ones(B, txt, dtype=torch.int64, device=device)
    return st, te, am


class Opts:
    tasks = ["a", "b", "c"]
    batch_size = 64
    seed = 42
    pin_mem = False
    accumulate_grad_batches = 1


class Out:
    def __init__(self, hs):
        self.hidden_states = hs


def make_method(n_sel):
    from mafed_b200.methods import CLMethod
    return CLMethod["featdistill"](
        memory_size=8, opts=Opts(), model_type="vlpythia",
        distillation_modality_weighing_strategy=RECIPE["modality"],
        distillation_layer_weighing_strategy=RECIPE["layer_strategy"], distillation_coeff=1.0,
        distillation_layer=None, distillation_loss=RECIPE["loss"], gamma=RECIPE["gamma"], num_hidden_layers=n_sel)


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline(wl, sample_B=None, iters=3, warmup=1):
    """The reference's op chain (oracle port, torch CPU ops + autograd) on the host cores."""
    from oracle import distill_oracle as O
    _, n_tuple, n_sel, B, txt, D, dt = WORKLOADS[wl]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if sample_B is None:
        sample_B = max(1, min(B, int(16 * 2048 / D)))
    st, te, am = O.make_inputs(n_sel, sample_B, txt, D, n_vis=N_VIS, dtype=torch_dtype(dt), seed=1234,
                               teacher="close", mask="full")
    cfg = O.OracleConfig(modality_strategy=RECIPE["modality"], layer_strategy=RECIPE["layer_strategy"],
                         gamma=RECIPE["gamma"], num_hidden_layers=n_sel, distillation_layer=None, loss=RECIPE["loss"],
                         num_vision_tokens=N_VIS)
    times = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        O.forward_backward(st, te, am, cfg)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    units = sample_B * (N_VIS + txt) * n_sel
    return {"value": units / min(times), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{wl} shape, B={sample_B} of {B} samples, {n_sel} layers, {dt}, torch CPU ops + autograd "
                      f"(oracle port of distillation.py:105-166), best of {iters}",
            "ms_per_sample_step": 1e3 * min(times)}


def eager_gpu_baseline(wl, st, te, am, iters=5, warmup=2):
    """The reference's own op chain (oracle port: ATen kernels + autograd, distillation.py:105-166 under
    autocast) run on the SAME GPU on the same tensors -- what the unmodified reference would spend on this
    path on a B200.  A baseline leg like `cpu_baseline`; the product never calls it."""
    from oracle import distill_oracle as O
    _, n_tuple, n_sel, B, txt, D, dt = WORKLOADS[wl]
    cfg = O.OracleConfig(modality_strategy=RECIPE["modality"], layer_strategy=RECIPE["layer_strategy"],
                         gamma=RECIPE["gamma"], num_hidden_layers=n_sel, distillation_layer=None, loss=RECIPE["loss"],
                         num_vision_tokens=N_VIS)
    leaves = [s.detach().requires_grad_(True) for s in st]

    def step():
        for s in leaves:
            s.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            total, per_layer, _ = O.distill(leaves, te, am, cfg)
            _ = [float(v) for v in per_layer.values()]      # the reference's per-layer .item() for W&B (:165)
        total.backward()

    for _ in range(warmup):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    units = B * (N_VIS + txt) * n_sel
    return {"value": units / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "what": "reference op chain (oracle port, eager PyTorch ATen kernels + autograd, autocast bf16) on this GPU"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    _, n_tuple, n_sel, B, txt, D, dt = WORKLOADS[wl]
    base = cpu_baseline(wl, iters=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_sample_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dt, "data": "synthetic",
        "config": workload_config(wl, args.gpus),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(wl, n_gpus):
    desc, n_tuple, n_sel, B, txt, D, dt = WORKLOADS[wl]
    return {"workload": desc, "id": wl, "layers_distilled": n_sel, "hidden_state_tuple": n_tuple, "D": D,
            "T": N_VIS + txt, "n_vis": N_VIS, "per_gpu_batch": B, "global_batch": B * n_gpus,
            "recipe": "mse/balanced/discounted gamma=0.5", "mask": "all-ones",
            "l2": "inputs (student+teacher) far larger than the 126 MB L2; no explicit flush",
            "parallelism": f"batch-sharded dp{n_gpus}, one allreduce of 2L+2 fp64"}


# ----------------------------------------------------------------------------- main arm
_REAL_STDOUT = None


def claim_stdout():
    """Route everything libraries print to stdout (NCCL's version banner, warnings) to stderr so that
    stdout carries exactly one line: the JSON result."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4", choices=list(WORKLOADS))
    ap.add_argument("--variant", default="default", choices=["default", "ldg", "tma"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # plain `python bench.py --gpus N`: relaunch as one process per GPU (what the driver does itself)
        os.dup2(_REAL_STDOUT.fileno(), 1)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
                                   str(29000 + os.getpid() % 2000), os.path.abspath(__file__), *sys.argv[1:]])
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the distillation path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    n_gpus = world

    from mafed_b200 import build, cabi
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    lib = cabi.load()
    from mafed_b200.distill_op import distill_backward, distill_forward

    wl = args.workload
    desc, n_tuple, n_sel, B, txt, D, dt = WORKLOADS[wl]
    T = N_VIS + txt
    esize = torch.finfo(torch_dtype(dt)).bits // 8
    st, te, am = make_device_inputs(wl, rank, device)
    fd = make_method(n_sel)
    fd.populate_batch_masks = True
    fd.past_model = lambda **kw: Out(tuple(te))
    leaves = [s.detach().requires_grad_(True) for s in st]
    units_per_step = B * T * n_sel * n_gpus

    def api_step():
        """The call a user makes: FeatureDistillation.distill(...) then loss.backward()."""
        for s in leaves:
            s.grad = None
        loss = fd.distill(Out(tuple(leaves)), {"attention_mask": am})
        loss.backward()
        return loss

    # ---- (1) kernel-level steps with per-stage events (for the roofline of each kernel)
    from mafed_b200.distill_op import distill_fused
    layers = list(range(n_sel))
    coeffs, modality_kind, lang_weights = fd._tables(layers)
    plan = fd._plan(layers, coeffs, fd.distillation_coeff, modality_kind, lang_weights)
    grads = [torch.empty_like(s) for s in st]
    gout = torch.ones((), dtype=torch.float32, device=device)

    def two_pass_step(ev=None):
        if ev:
            ev[0].record()
        out, scale, ln = distill_forward(st, te, am, plan, group=None)      # fwd kernel (losses in its last CTA)
        if ev:
            ev[1].record()
        distill_backward(ln, grads, scale, gout)                            # bwd kernel
        if ev:
            ev[2].record()
        return out

    def one_pass_step(ev=None):
        if ev:
            ev[0].record()
        out, scale, ln = distill_fused(st, te, grads, am, plan, group=None)  # the whole step: one launch
        if ev:
            ev[1].record()
        distill_backward(ln, grads, scale, gout, skip_if_equals=plan.assumed_grad_out)  # the gate: returns at once
        if ev:
            ev[2].record()
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    host_us = {}
    peer, trace_marks = None, []
    if world > 1:
        from mafed_b200.comm import get_peer_comm
        peer = get_peer_comm(None)

    def api_loop(single_pass):
        fd.single_pass = single_pass
        for _ in range(args.warmup):
            api_step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank)
        sync_all()
        if single_pass and peer is not None:
            trace_marks.append(peer.trace())     # exchange cycles before the timed steps (warm-up excluded)
        with sampler:
            t0 = time.perf_counter()
            e0.record()
            for _ in range(args.steps):
                loss = api_step()
            e1.record()
            host_us[single_pass] = (time.perf_counter() - t0) / args.steps * 1e6   # CPU time to enqueue one step
            sync_all()
        return e0.elapsed_time(e1), sampler, loss

    step_stats = {}

    def stage_loop(step_fn):
        for _ in range(args.warmup):
            step_fn()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        sync_all()
        for i in range(args.steps):
            step_fn(evs[i])
        sync_all()
        a = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
        b = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
        per_step = [e[0].elapsed_time(e[2]) for e in evs]
        step_stats[step_fn.__name__] = {"median_ms": statistics.median(per_step), "best_ms": min(per_step)}
        return a, b, evs[0][0].elapsed_time(evs[-1][2])

    fwd_ms, bwd_ms, two_raw_ms = stage_loop(two_pass_step)
    fused_ms, fixup_ms, one_raw_ms = stage_loop(one_pass_step)
    two_api_ms, _, _ = api_loop(False)
    api_ms, sampler, loss = api_loop(True)          # the product default: the headline `value`
    trace0 = trace_marks[-1] if trace_marks else None
    trace1 = peer.trace() if peer is not None else None
    uncoupled = None
    if world > 1:
        # the same API step with the exchange switched off (per-rank loss), all ranks at once: what every GPU does
        # on its own.  The coupled step cannot be faster than the slowest of these.
        fd.process_group = False
        own_ms, _, _ = api_loop(True)
        fd.process_group = None
        own = torch.tensor([own_ms / args.steps], dtype=torch.float64, device=device)
        every = [torch.zeros_like(own) for _ in range(world)]
        dist.all_gather(every, own)
        uncoupled = [float(x) for x in every]
    t = torch.tensor([api_ms, two_api_ms, one_raw_ms, two_raw_ms, fwd_ms, bwd_ms, fused_ms, fixup_ms],
                     dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    api_ms, two_api_ms, one_raw_ms, two_raw_ms, fwd_ms, bwd_ms, fused_ms, fixup_ms = t.tolist()
    ms_per_step = api_ms / args.steps
    value = units_per_step / (ms_per_step * 1e-3)

    peak, peak_src = peaks()
    peer_path = False
    if world > 1:
        from mafed_b200.comm import get_peer_comm
        peer_path = get_peer_comm(None) is not None
    row_bytes = D * esize
    per_gpu_units = B * T * n_sel
    gbs = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9
    fused_bytes = 3 * row_bytes * per_gpu_units
    bwd_bytes = 3 * row_bytes * per_gpu_units
    fwd_bytes = 2 * row_bytes * per_gpu_units
    two_ms = two_api_ms / args.steps

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dt,
        "data": "synthetic", "config": workload_config(wl, n_gpus),
        "arithmetic": f"{dt} hidden states widened exactly to fp32; fp32 accumulation, fp64 final reduction; "
                      f"gradients rounded once to {dt}",
        "mode": "one-pass (loss sums + gradients from a single read of student and teacher; 3*D*e bytes per "
                "token*layer; upstream gradient checked on the device in backward)",
        "roofline": {"bound": "hbm", "kernel": "k_bwd_tma<kFused> (whole step: masks, scale table, sums, gradients, losses): "
                                                "2 reads + 1 write",
                     "achieved": gbs(fused_bytes, fused_ms), "peak": peak, "unit": "GB/s",
                     "frac": gbs(fused_bytes, fused_ms) / peak, "traffic": measured_traffic(wl), "peak_source": peak_src,
                     "bytes_per_launch": fused_bytes, "bytes_per_unit": 3 * row_bytes, "ms_per_launch": fused_ms,
                     "fixup_launch_ms": fixup_ms},
        "roofline_step": {"achieved": gbs(fused_bytes, ms_per_step), "peak": peak, "unit": "GB/s",
                          "frac": gbs(fused_bytes, ms_per_step) / peak,
                          "frac_of_nominal_8000": gbs(fused_bytes, ms_per_step) / 8000.0, "bytes_per_unit": 3 * row_bytes},
        "kernel_value": units_per_step / (one_raw_ms / args.steps * 1e-3),
        "host_us_per_step": host_us.get(True),
        "kernel_level_step_ms": {"one_pass": step_stats.get("one_pass_step"), "two_pass": step_stats.get("two_pass_step")},
        "two_pass": {
            "note": "north_star's two-kernel form (fused forward, then fused backward): 5*D*e bytes per token*layer",
            "value": units_per_step / (two_ms * 1e-3), "ms_per_step": two_ms,
            "kernel_value": units_per_step / (two_raw_ms / args.steps * 1e-3),
            "roofline_fwd": {"kernel": "k_fwd_tma (losses + scale table in its last CTA): 2 reads", "achieved": gbs(fwd_bytes, fwd_ms), "unit": "GB/s",
                             "frac": gbs(fwd_bytes, fwd_ms) / peak, "bytes_per_launch": fwd_bytes, "ms_per_launch": fwd_ms},
            "roofline_bwd": {"kernel": "k_bwd_* <kBackward>: 2 reads + 1 write", "achieved": gbs(bwd_bytes, bwd_ms),
                             "unit": "GB/s", "frac": gbs(bwd_bytes, bwd_ms) / peak, "bytes_per_launch": bwd_bytes,
                             "ms_per_launch": bwd_ms},
            "roofline_step": {"achieved": gbs(fwd_bytes + bwd_bytes, two_ms), "unit": "GB/s",
                              "frac": gbs(fwd_bytes + bwd_bytes, two_ms) / peak,
                              "frac_of_nominal_8000": gbs(fwd_bytes + bwd_bytes, two_ms) / 8000.0,
                              "bytes_per_unit": 5 * row_bytes},
        },
        # per step: the fused kernel (modality masks, scale table, loss sums + gradients, and the loss algebra in its
        # last CTA) and the backward fix-up; batch-sharded over peer memory: the same two (both exchanges inside the
        # fused kernel); NCCL fallback: masks, counts, prologue, fused, reduce, finalize, fix-up
        "gpu_launches": args.steps * (2 if (world == 1 or peer_path) else 7),
        "exchange": "none" if world == 1 else ("nvlink peer-memory mailboxes inside the fused kernel: counts at its start, "
                                               "sums in its last CTA" if peer_path else "nccl allreduce"),
        "loss": float(loss.detach()),
    }
    if trace0 is not None:
        # SM cycles the in-kernel exchanges took on rank 0 (mafed_comm_trace), per step, in us at the sampled SM clock
        mhz = float(sampler.summary().get("sm_mhz") or 1900.0)
        calls = max(1, trace1[3] - trace0[3])
        line["exchange_trace_us"] = {
            "counts_exchange_in_fused_kernel": (trace1[0] - trace0[0]) / calls / mhz,
            "sums_publish_in_tail": (trace1[1] - trace0[1]) / calls / mhz,
            "sums_wait_for_peers_in_tail": (trace1[2] - trace0[2]) / calls / mhz,
            "steps_traced": calls}
    if uncoupled is not None:
        line["uncoupled_ms_per_rank"] = uncoupled
        line["uncoupled_note"] = ("API step with the exchange off, all ranks running at once; the coupled step waits for "
                                  "the slowest GPU every step: ms_per_step vs max(uncoupled) is the cost of the exchange "
                                  "itself, max(uncoupled) vs the 1-GPU run is GPU-to-GPU variation")
    line["clocks"] = sampler.summary()

    # ---- (2) end to end with HOST buffers (pinned): H2D inputs, step, D2H gradients + loss
    if world == 1 and not args.no_other_workloads:
        try:
            line["other_workloads"] = other_workloads(wl, device, peak)
        except Exception as exc:
            line["other_workloads"] = {"error": repr(exc)}
    if not args.no_e2e:
        try:
            line["e2e"] = run_e2e(args, fd, st, te, am, device, world, units_per_step)
        except Exception as exc:  # keep the headline line even if the host path cannot allocate
            line["e2e"] = {"error": repr(exc)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["eager_torch_gpu"] = eager_gpu_baseline(wl, st, te, am)
        except Exception as exc:
            line["eager_torch_gpu"] = {"error": repr(exc)}
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            base = cpu_baseline(wl)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:
            line["cpu_baseline"] = {"error": repr(exc)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def other_workloads(skip, device, peak, steps=100, warmup=20):
    """The other BASELINE.json configurations that fit one GPU, through the public API (one-pass step):
    tokens*layers/s and the step's algorithmic GB/s.  Informational; the headline stays `value`."""
    out = {}
    for wl in ("C2", "C3", "C1"):
        if wl == skip:
            continue
        desc, n_tuple, n_sel, B, txt, D, dt = WORKLOADS[wl]
        st, te, am = make_device_inputs(wl, 0, device)
        fd = make_method(n_sel)
        fd.past_model = lambda **kw: Out(tuple(te))
        leaves = [s.detach().requires_grad_(True) for s in st]

        def step():
            for s in leaves:
                s.grad = None
            loss = fd.distill(Out(tuple(leaves)), {"attention_mask": am})
            loss.backward()

        for _ in range(warmup):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        units = B * (N_VIS + txt) * n_sel
        esize = torch.finfo(torch_dtype(dt)).bits // 8
        gbs = 3 * D * esize * units / (ms * 1e-3) / 1e9
        out[wl] = {"workload": desc, "value": units / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "step_gbs": gbs,
                   "frac": gbs / peak, "bytes_per_unit": 3 * D * esize,
                   "note": "L2-assisted: student+teacher+gradient = 234 MB vs 126 MB L2" if wl == "C1" else ""}
        del st, te, leaves, fd
    return out


def run_e2e(args, fd, st, te, am, device, world, units_per_step):
    """Same metric through the public API with host-resident inputs and outputs.

    Every step: pinned host student/teacher/mask -> device (H2D), distill + backward, gradients and the
    loss -> pinned host (D2H).  Layers are pipelined over three streams so copies overlap the kernels.
    """
    import torch.distributed as dist
    from mafed_b200.host_step import CHostStep, HostStep
    # single GPU: the C ABI's own host-buffer entry (mafed_host_step_run); batch-sharded runs use the Python
    # pipeline over the same kernels because it carries the cross-rank exchange
    hs = (CHostStep if world == 1 else HostStep)(fd, st, te, am, device)
    for _ in range(2):
        hs.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        loss = hs.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / args.e2e_steps
    return {"value": units_per_step / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": hs.h2d_bytes,
            "d2h_bytes_per_step": hs.d2h_bytes, "ms_per_step": ms, "steps": args.e2e_steps,
            "loss": float(loss), "note": hs.note}


if __name__ == "__main__":
    main()
