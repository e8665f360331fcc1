"""bench.py — headline metric of the Sepformer hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

metric : audio-seconds separated per wall-second (BASELINE.json), 8 kHz
step   : one forward of the ContSep 2-spk model over one batch (configs[1]: 16 mixtures x 4 s,
         synthetic 4096-d context embeddings, bf16 tensor-core mode) per GPU; weak scaling —
         every rank separates its own batch, no data-path collective (mixtures are independent).
value  : whole-job audio-s / max-over-ranks device time, inputs already resident in HBM.
e2e    : same metric through the host-buffer C-ABI entry: pinned host mixtures -> H2D -> forward -> D2H of
         the separated waveforms, every step; `value` = the pipelined entry (cse_pipeline_*, two steps in
         flight), `blocking_call_value` = one blocking cse_forward_host call per step.
parity : mixtures 0 and 15 of the timed batch checked against the CPU oracle outside the timed region.
train  : BASELINE configs[2] measured in the same process (ContExt forward + -SI-SNR + backward + DDP gradient
         all-reduce + clip + AdamW under autocast, 2 mixtures x 4 s per GPU) so that the driver's 1 -> 8 GPU runs
         time the one collective this path has.  The whole step replays as ONE CUDA graph per rank
         (cse_b200.runtime.GraphedStep; at N > 1 with DistributedDataParallel's NCCL all-reduce captured inside, run
         last and under a watchdog so that a capture problem costs this sub-object, not the line).
roofline / cpu_baseline: see DESIGN.md §Measurement.
--impl reference: the reference's CPU path — its op sequence on the stock torch modules it instantiates
(oracle/eager_reference.py; the Python reference itself cannot travel to the GPU box), all host threads, one 4 s
mixture per step.
--workload train (not the driver's default): BASELINE configs[2] on its own line — ContExt forward + -SI-SNR loss +
backward + gradient all-reduce (stock DDP over NCCL) + clip + AdamW, 2 mixtures x 4 s per GPU, under torch.autocast
(bf16 tensor-core forward / recompute / dgrad / wgrad / attention backward; --train-precision fp32 = the parity
kernels); --train-graph off | auto | ddp, --train-ragged, --train-seconds, --train-loss pit.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"
BATCH, SECONDS, SR, SPK, CTX_TOKENS = 16, 4, 8000, 2, 1
T = SECONDS * SR


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"],
                    tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML in-process (one light query every 100 ms);
    `nvidia-smi -lms` as the fallback.  (A looping nvidia-smi beside a launch-bound step — the training leg issues
    ~1000 launches per 18 ms — slowed that step from 18 to 32 ms: its full-device queries contend for the driver
    lock the kernel launches need.  The forward legs replay one CUDA graph and were not affected.)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self._stop = None, threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001  (no NVML binding: fall back to the command-line tool)
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        flags = [("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)]
        try:
            smax = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            smax = ""
        while not self._stop.wait(0.1):
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:  # noqa: BLE001
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.lines.append(",".join([str(sm), str(smax), ""] + ["Active" if mask & bit else "Not Active" for _, bit in flags]))
            except Exception:  # noqa: BLE001
                pass

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def sample_once(self):
        """One NVML query from the calling thread (the training leg: a polling thread beside its ~1000 launches per
        step — in-process NVML as much as a looping nvidia-smi — slowed the step several-fold)."""
        try:
            import pynvml as n
            if self.nvml is None:
                n.nvmlInit()
                visible = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
                self.handle = n.nvmlDeviceGetHandleByIndex(phys)
                self.nvml = n
                self._smax = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
            sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            try:
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:  # noqa: BLE001
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            flags = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                     n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
            self.lines.append(",".join([str(sm), str(getattr(self, "_smax", "")), ""] +
                                       ["Active" if mask & bit else "Not Active" for bit in flags]))
        except Exception:  # noqa: BLE001
            pass

    def summary(self):
        """Summary of the samples taken with sample_once() (no polling thread to stop)."""
        return self._summarise()

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        return self._summarise()

    def _summarise(self):
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(steps, warmup):
    """The reference's own CPU path: its op sequence on the stock torch modules it instantiates (nn.MultiheadAttention,
    nn.LayerNorm, nn.GroupNorm, nn.Conv1d, ... — `oracle/eager_reference.py`, bit-identical to the reference modules'
    fp32 output, tests/test_oracle.py), one 4 s mixture per step, fp32, all host threads.  The parameters sit in our
    module mirror used as a plain container on the CPU; no kernel of ours runs."""
    import torch
    import cse_b200  # noqa: F401
    from cse_b200 import synth
    from cse_b200.models.ContSep import Sepformer
    from oracle import eager_reference as ER
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = Sepformer(SPK, add_mt=True)
    model.add_mt_pipeline()
    model.load_state_dict(synth.make_state_dict("contsep", SPK, seed=0))
    model.eval()
    mix, _ = synth.make_mixture(1, T, SPK, seed=1234)
    ctx = synth.make_context(1, CTX_TOKENS, seed=1234)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            ER.eager_forward(model, mix, ctx)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    return SECONDS / per_step, per_step, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    if getattr(args, "workload", "forward") == "train":
        value, per_step, cores = cpu_train_rate(steps, warmup, args.train_seconds, args.train_loss)
        sample = (f"1 mixture x {args.train_seconds} s forward + loss ({args.train_loss}) + backward under autograd per step ({steps} timed steps, "
                  f"{warmup} warm-up), fp32, oracle port of the reference modules")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ContExt 2-spk forward+backward (BASELINE.json configs[2]); bounded sample: 1 of the 2 mixtures per step, no optimizer step",
                       "batch_per_step": 1, "seconds": args.train_seconds, "sample_rate": SR, "ctx_tokens": CTX_TOKENS,
                       "loss": args.train_loss},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }), flush=True)
        return
    value, per_step, cores = cpu_reference_rate(steps, warmup)
    sample = f"1 mixture x {SECONDS} s per step ({steps} timed steps, {warmup} warm-up), fp32, the reference's op sequence on its own stock torch modules (oracle/eager_reference.py, bit-identical to the reference output)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ContSep 2-spk SpokenWoz-shape forward (configs[1]); bounded sample: 1 of 16 mixtures per step",
                   "batch_per_step": 1, "seconds": SECONDS, "sample_rate": SR, "ctx_tokens": CTX_TOKENS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def _op_point_dsisnr(est, gold):
    """max |SI-SNR(est, target) - SI-SNR(gold, target)| in dB over streams, targets = gold + seeded noise such that
    SI-SNR(gold, target) = 0 / 10 / 15 dB (the range the metric is used at)."""
    import torch
    from oracle import sepformer_oracle as O
    g = torch.Generator().manual_seed(99)
    noise = torch.randn(gold.shape, generator=g, dtype=torch.float64)
    gz = gold.double() - gold.double().mean(1, keepdim=True)
    worst = 0.0
    for level_db in (0.0, 10.0, 15.0):
        scale = (gz.pow(2).sum(1, keepdim=True) / noise.pow(2).sum(1, keepdim=True)).sqrt() * 10 ** (-level_db / 20)
        target = gold.double() + noise * scale
        for s in range(gold.shape[2]):
            a = O.tm_si_snr(est[:, :, s].double(), target[:, :, s])
            b = O.tm_si_snr(gold[:, :, s].double(), target[:, :, s])
            worst = max(worst, (a - b).abs().max().item())
    return worst


def parity_check(model, sd, mix_h, ctx_h, est_bf16, pred_bf16, dev):
    """Outside the timed region: mixtures 0 and 15 of the timed batch against the CPU oracle (fp32), for the bf16
    output that was just timed and for one fp32-mode forward of the same two mixtures."""
    import numpy as np
    import torch
    from oracle import sepformer_oracle as O
    idx = [0, BATCH - 1]
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = [O.sepformer_forward(sd, mix_h[i:i + 1], ctx_h[i:i + 1], "contsep", SPK) for i in idx]
    ref_est = torch.cat([r[0] for r in ref], 0)
    ref_pred = torch.cat([r[1] for r in ref], 0)

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm()).item()

    model.precision, graph = "fp32", model.use_cuda_graph
    model.use_cuda_graph = False
    with torch.no_grad():
        est32, pred32 = model(mix_h[idx].to(dev), ctx_h[idx].to(dev))
    model.precision, model.use_cuda_graph = "bf16", graph
    e16, p16 = est_bf16[idx].cpu(), pred_bf16[idx].cpu()
    out = {"checked_mixtures": idx, "against": "CPU oracle (fp32), live",
           "fp32_rel_l2": rel(est32.cpu(), ref_est), "fp32_context_pred_rel_l2": rel(pred32.cpu(), ref_pred),
           "bf16_rel_l2": rel(e16, ref_est), "bf16_context_pred_rel_l2": rel(p16, ref_pred),
           "bf16_dsisnr_db": _op_point_dsisnr(e16, ref_est),
           "tolerance": "fp32 <= 1e-4 rel-L2; bf16 <= 0.05 dB SI-SNR at 0/10/15 dB operating points and rel-L2 <= the reference's own bf16 drift"}
    drift = []
    for i in idx:
        f = os.path.join(ROOT, "tests", "golden", f"baseline_cfg2_mix{i}.npz")
        if os.path.isfile(f):
            with np.load(f) as z:
                drift.append(float(z["bf16_ref_rel_l2"]))
    if drift:
        out["reference_bf16_rel_l2"] = max(drift)
    out["ok"] = bool(out["fp32_rel_l2"] <= 1e-4 and out["bf16_dsisnr_db"] <= 0.05
                     and (not drift or out["bf16_rel_l2"] <= max(drift)))
    return out


def gemm_flops_per_forward(ps):
    """FLOPs issued by the tcgen05 GEMM launches of one forward (2 per multiply-add)."""
    N, F = 256, 1024
    rows = ps.rows_intra + ps.rows_inter
    per_row = 2 * (3 * N * N + N * N + 2 * N * F)
    stacks = 16 * rows * per_row                       # 8 layers x 2 blocks over both stacks
    BL = ps.B * ps.L
    head = 2 * BL * N * N * (1 + ps.spk + 3 * ps.spk)  # conv1d, conv2d, output, gate, end_conv
    return float(stacks + head)


def attention_flops_per_forward(ps):
    N = 256
    tok_i, tok_e = ps.B * ps.S * ps.n_intra, ps.B * 250 * ps.n_inter
    return float(16 * (tok_i * 4 * ps.n_intra * N + tok_e * 4 * ps.n_inter * N))


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel class (tcgen05 GEMMs), from the
# ncu --set full capture in profiles/r02_ncu_kernels.txt section (a): mean over the three GEMM-class launches of an
# intra and an inter layer (QKV 222 / 229 MB, out-proj 293 / 299 MB, feed-forward kernel with the layer's LayerNorms
# 467 / 480 MB; algorithmic 280 + 350 + 350 MB of GEMM traffic + 2 x 210 MB of LayerNorm traffic, part of it L2-only).
# A committed measurement, not something bench.py can re-measure.
GEMM_CLASS_DRAM_BYTES_PER_LAUNCH = 3.317e8


def run_ours(args):
    import torch
    import torch.distributed as dist
    import cse_b200  # noqa: F401
    from cse_b200 import _lib, shapes, synth
    from cse_b200.models.ContSep import Sepformer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if not args.no_train and args.train_graph == "auto":   # the DDP step is captured into a CUDA graph (see below)
            os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "0"
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    sd = synth.make_state_dict("contsep", SPK, seed=0)
    model = Sepformer(SPK, add_mt=True)
    model.add_mt_pipeline()
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    model.precision = "bf16"
    model.use_cuda_graph = True     # the ~255 launches of a forward replay as one CUDA graph

    # every rank separates its own batch (weak scaling); seeds differ per rank
    mix_h, _ = synth.make_mixture(BATCH, T, SPK, seed=1234 + rank)
    ctx_h = synth.make_context(BATCH, CTX_TOKENS, seed=1234 + rank)
    mix_h, ctx_h = mix_h.pin_memory(), ctx_h.pin_memory()
    mix_d, ctx_d = mix_h.to(dev), ctx_h.to(dev)
    est_h = torch.empty(BATCH, T, SPK).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        sync_all()
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = t[0].item(), t[1].item() / 1e3
        return ms, wall

    def step_device():
        with torch.no_grad():
            return model(mix_d, ctx_d)

    def step_host():
        model.separate_host(mix_h, ctx_h, est_host=est_h)

    pipe = model.host_pipeline(BATCH, T, c=CTX_TOKENS, depth=2)
    in_flight = []

    def step_pipelined():      # submit step i (H2D + forward + D2H enqueued), then collect step i-1
        in_flight.append(pipe.submit(mix_h, ctx_h))
        if len(in_flight) > 1:
            pipe.wait(in_flight.pop(0))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    step_host()
    for _ in range(3):
        step_pipelined()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.lines.clear()          # keep only samples taken during the timed regions
    # kernels per step: counted on one eager forward (the timed steps replay the same launches as a
    # captured CUDA graph, which the host-side counter does not see)
    model.use_cuda_graph = False
    n0 = lib.cse_launch_count()
    step_device()
    launches = lib.cse_launch_count() - n0
    model.use_cuda_graph = True
    ms_total, _ = timed(step_device, args.steps)
    ms_step = ms_total / args.steps
    audio_s = BATCH * SECONDS * world
    value = audio_s / (ms_step / 1e3)

    # end to end through the host-buffer entry: device time is not enough here (the call blocks on
    # the D2H), so use the max-over-ranks wall clock around the loop
    _, wall = timed(step_host, args.steps)
    e2e_blocking = audio_s / (wall / args.steps)
    # pipelined: K submits, K results; the loop's trailing in-flight step completes inside timed()'s synchronize
    _, wall = timed(step_pipelined, args.steps)
    e2e_value = audio_s / (wall / args.steps)
    while in_flight:
        pipe.wait(in_flight.pop(0))
    clocks = sampler.stop() if rank == 0 else None
    h2d = BATCH * T * 4 + BATCH * CTX_TOKENS * 4096 * 4
    d2h = BATCH * T * SPK * 4 + BATCH * 256 * 4

    # per-kernel-class device time over the same K steps (CUDA event pair around every launch of
    # the class, on the launching stream)
    roofline = None
    if rank == 0:
        ps = shapes.path_shape(BATCH, T, CTX_TOKENS, SPK)
        peaks = load_peaks()
        model.use_cuda_graph = False      # event pairs are recorded around eager launches
        lib.cse_profile_enable(1)
        for _ in range(args.steps):
            step_device()
        torch.cuda.synchronize()
        ms_cls = (C.c_double * 5)()
        n_cls = (C.c_longlong * 5)()
        _lib.call("cse_profile_collect", ms_cls, n_cls, 5)
        lib.cse_profile_enable(0)
        # the GEMM class = gemm_tc_kernel launches (slot 0) + the fused feed-forward kernel (slot 4)
        f_ms, f_n = ms_cls[4] / args.steps, n_cls[4] // args.steps
        ms_cls[0] += ms_cls[4]
        n_cls[0] += n_cls[4]
        names = ["tcgen05_gemm", "attention", "layernorm", "simt_gemm"]
        share = {names[i]: {"ms_per_step": ms_cls[i] / args.steps, "launches_per_step": n_cls[i] // args.steps}
                 for i in range(4) if n_cls[i]}
        g_ms, g_n = ms_cls[0] / args.steps, max(1, n_cls[0] // args.steps)
        flops_per_launch = gemm_flops_per_forward(ps) / g_n
        achieved = flops_per_launch / (g_ms / g_n * 1e-3) / 1e12
        dominant = None
        if f_n:      # the step's dominant kernel on its own: Linear(256,1024) -> ReLU -> Linear(1024,256) (+ the layer's LayerNorms)
            rows = ps.B * ps.S * ps.n_intra + ps.B * 250 * ps.n_inter            # rows of an intra + an inter stack
            f_flops = 16 * rows * 4.0 * 256 * 1024 / f_n                         # 2 blocks x 8 layers per stack; 2 GEMMs x 2 FLOP per MAC
            f_ach = f_flops / (f_ms / f_n * 1e-3) / 1e12
            dominant = {"kernel": "cse::ffn_tc_kernel (fused feed-forward sub-block; from 16 k rows per stack up it also carries the layer's two LayerNorms)",
                        "launches_per_step": int(f_n), "avg_launch_ms": f_ms / f_n, "share_of_step": f_ms / ms_step,
                        "algorithmic_flops_per_launch": f_flops, "achieved": f_ach, "unit": "TFLOP/s",
                        "frac": f_ach / peaks["tflops_sustained"]}
        a_ms = ms_cls[1] / args.steps
        roofline = {
            "kernel": "cse::gemm_tc_kernel + cse::ffn_tc_kernel (tcgen05 + TMA bf16 GEMMs: all Linear / 1x1-conv layers; the fused "
                      "FFN kernel also carries the layer's two LayerNorms, whose time counts here while their work is not FLOPs)",
            "bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
            "frac": achieved / peaks["tflops_sustained"], "peak_source": f"{peaks['source']} bf16 sustained (cuBLAS)",
            "traffic": GEMM_CLASS_DRAM_BYTES_PER_LAUNCH,
            "algorithmic_flops_per_launch": flops_per_launch,
            "avg_launch_ms": g_ms / g_n,
            "kernel_share_of_step": g_ms / ms_step,
            "classes": share,
            "dominant_kernel": dominant,
            "attention": {"achieved": attention_flops_per_forward(ps) / (a_ms * 1e-3) / 1e12 if a_ms else None,
                          "unit": "TFLOP/s (QK^T + PV only; the kernels are bound by softmax / tcgen05.ld, not by the MMAs)"},
            "whole_step": {"achieved": shapes.algorithmic_flops(ps) / (ms_step * 1e-3) / 1e12,
                           "unit": "TFLOP/s", "frac": shapes.algorithmic_flops(ps) / (ms_step * 1e-3) / 1e12 / peaks["tflops_sustained"]},
        }

    # N = 1: the training leg runs BEFORE the legs that use the host cores (parity oracle, CPU baseline: 16 intra-op
    # threads whose pool keeps spinning afterwards); an eagerly launched step measured 27 ms instead of 18 ms after them.
    # N > 1: it runs LAST, as one captured CUDA graph per rank with DistributedDataParallel's all-reduce inside (a
    # replay does not care about the host) and under a watchdog: everything else of the line is measured by then, and if
    # the capture or a replay hangs at this N the line is printed without the training measurement instead of not at all.
    train = None
    train_last = world > 1 and not args.no_train and args.train_graph == "auto"
    if not args.no_train and not train_last:
        train = measure_train(args, world, rank, local, dev, steps=max(3, min(args.steps, 10)), quiet=True)

    parity = None
    if rank == 0:
        est_bf16, pred_bf16 = step_device()
        torch.cuda.synchronize()
        parity = parity_check(model, sd, mix_h, ctx_h, est_bf16, pred_bf16, dev)
    pipe.close()
    del pipe

    cpu = None
    if rank == 0 and world == 1:
        v, per_step, cores = cpu_reference_rate(3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"1 of the 16 mixtures ({SECONDS} s) per forward, 3 timed forwards + 1 warm-up, fp32, the reference's op sequence on stock torch modules, {cores} threads"}

    def emit(train_obj):
        if rank != 0:
            return
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ContSep 2-spk SpokenWoz-shape forward (BASELINE.json configs[1])",
                       "batch_per_gpu": BATCH, "seconds": SECONDS, "sample_rate": SR, "ctx_tokens": CTX_TOKENS,
                       "num_spks": SPK, "weights": "random-init (seeded)", "sharding": f"dp{world}: independent mixtures, no collective",
                       "l2": "no flush: one step streams ~1.4 GB of activations through a 126 MB L2"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                      "api": "Sepformer.host_pipeline -> cse_pipeline_submit / cse_pipeline_wait (pinned host buffers, 2 steps in flight, every step's H2D and D2H inside the timed region)",
                                      "blocking_call_value": e2e_blocking,
                                      "blocking_call_api": "Sepformer.separate_host -> cse_forward_host (one blocking call per step)"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "train": train_obj,
        }), flush=True)

    if train_last:
        import threading
        dist.barrier()
        torch.cuda.synchronize()
        finished = threading.Event()

        def watchdog():
            if not finished.wait(TRAIN_WATCHDOG_S):
                emit({"error": f"the captured DistributedDataParallel step did not finish within {TRAIN_WATCHDOG_S} s at N = {world}: "
                               "no training measurement on this line"})
                sys.stdout.flush()
                os._exit(0)

        threading.Thread(target=watchdog, daemon=True).start()
        failed = None
        try:
            args.train_graph = "ddp"
            train = measure_train(args, world, rank, local, dev, steps=max(3, min(args.steps, 10)), quiet=True)
        except Exception as e:      # (a rank that fails alone leaves the others in a collective: their watchdogs end them)
            failed = repr(e)
        finished.set()
        if failed is not None:
            emit({"error": "training leg failed: " + failed})
            sys.stdout.flush()
            os._exit(0)
    emit(train)
    if world > 1:
        dist.destroy_process_group()


TRAIN_WATCHDOG_S = 300
TRAIN_BATCH = 2


def cpu_train_rate(steps, warmup, seconds=SECONDS, loss_kind="sisnr"):
    """Oracle port, forward + backward of the training loss under autograd, one mixture per step."""
    import torch
    import cse_b200  # noqa: F401
    from cse_b200 import synth
    from oracle import sepformer_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    variant = "contsep" if loss_kind == "pit" else "context"
    sd = {k: (v.requires_grad_(True) if "pos_enc" not in k else v)
          for k, v in synth.make_state_dict(variant, SPK, seed=0).items()}
    mix, src = synth.make_mixture(1, seconds * SR, SPK, seed=1234)
    ctx = synth.make_context(1, CTX_TOKENS, seed=1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if loss_kind == "pit":
            est, pred = O.sepformer_forward(sd, mix, ctx, "contsep", SPK)
            pit, _ = O.pit_si_snr(est, src)
            loss = pit.mean() + torch.nn.functional.cross_entropy(pred, torch.zeros(1, dtype=torch.long))
        else:
            est = O.sepformer_forward(sd, mix, ctx, "context", SPK)
            loss = -O.tm_si_snr(est[:, :, 0], src[:, :, 0]).mean()
        loss.backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    return seconds / per_step, per_step, cores


def measure_train(args, world, rank, local, dev, steps, quiet=False):
    """BASELINE configs[2]: ContExt forward + loss + backward + DDP gradient all-reduce + clip + AdamW(amsgrad),
    batch 2 per GPU (weak scaling), under torch.autocast like the reference (train_ContExt.py:365) unless
    --train-precision fp32.  Returns the result dict on rank 0 (None elsewhere).  The process group must exist when
    world > 1."""
    import torch
    import torch.distributed as dist
    import cse_b200  # noqa: F401
    from cse_b200 import _lib, losses, shapes, synth
    from cse_b200.models.ContExt import Sepformer
    from cse_b200.models.ContSep import Sepformer as ContSep

    seconds, loss_kind = args.train_seconds, args.train_loss
    amp = args.train_precision != "fp32"
    Tt = seconds * SR
    lib = _lib.load()
    if loss_kind == "pit":       # train_ContSep.py: PIT SI-SNR + cross-entropy on the context selector
        model = ContSep(SPK, add_mt=True)
        model.add_mt_pipeline()
        model.load_state_dict(synth.make_state_dict("contsep", SPK, seed=0))
    else:                        # train_ContExt.py: -SI-SNR of the extracted stream
        model = Sepformer(SPK, add_ctx=True)
        model.add_ctx_pipeline()
        model.load_state_dict(synth.make_state_dict("context", SPK, seed=0))
    model = model.to(dev).train()
    model.precision = None if amp else "fp32"
    net = model
    cap_stream = None
    if world > 1:   # the reference's own wrapper (train_ContExt.py:269-273): bucketed NCCL all-reduce overlapped with backward
        ddp_kw = {}
        if args.ddp_bucket_view:
            ddp_kw["gradient_as_bucket_view"] = True
        if args.ddp_bucket_mb:
            ddp_kw["bucket_cap_mb"] = args.ddp_bucket_mb
        if args.ddp_static_graph:
            ddp_kw["static_graph"] = True
        if args.train_graph == "ddp":   # (opt-in) DDP built on the stream its warm-up iterations and the capture use
            cap_stream = torch.cuda.Stream()
            cap_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap_stream):
                net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], output_device=local,
                                                                find_unused_parameters=False, **ddp_kw)
            torch.cuda.current_stream().wait_stream(cap_stream)
        else:
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], output_device=local,
                                                            find_unused_parameters=False, **ddp_kw)
    fused_opt = args.train_optim == "fused"
    if fused_opt:   # clip_grad_norm_ + AdamW(amsgrad) in three launches (cse_optim_step), no host sync on the norm
        from cse_b200.optim import AdamW as FusedAdamW
        opt = FusedAdamW(model.parameters(), lr=1e-4, amsgrad=True)         # train_ContSep.py:233
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, amsgrad=True)
    sisnr = losses.ScaleInvariantSignalNoiseRatio()
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)

    if args.train_ragged:        # DailyTalk-like lengths U(1.5 s, 8 s), right-padded to the batch max (dataset_train_CSE.py:532-569)
        g = torch.Generator().manual_seed(777 + rank)
        lens = (torch.rand(TRAIN_BATCH, generator=g) * 6.5 + 1.5) * SR
        lens = [int(x) // 8 * 8 for x in lens.tolist()]
        Tt = max(lens)
    mix_h, src_h = synth.make_mixture(TRAIN_BATCH, Tt, SPK, seed=4321 + rank)
    if args.train_ragged:
        for i, n in enumerate(lens):
            mix_h[i, n:] = 0
            src_h[i, n:] = 0
    ctx_h = synth.make_context(TRAIN_BATCH, CTX_TOKENS, seed=4321 + rank)
    tgt_h = (src_h if loss_kind == "pit" else src_h[:, :, 0]).contiguous().pin_memory()
    label = torch.zeros(TRAIN_BATCH, dtype=torch.long, device=dev)
    mix_h, ctx_h = mix_h.pin_memory(), ctx_h.pin_memory()
    mix_d, ctx_d, tgt_d = mix_h.to(dev), ctx_h.to(dev), tgt_h.to(dev)

    def fwd_bwd(module, mix, ctx, tgt):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):      # train_ContExt.py:365 (--bf16)
            if loss_kind == "pit":
                est, pred = module(mix, ctx)
                loss = (losses.get_si_snr_with_pitwrapper(est, tgt).mean()   # train_ContSep.py:391-394
                        + torch.nn.functional.cross_entropy(pred.float(), label))
            else:
                est = module(mix, ctx)
                loss = -sisnr(est[:, :, 0], tgt)                            # train_ContExt.py:366-367
        loss.backward()
        return loss

    def step(mix, ctx, tgt, module=None):
        opt.zero_grad(set_to_none=True)
        loss = fwd_bwd(module or net, mix, ctx, tgt)
        if fused_opt:
            opt.step(max_norm=5.0)                                          # train_ContSep.py:411-416 in one call
        else:
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)         # train_ContSep.py:411
            opt.step()
        return loss

    def step_device():
        return step(mix_d, ctx_d, tgt_d)

    def step_no_allreduce():     # the same step without the collective (DDP.no_sync): what the all-reduce adds
        if world > 1:
            with net.no_sync():
                return step(mix_d, ctx_d, tgt_d)
        return step(mix_d, ctx_d, tgt_d)

    def step_host():
        loss = step(mix_h.to(dev, non_blocking=True), ctx_h.to(dev, non_blocking=True),
                    tgt_h.to(dev, non_blocking=True))
        return loss.item()                                                  # D2H of the step's loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k, sample_clocks=False):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        if sample_clocks and rank == 0:
            # the device is still executing the last enqueued step(s): clocks under load, while the query — slow when
            # it has to wait for the driver lock (it cost ~2.5 ms per step when issued between steps) — delays only
            # the host, after the closing event has been enqueued
            sampler.sample_once()
            sampler.sample_once()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        sync_all()
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = t[0].item(), t[1].item() / 1e3
        return ms, wall

    sampler = ClockSampler(local)   # sampled inline (sample_once) between timed steps, not by a polling thread
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.sample_once()       # NVML initialisation happens here, outside the timed region
        sampler.lines.clear()
    n0 = lib.cse_launch_count()
    loss0 = step_device()
    launches = lib.cse_launch_count() - n0
    loss0 = float(loss0.item())
    # One process, fused optimiser: the whole step (forward, loss, backward, clip + AdamW) replays as ONE CUDA graph
    # (runtime.GraphedStep) — run eagerly its ~1000 short launches are bound by the launching host thread.  A
    # DistributedDataParallel step (N > 1) stays eager.
    graphed = None
    ms_eager = None
    if world > 1 and args.train_graph == "ddp":
        ms_e, _ = timed(step_device, steps)          # the eagerly launched DDP step, for the comm block below
        ms_eager = ms_e / steps
    if fused_opt and ((world == 1 and args.train_graph != "off") or (world > 1 and args.train_graph == "ddp")):
        from cse_b200.runtime import GraphedStep
        # (DDP: eleven warm-up iterations before the capture, as torch's CUDA-graph notes ask for)
        graphed = GraphedStep(lambda m, c, t: step(m, c, t), mix_d, ctx_d, tgt_d, warmup=3 if world == 1 else 11,
                              stream=cap_stream)

        def step_device():                                                   # noqa: F811
            return graphed(mix_d, ctx_d, tgt_d)

        def step_host():                                                     # noqa: F811
            return graphed(mix_h, ctx_h, tgt_h).item()                       # H2D into the static inputs, replay, D2H of the loss

        for _ in range(2):
            step_device()
        torch.cuda.synchronize()
    ms_total, _ = timed(step_device, steps, sample_clocks=True)
    ms_step = ms_total / steps
    audio_s = TRAIN_BATCH * (Tt / SR) * world
    value = audio_s / (ms_step / 1e3)
    _, wall = timed(step_host, steps)
    e2e_value = audio_s / (wall / steps)
    comm = None
    if world > 1:
        # (with a captured DDP step the no_sync() comparison is not taken: DDP's eager no_sync bookkeeping after a
        # capture is not what the graph replays; the eager step measured before the capture is reported beside it)
        ms_nosync = None
        if graphed is None:
            ms_nosync, _ = timed(step_no_allreduce, steps)
        flat = torch.zeros(n_params, dtype=torch.float32, device=dev)

        def allreduce_alone():   # the same bytes in DDP's default 25 MB buckets, nothing else running
            for chunk in flat.split(25 * 1024 * 1024 // 4):
                dist.all_reduce(chunk)

        ms_ar, _ = timed(allreduce_alone, steps)
        alone = ms_ar / steps
        exposed = max(0.0, ms_step - ms_nosync / steps) if ms_nosync is not None else None
        comm = {"collective": "NCCL all-reduce of fp32 gradients (stock DistributedDataParallel"
                              + (f", {ddp_kw}" if ddp_kw else ", default arguments as train_ContExt.py:269-273: 25 MB buckets") + ")",
                "allreduce_bytes_per_step": n_params * 4, "step_ms": ms_step, "step_ms_eager": ms_eager,
                "step_ms_without_allreduce": ms_nosync / steps if ms_nosync is not None else None,
                "exposed_allreduce_ms": exposed, "allreduce_alone_ms": alone,
                "overlap_fraction": max(0.0, min(1.0, 1.0 - exposed / alone)) if (alone > 0 and exposed is not None) else None}
    clocks = sampler.summary() if rank == 0 else None
    if rank != 0:
        return None
    ps = shapes.path_shape(TRAIN_BATCH, Tt, CTX_TOKENS, SPK if loss_kind == "pit" else 1)
    peaks = load_peaks()
    # forward + recomputed forward (layer checkpointing) + dgrad + wgrad = 4x the forward contractions
    flops = 4.0 * shapes.algorithmic_flops(ps)
    achieved = flops / (ms_step * 1e-3) / 1e12
    kernels = ("cse::gemm_tc_kernel / ffn-free bf16 layers + attention (tcgen05 forward, recompute, dgrad, wgrad)" if amp
               else "cse::gemm_simt_kernel / wgrad_kernel / attention_f32 + attention_bwd (fp32 FFMA parity mode)")
    roofline = {"kernel": "whole step: " + kernels, "bound": "tensor", "achieved": achieved,
                "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"],
                "peak_source": f"{peaks['source']} bf16 sustained (cuBLAS)", "traffic": None}
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if amp else "f32", "data": "synthetic",
        "config": {"workload": ("ContSep 2-spk forward + PIT SI-SNR + selector CE" if loss_kind == "pit" else
                                "ContExt 2-spk forward + -SI-SNR loss")
                   + " + backward + grad all-reduce + clip + AdamW (BASELINE.json configs[2])",
                   "batch_per_gpu": TRAIN_BATCH, "seconds": Tt / SR, "ragged_lengths": bool(args.train_ragged),
                   "sample_rate": SR, "ctx_tokens": CTX_TOKENS, "num_spks": SPK, "weights": "random-init (seeded)",
                   "precision": "torch.autocast(bfloat16): bf16 tensor-core transformer layers, fp32 elsewhere" if amp else "fp32",
                   "sharding": f"dp{world}: stock DistributedDataParallel, NCCL gradient all-reduce overlapped with backward",
                   "l2": "no flush: one step streams several GB of activations through a 126 MB L2"},
        "clocks": clocks, "loss": loss0, "n_params": n_params, "comm": comm,
        "step_execution": ("one CUDA graph replay per step (cse_b200.runtime.GraphedStep)"
                           + (", DistributedDataParallel's NCCL all-reduce captured inside" if world > 1 else "")
                           if graphed is not None
                           else "eager launches (DistributedDataParallel step)" if world > 1 else "eager launches"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(mix_h.numel() + ctx_h.numel() + tgt_h.numel()) * 4,
                "d2h_bytes_per_step": 4, "api": "model(mix, ctx) -> loss.backward() -> optimizer.step() from pinned host buffers"},
        "gpu_launches": int(launches), "roofline": roofline,
    }


def run_train(args):
    """--workload train: the configs[2] measurement on its own line."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if args.train_graph == "ddp":   # whole-step capture with the NCCL all-reduce inside (torch's CUDA-graph notes)
            os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "0"
        dist.init_process_group("nccl", device_id=dev)
    res = measure_train(args, world, rank, local, dev, steps=args.steps)
    if rank == 0:
        if world == 1:
            v, per_step, cores = cpu_train_rate(2, 1, args.train_seconds, args.train_loss)
            res["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"1 mixture ({args.train_seconds} s) forward+backward per step under autograd, 2 timed steps + 1 warm-up, fp32 oracle port, {cores} threads"}
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="forward", choices=["forward", "train"],
                    help="forward = BASELINE configs[1] (the driver's contract); train = configs[2]")
    ap.add_argument("--train-seconds", type=int, default=SECONDS, help="--workload train: mixture length (4; 16 = max_sp_len cap)")
    ap.add_argument("--train-loss", default="sisnr", choices=["sisnr", "pit"],
                    help="--workload train: ContExt -SI-SNR (train_ContExt.py:367) or ContSep PIT + selector CE (train_ContSep.py:391-394)")
    ap.add_argument("--train-optim", default="fused", choices=["fused", "torch"],
                    help="fused: cse_b200.optim.AdamW (clip + AdamW(amsgrad) in three launches); torch: stock "
                         "clip_grad_norm_ + torch.optim.AdamW as the reference calls them")
    ap.add_argument("--train-precision", default="autocast", choices=["autocast", "fp32"],
                    help="training leg: torch.autocast(bf16) like the reference's --bf16 (tensor-core layers), or the fp32 parity kernels")
    ap.add_argument("--train-ragged", action="store_true",
                    help="training leg: DailyTalk-like lengths U(1.5 s, 8 s) right-padded to the batch max instead of fixed --train-seconds")
    ap.add_argument("--train-graph", default="auto", choices=["auto", "off", "ddp"],
                    help="training leg: auto = the whole step replays as one CUDA graph (N = 1 always; N > 1: in the default "
                         "forward workload's `train` sub-object, captured with DistributedDataParallel's all-reduce inside "
                         "and run last under a watchdog; `--workload train` at N > 1 stays eager so that its no_sync() "
                         "comparison can be taken); off = always eager; ddp = capture the DDP step in `--workload train` too")
    ap.add_argument("--no-train", action="store_true", help="forward workload: skip the `train` sub-object")
    ap.add_argument("--ddp-bucket-view", action="store_true", help="training leg, N > 1: DDP(gradient_as_bucket_view=True) (A/B; the reference uses the defaults)")
    ap.add_argument("--ddp-bucket-mb", type=int, default=0, help="training leg, N > 1: DDP(bucket_cap_mb=...) (A/B; default 25)")
    ap.add_argument("--ddp-static-graph", action="store_true", help="training leg, N > 1: DDP(static_graph=True) (A/B)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
