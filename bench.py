#!/usr/bin/env python
"""Benchmark of the DepthG hot path on B200 (see the contract in DESIGN.md section "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our sm_100a path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU

One "step" = ContrastiveCorrelationLoss forward + backward (depth-guided FPS, bilinear
gathers, fused correlation loss, scatter backward) on one batch of the cocostuff27
ViT-B/8 training shape (BASELINE.json configs[1]): B=32 per GPU, C=768, 28x28, dim=90,
feature_samples=11, fps sampling, pointwise, depth term on.  Prints ONE JSON line.

Keys beyond the base contract: ``roofline`` (SURVEY 8(d)'s algorithmic bytes of the step over the dominant kernel's
live CUDA-event time, with the kernel's ncu DRAM traffic and traffic / algorithmic beside it), ``roofline_step``
(the same bytes over the whole step), ``extra_configs`` (BASELINE configs[3], the S = 12 paper value and the dense
configs[4] as side measurements), ``breakdown_us`` (library-side per-kernel events), ``e2e`` (pinned host buffers, the faster of a
double-buffered and a serialised loop, both reported), ``cpu_baseline`` (oracle port on the host cores),
``reference_ops_on_gpu`` (the reference's op sequence as stock torch ops on this GPU), ``cuda_graph`` (same step,
no host work) and ``torch_negative_sampler`` (same step with the reference's randperm stream), ``knn`` (the precompute_knns build, query-sharded at N > 1; ``parity_checked`` = the timed
result against the oracle on sampled rows and, at N > 1, against a one-GPU build) and ``probes`` (fused probe losses vs
the trainer's torch op sequence).

Sampling schedule: by default the loss is handed batch i+1's depth maps before the forward of batch i
(``queue_next_sampling``), so batch i+1's farthest-point sampling runs as extra CTAs of forward i's correlation kernel and
every timed step still computes exactly one batch's sampling inside the timed region; ``no_lookahead`` is the same loop
with the sampling at the head of each step (DEPTHG_BENCH_LOOKAHEAD=0 makes that the main value and reports the default
as ``lookahead``).

At N > 1 every step is followed by the all-reduce of the trainable-head gradient (729 012 floats)
(DEPTHG_BENCH_ALLREDUCE = symm | symm_inline | fps | inline | graph | graph_hp | async | none; symm (default) = torch's
symmetric-memory multimem / two-shot all-reduce on a side stream; with look-ahead sampling it runs free beside the next
step's gathers (DEPTHG_BENCH_AR_WAIT=gather: the next forward waits for it first), with in-step sampling underneath the
next step's FPS kernel.  Measured, ms/step at N = 1 / 2 / 8: look-ahead 0.179 / 0.220 / 0.237, in-step 0.208 / 0.2155 /
0.347).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: the contract is ONE JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # (the banner itself still prints at WARN: send it to stderr)
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "corr_loss_fwd_bwd_samples_per_s"
UNIT = "samples/s"
# cocostuff27 ViT-B/8 (paper_reproduction.sh:8): shifts and loss weights of the paper run
CFG2 = dict(B=32, C=768, D=90, H=28, W=28, Hd=224, Wd=224, S=11, neg_samples=5,
            pos_intra_shift=0.2103, pos_inter_shift=0.1233, neg_inter_shift=0.9748, depth_feat_shift=0.0359)
WEIGHTS = dict(pos_inter=1.0501, pos_intra=0.2305, neg_inter=0.2485, depth_feat=0.1603)
HEAD_GRAD_FLOATS = 729_012   # trainable 1x1-conv head of ViT-B / dim 90: the DDP all-reduce payload
KNN = dict(N=49_629, F=768, k=30)


def make_cfg(S=CFG2["S"]):
    from types import SimpleNamespace
    return SimpleNamespace(feature_samples=S, use_salience=False, depth_sampling="fps", fps_gpu=False, pointwise=True,
                           zero_clamp=True, stabalize=False, neg_samples=CFG2["neg_samples"],
                           pos_intra_shift=CFG2["pos_intra_shift"], pos_inter_shift=CFG2["pos_inter_shift"],
                           neg_inter_shift=CFG2["neg_inter_shift"], depth_feat_correlation_loss=True,
                           depth_feat_shift=CFG2["depth_feat_shift"])


def algorithmic_bytes(B, C=CFG2["C"], D=CFG2["D"], HW=784, Hd=224, Wd=224):
    """inputs read once, code gradients written once, no 5-D tensors."""
    return 4 * (2 * B * C * HW + 2 * B * D * HW + 2 * B * D * HW + 2 * B * Hd * Wd)


def algorithmic_flops(B, S=CFG2["S"], C=CFG2["C"], D=CFG2["D"], N=CFG2["neg_samples"]):
    P = S * S
    return (N + 2) * B * P * P * 2 * (C + 3 * D)


def weighted(out):
    return (WEIGHTS["pos_intra"] * out[0] + WEIGHTS["pos_inter"] * out[2] + WEIGHTS["neg_inter"] * out[4].mean()
            + WEIGHTS["depth_feat"] * out[6])


_GRADW = {}


def backprop(out):
    """d/d(code) of the trainer's weighted sum (train_segmentation.py:331-334): the four scalar losses are
    backpropagated with their weights as grad_tensors, i.e. exactly weighted(out).backward() without the
    seven scalar mul/add kernels of building that sum."""
    dev = out[0].device
    if dev not in _GRADW:
        _GRADW[dev] = [torch.tensor(WEIGHTS[k], device=dev) for k in ("pos_intra", "pos_inter", "neg_inter",
                                                                       "depth_feat")]
    neg = out[4] if out[4].dim() == 0 else out[4].mean()
    # single-threaded engine: the graph is one node deep, handing it to the device worker thread costs more
    # (~80 us) than running it inline
    with torch.autograd.set_multithreading_enabled(False):
        torch.autograd.backward([out[0], out[2], neg, out[6]], grad_tensors=_GRADW[dev])


def synth_inputs(B, gen, device, channels_last=True):
    """Synthetic batch of the cfg2 shape.  Features arrive channels-last on the live path
    (permuted [B,HW,C] views, SURVEY.md 7 hard part 6); depth is integer-valued like the uint8 PNGs."""
    c = CFG2

    def feat(ch):
        x = torch.randn((B, c["H"], c["W"], ch), generator=gen, device=device)
        return x.permute(0, 3, 1, 2) if channels_last else x.permute(0, 3, 1, 2).contiguous()

    depth = lambda: torch.randint(0, 256, (B, 1, c["Hd"], c["Wd"]), generator=gen, device=device).float()  # noqa: E731
    return dict(feats=feat(c["C"]), feats_pos=feat(c["C"]), code=feat(c["D"]), code_pos=feat(c["D"]), depth=depth(),
                depth_pos=depth())


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    _NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
              0x80: "hw_power_brake", 0x2: "applications_clocks", 0x10: "sync_boost", 0x100: "display_clocks"}

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ----------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """The reference algorithm (oracle port: torch-CPU fp32 + NumPy FPS, the reference's own
    libraries) on the host cores.  /root/reference is not on the GPU box, so the committed
    restatement (pinned bit-for-bit to the real reference, tests/test_oracle.py) is what runs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import depthg_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = make_cfg(CFG2["S"])
    gen = torch.Generator().manual_seed(0)

    def one_step(inp):
        code = inp["code"].detach().requires_grad_(True)
        code_pos = inp["code_pos"].detach().requires_grad_(True)
        out = O.ContrastiveCorrelationLoss(cfg)(inp["feats"], inp["feats_pos"], None, None, code, code_pos,
                                                inp["depth"], inp["depth_pos"])
        backprop(out)
        return out[0].item()

    B = CFG2["B"]
    inp = synth_inputs(B, gen, "cpu")
    # Always the full cfg2 batch (the config the GPU arm runs per GPU).  One CPU process runs it whatever --gpus says:
    # the printed config names the batch that actually ran, never B x N.  A step is ~0.4 s on 16 cores, so the
    # requested steps fit in a few minutes; on a much slower host the STEP COUNT is cut, not the batch.
    one_step(inp)                                   # untimed: thread pool, allocator, first-touch
    t0 = time.perf_counter()
    one_step(inp)
    first = time.perf_counter() - t0
    budget = 240.0
    steps, warmup = args.steps, args.warmup
    if first * (steps + warmup) > budget:
        warmup = 1
        steps = max(3, int(budget / first) - warmup)
    for _ in range(warmup):
        one_step(inp)
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step(inp)
    dt = time.perf_counter() - t0
    value = B * steps / dt
    cfgd = workload_config(B, 1, extra={"device": "cpu", "parallelism": "one host process, %d threads" % cores,
                                        "gpus_requested": args.gpus})
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfgd,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{steps} steps of {B} samples of the cfg2 shape (fwd+bwd incl. NumPy FPS)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(B, n_gpus, extra=None):
    c = {"workload": "cocostuff27 ViT-B/8 ContrastiveCorrelationLoss fwd+bwd (BASELINE configs[1])",
         "per_gpu_batch": B, "global_batch": B * n_gpus, "C": CFG2["C"], "dim": CFG2["D"], "grid": "28x28",
         "feature_samples": CFG2["S"], "neg_samples": CFG2["neg_samples"], "depth_sampling": "fps", "pointwise": True,
         "depth_term": True, "layout": "channels_last (live trainer layout)", "parallelism": f"dp{n_gpus}"}
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------- our arm
def knn_cpu_baseline(N, F, k, chunks=2):
    """The reference's KNN loop (src/precompute_knns.py:99-113, oracle port) on the host cores, on a bounded sample:
    `chunks` of its N // 64 query-row chunks against the full database."""
    from oracle import depthg_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(7)
    db = torch.nn.functional.normalize(torch.randn((N, F), generator=g), dim=1)
    step = N // 64
    rows = min(N, chunks * step)
    O.knn_rows(db[:step // 8], db, k)                         # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    for i in range(0, rows, step):
        O.knn_rows(db[i:i + step], db, k)
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": f"{rows} of {N} query rows ({chunks} of the reference's 64 chunks) against the full database",
            "seconds": dt, "full_build_s_extrapolated": dt * N / rows}


def probe_bench(dev, B, n=30):
    """Linear-probe CE and cluster-probe loss fwd+bwd at the cfg2 shapes (D=90, 28x28 code, 27 classes,
    224x224 labels): depthg_b200.probes vs the trainer's torch op sequence (src/train_segmentation.py:419-441)."""
    import torch.nn.functional as F
    from depthg_b200.probes import ClusterLookup, linear_probe_loss
    D, K, h, Hl = CFG2["D"], 27, 28, 224
    g = torch.Generator(device=dev).manual_seed(7)
    code = torch.randn(B, D, h, h, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    label = torch.randint(-1, K, (B, Hl, Hl), device=dev, generator=g)
    weight = (torch.randn(K, D, 1, 1, device=dev, generator=g) / D ** 0.5).requires_grad_(True)
    bias = torch.zeros(K, device=dev, requires_grad=True)
    probe = ClusterLookup(D, K).to(dev)

    def fused_linear():
        weight.grad = bias.grad = None
        linear_probe_loss(code, weight, bias, label).backward()

    def torch_linear():
        weight.grad = bias.grad = None
        flat = label.reshape(-1)
        mask = (flat >= 0) & (flat < K)
        lg = F.interpolate(F.conv2d(torch.clone(code.detach()), weight, bias), label.shape[-2:], mode="bilinear",
                           align_corners=False)
        lg = lg.permute(0, 2, 3, 1).reshape(-1, K)
        F.cross_entropy(lg[mask], flat[mask]).mean().backward()

    def fused_cluster():
        probe.clusters.grad = None
        probe(code, None)[0].backward()

    def torch_cluster():
        probe.clusters.grad = None
        ip = torch.einsum("bchw,nc->bnhw", F.normalize(code, dim=1), F.normalize(probe.clusters, dim=1))
        cp = F.one_hot(torch.argmax(ip, dim=1), K).permute(0, 3, 1, 2).to(torch.float32)
        (-(cp * ip).sum(1).mean()).backward()

    def timeit(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    out = {"shapes": {"B": B, "D": D, "classes": K, "code": "28x28", "label": "224x224"}}
    for name, fn in (("linear_probe_ce_us", fused_linear), ("linear_probe_ce_torch_ops_us", torch_linear),
                     ("cluster_probe_us", fused_cluster), ("cluster_probe_torch_ops_us", torch_cluster)):
        out[name] = round(timeit(fn), 1)
    out["note"] = "fwd+bwd per call incl. host overhead; torch_ops = the trainer's own op sequence on the same GPU"
    return out


def side_config(dev, name, B, C, D, S, sampling, pointwise, steps, peaks_):
    """Device-resident fwd+bwd of one of the OTHER BASELINE configs (cfg4 Cityscapes, cfg5 dense, the S = 12 paper
    value): ms/step, library-side per-kernel breakdown and the roofline fraction of the bound that applies."""
    import ctypes
    from depthg_b200 import _lib
    from depthg_b200.modules import ContrastiveCorrelationLoss
    hbm_peak, tf_burst, tf_sust, _ = peaks_
    cfg = make_cfg(S)
    cfg.depth_sampling, cfg.pointwise = sampling, pointwise
    g = torch.Generator(device=dev).manual_seed(99)
    H = CFG2["H"]

    def feat(ch):
        return torch.randn((B, H, H, ch), generator=g, device=dev).permute(0, 3, 1, 2)

    nsets = 2 if S < 20 else 1          # rotate inputs where one set does not already exceed L2
    sets = []
    for _ in range(nsets):
        sets.append(dict(f=feat(C), fp=feat(C), c=feat(D).requires_grad_(True), cp=feat(D).requires_grad_(True),
                         d=torch.randint(0, 256, (B, 1, 8 * H, 8 * H), generator=g, device=dev).float(),
                         dp=torch.randint(0, 256, (B, 1, 8 * H, 8 * H), generator=g, device=dev).float()))
    fn = ContrastiveCorrelationLoss(cfg)

    def step(i):
        t = sets[i % nsets]
        t["c"].grad = None
        t["cp"].grad = None
        backprop(fn(t["f"], t["fp"], None, None, t["c"], t["cp"], t["d"], t["dp"]))

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib = _lib.lib()
    lib.dg_profile_enable(1)
    reps = min(steps, 5)
    for i in range(reps):
        step(i)
    n = lib.dg_profile_collect(None, 0)
    buf = ctypes.create_string_buffer(n + 16)
    lib.dg_profile_collect(buf, n + 16)
    lib.dg_profile_enable(0)
    br = {}
    for ln in buf.value.decode().strip().split("\n"):
        if ln:
            nm, _, us = ln.split("\t")
            br[nm] = round(float(us) / reps, 2)
    P = S * S
    nb = algorithmic_bytes(B, C, D)
    fl = (CFG2["neg_samples"] + 2) * B * P * P * 2 * (C + 3 * D)
    corr_us = br.get("corr_pipe_kernel", br.get("corr_umma_kernel", 0.0))
    out = {"config": name, "B": B, "C": C, "dim": D, "S": S, "points": P, "sampling": sampling, "pointwise": pointwise,
           "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "breakdown_us": br,
           "algorithmic_bytes": nb, "algorithmic_flops": fl,
           "hbm_frac_step": nb / (ms * 1e-3) / 1e9 / hbm_peak,
           "mem_GB": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2)}
    if corr_us:
        out["corr_kernel"] = {"us": corr_us, "useful_tflops": fl / (corr_us * 1e-6) / 1e12,
                              "frac_of_bf16_sustained_div3": fl / (corr_us * 1e-6) / 1e12 / (tf_sust / 3.0),
                              "frac_of_bf16_sustained": fl / (corr_us * 1e-6) / 1e12 / tf_sust,
                              "hbm_frac": nb / (corr_us * 1e-6) / 1e9 / hbm_peak,
                              "note": "useful flops (N+2) B P^2 2 (C + 3 D) counted once; every product is issued as a "
                                      "3-term hi/lo split, so the issued tensor work is 3x this"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-knn", action="store_true", help="skip the KNN-build side measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg4 / S=12 / dense side configurations")
    ap.add_argument("--nchw", action="store_true", help="NCHW-contiguous inputs instead of channels-last")
    ap.add_argument("--feature-samples", type=int, default=CFG2["S"],
                    help="S of the workload (11 = BASELINE configs[1]; 12 = the ViT-B paper script's value)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    CFG2["S"] = args.feature_samples
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from depthg_b200 import _lib
    from depthg_b200.modules import ContrastiveCorrelationLoss
    from depthg_b200.precompute_knns import knn_topk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lib = _lib.lib()
    hbm_peak, tf_burst, tf_sust, peak_src = peaks()

    B = CFG2["B"]
    cfg = make_cfg(CFG2["S"])
    loss_fn = ContrastiveCorrelationLoss(cfg)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # rotate over input sets totalling > L2 (126 MB) so no step finds its inputs cached
    NSETS = 3
    sets = [synth_inputs(B, gen, dev, channels_last=not args.nchw) for _ in range(NSETS)]
    for s in sets:
        s["code"].requires_grad_(True)
        s["code_pos"].requires_grad_(True)
    head_grad = torch.zeros(HEAD_GRAD_FLOATS, device=dev) if world > 1 else None
    pending = [None]
    # how the per-step head-gradient all-reduce is issued: "async" = dist.all_reduce(async_op=True) (what DDP's
    # reducer does), "graph"/"graph_hp" = the same NCCL all-reduce captured once and replayed on a (high-priority)
    # side stream, "inline" = replayed on the compute stream right after the backward kernel, "none" = off.
    # Measured at N=8 (ms/step): graph 0.505, graph_hp 0.432, inline 0.344 - the all-reduce is latency-bound
    # (2.9 MB over NVSwitch) and its spinning CTAs fight the SM-filling step kernels when overlapped, so the
    # default runs it in stream order.
    # "symm" (default) = the symmetric-memory all-reduce, else "fps" = the captured NCCL one, on a side stream after backward i; step i+1's forward waits for it between its FPS
    # kernel and its gathers (loss_fn.wait_after_fps): the all-reduce runs underneath step i+1's FPS kernel (2B CTAs on
    # a 148-SM GPU) and the sampler draw, and is over before the SM-filling gathers start - overlap without contention.
    ar_mode = os.environ.get("DEPTHG_BENCH_ALLREDUCE", "symm") if world > 1 else "none"
    ar_stream = ar_graph = None
    ar_done = torch.cuda.Event()
    ar_inline = ar_mode == "inline"          # replay on the compute stream: no overlap, no SM contention
    ar_fps = ar_mode == "fps"                # replay on the sampler's side stream, under the next step's FPS
    ar_hp = ar_mode in ("graph_hp",)         # side stream with high priority
    ar_symm = None
    if ar_mode in ("symm", "symm_inline"):
        # the same all-reduce on torch's symmetric-memory kernels (NVLS multimem when the fabric has it, else two-shot
        # over peer pointers): one small kernel instead of NCCL's protocol - latency is what matters at 2.9 MB
        ar_fps = ar_mode == "symm"
        try:
            import torch.distributed._symmetric_memory as symm
            gname = dist.group.WORLD.group_name
            hg = symm.empty(HEAD_GRAD_FLOATS, dtype=torch.float32, device=dev)
            hg.zero_()
            hdl = symm.rendezvous(hg, gname)
            opname = "multimem_all_reduce_" if hdl.has_multicast_support else "two_shot_all_reduce_"
            op = getattr(torch.ops.symm_mem, opname)
            op(hg, "sum", gname)
            torch.cuda.synchronize()
            head_grad, ar_symm = hg, (lambda: op(hg, "sum", gname))
            ar_stream = torch.cuda.Stream(device=dev)
            ok = 1.0
        except Exception:  # noqa: BLE001
            ok = 0.0
        flags = torch.tensor([ok], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if flags.item() == 0.0:
            ar_symm, ar_mode, ar_fps = None, "fps", True
        else:
            ar_mode = "symm:" + opname
    if ar_mode in ("graph", "graph_hp", "inline", "fps"):
        ar_mode = "graph"
        try:
            ar_stream = torch.cuda.Stream(device=dev, priority=-1 if ar_hp else 0)
            dist.all_reduce(head_grad)                     # communicator warm-up outside capture
            torch.cuda.synchronize()
            ar_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ar_graph, stream=ar_stream):
                dist.all_reduce(head_grad)
        except Exception:  # noqa: BLE001  (capture unsupported on this stack: fall back to the plain async call)
            ar_mode, ar_stream, ar_graph = "async", None, None
        flags = torch.tensor([1.0 if ar_mode == "graph" else 0.0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)       # all ranks must agree on the mode
        if flags.item() == 0.0:
            ar_mode, ar_stream, ar_graph = "async", None, None

    # Sampling one batch ahead (default): before the forward of batch i the loss is handed batch i+1's depth maps
    # (queue_next_sampling); their farthest-point sampling then runs as extra CTAs of forward i's correlation kernel,
    # on SMs that kernel's item list leaves idle, and forward i+1 starts at its gathers.  Every timed step still
    # computes exactly one batch's sampling inside the timed region (step i computes batch i+1's).
    # DEPTHG_BENCH_LOOKAHEAD=0 turns it off; the plain step is always reported as the side key `no_lookahead`.
    # Default at every N.  At N > 1 the step also carries the head-gradient all-reduce (2.9 MB: latency and rank skew,
    # not bandwidth).  Measured, ms/step (one B200 box, symmetric-memory all-reduce):
    #              N = 1     N = 2     N = 8
    #   look-ahead 0.179     0.220     0.237    all-reduce free-running beside the next step's gathers
    #   in-step    0.208     0.2155    0.347    all-reduce under the next step's FPS kernel, forward waits for it
    # The in-step schedule couples the ranks every step (the forward waits for the all-reduce, i.e. for the slowest
    # rank); since the FPS kernel got shorter than the 8-rank all-reduce it no longer hides it.
    lookahead = [os.environ.get("DEPTHG_BENCH_LOOKAHEAD", "1") != "0"]
    # at N > 1 the head-gradient all-reduce of step i runs on a side stream; with look-ahead there is no FPS kernel at
    # the start of step i+1 to hide it under: "gather" makes step i+1 wait for it before its gathers, "free" (default)
    # lets it run beside them (it only has to land before the next all-reduce / the end of the timed region)
    ar_wait = os.environ.get("DEPTHG_BENCH_AR_WAIT", "free")

    def step(i):
        s = sets[i % NSETS]
        s["code"].grad = None
        s["code_pos"].grad = None
        if lookahead[0]:
            nx = sets[(i + 1) % NSETS]
            loss_fn.queue_next_sampling(nx["depth"], nx["depth_pos"])
        out = loss_fn(s["feats"], s["feats_pos"], None, None, s["code"], s["code_pos"], s["depth"], s["depth_pos"])
        backprop(out)
        if ar_mode == "async":   # DDP semantics: the only exchange is the head-gradient all-reduce; like DDP it is
            if pending[0] is not None:   # asynchronous and only has to land before the next optimiser step
                pending[0].wait()
            pending[0] = dist.all_reduce(head_grad, async_op=True)
        elif ar_mode == "graph" and ar_inline:
            ar_graph.replay()
        elif ar_symm is not None and ar_fps:
            ar_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(ar_stream):
                ar_symm()
                ar_done.record(ar_stream)
            if not (lookahead[0] and ar_wait == "free"):
                loss_fn.wait_after_fps = ar_done
        elif ar_symm is not None:
            ar_symm()
        elif ar_mode == "graph" and ar_fps:
            ar_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(ar_stream):
                ar_graph.replay()
                ar_done.record(ar_stream)
            if not (lookahead[0] and ar_wait == "free"):
                loss_fn.wait_after_fps = ar_done    # the next forward waits for it between its FPS and its gathers
        elif ar_mode == "graph":
            ar_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(ar_stream):
                ar_graph.replay()
        return out

    def drain():   # the last all-reduce must land inside the timed region
        if pending[0] is not None:
            pending[0].wait()
            pending[0] = None
        if ar_stream is not None:
            torch.cuda.current_stream().wait_stream(ar_stream)

    def barrier():
        if world > 1:
            drain()
            dist.barrier()
        torch.cuda.synchronize()

    # Start-up priming, untimed and before the W warm-up steps: the first ~10 steps of a fresh process are not steady
    # state (caching-allocator growth for the 151 MB arenas and the gradient / sampling buffers, lazy kernel-attribute
    # set-up, clock ramp) - measured: 0.246 ms/step over 50 steps after 5 warm-up steps against 0.179 after 10+.
    PRIME = 12
    for i in range(PRIME):
        step(i)
    barrier()
    for i in range(args.warmup):        # (step indices keep counting, so the batch the last warm-up step sampled ahead
        step(PRIME + i)                 #  IS the first timed step's batch)
    barrier()
    l0 = lib.dg_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for i in range(args.steps):
            step(PRIME + args.warmup + i)
        drain()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib.dg_kernel_launches() - l0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)

    # the other schedule as a side key: the plain step (every forward samples its own batch, FPS at the head of the
    # step) when look-ahead is the default, and vice versa
    look_default = lookahead[0]
    lookahead[0] = not look_default
    loss_fn.queue_next_sampling(None, None)
    loss_fn._presampled = None
    for i in range(args.warmup):
        step(i)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    drain()
    ev1.record()
    barrier()
    ms_plain = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_plain], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_plain = float(t.item())

    lookahead[0] = False
    loss_fn.queue_next_sampling(None, None)
    loss_fn._presampled = None
    # same step with negative_sampler="torch": neg_samples x torch.randperm replayed as a CUDA graph on a side stream
    # (the reference's exact RNG stream; more host work per step)
    loss_fn.negative_sampler = "torch"
    for i in range(args.warmup):
        step(i)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    ms_fused = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_fused], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_fused = float(t.item())
    loss_fn.negative_sampler = "fused"

    # ---- the same step captured once per input set into a CUDA graph (forward + backward + the torch.randperm
    #      sampler) and replayed: what the path does when the host is out of the way
    graphed = None
    if world == 1:
        try:
            graphs = []
            for si in range(NSETS):
                for _ in range(2):
                    step(si)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step(si)
                graphs.append(g)
            for i in range(args.warmup):
                graphs[i % NSETS].replay()
            torch.cuda.synchronize()
            ev0.record()
            for i in range(args.steps):
                graphs[i % NSETS].replay()
            ev1.record()
            torch.cuda.synchronize()
            gms = ev0.elapsed_time(ev1) / args.steps
            graphed = {"value": B / (gms / 1e3), "unit": UNIT, "ms_per_step": gms,
                       "note": "forward+backward (incl. the sampler) captured per input set with "
                               "torch.cuda.graph and replayed; same kernels, no per-step host work"}
            del graphs
        except Exception as e:  # noqa: BLE001
            graphed = {"error": repr(e)[:200]}

    # ---- per-kernel breakdown: the library brackets each of its kernels with CUDA events on the launching stream
    import ctypes
    reps = min(args.steps, 20)
    lib.dg_profile_enable(1)
    for i in range(reps):
        step(i)
    need = lib.dg_profile_collect(None, 0)
    buf = ctypes.create_string_buffer(need + 16)
    lib.dg_profile_collect(buf, need + 16)
    lib.dg_profile_enable(0)
    breakdown, calls = {}, {}
    for ln in buf.value.decode().strip().split("\n"):
        if ln:
            name, n, us = ln.split("\t")
            breakdown[name] = float(us) / reps
            calls[name] = int(n) / reps

    # Dominant kernel and its roofline.  Numerator = SURVEY.md 8(d)'s ALGORITHMIC bytes of the step (inputs read once,
    # code gradients written once: 6.35 MB per sample x the B samples one launch processes = 203.1 MB at B = 32) over
    # the dominant kernel's live CUDA-event time.  `traffic` is what that kernel really moved (ncu dram bytes, committed
    # capture), and traffic / algorithmic shows the re-read factor.  The whole-step figure is `roofline_step`.
    step_bytes = algorithmic_bytes(B)
    ours = [k for k in breakdown if k.endswith("_kernel")]
    dom = max(ours, key=breakdown.get)
    dom_us = breakdown[dom]
    achieved = step_bytes / (dom_us * 1e-6) / 1e9
    traffic = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    step_traffic = None
    if os.path.isfile(tpath):   # dram__bytes_read+write per launch from the committed ncu --set full capture
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["dram_bytes"].get(dom), tj.get("source")
        step_traffic = sum(v for k, v in tj["dram_bytes"].items() if k in breakdown)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes": step_bytes, "algorithmic_bytes_per_sample": step_bytes // B,
                "traffic_over_algorithmic": (traffic / step_bytes) if traffic else None, "traffic_source": traffic_src,
                "us_per_step": dom_us, "launches_per_step": calls.get(dom),
                "share_of_gpu_time": dom_us / max(sum(breakdown[k] for k in ours), 1e-9)}
    if dom in ("corr_umma_kernel", "corr_pipe_kernel"):
        fl = algorithmic_flops(B)
        roofline["tensor"] = {"flops": fl, "achieved_tflops": fl / (dom_us * 1e-6) / 1e12,
                              "frac_of_bf16_sustained_div3": fl / (dom_us * 1e-6) / 1e12 / (tf_sust / 3.0),
                              "note": "fd and cd run as 3-term fp16 hi/lo products (fp32-grade), the gradient GEMMs likewise"}
    roofline_step = {"bound": "hbm", "algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                     "peak": hbm_peak, "unit": "GB/s", "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                     "traffic": step_traffic,
                     "traffic_over_algorithmic": (step_traffic / step_bytes) if step_traffic else None,
                     "gpu_time_us": sum(breakdown[k] for k in ours),
                     "frac_of_gpu_time": step_bytes / (sum(breakdown[k] for k in ours) * 1e-6) / 1e9 / hbm_peak,
                     "tensor_flops": algorithmic_flops(B),
                     "tensor_frac_bf16_sustained": algorithmic_flops(B) / (ms_per_step * 1e-3) / 1e12 / tf_sust}

    # ---- end-to-end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    feat_keys = ("feats", "feats_pos", "code", "code_pos")

    def to_host(k, v):
        v = v.detach()
        if k in feat_keys and not args.nchw:
            v = v.permute(0, 2, 3, 1)          # the NHWC-contiguous base of the channels-last view
        return v.contiguous().cpu().pin_memory()

    host = {k: to_host(k, v) for k, v in sets[0].items()}
    h2d = sum(v.numel() * 4 for v in host.values())
    res_host = torch.empty(4, dtype=torch.float32).pin_memory()
    gh = [torch.empty_like(host["code"]).pin_memory(), torch.empty_like(host["code_pos"]).pin_memory()]
    d2h = res_host.numel() * 4 + sum(g.numel() * 4 for g in gh)

    def e2e_step():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        if not args.nchw:
            for k in feat_keys:
                d[k] = d[k].permute(0, 3, 1, 2)
        code = d["code"].requires_grad_(True)
        code_pos = d["code_pos"].requires_grad_(True)
        out = loss_fn(d["feats"], d["feats_pos"], None, None, code, code_pos, d["depth"], d["depth_pos"])
        backprop(out)
        res_host.copy_(torch.stack([out[0], out[2], out[4], out[6]]).detach(), non_blocking=True)
        g0, g1 = code.grad, code_pos.grad
        if not args.nchw:
            g0, g1 = g0.permute(0, 2, 3, 1), g1.permute(0, 2, 3, 1)
        gh[0].copy_(g0, non_blocking=True)
        gh[1].copy_(g1, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        e2e_step()
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ev1.record()
    barrier()
    serial_ms = ev0.elapsed_time(ev1)

    # The same end-to-end step as an input pipeline would run it: double-buffered device inputs, the H2D copy of
    # step i+1 on a copy stream under the compute of step i, the D2H of step i's results on a third stream.  Every
    # step still moves all of its inputs from pinned host memory and lands all of its results in host memory.
    cur = torch.cuda.current_stream()
    h2d_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    slots = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host.items()} for _ in range(2)]
    loaded = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    res_hosts = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
    gh2 = [[torch.empty_like(g).pin_memory() for g in gh] for _ in range(2)]

    def issue_h2d(i):
        slot = i % 2
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(consumed[slot])          # the slot's previous step has finished reading it
            for k, v in host.items():
                slots[slot][k].copy_(v, non_blocking=True)
            loaded[slot].record(h2d_stream)

    def e2e_pipelined(n):
        for ev in consumed:
            ev.record(cur)
        issue_h2d(0)
        for i in range(n):
            slot = i % 2
            if i + 1 < n:
                issue_h2d(i + 1)
            cur.wait_event(loaded[slot])
            d = dict(slots[slot])
            if not args.nchw:
                for k in feat_keys:
                    d[k] = d[k].permute(0, 3, 1, 2)
            code = d["code"].detach().requires_grad_(True)
            code_pos = d["code_pos"].detach().requires_grad_(True)
            out = loss_fn(d["feats"], d["feats_pos"], None, None, code, code_pos, d["depth"], d["depth_pos"])
            backprop(out)
            res = torch.stack([out[0], out[2], out[4], out[6]]).detach()
            g0, g1 = code.grad, code_pos.grad
            if not args.nchw:
                g0, g1 = g0.permute(0, 2, 3, 1), g1.permute(0, 2, 3, 1)
            consumed[slot].record(cur)
            d2h_stream.wait_event(consumed[slot])
            with torch.cuda.stream(d2h_stream):
                res_hosts[slot].copy_(res, non_blocking=True)
                gh2[slot][0].copy_(g0, non_blocking=True)
                gh2[slot][1].copy_(g1, non_blocking=True)
            for t in (res, g0, g1):
                t.record_stream(d2h_stream)
        torch.cuda.synchronize()                            # every step's results are in host memory

    e2e_pipelined(3)
    barrier()
    ev0.record()
    e2e_pipelined(e2e_steps)
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([e2e_ms, serial_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, serial_ms = float(t[0].item()), float(t[1].item())
    # both loops move the same bytes per step; which one wins depends on the box (PCIe topology, how many ranks share
    # the host bridge), so report the faster one as the end-to-end number and keep both
    pipelined = {"value": world * B * e2e_steps / (e2e_ms / 1e3), "ms_per_step": e2e_ms / e2e_steps}
    unpipelined = {"value": world * B * e2e_steps / (serial_ms / 1e3), "ms_per_step": serial_ms / e2e_steps}
    best = pipelined if pipelined["value"] >= unpipelined["value"] else unpipelined
    e2e = {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": best["ms_per_step"], "steps": e2e_steps,
           "mode": "pipelined" if best is pipelined else "unpipelined",
           "pipelined": pipelined, "unpipelined": unpipelined,
           "note": "pinned host buffers -> H2D -> FPS/gather/loss/backward -> D2H of 4 losses + both code gradients; "
                   "pipelined = double-buffered, the H2D of step i+1 and the D2H of step i overlap compute; "
                   "unpipelined = copies and compute serialised, one synchronize per step"}

    # ---- KNN build side metric (query-row sharded; all-gather of the database when N > 1)
    knn = None
    if not args.no_knn:
        from depthg_b200.distributed import allgather_rows, knn_shard_bounds, sharded_knn_build
        N, F, k = KNN["N"], KNN["F"], KNN["k"]
        g2 = torch.Generator(device=dev).manual_seed(7)
        allf = torch.nn.functional.normalize(torch.randn((N, F), generator=g2, device=dev), dim=1)
        lo, hi = knn_shard_bounds(N, world, rank)
        local = allf[lo:hi].contiguous()
        knn_mode = os.environ.get("DEPTHG_BENCH_KNN", "overlap")   # overlap (two-phase, default) | serial (round 1)

        def knn_step():
            if world > 1 and knn_mode == "serial":
                per = -(-N // world)
                src = local if hi - lo == per else torch.cat([local, local.new_zeros((per - (hi - lo), F))])
                gathered = torch.empty((world * per, F), device=dev)
                dist.all_gather_into_tensor(gathered, src)
                return knn_topk(local, gathered[:N], k, return_stats=True)
            return sharded_knn_build(local, N, k, return_stats=True)

        knn_step()
        barrier()
        ev0.record()
        reps_k = 2
        for _ in range(reps_k):
            kidx, kstats = knn_step()
        ev1.record()
        barrier()
        kms = ev0.elapsed_time(ev1) / reps_k
        if world > 1:
            t = torch.tensor([kms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            kms = float(t.item())
        # Parity of the TIMED result, outside the timed region: sampled rows of this rank's block against the oracle
        # (fp32 einsum + topk on the host, src/precompute_knns.py:99-113) under the 1e-6 tie rule; at N > 1 rank 0 also
        # compares the concatenated sharded result with a one-GPU build of the whole index.
        from oracle import depthg_oracle as O
        rs_k = np.random.RandomState(11 + rank)
        rows = np.unique(np.concatenate([rs_k.randint(0, hi - lo, 192), np.arange(hi - lo - 32, hi - lo)]))
        allc = allf.cpu()
        _, want = O.knn_rows(allc[lo:hi][rows], allc, k)
        got = kidx[torch.from_numpy(rows).to(dev)].cpu()
        sims_s = allc[lo:hi][rows] @ allc.T
        mism = got != want
        gap = (torch.gather(sims_s, 1, got) - torch.gather(sims_s, 1, want)).abs()
        worst = float(gap[mism].max()) if bool(mism.any()) else 0.0
        parity = {"rows_checked": int(len(rows)), "slots_differing": int(mism.sum()), "worst_fp32_gap": worst,
                  "ok": bool(worst < 1e-6), "fallback_rows": kstats["fallback_rows"],
                  "pipeline_error": kstats["pipeline_error"],
                  "against": "oracle (fp32 einsum + topk on the host) on sampled query rows of the timed result"}
        if world > 1:
            per = -(-N // world)
            pad = kidx if hi - lo == per else torch.cat([kidx, kidx.new_zeros((per - (hi - lo), k))])
            full_sharded = torch.empty((world * per, k), device=dev, dtype=kidx.dtype)
            dist.all_gather_into_tensor(full_sharded, pad)
            full_sharded = full_sharded[:N]
            flag = torch.tensor([1.0 if parity["ok"] else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            parity["ok_all_ranks"] = bool(flag.item() == 1.0)
            if rank == 0:
                one = knn_topk(allf, allf, k)
                diff = one != full_sharded
                nd = int(diff.sum())
                worst1 = 0.0
                if nd:
                    r_bad = torch.nonzero(diff.any(1)).flatten()[:4096]
                    sb = allf[r_bad] @ allf.T
                    g1 = (torch.gather(sb, 1, one[r_bad]) - torch.gather(sb, 1, full_sharded[r_bad])).abs()
                    worst1 = float(g1[diff[r_bad]].max())
                parity["sharded_vs_one_gpu"] = {"slots_differing": nd, "worst_fp32_gap": worst1, "ok": worst1 < 1e-6}
                del one
            del full_sharded
        del allf, allc, sims_s
        flops = 2.0 * N * N * F
        per_gpu_tflops = flops / (kms / 1e3) / 1e12 / world
        knn = {"metric": "knn_build_img_per_s", "value": N / (kms / 1e3), "unit": "img/s", "ms": kms, "N": N, "F": F,
               "k": k, "scaling": "strong",
               "sharding": "query rows; all-gather of the feature database" +
                           (" overlapped with the tensor pass over the local rows (two-phase build)"
                            if world > 1 and knn_mode != "serial" else ""),
               "parity_checked": parity,
               "roofline": {"bound": "tensor", "achieved": per_gpu_tflops, "peak": tf_burst,
                            "unit": "TFLOP/s per GPU", "frac": per_gpu_tflops / tf_burst, "n_gpus": world,
                            "note": "useful flops 2 N^2 F counted once over the whole build (panel split, tensor "
                                    "pass, exact fp32 re-rank), divided by N_gpus x the measured 16-bit burst peak of "
                                    "ONE GPU; the default tensor pass is one fp16 product per pair (the fast pass of the "
                                    "precision ladder) and is bound by the L2->SM operand stream (~9 TB/s aggregate), "
                                    "not by the tensor pipe; DEPTHG_B200_KNN_PASS=split3 is the 3-term bf16 pass"},
               "tensor_pass": os.environ.get("DEPTHG_B200_KNN_PASS", "fast fp16 (default)")}

    if knn is not None and rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            knn["cpu_baseline"] = knn_cpu_baseline(KNN["N"], KNN["F"], KNN["k"])
        except Exception as e:  # noqa: BLE001  (auxiliary)
            knn["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- CPU baseline (oracle port) on this box's host cores, rank 0 at N=1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import depthg_oracle as O
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cin = {k: v.detach().cpu().contiguous() for k, v in sets[0].items()}
        times = []
        for _ in range(3):
            code = cin["code"].clone().requires_grad_(True)
            code_pos = cin["code_pos"].clone().requires_grad_(True)
            t0 = time.perf_counter()
            out = O.ContrastiveCorrelationLoss(cfg)(cin["feats"], cin["feats_pos"], None, None, code, code_pos,
                                                    cin["depth"], cin["depth_pos"])
            backprop(out)
            times.append(time.perf_counter() - t0)
            if sum(times) > 25:
                break
        cpu_baseline = {"value": B / min(times), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{len(times)} full steps of the same cfg2 batch (32 samples), best time; "
                                  f"torch-CPU fp32 + NumPy FPS as the reference runs it",
                        "ms_per_step": min(times) * 1e3}

    # ---- the reference's own ops run eagerly on this GPU (oracle port with CUDA tensors; FPS stays NumPy-on-host with a
    #      device->host sync per image, as the reference ships it): the GPU-vs-GPU comparison of SURVEY 8(d)
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import depthg_oracle as O
        gin = {k: v.detach() for k, v in sets[0].items()}
        times = []
        try:   # an auxiliary comparison: it must never cost the bench line
            for _ in range(3):
                code = gin["code"].clone().requires_grad_(True)
                code_pos = gin["code_pos"].clone().requires_grad_(True)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = O.ContrastiveCorrelationLoss(cfg)(gin["feats"], gin["feats_pos"], None, None, code, code_pos,
                                                        gin["depth"], gin["depth_pos"])
                backprop(out)
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
            ref_gpu = {"value": B / min(times), "unit": UNIT, "ms_per_step": min(times) * 1e3,
                       "note": "reference algorithm as PyTorch eager ops on the same B200, FPS on the host in NumPy as shipped"}
        except Exception as e:  # noqa: BLE001
            ref_gpu = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- probe losses: fused kernels vs the reference's stock-torch op sequence on this GPU
    probes = None
    if rank == 0 and world == 1 and not args.no_knn:
        try:
            probes = probe_bench(dev, B)
        except Exception as e:  # noqa: BLE001  (auxiliary: must never cost the bench line)
            probes = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- the other BASELINE configs as side keys (rank 0 at N = 1): configs[3] Cityscapes B = 64 / dim 100 / random
    #      coordinates / pointwise off; the paper script's S = 12; configs[4] dense 784 x 784 at B = 64
    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        extra = {}
        for nm, kw in (("cfg4_cityscapes", dict(B=64, C=768, D=100, S=11, sampling="none", pointwise=False, steps=20)),
                       ("cfg2_s12", dict(B=32, C=768, D=90, S=12, sampling="fps", pointwise=True, steps=20)),
                       ("cfg5_dense", dict(B=64, C=768, D=90, S=28, sampling="fps", pointwise=True, steps=5))):
            try:
                torch.cuda.empty_cache()
                extra[nm] = side_config(dev, nm, peaks_=(hbm_peak, tf_burst, tf_sust, peak_src), **kw)
            except Exception as e:  # noqa: BLE001  (auxiliary: must never cost the bench line)
                extra[nm] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(B, world, extra={
                    "l2_policy": f"rotating {NSETS} input sets ({NSETS * step_bytes / 1e6:.0f} MB) > 126 MB L2",
                    "priming_steps": PRIME,   # untimed start-up steps before the W warm-up steps
                    "negative_sampler": "fused (module default: drawn by one CTA of the sampling launch)",
                    "sampling_schedule": ("one batch ahead: batch i+1's FPS / depth signs / permutations ride as extra "
                                          "CTAs of forward i's correlation kernel (queue_next_sampling); every timed "
                                          "step computes one batch's sampling" if look_default else
                                          "in step: FPS is the first launch of every forward"),
                    "allreduce_floats_per_step": HEAD_GRAD_FLOATS if ar_mode != "none" else 0,
                    "allreduce_issue": ar_mode + ("_inline" if ar_inline else "_hp" if ar_hp else
                                                  ("_beside_next_gathers" if (look_default and ar_wait == "free")
                                                   else "_under_next_fps") if ar_fps else ""),
                    "layout": "nchw" if args.nchw else "channels_last (live trainer layout)"}),
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "launches_per_step": launches / args.steps, "roofline": roofline, "roofline_step": roofline_step,
                "breakdown_us": {k: round(v, 2) for k, v in breakdown.items()}, "cpu_baseline": cpu_baseline,
                ("no_lookahead" if look_default else "lookahead"): {
                    "value": world * B * args.steps / (ms_plain / 1e3), "unit": UNIT, "ms_per_step": ms_plain / args.steps,
                    "note": ("the same step with the sampling inside it (FPS first, then the gathers): what a caller "
                             "gets without queue_next_sampling / prefetch_sampling" if look_default else
                             "the same step with the sampling one batch ahead (queue_next_sampling): bench.py's "
                             "default schedule (this run was started with DEPTHG_BENCH_LOOKAHEAD=0)")},
                "cuda_graph": graphed, "reference_ops_on_gpu": ref_gpu,
                "torch_negative_sampler": {"value": world * B * args.steps / (ms_fused / 1e3), "unit": UNIT,
                                           "ms_per_step": ms_fused / args.steps,
                                           "note": "negative_sampler='torch': neg_samples x torch.randperm (the reference's "
                                                   "exact RNG stream, replayed as a CUDA graph on a side stream) instead of "
                                                   "the default one-launch dg_super_perms (same distribution)"},
                "knn": knn, "probes": probes, "extra_configs": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        # a captured NCCL graph can deadlock communicator teardown; everything is measured and printed, so leave
        # without the destructor chain (torchrun only needs exit code 0)
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
