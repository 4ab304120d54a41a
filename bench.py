"""Benchmark of the EEG-CLIP train step (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm's CPU port (oracle/) on the host cores

A step is the reference hot loop body (train_clip_final.py:484-492): forward of both towers, symmetric InfoNCE,
zero_grad, backward, AdamW -- train mode (dropout on), default towers (EEGConformerInterleaved depth 10, convLSTM
speech tower), 64-channel 64 Hz 5 s windows (T=320), wav2vec2-shaped speech features (1024), batch 256 per GPU.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "EEG-speech CLIP train samples/sec"
CONV_MATH_NOTE = "bf16x3 executes 3 MMAs per algorithmic MAC: ceiling of frac is 1/3"
UNIT = "samples/s"
T_WIN, F_SPEECH, DEPTH = 320, 1024, 10
CONV_FLOP_PER_SAMPLE = 2 * T_WIN * 64 * 64 * 64          # one Conv1d(64,64,k=64) pass over one window (fwd == dgrad == wgrad)
CONV32_FLOP_PER_SAMPLE = 2 * T_WIN * 64 * 64 * 32        # the speech tower's BasicBlock(k=32) conv (clip_model.py:285-291)
# class-0 launches per step: fwd + dgrad of the 10 EEG BasicBlocks (k=64) and of the speech tower's one BasicBlock (k=32)
CONV_CLASS0_FLOP_PER_SAMPLE_STEP = 2 * DEPTH * CONV_FLOP_PER_SAMPLE + 2 * CONV32_FLOP_PER_SAMPLE
# whole-step algorithmic work per sample: EEG tower fwd+bwd 6769.5 MFLOP + convLSTM speech tower 771 MFLOP
STEP_MFLOP_PER_SAMPLE = 6769.5 + 771.0
VLAAI_FWDBWD_FLOP = 98758e6                               # per sample


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=256, help="windows per GPU")
    p.add_argument("--cpu_batch", type=int, default=64, help="windows per CPU-baseline step (BASELINE config 1: batch 64)")
    p.add_argument("--no_extras", action="store_true", help="skip the head / scoring / VLAAI / config-3 side measurements")
    p.add_argument("--math", type=str, default=None, choices=["fp32", "bf16x3", "bf16"])
    p.add_argument("--no_cpu_baseline", action="store_true")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        # started before the warm-up steps (nvidia-smi takes ~0.2 s to come up); only the samples that arrived between
        # mark_begin() and mark_end() -- the timed region -- count (a sample is printed up to one period after it was taken)
        lo, hi = (self.t0 or 0.0), (self.t1 or time.time()) + 0.02
        self.rows = [r for t, r in self.rows if lo <= t <= hi]
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = max([int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU baseline: the reference algorithm restated in oracle/ (the reference itself is Python and is not on the GPU box)
# ---------------------------------------------------------------------------------------------------
def cpu_port_step_fn(batch, train=True):
    from oracle import eegclip_oracle as O, synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {}
    sd.update(synth.make_state_dict(synth.interleaved_shapes(DEPTH, T_WIN), 1, "eegModel."))
    sd.update(synth.make_state_dict(synth.conv_lstm_shapes(T_WIN), 2, "speechModel."))
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    tau = torch.tensor(0.075, requires_grad=True)
    tau_e = torch.tensor(0.075, requires_grad=True)
    params = list(sd.values()) + [tau, tau_e]
    opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01)
    mem = torch.rand(10001, T_WIN * 8)
    drop = O.Drop(train=train, seed=0, native=True)  # torch's own dropout: the reference's actual cost (72 % bernoulli_)
    eeg, sp = synth.randn(3, batch, T_WIN, 64), synth.randn(4, batch, T_WIN, F_SPEECH)
    ids = torch.arange(1, batch + 1)

    def step():
        ef = O.eeg_conformer_interleaved(sd, eeg, DEPTH, drop, pre="eegModel.")
        sf = O.eeg_conv_lstm(sd, sp, drop=drop, pre="speechModel.", fast_lstm=True)
        l_ce, l_avg, l_tot = O.clip_sim_no_latent_proj(ef, sf, ids, mem, tau, tau_e, 1.0, 0.0)
        opt.zero_grad()
        l_tot.backward()
        opt.step()
        return float(l_ce.detach())
    return step


def time_cpu_port(batch, steps, warmup, train=True):
    step = cpu_port_step_fn(batch, train)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, ms = time_cpu_port(args.cpu_batch, args.steps, max(args.warmup, 1))
    cores = os.cpu_count() or 1
    sample = (f"{args.cpu_batch} windows per step (BASELINE config 1's batch; a bounded sample of the batch-{args.batch} workload: the "
              "per-window cost of the CPU path does not depend on the batch), train mode, fwd+bwd+AdamW")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "reference_arm": "CPU port of the reference algorithm (oracle/), "
                   "torch CPU ops, all host threads; the Python reference itself is not present on the GPU box"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


def workload_name(batch):
    return (f"full CLIP train step (fwd+bwd+AdamW, train mode): EEGConformerInterleaved depth {DEPTH} + convLSTM speech tower + "
            f"CLIPSimNoLatentProj, 64-ch 64 Hz 5 s windows (T={T_WIN}), wav2vec2-shaped speech ({F_SPEECH}), batch {batch} per GPU")


# ---------------------------------------------------------------------------------------------------
# side measurements carried in the same JSON line (BASELINE metric's "fused-loss % of roofline", configs 3-5)
# ---------------------------------------------------------------------------------------------------
def _timed_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure_head(dev, pk, B=4096, D=2560):
    """Fused contrastive head at BASELINE config 3's size on ONE GPU: l2-normalise + S.E^T.e^tau + row/column CE + backward
    (all gradients incl. normalisation), CUDA events over 5 passes; algorithmic work 6.B^2.D (fwd 2, bwd 4)."""
    from transformer_clip_eeg_b200.parallel import infonce_loss
    g = torch.Generator().manual_seed(1)
    E = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
    S = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
    tau = torch.tensor(0.075, device=dev, requires_grad=True)

    def fb():
        E.grad = S.grad = tau.grad = None
        infonce_loss(E, S, tau).backward()
    ms = _timed_ms(fb)
    tf = 6.0 * B * B * D / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "symmetric InfoNCE head fwd+bwd (eegclip_l2norm_* + eegclip_infonce_lse/_loss/_backward), one GPU",
            "B": B, "D": D, "ms": ms, "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_sustained"],
            "algorithmic_flop": 6.0 * B * B * D}


def measure_scoring(dev, pk, N=4096, D=2560):
    """BASELINE config 4 on one GPU: candidate row-dots (HBM-bound) and bank similarity + top-100 (tensor-bound GEMM)."""
    from transformer_clip_eeg_b200 import train_clip_helper_functions as H
    g = torch.Generator().manual_seed(2)
    E = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1).to(dev)
    out = {"N": N, "D": D, "rowdots": {}, "bank": {}}
    for K in (2, 5, 100):
        C = torch.randn(N, K, D, device=dev)
        ms = _timed_ms(lambda: H.mm_scores(E, C), reps=3, warm=1)
        gbs = N * K * D * 4 / (ms * 1e-3) / 1e9
        out["rowdots"][f"K={K}"] = {"ms": ms, "GB/s": gbs, "frac_hbm": gbs / pk["hbm_gbs"]}
        del C
    for M in (1000, 10000, 100000):
        Bk = torch.randn(M, D, device=dev)
        ms = _timed_ms(lambda: H.bank_logits(E, Bk), reps=3, warm=1)
        ms_k = _timed_ms(lambda: H.bank_topk(E, Bk, 100), reps=3, warm=1)
        out["bank"][f"M={M}"] = {"logits_ms": ms, "TFLOP/s": 2.0 * N * M * D / (ms * 1e-3) / 1e12, "with_top100_ms": ms_k,
                                 "top100_over_gemm": ms_k / ms}
        del Bk
    return out


def measure_vlaai(dev, pk):
    """BASELINE config 5: VLAAI forward+backward on synthetic EEG, one GPU."""
    from transformer_clip_eeg_b200 import vlaai
    torch.manual_seed(0)
    model = vlaai.VLAAI().to(dev).train()
    out = {}
    for B in (64, 256):
        x = torch.randn(B, 320, 64, device=dev)

        def fb():
            model.zero_grad(set_to_none=True)
            model(x).sum().backward()
        ms = _timed_ms(fb, reps=3, warm=2)
        tf = B * VLAAI_FWDBWD_FLOP / (ms * 1e-3) / 1e12
        out[f"B={B}"] = {"ms_fwd_bwd": ms, "samples/s": B / (ms * 1e-3), "TFLOP/s": tf, "frac_tensor": tf / pk["bf16_sustained"]}
    return out


# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    import transformer_clip_eeg_b200 as pkg
    from transformer_clip_eeg_b200 import _lib, train_clip_final as tcf
    from transformer_clip_eeg_b200.optim import AdamW
    from transformer_clip_eeg_b200.parallel import bind_to_gpu_numa_node, broadcast_parameters

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback; use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    if args.math:
        _lib.set_default_math(args.math)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None   # before the pinned staging buffers are allocated
    lib = _lib.load()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                              # (nvidia-smi is up long before the timed region; only that region's samples count)
    B = args.batch
    torch.manual_seed(0)
    cli = tcf.build_parser().parse_args([])          # reference defaults: depth 10, convLSTM, latent 8, tau 0.075 ...
    model = tcf.build_model(cli, T_WIN, 10000, dev)
    model.shard_group = group
    broadcast_parameters(model, group)
    opt = AdamW(model.parameters(), betas=(cli.beta1, cli.beta2), weight_decay=cli.weight_decay, lr=cli.learning_rate)
    model.train()

    # synthetic inputs: NBUF distinct batches in pinned host memory (+ resident device copies for the kernel-side number)
    NBUF = 2
    g = torch.Generator().manual_seed(rank)
    host = [(torch.randn(B, T_WIN, 64, generator=g).pin_memory(), torch.randn(B, T_WIN, F_SPEECH, generator=g).pin_memory(),
             (torch.randperm(10000, generator=g)[:B] + 1).pin_memory()) for _ in range(NBUF)]
    resident = [tuple(t.to(dev) for t in h) for h in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident(i):
        eeg, sp, ids = resident[i % NBUF]
        return tcf.train_step(model, opt, eeg, sp, ids, group=group)

    class _HostBatches:
        """the pinned host batches, in the reference's batch-tuple form, for the package's input pipeline"""
        def __init__(self, n):
            self.n = n

        def __iter__(self):
            for i in range(self.n):
                e, s_, ids_ = host[i % NBUF]
                yield e, [s_], ids_, None

    def run_e2e(steps, timed_from=None):
        # public API end to end: DevicePrefetcher (pinned H2D on a copy stream, double-buffered) -> train_step -> loss.item()
        # timed_from = W: ONE continuous run of W + K steps; the K timed steps start once the pipeline is in steady state (the first
        # batch of a run has nothing to hide its H2D copy under: at K = 10 that fill alone was 0.6 ms per timed step).  Every timed
        # step still has its own H2D copy (of the next batch, on the copy stream) and its own loss read inside the region.
        reader = tcf.LossReader(dev)
        got = []
        e0 = e1 = None
        for k, (eeg, sp, ids) in enumerate(tcf.DevicePrefetcher(_HostBatches(steps), dev)):
            if timed_from is not None and k == timed_from:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            loss_ce, _, _ = tcf.train_step(model, opt, eeg, sp, ids, group=group)
            got += reader.push(loss_ce)              # device -> host read of every step's loss (pinned, read one step later)
        if e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
        got += reader.drain()
        assert len(got) == steps and all(v == v for v in got), got
        if e0 is None:
            return got[-1]
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / (steps - timed_from)

    def timed(fn, steps, whole=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        if whole:
            fn(steps)
        else:
            for i in range(steps):
                fn(i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    l0 = lib.eegclip_launch_count()
    # timed region: CUDA events around the roofline kernel's launches only (class 0); events around every launch would sit
    # between all kernels and switch off programmatic dependent launch for the whole step
    _lib.call("eegclip_tune_set", 8, 1)
    _lib.call("eegclip_profile_begin")
    sampler.mark_begin()
    ms_step = timed(step_resident, args.steps)
    sampler.mark_end()
    prof_ms = (ctypes.c_double * 12)()
    prof_n = (ctypes.c_longlong * 12)()
    _lib.call("eegclip_profile_end", ctypes.cast(prof_ms, ctypes.c_void_p), ctypes.cast(prof_n, ctypes.c_void_p), 12)
    launches = lib.eegclip_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    conv_ms_total, conv_launches_timed = prof_ms[0], prof_n[0]
    # per-class breakdown: a second pass of the same K steps with events around every launch (not part of `value`)
    # (single stream: with the speech tower on its side stream the per-launch events of the two towers would overlap)
    _lib.call("eegclip_tune_set", 8, 0)
    two_streams = os.environ.get("EEGCLIP_TWO_STREAMS", "1")
    os.environ["EEGCLIP_TWO_STREAMS"] = "0"
    step_resident(0)
    _lib.call("eegclip_profile_begin")
    ms_step_profiled = timed(step_resident, args.steps)
    _lib.call("eegclip_profile_end", ctypes.cast(prof_ms, ctypes.c_void_p), ctypes.cast(prof_n, ctypes.c_void_p), 12)
    os.environ["EEGCLIP_TWO_STREAMS"] = two_streams
    run_e2e(max(args.warmup, 3) + 3)                 # warm-up: also lets the copy stream's allocator pool reach steady state
    sync_all()
    ms_e2e = run_e2e(max(args.warmup, 3) + args.steps, timed_from=max(args.warmup, 3))

    pk = peaks()
    extras = None
    if not args.no_extras:
        extras = {}
        if world > 1 and B != 512:
            # BASELINE config 3 proper: 512 windows per rank (global 4096 at 8 GPUs), timed beside the 256-per-rank weak-scaling number
            B3 = 512
            g3 = torch.Generator().manual_seed(100 + rank)
            batch3 = (torch.randn(B3, T_WIN, 64, generator=g3).to(dev), torch.randn(B3, T_WIN, F_SPEECH, generator=g3).to(dev),
                      (torch.randperm(10000, generator=g3)[:B3] + 1).to(dev))
            fn3 = lambda i: tcf.train_step(model, opt, *batch3, group=group)
            for i in range(3):
                fn3(i)
            ms3 = timed(fn3, max(3, args.steps // 2))
            extras["config3_512_per_rank"] = {"global_batch": world * B3, "ms_per_step": ms3, "value": world * B3 / (ms3 * 1e-3), "unit": UNIT}
            del batch3
        if world == 1:
            torch.cuda.empty_cache()
            extras["roofline_head"] = measure_head(dev, pk)
            extras["scoring"] = measure_scoring(dev, pk)
            extras["vlaai"] = measure_vlaai(dev, pk)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # class 5 holds the head's tcgen05 similarity launches (and the exact-fp32 GEMMs, of which the bf16x3 step launches none)
    names = ["conv_fwd_dgrad_tc", "conv_wgrad_tc", "attn_fwd", "attn_bwd", "ln_ct", "head_similarity_tc", "lin_tc", "lin_wgrad_tc", "lstm_recurrence"]
    kern = {n: {"ms_per_step": prof_ms[i] / args.steps, "launches_per_step": prof_n[i] / args.steps} for i, n in enumerate(names)}
    conv_launches = max(1, conv_launches_timed)
    conv_ms = conv_ms_total / conv_launches
    # class 0 = 20 Conv1d(k=64) fwd/dgrad launches + 2 launches of the speech tower's k=32 conv per step: credit each its own FLOPs
    conv_flop_timed = args.steps * B * CONV_CLASS0_FLOP_PER_SAMPLE_STEP
    achieved = conv_flop_timed / (conv_ms_total * 1e-3) / 1e12 if conv_launches_timed else 0.0
    step_flop = B * STEP_MFLOP_PER_SAMPLE * 1e6 + 6.0 * B * (world * B) * (T_WIN * 8)
    step_tf = step_flop / (ms_step * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv_tc_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    math_name = {0: "f32", 1: "bf16x3-fp32acc", 2: "bf16-fp32acc"}[_lib.default_math()]
    line = {
        "metric": METRIC, "value": world * B / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": math_name, "data": "synthetic",
        "config": {"workload": workload_name(B), "global_batch": world * B,
                   "parallelism": f"dp{world}: per-rank towers, NCCL all-gather of embeddings, sharded InfoNCE, SUM all-reduce of grads",
                   "numa": f"rank 0 bound to NUMA node {numa_node} of its GPU (every rank binds to its own GPU's node)" if numa_node is not None else "no NUMA binding",
                   "l2": f"per-step inputs ({h2d_bytes / 1e6:.0f} MB) and activations (>2 GB) exceed the 126 MB L2; {NBUF} batches rotate",
                   "speech_tower": "1x1 conv, BasicBlock(k=32) and both bi-LSTMs (input GEMMs + recurrence kernels) on eegclip kernels",
                   "instrumentation": "the timed region of `value` records CUDA events around the roofline kernel's 22 launches per step "
                                      "(programmatic dependent launch overlap ends at an event: ~0.2 ms per step); the e2e region carries none",
                   "streams": ("speech tower on a side stream next to the EEG tower, forward and backward (clip_model.run_towers)"
                               if two_streams != "0" else "both towers on one stream (EEGCLIP_TWO_STREAMS=0)")},
        "clocks": clocks,
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4,
                "pipeline": "DevicePrefetcher: pinned double-buffered H2D on a copy stream overlapping the previous step; every step's loss is "
                            "copied to pinned host memory and read on the host one step later (LossReader), all inside the timed region; "
                            "the K timed steps are the last K of one continuous run of W + K steps (pipeline in steady state)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "conv64_tc_kernel (Conv1d k=64 forward + data-gradient launches)",
                     "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                     "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
                     "algorithmic_flop_per_launch": conv_flop_timed / conv_launches, "avg_launch_ms": conv_ms,
                     "launches_timed": int(conv_launches_timed),
                     "frac_single_stream": (args.steps * B * CONV_CLASS0_FLOP_PER_SAMPLE_STEP / (prof_ms[0] * 1e-3) / 1e12 / pk["bf16_sustained"])
                                           if prof_ms[0] > 0 else None,
                     "note": "per step 20 launches of the k=64 conv (B x 167.77 MFLOP each) + 2 of the speech tower's k=32 conv (B x 83.89 MFLOP "
                             "each), each credited its own FLOPs; " + CONV_MATH_NOTE +
                             "; in the timed region the speech tower runs on a side stream, so a conv launch shares the SMs with the other "
                             "tower's kernels for part of its duration (frac_single_stream: the same launches in the single-stream "
                             "per-class pass)",
                     "traffic": traffic},
        "roofline_whole_step": {"bound": "tensor", "achieved": step_tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                                "frac": step_tf / pk["bf16_sustained"], "algorithmic_flop_per_step": step_flop,
                                "note": "per sample 6769.5 MFLOP (EEG tower fwd+bwd) + 771 MFLOP (convLSTM speech tower) + head 6.b.B.D, "
                                        "divided by ms_per_step; the step is a chain of HBM-bound token kernels around the tensor-bound conv"},
        "kernels": kern,
        "kernels_note": f"per-class CUDA-event times from a second pass of the same {args.steps} steps with events around every launch "
                        f"and both towers on one stream ({ms_step_profiled:.2f} ms/step: the events serialise the launches); the timed "
                        "region records events around the roofline kernel's launches only",
    }
    if extras is not None:
        line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        val, ms = time_cpu_port(args.cpu_batch, 2, 1)
        val_e, ms_e = time_cpu_port(args.cpu_batch, 1, 1, train=False)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": f"{args.cpu_batch} windows per step (BASELINE config 1's batch), 2 timed steps after 1 warm-up "
                                          f"({ms:.0f} ms/step), train mode (torch's own dropout), same model/step as the GPU arm, torch CPU "
                                          "ops on all host threads",
                                "value_eval_mode": val_e, "ms_per_step_eval_mode": ms_e}
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """Keep stdout clean for the ONE JSON line: libraries (NCCL's version banner, cuDNN warnings) write to fd 1 too."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


if __name__ == "__main__":
    a = parse()
    _OUT = _claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
