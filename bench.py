#!/usr/bin/env python
"""bench.py -- volumes/sec of the MicFormer dual-stream hot path (fwd + MDiceLoss + bwd + optimizer step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config train|w7] [--batch B]

N>1 is launched by the driver with torch.distributed.run (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.
Workload (BASELINE.json configs[1]): MicFormer train config Head(embed_dim=48, num_classes=8) -- what
MicFormer/train_mmwhs_noPad.py:92 builds -- batch 2 per GPU of synthetic 2x(1,128,128,128) CT/MR volumes with
random one-hot labels, fp32.  `value` = device-timed throughput with inputs resident in HBM; `e2e` = the same
step through the public nn.Module API with pinned-host inputs copied in and the loss read back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="train", choices=["train", "w7"])
    ap.add_argument("--batch", type=int, default=2, help="volumes per GPU per step")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--eval-mode", action="store_true", help="disable DropPath (default: train mode like the script)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-pass", action="store_true")
    ap.add_argument("--no-attn-isolation", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the window-7 step and the eager-CUDA context block")
    ap.add_argument("--arena", type=int, default=int(os.environ.get("MICFORMER_GRAD_ARENA", "-1")),
                    help="1: gradients accumulate directly into one flat buffer (micformer_b200.arena.GradArena) and the "
                         "all-reduce runs in place on it; default: on for N > 1 (measured neutral at N = 1)")
    ap.add_argument("--gemm-mode", type=int, default=int(os.environ.get("MICFORMER_GEMM_MODE", "1")),
                    help="1 (default): tcgen05 TF32 GEMMs/convs, logits within 1e-3 of the fp32 CPU path; 0: exact fp32")
    ap.add_argument("--profile-step", action="store_true",
                    help="warm up, then run ONE step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    ap.add_argument("--dump-kernels", default=None, help="write the full per-kernel table (JSON) here")
    ap.add_argument("--graph", type=int, default=int(os.environ.get("MICFORMER_CUDA_GRAPH", "1")),
                    help="1: capture the whole step (fwd+loss+bwd+all-reduce+Adam) in ONE CUDA graph and replay it")
    return ap.parse_args()


def cfg_of(name):
    from oracle import micformer_oracle as O      # shapes/seeded inputs only; the product never imports it
    return {"train": O.TRAIN, "w7": O.W7}[name]


WORKLOADS = {"train": "MicFormer train config Head(embed_dim=48,num_classes=8,window=2^3,depths 2-2-6-2)",
             "w7": "MicFormer Head(embed_dim=96,num_classes=8,window=7^3,depths 2-2-6-2)"}
DTYPE = ("f32 storage; tensor-core products: split-bf16 (hi+lo, 3 MMAs, ~2^-17) in the fused stage-0 blocks and the forward GEMMs "
         "(3xTF32 where a weight is read transposed), single-pass TF32 in the backward GEMMs and the 3x3x3 convs; fp32 accumulate, "
         "fp32 LayerNorm / softmax / GELU / loss / Adam")


def config_of(args):
    """the ONE config dict both arms print (the driver compares them key by key)"""
    return {"workload": WORKLOADS[args.config], "volumes_per_gpu_per_step": args.batch, "volume": f"2x(1,{args.size}^3)",
            "step": "fwd + MDiceLoss + bwd + grad all-reduce (N>1) + Adam", "mode": "eval" if args.eval_mode else "train",
            "l2": "per-step working set (activations + 247 MB weights/grads) >> 126 MB L2; no explicit flush"}


def family(kernel_key: str) -> str:
    """kernel-table key (C-ABI name + [shape]) -> kernel family, the granularity the roofline is reported at"""
    n = kernel_key.split("[")[0]
    if n.startswith("mic_linear_"):
        return "gemm (mic_linear_fwd/bwd_data/bwd_weight: tcgen05 TF32 GEMMs of the unfused stages 1-3, patch ops, tail)"
    if n.startswith("mic_conv3"):
        return "conv3 (3x3x3 convs: offset nets, out_conv)"
    if n.startswith("mic_mlp_block") or n.startswith("mic_attn_block"):
        return "fused_block (mic_attn_block_* / mic_mlp_block_*: fused tcgen05 half-blocks of stage 0)"
    if n.startswith("mic_layernorm"):
        return "layernorm"
    if n.startswith("mic_window_attn"):
        return "window_attn (unfused small-window attention, stages 1-3)"
    return n


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- reference arm
def oracle_step_fn(cfg, B, S, train_mode, threads):
    """The reference's CPU implementation of the path, restated (oracle/micformer_oracle.py, kind 'port')."""
    from oracle import micformer_oracle as O
    torch.set_num_threads(threads)
    sd = O.synth_state_dict(cfg, seed=0)
    params = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items()}
    opt = torch.optim.Adam(params.values(), lr=1e-4, weight_decay=0.0)      # train_mmwhs_noPad.py:114
    x, lab = O.synth_inputs(B, S, cfg.num_classes, seed=1)
    gen = torch.Generator().manual_seed(0)

    def step():
        opt.zero_grad(set_to_none=True)
        logits = O.head_forward(x, params, cfg, training=train_mode, gen=gen)
        loss = O.mdice_loss(logits, lab)
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = cfg_of(args.config)
    threads = os.cpu_count() or 1
    B = args.batch
    step = oracle_step_fn(cfg, B, args.size, not args.eval_mode, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    line = {
        "impl": "reference", "metric": "volumes/sec (2-modal 128^3) fwd+bwd", "value": val, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (torch CPU ops)", "data": "synthetic",
        "config": config_of(args),
        "cpu_baseline": {"value": val, "unit": "volumes/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of batch {B} at {args.size}^3, fwd+MDiceLoss+bwd+Adam, "
                                   f"oracle port of the reference's torch-CPU path, {threads} threads"},
        "e2e": {"value": val, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- secondary blocks
def secondary_w7(dev, args, steps=10):
    """Head(embed_dim=96, window 7^3) -- the model family BASELINE config 4's attention shape comes from -- same step
    (fwd + MDiceLoss + bwd + fused Adam, batch 2 of 2x(1,128^3)), eager launches, CUDA events, `steps` timed steps."""
    from micformer_b200.models.MICFormer_self import Head
    from micformer_b200.loss.dice import MDiceLoss
    from micformer_b200.optim import FusedAdam
    from oracle import micformer_oracle as O
    cfg = O.W7
    torch.manual_seed(0)
    model = Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size).to(dev)
    model.train(not args.eval_mode)
    crit = MDiceLoss()
    opt = FusedAdam(model.parameters(), lr=1e-4, weight_decay=0.0)
    x, lab = O.synth_inputs(args.batch, args.size, cfg.num_classes, seed=1)
    x, lab = x.to(dev), lab.to(dev)

    def one():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), lab)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        one()
    torch.cuda.synchronize()
    # the same treatment as the headline step: the whole step in ONE CUDA graph (eager, this model's ~3400 launches cost
    # ~30 ms of host time per step: 99 ms eager vs 68 ms replayed, gpurun_out r2bd / r2bf)
    graph, loss = None, None
    if args.graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        opt.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = crit(model(x), lab)
            loss.backward()
            opt.step()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        if graph is not None:
            graph.replay()
        else:
            loss = one()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"workload": WORKLOADS["w7"], "volumes_per_step": args.batch, "steps": steps, "warmup": 3, "ms_per_step": round(ms, 3),
           "volumes_per_s": round(args.batch / ms * 1e3, 2), "loss": round(float(loss.detach()), 6), "cuda_graph": graph is not None,
           "note": "window attention forward and backward on tcgen05 (343-token windows)"}
    del graph
    return out


def eager_cuda(dev, cfg, B, S, train_mode, steps=3):
    """The reference's own ops (the oracle port: F.linear / bmm / softmax / conv3d / layer_norm / grid_sample ...) executed
    EAGERLY on this GPU by torch -- what the unmodified reference would run on a B200 -- with the script's settings
    (cudnn off, fp32 matmuls: train_mmwhs_noPad.py:28-29) and with cuDNN + TF32 enabled.  Informational context only."""
    from oracle import micformer_oracle as O
    out = {}
    x, lab = O.synth_inputs(B, S, cfg.num_classes, seed=1)
    x, lab = x.to(dev), lab.to(dev)
    for name, cudnn_on, tf32 in (("script_settings_cudnn_off_fp32", False, False), ("cudnn_tf32", True, True)):
        try:
            torch.backends.cudnn.enabled = cudnn_on
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            params = {k: torch.nn.Parameter(v.to(dev)) for k, v in O.synth_state_dict(cfg, seed=0).items()}
            opt = torch.optim.Adam(params.values(), lr=1e-4, weight_decay=0.0)
            gen = torch.Generator().manual_seed(0)

            def one():
                opt.zero_grad(set_to_none=True)
                loss = O.mdice_loss(O.head_forward(x, params, cfg, training=train_mode, gen=gen), lab)
                loss.backward()
                opt.step()

            one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": round(ms, 2), "volumes_per_s": round(B / ms * 1e3, 2), "steps": steps}
            del params, opt
            torch.cuda.empty_cache()
        except Exception as e:      # informational block: never fail the bench line
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
    torch.backends.cudnn.enabled = True
    torch.backends.cuda.matmul.allow_tf32 = False
    out["note"] = ("oracle port of the reference ops, eager torch on the same GPU, same batch and step; not the product path")
    return out


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from micformer_b200 import _native
    from micformer_b200.models.MICFormer_self import Head
    from micformer_b200.loss.dice import MDiceLoss
    from micformer_b200.parallel import GradSync
    from oracle import micformer_oracle as O      # seeded synthetic inputs + CPU baseline only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout must carry exactly ONE JSON line: native libraries (the NCCL version banner) write to file descriptor 1
    # directly, so fd 1 is pointed at stderr for the run and the line goes out through a saved duplicate
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"     # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)
    _native.set_gemm_mode(args.gemm_mode)

    cfg = cfg_of(args.config)
    B, S = args.batch, args.size
    torch.manual_seed(0)                          # identical replicas on every rank
    model = Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size).to(dev)
    model.train(not args.eval_mode)
    crit = MDiceLoss()
    from micformer_b200.optim import FusedAdam
    opt = FusedAdam(model.parameters(), lr=1e-4, weight_decay=0.0)          # train_mmwhs_noPad.py:114
    arena = None
    if args.arena == 1 or (args.arena < 0 and world > 1):
        from micformer_b200.arena import GradArena
        arena = GradArena.for_model(model)         # .grad slices of one flat buffer ([decoder | encoder]): 1 memset, in-place all-reduce
        opt.attach_arena(arena)
    sync = GradSync(list(model.parameters()), arena=arena)
    if arena is not None and world > 1 and os.environ.get("MICFORMER_OVERLAP", "1") != "0":
        sync.enable_overlap(model)                 # decoder gradients are exchanged under the encoder's backward
    x_h, lab_h = O.synth_inputs(B, S, cfg.num_classes, seed=1 + rank)
    x_h, lab_h = x_h.pin_memory(), lab_h.pin_memory()
    x_d, lab_d = x_h.to(dev), lab_h.to(dev)
    torch.manual_seed(1234 + rank)                # DropPath masks differ per rank like independent data loaders

    def step(x, lab):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), lab)
        loss.backward()
        sync.sync()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(3, args.warmup)):
        step(x_d, lab_d)
    # --- whole-step CUDA graph: ~3000 kernel launches per step collapse into one cudaGraphLaunch -------------
    graph, loss_static, launches_per_step = None, None, None
    if args.graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step(x_d, lab_d)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        opt.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        _native.reset_launch_count()
        cap_stream = torch.cuda.Stream(priority=-1) if os.environ.get("MICFORMER_STREAM_PRIO", "0") == "1" else None
        with torch.cuda.graph(graph, stream=cap_stream):
            if arena is not None:
                arena.zero()                       # one memset node; without the arena the grads are re-created per replay
            loss_static = crit(model(x_d), lab_d)
            loss_static.backward()
            sync.sync()
            opt.step()
        launches_per_step = _native.launch_count()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()

    def run_step():
        if graph is not None:
            graph.replay()
            return loss_static
        return step(x_d, lab_d)

    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(x_d, lab_d)          # eager on purpose: ncu attributes time to individual kernels
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # --- device-resident throughput ---------------------------------------------------------------------
    clocks = ClockSampler(local)
    clocks.start()
    _native.reset_launch_count()
    ms = timed(run_step, args.steps)
    launches = _native.launch_count() if graph is None else launches_per_step * args.steps
    clk = clocks.stop()
    value = world * B * args.steps / (ms / 1e3)

    # --- end to end through the public API: every step copies its inputs from pinned host memory and reads the loss
    #     back.  The copies are issued the way a prefetching loader does it (pin_memory + non_blocking on a copy
    #     stream, one step ahead, into staging buffers); the step itself starts with a device-to-device move into the
    #     graph's static inputs.  One full H2D transfer of x and the labels is inside every timed step.
    copy_stream = torch.cuda.Stream()
    x_stage, lab_stage = torch.empty_like(x_d), torch.empty_like(lab_d)
    copied, stage_free = torch.cuda.Event(), torch.cuda.Event()

    def prefetch():
        copy_stream.wait_event(stage_free)
        with torch.cuda.stream(copy_stream):
            x_stage.copy_(x_h, non_blocking=True)
            lab_stage.copy_(lab_h, non_blocking=True)
            copied.record(copy_stream)

    def e2e_step():
        cur = torch.cuda.current_stream()
        cur.wait_event(copied)                     # this step's inputs have landed in the staging buffers
        if graph is not None:
            x_d.copy_(x_stage); lab_d.copy_(lab_stage)
            stage_free.record(cur)
            prefetch()                             # next step's H2D overlaps this step's compute
            graph.replay()
            return float(loss_static.detach())     # D2H read of the loss (train_mmwhs_noPad.py:189-197)
        x, lab = x_stage.clone(), lab_stage.clone()
        stage_free.record(cur)
        prefetch()
        return float(step(x, lab).detach())

    stage_free.record(torch.cuda.current_stream())
    prefetch()
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_val = world * B * args.steps / (ms_e2e / 1e3)

    # --- per-kernel pass (CUDA events around every C-ABI launch, same steps) -> roofline of the dominant kernel FAMILY
    roofline, shares, step_roofline = None, None, None
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk)); peaks["src"] = "measured"
    if rank == 0 and world == 1 and not args.no_kernel_pass:      # N = 1 only: the pass steps the model on this rank alone
        _native.profile_begin()
        ksteps = min(args.steps, 3)
        for _ in range(ksteps):
            step(x_d, lab_d)
        prof = _native.profile_end()
        tot = sum(v["ms"] for v in prof.values())
        top = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
        fam = {}
        for k, v in prof.items():
            f = fam.setdefault(family(k), {"ms": 0.0, "calls": 0, "bytes": 0, "flops": 0})
            for key in f:
                f[key] += v[key]
        ftop = sorted(fam.items(), key=lambda kv: -kv[1]["ms"])
        shares = [{"family": k, "share": round(v["ms"] / tot, 4), "ms_per_step": round(v["ms"] / ksteps, 4),
                   "calls_per_step": v["calls"] // ksteps,
                   "gbs": round(v["bytes"] / (v["ms"] * 1e6), 1) if v["ms"] > 0 else None,
                   "tflops": round(v["flops"] / (v["ms"] * 1e9), 2) if v["ms"] > 0 else None} for k, v in ftop[:8]]
        if args.dump_kernels:
            os.makedirs(os.path.dirname(os.path.abspath(args.dump_kernels)), exist_ok=True)
            with open(args.dump_kernels, "w") as f:
                json.dump({"steps": ksteps, "total_ms_per_step": tot / ksteps,
                           "kernels": [{"kernel": k, "ms_per_step": v["ms"] / ksteps, "calls_per_step": v["calls"] / ksteps,
                                        "us_per_call": 1e3 * v["ms"] / v["calls"], "share": v["ms"] / tot,
                                        "gbs": v["bytes"] / (v["ms"] * 1e6) if v["ms"] > 0 else None,
                                        "tflops": v["flops"] / (v["ms"] * 1e9) if v["ms"] > 0 else None}
                                       for k, v in top]}, f, indent=1)
        name, v = ftop[0]
        ach = v["bytes"] / (v["ms"] * 1e6)        # GB/s: algorithmic bytes / event-timed duration, summed over the family
        # measured DRAM traffic of the same family from the committed ncu pass over one step (profiles/r02_traffic.json:
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, made by scripts/traffic_pass.sh)
        traffic = None
        tj = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tj):
            ent = json.load(open(tj)).get("families", {}).get(name.split(" ")[0])
            if ent:
                traffic = ent.get("dram_bytes_per_launch")
        roofline = {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"],
                    "peak_source": peaks["src"], "unit": "GB/s", "frac": round(ach / peaks["hbm_gbs"], 4),
                    "traffic": traffic, "launches_per_step": v["calls"] // ksteps,
                    "avg_launch_ms": round(v["ms"] / v["calls"], 5),
                    "algorithmic_bytes_per_launch": v["bytes"] // v["calls"],
                    "algorithmic_tflops": round(v["flops"] / (v["ms"] * 1e9), 2),
                    "share_of_kernel_time": round(v["ms"] / tot, 4),
                    "note": "family = all launches of these entry points in one step; achieved = summed algorithmic bytes / summed "
                            "CUDA-event time (per-launch, serialised)"}
        # whole step: algorithmic work of every kernel of one step against the step time of the timed region
        gflop = sum(v["flops"] for v in prof.values()) / ksteps / 1e9
        gbyte = sum(v["bytes"] for v in prof.values()) / ksteps / 1e9
        step_ms = ms / args.steps
        step_roofline = {"gflop_per_step": round(gflop, 1), "algorithmic_gb_per_step": round(gbyte, 2),
                         "tflops": round(gflop / step_ms, 2), "gbs": round(gbyte / step_ms * 1e3, 1),
                         "frac_of_hbm_peak": round(gbyte / step_ms * 1e3 / peaks["hbm_gbs"], 4),
                         "frac_of_tf32_peak": round(gflop / step_ms / (peaks.get("bf16_tflops", 1590.0) / 2.0), 4),
                         "serialised_kernel_ms_per_step": round(tot / ksteps, 3),
                         "note": "sum of per-kernel algorithmic bytes / flops (micformer_b200/_native.COST) over one step / "
                                 "graph-replayed step time; tf32 peak = half the measured bf16 burst"}
    # --- BASELINE.json's second metric: the cross-modal attention kernel in isolation (config 4: 4096 windows of 343
    #     tokens, 96 channels, 3 heads of 32; fp32 I/O, TF32 tensor cores), timed alone with CUDA events
    attn = None
    if rank == 0 and world == 1 and not args.no_kernel_pass and not args.no_attn_isolation:
        from micformer_b200 import ops as _ops
        torch.cuda.empty_cache()
        Bw, Ca, Ha = 4096, 96, 3
        qkv = torch.randn(Bw * 343, 3 * Ca, device=dev)
        for _ in range(3):
            _ops.window_attn_fwd(qkv, Ca, Ha, Bw, (7, 7, 7), (7, 7, 7))
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            _ops.window_attn_fwd(qkv, Ca, Ha, Bw, (7, 7, 7), (7, 7, 7))
        a1.record(); torch.cuda.synchronize()
        ams = a0.elapsed_time(a1) / 5
        aflops = 4.0 * Bw * Ha * 343 * 343 * 32
        abytes = 4.0 * (4 * Bw * 343 * Ca + Bw * 343 * Ha)
        pk = peaks
        tf32_peak = pk.get("bf16_tflops", 1590.0) / 2.0          # dense TF32 = half the measured bf16 cuBLAS burst figure
        attn = {"workload": "cross-modal window attention forward, 4096 windows x 343 tokens x 96 ch x 3 heads (config 4), fp32 I/O",
                "ms": round(ams, 4), "tflops": round(aflops / ams / 1e9, 1), "tf32_peak_tflops": round(tf32_peak, 1),
                "frac_of_tf32_peak": round(aflops / ams / 1e9 / tf32_peak, 4),
                "frac_of_bf16_peak": round(aflops / ams / 1e9 / pk.get("bf16_tflops", 1590.0), 4),
                "algorithmic_gbs": round(abytes / ams / 1e6, 1),
                "frac_of_hbm_peak": round(abytes / ams / 1e6 / pk["hbm_gbs"], 4),
                "note": "fp32 Q/K/V/O make this shape HBM-bound below the tensor ridge: 2.16 GB / measured copy bandwidth = 0.33 ms floor"}
        # the same shape's backward (tcgen05: csrc/window_attn_tc_bwd.cu): dQ, dK, dV from Q, K, V, O, dO, lse
        o_a, lse_a = _ops.window_attn_fwd(qkv, Ca, Ha, Bw, (7, 7, 7), (7, 7, 7))
        do_a = torch.randn_like(o_a)
        for _ in range(3):
            _ops.window_attn_bwd(qkv, o_a, do_a, lse_a, Ca, Ha, Bw, (7, 7, 7), (7, 7, 7))
        torch.cuda.synchronize()
        a0.record()
        for _ in range(5):
            _ops.window_attn_bwd(qkv, o_a, do_a, lse_a, Ca, Ha, Bw, (7, 7, 7), (7, 7, 7))
        a1.record(); torch.cuda.synchronize()
        bms = a0.elapsed_time(a1) / 5
        bflops = 2.5 * aflops                                     # five 343 x 343 x 32 products per (window, head)
        bbytes = 4.0 * (8 * Bw * 343 * Ca + Bw * 343 * Ha)       # q, k, v, o, do in; dq, dk, dv out; lse
        attn["backward"] = {"ms": round(bms, 4), "tflops": round(bflops / bms / 1e9, 1),
                            "frac_of_tf32_peak": round(bflops / bms / 1e9 / tf32_peak, 4),
                            "algorithmic_gbs": round(bbytes / bms / 1e6, 1),
                            "frac_of_hbm_peak": round(bbytes / bms / 1e6 / pk["hbm_gbs"], 4),
                            "note": "algorithmic flops (5 products); the kernel runs 8 (scores in both orientations) -- "
                                    "A operands of the accumulating products come from TMEM"}
        del qkv, o_a, lse_a, do_a
    if world > 1:
        dist.barrier()

    # --- secondary configs (N = 1): the window-7 model (the shape family of BASELINE config 4) through the same step, and the
    #     reference ops eager on this GPU -- the only GPU-vs-GPU context this project has: informational
    secondary, eager = None, None
    run_info = {"parallelism": f"dp{world}", "gemm_mode": args.gemm_mode, "cuda_graph": bool(args.graph),
                "fused_blocks": os.environ.get("MICFORMER_FUSED", "1") != "0", "grad_arena": arena is not None}
    if rank == 0 and world == 1 and not args.no_secondary and args.config == "train":
        del graph
        model = opt = arena = sync = None
        torch.cuda.empty_cache()
        secondary = {"w7_step": secondary_w7(dev, args), }
        torch.cuda.empty_cache()
        eager = eager_cuda(dev, cfg, B, S, not args.eval_mode)
        torch.cuda.empty_cache()

    # --- CPU baseline on the host cores (rank 0, N=1 only): oracle port, bounded sample -------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cstep = oracle_step_fn(cfg, 1, S, not args.eval_mode, threads)
        cstep()
        best = 1e30
        for _ in range(2):
            t0 = time.perf_counter(); cstep(); best = min(best, time.perf_counter() - t0)
        cpu = {"value": 1.0 / best, "unit": "volumes/s", "cores": threads, "kind": "port",
               "sample": f"1 volume (batch 1) at {S}^3, fwd+MDiceLoss+bwd+Adam, best of 2 after 1 warm-up, "
                         f"oracle port of the reference torch-CPU path on {threads} threads"}

    if rank == 0:
        h2d = x_h.numel() * 4 + lab_h.numel() * 4
        line = {
            "metric": "volumes/sec (2-modal 128^3) fwd+bwd", "value": value, "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": config_of(args),
            "run": run_info,
            "clocks": clk,
            "e2e": {"value": e2e_val, "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline, "step_roofline": step_roofline, "cpu_baseline": cpu, "attention_kernel": attn,
            "kernel_shares": shares, "secondary": secondary, "eager_cuda": eager,
        }
        sys.stdout.flush()
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # NCCL communicators referenced by a captured CUDA graph do not tear down cleanly: drain, rendezvous once
        # more and leave without running destructors (every rank exits 0; rank 0 has already printed its line)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
