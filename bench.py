#!/usr/bin/env python
"""Benchmark of the hot path: frozen INT8 DeiT forward (BASELINE.json: images/sec DeiT-B INT8 bs=256).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model NAME] [--batch B]

One "step" = one forward of the synthetic frozen model over one batch of B synthetic 3x224x224
fp32 images per GPU (batch sharded across ranks, weak scaling, no forward collective; the frozen
parameter pack is NCCL-broadcast once from rank 0).  Prints ONE JSON line (rank 0):

  value        whole-job images/s, inputs already resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the public API with HOST buffers: pinned fp32 images -> H2D ->
               engine -> D2H logits, every step
  roofline     dominant kernel (tcgen05 INT8 GEMM): algorithmic int-ops / measured launch time vs the
               measured tensor peak (MEASURED_PEAKS.json bf16 x 2)
  cpu_baseline the oracle port (oracle/, pinned to the reference) timed on this box's host cores
  --impl reference   the reference's CPU integer path (oracle port; the reference itself is Python
               and cannot travel to the GPU box) on rank 0 only, same metric / config
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
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "images/sec DeiT-B INT8 bs=256"
MODEL = "deit_base_patch16_224"


def int_ops_per_image(meta) -> float:
    """Algorithmic integer ops (2 x MAC) per image, SURVEY.md section 8(d)."""
    C, N, H, D, Hd = meta["embed_dim"], meta["n_tok"], meta["num_heads"], meta["head_dim"], meta["mlp_hidden"]
    pe = (N - 1) * C * meta["in_chans"] * meta["patch"] ** 2
    blk = N * C * 3 * C + 2 * H * N * N * D + N * C * C + 2 * N * C * Hd
    return 2.0 * (pe + meta["depth"] * blk + C * meta["num_classes"])


def gemm_shapes(meta, B):
    """(name, M, N, K, launches per forward) of every tcgen05 GEMM launch."""
    C, N, Hd, L = meta["embed_dim"], meta["n_tok"], meta["mlp_hidden"], meta["depth"]
    M = B * N
    return [("patch_embed", B * (N - 1), C, meta["in_chans"] * meta["patch"] ** 2, 1),
            ("qkv", M, 3 * C, C, L), ("proj", M, C, C, L), ("fc1", M, Hd, C, L), ("fc2", M, C, Hd, L),
            ("head", B, meta["num_classes"], C, 1)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under the benchmark's load (B200_PROFILING.md): one
    streaming `nvidia-smi -lms 50` process, read by a thread, from the warm-up to the end of the timed regions."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                line = line.strip()
                if line:
                    self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        try:
            if self.proc is not None:
                self.proc.terminate()                      # the exact process started above
        except Exception:
            pass
        self.thread.join(timeout=3)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        pw = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        # "under load": samples whose power draw is above the midpoint between idle and the maximum seen
        if pw and len(pw) == len(sm) and max(pw) > 1.3 * min(pw):
            thr = 0.5 * (max(pw) + min(pw))
            load = [c for c, w in zip(sm, pw) if w >= thr] or sm
        else:
            load = sm
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "samples_under_load": len(load),
                "power_w_max": max(pw) if pw else None}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(kernel_substr):
    """Average DRAM bytes per launch of a kernel from the newest committed ncu --set full summary
    (profiles/ncu_full_*.json, written by tools/summarize_ncu.py); None if there is none."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_full_*.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    tot_b = tot_n = 0
    for name, e in d["kernels"].items():
        if kernel_substr in name:
            tot_b += e["dram_bytes"]
            tot_n += e["launches"]
    return (tot_b / tot_n if tot_n else None), os.path.basename(files[-1])


def build_pack(model_name):
    from ivit_b200.calib import build_synthetic
    from ivit_b200.pack import export_deit
    return export_deit(build_synthetic(model_name))


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's integer path (all host threads)."""
    if rank != 0:
        return
    import oracle as O
    import oracle.model as OM
    from ivit_b200.synth import synth_images
    cores = os.cpu_count() or 1
    O.set_threads(1)                                  # one image per host thread (images are independent)
    pack = build_pack(args.model)
    imgs_per_step = args.ref_images or min(cores, 128)
    x = synth_images(imgs_per_step, seed=11).numpy()
    for _ in range(min(args.warmup, 1)):
        OM.deit_forward_parallel(pack, x, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        OM.deit_forward_parallel(pack, x, threads=cores)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = imgs_per_step / dt
    line = {"metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic", "impl": "reference",
            "config": {"workload": "%s frozen INT8 forward, 3x224x224 synthetic, batch 256 per GPU" % args.model,
                       "note": "reference is pure Python/PyTorch and cannot travel to the GPU box; its CPU integer path "
                               "is timed through the oracle port (oracle/, bit-pinned to the reference)"},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": "%d images of the batch per step, one image per host thread (%d threads)" % (imgs_per_step, cores)},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def time_gemms(eng, B, iters=10):
    """Average device time of each GEMM shape of the forward, measured with CUDA events on the
    launching stream (the dominant kernel of the step)."""
    import ivit_b200.kernels as K
    meta, t = eng.meta, eng.t
    dev = eng.device
    names = {"patch_embed": ("patch_embed.proj", "patch_embed.qact", 16), "qkv": ("blocks.0.attn.qkv", "blocks.0.attn.qact1", 8),
             "proj": ("blocks.0.attn.proj", "blocks.0.attn.qact3", 16), "fc1": ("blocks.0.mlp.fc1", "blocks.0.mlp.qact_gelu", 8),
             "fc2": ("blocks.0.mlp.fc2", "blocks.0.mlp.qact2", 16)}
    res = []
    for name, M, N, Kd, count in gemm_shapes(meta, B):
        if name == "head":
            continue
        lin, qa, bits = names[name]
        a = torch.randint(-128, 128, (M, Kd), dtype=torch.int8, device=dev)
        out = torch.empty((M, N), dtype=torch.int8 if bits == 8 else torch.int16, device=dev)
        kw = {}
        if name in ("proj", "fc2"):
            r = torch.randint(-30000, 30000, (M, N), dtype=torch.int16, device=dev)
            blk = "blocks.0.qact2" if name == "proj" else "blocks.0.qact4"
            kw = dict(two_stage=True, me2=eng.s[blk + ".me"], residual=r, res_me=eng.s[blk + ".me_res"])
        run = lambda: K.gemm_i8(a, t[lin + ".weight_integer"], bias=t[lin + ".bias_integer"], mode="requant",
                                me=t[qa + ".me"], bits=bits, out=out, **kw)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res.append({"name": name, "M": M, "N": N, "K": Kd, "launches": count, "ms": ms,
                    "tops": 2.0 * M * N * Kd / (ms * 1e-3) / 1e12})
        del a, out
    return res


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from ivit_b200.dist import broadcast_pack
    from ivit_b200.engine import Engine
    from ivit_b200.synth import synth_images
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pack = build_pack(args.model) if rank == 0 else None
    if world > 1:
        pack = broadcast_pack(pack, src=0, device=dev)       # one-time NCCL broadcast of the frozen INT8 parameters
    eng = Engine(pack, dev)
    B = args.batch
    g = torch.Generator(device="cpu")
    g.manual_seed(1234 + rank)
    host = torch.randn((B, 3, eng.meta["img_size"], eng.meta["img_size"]), generator=g).pin_memory()
    x = host.to(dev, non_blocking=True)
    out_host = torch.empty((B, eng.meta["num_classes"]), dtype=torch.float32).pin_memory()

    def barrier():
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
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        eng(x)

    # end-to-end step: pinned host images -> H2D -> engine -> D2H logits, every step.  The H2D copy of step
    # i+1 runs on a copy stream while step i computes (two device staging buffers); every step still pays
    # its own H2D and D2H inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)

    def make_e2e(host_t):
        """Pipelined end-to-end step over `host_t` (pinned): H2D of step i+1 on the copy stream during step i."""
        stage = [torch.empty(host_t.shape, dtype=host_t.dtype, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"i": 0, "primed": False}

        def issue_h2d(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                stage[slot].copy_(host_t, non_blocking=True)
                ready[slot].record(copy_stream)

        def step():
            main = torch.cuda.current_stream(dev)
            slot = state["i"] & 1
            if not state["primed"]:
                consumed[0].record(main); consumed[1].record(main)
                issue_h2d(slot)
                state["primed"] = True
            issue_h2d(slot ^ 1)                             # prefetch the next step's images
            main.wait_event(ready[slot])
            logits = eng(stage[slot])
            consumed[slot].record(main)
            out_host.copy_(logits, non_blocking=True)
            state["i"] += 1
        return step

    step_e2e = make_e2e(host)
    # the same images as decoded uint8 pixels (the engine applies ToTensor + Normalize on the device): 4x fewer H2D bytes
    host_u8 = torch.randint(0, 256, host.shape, generator=g, dtype=torch.uint8).pin_memory()
    step_e2e_u8 = make_e2e(host_u8)

    with ClockSampler(local_rank) as clk:                  # sampled from the warm-up to the end of both timed regions
        for _ in range(max(args.warmup, 3)):
            step_resident()
        total_ms = timed(step_resident, args.steps)
        for _ in range(max(args.warmup, 6)):               # both staging buffers seen often enough to be graph-bound
            step_e2e()
        e2e_ms = timed(step_e2e, args.steps)
        for _ in range(3):
            step_e2e_u8()
        e2e_u8_ms = timed(step_e2e_u8, args.steps)
    clocks = clk.summary()

    ms_step = total_ms / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_val = world * B / (e2e_ms / args.steps * 1e-3)
    if rank != 0:
        return
    peaks, peak_src = measured_peaks()
    # dominant kernel: every tcgen05 GEMM launch of the step timed in place (CUDA events on the launching stream around
    # each launch of eager forwards), grouped by shape; the isolated L2-warm timing of time_gemms() is kept beside it
    per_launch = eng.time_gemms(x, forwards=3)
    groups = {}
    for name, M_, N_, K_, ms in per_launch:
        key = "head" if name == "head" else name.split(".")[-1] if name.startswith("blocks.") else "patch_embed"
        g_ = groups.setdefault(key, {"name": key, "M": M_, "N": N_, "K": K_, "launches": 0, "ms": 0.0})
        g_["launches"] += 1
        g_["ms"] += ms
    gem = []
    for g_ in groups.values():
        g_["ms"] /= g_["launches"]
        g_["tops"] = 2.0 * g_["M"] * g_["N"] * g_["K"] / (g_["ms"] * 1e-3) / 1e12
        gem.append(g_)
    iso = {g_["name"]: g_["ms"] for g_ in time_gemms(eng, B)}
    for g_ in gem:
        g_["ms_isolated_l2_warm"] = iso.get(g_["name"])
    ops = sum(2.0 * g_["M"] * g_["N"] * g_["K"] * g_["launches"] for g_ in gem)
    gms = sum(g_["ms"] * g_["launches"] for g_ in gem)
    achieved = ops / (gms * 1e-3) / 1e12
    peak = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    cpu = None
    if not args.no_cpu_baseline:
        import oracle as O
        import oracle.model as OM
        cores = os.cpu_count() or 1
        O.set_threads(1)
        n_img = args.ref_images or min(cores, 128)
        xs = synth_images(n_img, seed=11).numpy()
        t0 = time.perf_counter()
        reps = 0
        while reps < 1 or (time.perf_counter() - t0 < 10.0 and reps < 20):
            OM.deit_forward_parallel(pack, xs, threads=cores)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": n_img / dt, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "%d forward(s) of %d images (same synthetic %s pack), one image per host thread (%d threads)" % (
                   reps, n_img, args.model, cores)}
    act_mb = B * eng.meta["n_tok"] * eng.meta["mlp_hidden"] / 1e6
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "%s frozen INT8 forward, batch %d per GPU, 3x224x224 synthetic fp32 images" % (args.model, B),
                       "global_batch": world * B, "parallelism": "batch sharded dp%d, one-time NCCL weight broadcast" % world,
                       "l2": "no flush needed: per-step working set (fp32 input %.0f MB + fc1/GELU activations 2 x %.0f MB) exceeds the 126 MB L2" % (
                           B * 3 * 224 * 224 * 4 / 1e6, act_mb),
                       "int_ops_per_image": int_ops_per_image(eng.meta)},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("gemm_i8_tcgen05_kernel")[0], "traffic_source": ncu_traffic("gemm_i8_tcgen05_kernel")[1],
                         "timing": "CUDA events around each GEMM launch inside eager forwards (operands as produced by the preceding kernels)",
                         "algorithmic_bytes_per_launch": sum((g_["M"] * g_["K"] + g_["N"] * g_["K"] + g_["M"] * g_["N"] * (1 if g_["name"] in ("qkv", "fc1") else 4))
                                                             * g_["launches"] for g_ in gem) / sum(g_["launches"] for g_ in gem), "kernel": "gemm_i8_tcgen05_kernel (all %d GEMM launches of the step)" % sum(g_["launches"] for g_ in gem),
                         "peak_source": "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (%s); int8 tensor rate = 2 x bf16" % peak_src,
                         "per_shape": gem, "gemm_share_of_step": gms / ms_step,
                         "whole_step_int_tops": int_ops_per_image(eng.meta) * B / (ms_step * 1e-3) / 1e12},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": int(host.numel() * 4) * world,
                    "d2h_bytes_per_step": int(out_host.numel() * 4) * world, "ms_per_step": e2e_ms / args.steps,
                    "uint8_input": {"value": world * B / (e2e_u8_ms / args.steps * 1e-3), "unit": "images/s",
                                    "h2d_bytes_per_step": int(host_u8.numel()) * world, "ms_per_step": e2e_u8_ms / args.steps,
                                    "note": "decoded uint8 pixels in, ToTensor + Normalize fused into the stem kernel"}},
            "gpu_launches": (eng.launches_per_forward - 1) * args.steps,
            "clocks": clocks}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--ref-images", type=int, default=0, help="images per CPU-baseline step (bounded sample; 0 = one per host core, max 128)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
