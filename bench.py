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
  cpu_baseline the reference's own CPU integer path timed on this box's host cores: the UNMODIFIED reference
               model(x) (baseline/_ref staged by tools/fetch_ref.py, its own quantization_utils, fp32 carrier, frozen:
               kind "reference"), or the oracle port when the staged copy is absent (kind "port")
  configs      the other single-GPU BASELINE.json configs (DeiT-small bs=128, Swin-tiny bs=128), device-timed
  parity       logits of seeded images: every rank's sha256 all-gathered (N > 1), rank 0 against the CPU oracle
  --impl reference   that CPU arm alone, on rank 0 only, same metric / config
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


def int_ops_per_image_swin(meta) -> float:
    """Algorithmic integer ops (2 x MAC) per image of a Swin pack (8.98e9 for Swin-tiny)."""
    C, R = meta["embed_dim"], meta["grid"]
    mac = R * R * C * meta["in_chans"] * meta["patch"] ** 2
    for li, depth in enumerate(meta["depths"]):
        L, N, Hd = R * R, meta["window"][li] ** 2, meta["mlp_hidden"][li]
        mac += depth * (L * C * 3 * C + 2 * L * N * C + L * C * C + 2 * L * C * Hd)
        if li + 1 < len(meta["depths"]):
            mac += (L // 4) * (4 * C) * (2 * C)
            C, R = 2 * C, R // 2
    return 2.0 * (mac + C * meta["num_classes"])


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
    # the newest DeiT capture (ncu_full_r<round><tag>.json; the Swin captures are ncu_full_swin_*.json)
    files = sorted(f for f in glob.glob(os.path.join(ROOT, "profiles", "ncu_full_r*.json")))
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
    from ivit_b200.pack import export_deit, export_swin
    return (export_swin if model_name.startswith("swin") else export_deit)(build_synthetic(model_name))


def make_engine(pack, dev):
    from ivit_b200.engine import Engine
    from ivit_b200.swin_engine import SwinEngine
    return SwinEngine(pack, dev) if pack.meta["arch"] == "swin" else Engine(pack, dev)


def pack_int_ops(meta) -> float:
    return int_ops_per_image_swin(meta) if meta["arch"] == "swin" else int_ops_per_image(meta)


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(model_name, pack, n_img, reps_or_seconds, warmup=0):
    """Time the reference's CPU integer path on this box's host cores.  Preferred: the UNMODIFIED reference model(x)
    (staged under baseline/_ref, its own quantization_utils, fp32 carrier, frozen ranges from the golden calibration
    table -- quant_train.py:325-334 call shape), kind "reference".  Fallback: the oracle port, kind "port".
    reps_or_seconds: int = exactly that many timed forwards; float = repeat until that many seconds have passed.
    Returns (images/s, kind, cores, sample description, seconds per forward)."""
    from ivit_b200.synth import synth_images
    cores = os.cpu_count() or 1
    x = synth_images(n_img, seed=11)
    fwd, kind = None, None
    try:
        from ivit_b200.calib import apply_calibration, load_calibration
        from ivit_b200.dropin import cuda_calls_are_noops, load_reference_models
        from ivit_b200.synth import synth_parameters
        ref = load_reference_models(mirror=False)
        cal = load_calibration(model_name)
        model = getattr(ref, model_name)(pretrained=False).eval()
        if synth_parameters(model, cal["seed"]) != cal["weights_sha256"]:
            raise RuntimeError("reference parameter names differ from the calibration table's")
        apply_calibration(model, cal["ranges"])
        ref.freeze_model(model)
        torch.set_num_threads(cores)

        def fwd():
            with torch.no_grad(), cuda_calls_are_noops():
                return model(x)
        kind = "reference"
        what = "unmodified reference model(x) from baseline/_ref (fp32 carrier, frozen, torch CPU, %d threads)" % cores
    except FileNotFoundError:
        import oracle as O
        import oracle.model as OM
        O.set_threads(1)                                  # one image per host thread (images are independent)
        xn = x.numpy()
        if pack.meta["arch"] == "swin":
            def fwd():
                return OM.swin_forward(pack, xn)
        else:
            def fwd():
                return OM.deit_forward_parallel(pack, xn, threads=cores)
        kind = "port"
        what = "oracle port (baseline/_ref not staged), one image per host thread (%d threads)" % cores
    for _ in range(warmup):
        fwd()
    t0 = time.perf_counter()
    reps = 0
    if isinstance(reps_or_seconds, int):
        for _ in range(max(reps_or_seconds, 1)):
            fwd()
            reps += 1
    else:
        while reps < 1 or (time.perf_counter() - t0 < reps_or_seconds and reps < 20):
            fwd()
            reps += 1
    dt = (time.perf_counter() - t0) / reps
    return n_img / dt, kind, cores, "%d forward(s) of %d images: %s" % (reps, n_img, what), dt


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """CPU arm: the reference's own CPU integer path (all host threads), a bounded sample of the workload per step."""
    if rank != 0:
        return
    pack = build_pack(args.model)
    n_img = args.ref_images or 8
    val, kind, cores, sample, dt = cpu_reference_run(args.model, pack, n_img, int(args.steps), warmup=args.warmup)
    line = {"metric": METRIC if args.model == MODEL and args.batch == 256 else "images/sec %s INT8 bs=%d" % (args.model, args.batch),
            "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic", "impl": "reference",
            "config": {"workload": "%s frozen INT8 forward, 3x224x224 synthetic fp32 images, batch %d per GPU" % (args.model, args.batch),
                       "images_per_step": n_img,
                       "note": "CPU arm: each step is a bounded sample of %d images of the batch (throughput is batch-flat on "
                               "the CPU); nothing of this repository's kernels or engine is on this path" % n_img},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
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


def device_time(fn, steps, warmup):
    """Average ms per call of fn on the current stream (CUDA events, synchronise on both sides)."""
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def other_config(model_name, batch, dev, peaks, steps):
    """One of the other single-GPU BASELINE.json configs, device-timed (inputs resident, CUDA-graph replay)."""
    pack = build_pack(model_name)
    eng = make_engine(pack, dev)
    g = torch.Generator(device="cpu")
    g.manual_seed(4321)
    x = torch.randn((batch, 3, pack.meta["img_size"], pack.meta["img_size"]), generator=g).to(dev)
    ms = device_time(lambda: eng(x), steps, 5)
    ops = pack_int_ops(pack.meta) * batch
    tops = ops / (ms * 1e-3) / 1e12
    out = {"workload": "%s frozen INT8 forward, batch %d, 1 GPU" % (model_name, batch), "value": batch / (ms * 1e-3),
           "unit": "images/s", "ms_per_step": ms, "int_ops_per_image": pack_int_ops(pack.meta), "whole_step_int_tops": tops,
           "frac_of_sustained_tensor_peak": tops / (2.0 * float(peaks["bf16_tflops_sustained"])),
           "gpu_launches_per_step": eng.launches_per_forward}
    if hasattr(eng, "attention_fallbacks"):
        out["attention_fallbacks_to_mma_sync"] = eng.attention_fallbacks
    del eng, x
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world, local_rank):
    import hashlib
    import torch.distributed as dist
    from ivit_b200.dist import broadcast_pack
    from ivit_b200.synth import synth_images
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pack = build_pack(args.model) if rank == 0 else None
    if world > 1:
        pack = broadcast_pack(pack, src=0, device=dev)       # one-time NCCL broadcast of the frozen INT8 parameters
    eng = make_engine(pack, dev)
    is_deit = pack.meta["arch"] == "deit"
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

    # ---- parity outside the timed region: the same 4 seeded images on every rank (weights arrived over NCCL on ranks > 0)
    xs = synth_images(4, seed=5)
    logits4 = eng(xs.to(dev)).float().cpu().numpy().copy()
    digest = hashlib.sha256(logits4.tobytes()).digest()
    parity = {"images": 4}
    if world > 1:
        mine = torch.tensor(list(digest), dtype=torch.uint8, device=dev)
        allg = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allg, mine)
        parity["ranks_ok"] = bool(all(torch.equal(a, allg[0]) for a in allg))
        parity["ranks"] = world

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
    e2e_u8_ms = None
    with ClockSampler(local_rank) as clk:                  # sampled from the warm-up to the end of the timed regions
        for _ in range(max(args.warmup, 3)):
            step_resident()
        total_ms = timed(step_resident, args.steps)
        for _ in range(max(args.warmup, 6)):               # both staging buffers seen often enough to be graph-bound
            step_e2e()
        e2e_ms = timed(step_e2e, args.steps)
        if is_deit:
            # the same images as decoded uint8 pixels (the engine applies ToTensor + Normalize on the device): 4x fewer H2D bytes
            host_u8 = torch.randint(0, 256, host.shape, generator=g, dtype=torch.uint8).pin_memory()
            step_e2e_u8 = make_e2e(host_u8)
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
    peak_sus = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    peak_burst = 2.0 * float(peaks.get("bf16_tflops", peaks.get("bf16_tflops_sustained")))
    whole_tops = pack_int_ops(eng.meta) * B / (ms_step * 1e-3) / 1e12
    roof = {"bound": "tensor", "unit": "TFLOP/s", "peak": peak_sus, "peak_burst": peak_burst,
            "peak_source": "2 x bf16_tflops_sustained (peak) / 2 x bf16_tflops (peak_burst) of MEASURED_PEAKS.json (%s); "
                           "int8 tensor rate = 2 x bf16" % peak_src,
            "whole_step_int_tops": whole_tops, "whole_step_frac": whole_tops / peak_sus,
            "whole_step_frac_burst": whole_tops / peak_burst}
    if is_deit:
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
        nl = sum(g_["launches"] for g_ in gem)
        traffic, tsrc = ncu_traffic("gemm_i8_tcgen05_kernel")
        roof.update({"achieved": achieved, "frac": achieved / peak_sus, "frac_burst": achieved / peak_burst,
                     "traffic": traffic, "traffic_source": tsrc,
                     "timing": "CUDA events around each GEMM launch inside eager forwards (operands as produced by the preceding "
                               "kernels): closer to a kernel timed alone than to a long step, so both the sustained (frac) and the "
                               "burst (frac_burst) peak are stated",
                     "algorithmic_bytes_per_launch": sum((g_["M"] * g_["K"] + g_["N"] * g_["K"] + g_["M"] * g_["N"] * (1 if g_["name"] in ("qkv", "fc1") else 4))
                                                         * g_["launches"] for g_ in gem) / nl,
                     "kernel": "gemm_i8_tcgen05_kernel (all %d GEMM launches of the step)" % nl,
                     "per_shape": gem, "gemm_share_of_step": gms / ms_step})
    else:
        roof.update({"achieved": whole_tops, "frac": whole_tops / peak_sus, "traffic": None,
                     "kernel": "whole step (Swin: no single dominant kernel; see profiles/launches_swin_r2*.md)"})
    cpu = None
    if not args.no_cpu_baseline and world == 1:            # the CPU arm is timed at N = 1 only
        n_img = args.ref_images or 8
        v, kind, cores, sample, _ = cpu_reference_run(args.model, pack, n_img, 12.0)
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample}
    # rank 0 against the CPU oracle on the parity images (2 of the 4: the oracle port is slow)
    if not args.no_cpu_baseline and world == 1:
        import oracle.model as OM
        want = (OM.deit_forward if is_deit else OM.swin_forward)(pack, xs[:2].numpy())
        parity["oracle_ok"] = bool(np.array_equal(logits4[:2], want))
    configs = {}
    if world == 1 and not args.no_other_configs:
        for name, b in (("deit_small_patch16_224", 128), ("swin_tiny_patch4_window7_224", 128)):
            if name != args.model:
                configs[name] = other_config(name, b, dev, peaks, max(args.steps, 10))
    img_mb = B * 3 * eng.meta["img_size"] ** 2 * 4 / 1e6
    e2e = {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": int(host.numel() * 4) * world,
           "d2h_bytes_per_step": int(out_host.numel() * 4) * world, "ms_per_step": e2e_ms / args.steps}
    if e2e_u8_ms is not None:
        e2e["uint8_input"] = {"value": world * B / (e2e_u8_ms / args.steps * 1e-3), "unit": "images/s",
                              "h2d_bytes_per_step": int(host.numel()) * world, "ms_per_step": e2e_u8_ms / args.steps,
                              "note": "decoded uint8 pixels in, ToTensor + Normalize fused into the stem kernel"}
    line = {"metric": METRIC if args.model == MODEL and B == 256 else "images/sec %s INT8 bs=%d" % (args.model, B),
            "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "%s frozen INT8 forward, batch %d per GPU, 3x224x224 synthetic fp32 images" % (args.model, B),
                       "global_batch": world * B, "parallelism": "batch sharded dp%d, one-time NCCL weight broadcast" % world,
                       "l2": "no flush needed: per-step working set (fp32 input %.0f MB + activations) exceeds the 126 MB L2" % img_mb,
                       "int_ops_per_image": pack_int_ops(eng.meta)},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "configs": configs,
            "gpu_launches": eng.launches_per_forward * args.steps,
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
    ap.add_argument("--ref-images", type=int, default=0, help="images per CPU-baseline step (bounded sample; 0 = 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the DeiT-small / Swin-tiny lines of `configs`")
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
