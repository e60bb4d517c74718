#!/usr/bin/env python
"""bench.py -- region captions / second of the patch -> region -> caption hot path on B200.

    python bench.py --gpus 1 --steps 5 --warmup 3                    # our arm (libpio_sm100, CUDA)
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1   # the reference's CPU algorithm on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                        # one rank per GPU, images sharded, no collective

A *step* is one pass of the hot path over one batch of synthetic input of BASELINE.json configs[1]
("talk2dino_decap dense captioning: 518px images, 64 synthetic bboxes/image, batch 64"): DINOv2 ViT-B/14-reg
forward -> CLS attention map -> box pooling -> caption-memory projection (M = 591 753) -> 30-step greedy decode,
4096 region captions per GPU per step.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "region captions/sec"
UNIT = "captions/s"
BANK_ROWS = 591_753  # configs/mlp.k.yaml: support_memory_size


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--workload", default="dense", choices=["dense", "traces", "regionset", "regionset-viecap"],
                    help="dense = BASELINE configs[1] (the headline); traces = configs[2] (1 mouse trace / image, attention weighting, "
                         "batch 256); regionset = configs[3] (talk2dino_capdec, box sets -> one caption / image); regionset-viecap = configs[3] with "
                         "the ViECap captioner (mapping network + entity prompt + GPT-2 small, 64 tokens)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--boxes", type=int, default=None)
    ap.add_argument("--size", type=int, default=518)
    ap.add_argument("--bank-rows", type=int, default=BANK_ROWS)
    ap.add_argument("--pool", default="gauss", choices=["mean", "gauss", "attn"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-images", type=int, default=1)
    ap.add_argument("--cpu-sample-boxes", type=int, default=None, help="default: the workload's own boxes/image (same mix as the GPU arm)")
    ap.add_argument("--e2e-ids", action="store_true", help="e2e leg returns id tensors (return_ids=True) instead of the reference's strings")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the eager-PyTorch-on-GPU baseline (tools/gpu_eager.py)")
    ap.add_argument("--no-parity-sample", action="store_true", help="skip ids_vs_fp32_mode (bf16 captions vs the fp32 parity mode)")
    ap.add_argument("--no-sharded-bank", action="store_true", help="skip the config[4] sub-record at WORLD_SIZE > 1")
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 256 if a.workload == "traces" else 64
    if a.boxes is None:
        a.boxes = {"dense": 64, "traces": 1, "regionset": 8, "regionset-viecap": 8}[a.workload]
    if a.cpu_sample_boxes is None:
        a.cpu_sample_boxes = a.boxes
    if a.workload.startswith("regionset"):
        a.bank_rows = 0  # CapDec: no caption memory (configs/mlp_noise.k.yaml: support_memory_size 0)
    return a


def workload_flags(args):
    """forward() keywords of the workload (the eval drivers' flag mapping)."""
    if args.workload == "traces":   # eval_trace_captioning.py:309-324 (gaussian flags are passed and ignored by the trace branch)
        return dict(use_attention_tracing=True, gaussian_avg=True, gaussian_bbox_variance=1.0)
    if args.workload.startswith("regionset"):  # eval_region_set_captioning.py:322-337
        return dict(get_controllable_capts=True, gaussian_avg=True, gaussian_bbox_variance=1.0)
    return dict(gaussian_avg=args.pool == "gauss", gaussian_bbox_variance=1.0, use_attn_map_for_bboxes=args.pool == "attn")


def workload_batch(args, synth_mod, B, seed):
    """One synthetic host batch of the workload: dict of forward() inputs."""
    S = args.size
    batch = {"imgs": synth_mod.synth_images(B, S, seed=seed)}
    if args.workload == "traces":
        batch["traces"] = synth_mod.synth_traces(B, seed=seed)
    elif args.workload.startswith("regionset"):
        batch["bboxes"] = synth_mod.synth_boxes(B, args.boxes, S, seed=seed, pad="set")
    else:
        batch["bboxes"] = synth_mod.synth_boxes(B, args.boxes, S, seed=seed, pad="dense")
    return batch


def out_key(args):
    return {"dense": "bbox_capts", "traces": "trace_capts", "regionset": "set_controllable_capts",
            "regionset-viecap": "set_controllable_capts"}[args.workload]


def decode_steps(args):
    return 64 if args.workload == "regionset-viecap" else 30


def captions_per_step(args):
    return args.batch * args.boxes if args.workload == "dense" else args.batch


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """start of the window whose samples count (the sampler itself is started earlier: nvidia-smi needs ~1 s to come up)"""
        self.t0 = time.perf_counter()

    def report(self):
        """clock statistics of the window [mark(), now]; the sampler keeps running (starting a second nvidia-smi next to a
        timed region stalls the driver for tens of ms while it initialises)"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.perf_counter()
        time.sleep(0.15)  # let the sample that covers the end of the window arrive
        t0 = getattr(self, "t0", 0.0)
        lines = list(self.lines)
        window = [ln for (t, ln) in lines if t0 <= t <= t1 + 0.15]
        if not window:  # a window shorter than the sampling period: take the samples closest to it
            window = [ln for (_, ln) in lines[-2:]]
        sm, mx, reasons, power = [], [], set(), []
        for ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        out = self.report()
        if self.proc is not None:
            self.proc.terminate()
        return out


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_sample(args, steps: int, warmup: int):
    """The reference's own algorithm (oracle port: no KV cache, bank re-normalised per call, Python box loop) on the
    host cores, on a bounded sample of the same workload.  Returns (captions/s, seconds per step, description)."""
    from oracle import decap as o_decap
    from oracle import dinov2 as o_vit
    from oracle import pipeline as o_pipe

    torch.set_num_threads(os.cpu_count() or 1)
    if args.workload == "regionset-viecap":
        return cpu_sample_viecap(args, steps, warmup)
    B, R, S = args.cpu_sample_images, args.cpu_sample_boxes, args.size
    if args.workload != "dense":
        B, R = max(B, 4), args.boxes
    vit_w, dec_w = o_vit.make_weights(1234), o_decap.make_weights(1234)
    bank = o_pipe.synth_bank(args.bank_rows, 768, seed=7) if args.bank_rows > 0 else None
    model = o_pipe.OracleModel(vit_w, dec_w, bank)
    kw = workload_flags(args)
    sub = argparse.Namespace(**dict(vars(args), boxes=R))
    times = []
    for i in range(warmup + steps):
        batch = workload_batch(sub, o_pipe, B, 100 + i)
        t0 = time.perf_counter()
        model.forward(batch.pop("imgs"), get_cls_capt=False, use_cache=False, **batch, **kw)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    n = B * R if args.workload == "dense" else B
    sample = (f"{args.workload}: {B} x {S}px image(s), {R} box(es)/trace(s) each = {n} captions/step, bank M={args.bank_rows}, fp32, "
              f"reference algorithm (no KV cache, 30 steps), {len(times)} timed step(s) after {warmup} warm-up")
    return n / sec, sec, sample


def cpu_sample_viecap(args, steps: int, warmup: int):
    """configs[3] with the ViECap captioner on the host cores: oracle ViT + box-set pooling + oracle/viecap.py with the KV
    cache the reference's greedy_search uses, GPT-2 small, 64 tokens."""
    from oracle import dinov2 as o_vit
    from oracle import pipeline as o_pipe
    from oracle import pooling as o_pool
    from oracle import viecap as ov
    from patchioner_b200 import synth

    B, R, S = 2, args.boxes, args.size
    vit_w = o_vit.make_weights(1234)
    w = ov.make_weights(seed=4321, n_layer_gpt=12, n_layer_map=8)
    ents, ent_emb = synth.synth_entities(80)
    tok = ov.ToyTokenizer()
    kw = workload_flags(args)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            batch = workload_batch(args, o_pipe, B, 100 + i)
            t0 = time.perf_counter()
            d = o_vit.forward(vit_w, batch["imgs"])
            feats = o_pool.extract_bboxes_feats(d["x_norm_patchtokens"], batch["bboxes"], kw["gaussian_avg"], kw["gaussian_bbox_variance"],
                                                get_single_embedding_per_image=True, patch_size=14)
            ov.viecap_forward(w, feats.clone(), ents, ent_emb, tok, use_cache=True)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    sample = (f"regionset-viecap: {B} x {S}px images, sets of {R} boxes = {B} captions/step, fp32, reference algorithm (KV-cached "
              f"greedy search, 64 tokens, 12-layer GPT-2), {len(times)} timed step(s) after {warmup} warm-up")
    return B / sec, sec, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, sec, sample = cpu_sample(args, max(1, args.steps), max(0, args.warmup))
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample + "; each step of this arm is that bounded sample of the workload"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args):
    name = {"dense": f"talk2dino_decap dense captioning: {args.size}px images, {args.boxes} synthetic bboxes/image, batch {args.batch} per GPU "
                     f"(BASELINE.json configs[1])",
            "traces": f"talk2dino_decap trace captioning with attention weighting: {args.size}px images, 1 synthetic mouse trace/image, "
                      f"batch {args.batch} per GPU (BASELINE.json configs[2])",
            "regionset": f"talk2dino_capdec region-set captioning: {args.size}px images, sets of up to {args.boxes} boxes/image -> one "
                         f"caption per image, batch {args.batch} per GPU (BASELINE.json configs[3])",
            "regionset-viecap": f"viecap region-set captioning: {args.size}px images, sets of up to {args.boxes} boxes/image -> one caption per "
                                f"image by the ViECap captioner (8-layer mapping network, 80 synthetic entities, GPT-2 small, 64 greedy "
                                f"tokens), batch {args.batch} per GPU (BASELINE.json configs[3])"}[args.workload]
    cfg = {"workload": name,
           "images_per_gpu": args.batch, "boxes_per_image": args.boxes, "regions_per_gpu_per_step": captions_per_step(args),
           "image_size": args.size, "bank_rows": args.bank_rows, "pooling": args.pool, "decode_steps": decode_steps(args),
           "parallelism": f"dp{args.gpus} over images, no collective",
           "cache": f"per-step inputs ({args.batch * 3 * args.size * args.size * 4 / 1e6:.0f} MB of images) and activations "
                    f"({args.batch * ((args.size // 14) ** 2 + 5) * 768 * 4 * 5 / 1e6:.0f} MB per ViT layer) exceed the 126 MB L2 between "
                    "steps; no explicit flush"}
    return cfg


# ------------------------------------------------------------------------------------------------ our arm (CUDA)
def run_ours(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from patchioner_b200 import Patchioner, ops, synth

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # early: nvidia-smi takes about a second to deliver its first sample
    B, R, S = args.batch, args.boxes, args.size
    vit_w, dec_w = synth.make_vit_weights(1234), synth.make_decoder_weights(1234)
    bank = synth.synth_bank(args.bank_rows, 768, seed=7) if args.bank_rows > 0 else None
    cfg = {"decap_weights": dec_w, "prefix_size": 768, "support_memory_size": args.bank_rows,
           "dino_model": "dinov2_vitb14_reg", "normalize": True, "resize_dim": S, "crop_dim": S,
           "dino_weights": vit_w, "memory_bank": bank, "precision": args.precision}
    if args.workload == "regionset-viecap":  # configs/mlp.viecap.k.yaml
        ents, ent_emb = synth.synth_entities(80)
        cfg.update({"decap_weights": None, "normalize": False, "clip_model_name": "ViT-B/16",
                    "viecap": {"state_dict": synth.make_viecap_weights(), "entities_text": ents, "texts_embeddings": ent_emb,
                               "tokenizer": synth.WordTokenizer(), "clip_hidden_size": 768, "project_length": 10, "temperature": 0.01,
                               "top_k": 3, "threshold": 0.4, "using_hard_prompt": True, "soft_prompt_first": True,
                               "using_greedy_search": True}})
    model = Patchioner.from_config(cfg, device=dev)
    del cfg
    del bank, vit_w, dec_w
    kw = workload_flags(args)
    key = out_key(args)

    # distinct synthetic batches per rank (data-parallel shards), pinned on the host for the e2e leg
    n_sets = 2
    host = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in workload_batch(args, synth, B, 1000 * rank + i).items()}
            for i in range(n_sets)]
    resident = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values() if torch.is_tensor(v))
    if args.workload == "traces":  # trace points travel as [n, 2] doubles + offsets
        h2d_bytes += sum(len(t) for t in host[0]["traces"]) * 16 + (B + 1) * 4
    stream = torch.cuda.current_stream()

    def step_resident(i):
        return model(**resident[i % n_sets], get_cls_capt=False, return_ids=True, **kw)[key]

    e2e_marks = []

    def run_e2e(n, ids=False):
        """n steps through the public serving API: every step copies its pinned host inputs in (the copy of step i+1 is
        issued under step i's kernels: Patchioner.forward_pipelined) and hands the caller what the reference's forward
        returns -- caption STRINGS (ids -> pinned host buffer -> batched detokenisation, inside the timed region); with
        ids=True the extension return_ids=True (int32 ids read back to the host) is timed instead."""
        batches = (host[i % n_sets] for i in range(n))
        overlap = os.environ.get("PIO_E2E_OVERLAP", "1") != "0"  # A/B switch: second compute stream for batch i+1
        flags = dict(get_cls_capt=False, **kw)
        if ids:
            flags["return_ids"] = True
        for out in model.forward_pipelined(batches, overlap_compute=overlap, **flags):
            r = out[key]
            if ids:
                r.cpu()  # device -> host read of the step's result
            else:
                assert isinstance(r, list) and len(r) == B  # [B][R] strings (dense) or [B] strings
            e2e_marks.append(time.perf_counter())  # host time at which each step's result is on the host (diagnostic only)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, whole_run=False, mark=None):
        if whole_run:
            fn(warmup)
        else:
            for i in range(warmup):
                fn(i)
        barrier()
        if rank == 0:
            (mark or sampler).mark()  # clock samples count from here (the timed region of this leg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if whole_run:
            fn(steps)
        else:
            for i in range(steps):
                fn(warmup + i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    ops.reset_launch_count()
    ms_total = timed(step_resident, args.steps, args.warmup)
    launches = ops.launch_count()
    # launches counted include the warm-up steps: keep the timed share
    launches = launches * args.steps // (args.steps + args.warmup)
    clocks = sampler.report() if rank == 0 else None
    wu = max(2, args.warmup)  # >= 2 warm-up batches: both streams' scratch exists
    ms_e2e_ids = timed(lambda n: run_e2e(n, ids=True), args.steps, wu, whole_run=True)
    if args.e2e_ids:
        ms_e2e, e2e_returns = ms_e2e_ids, "int32 ids (return_ids=True)"
    else:
        del e2e_marks[:]
        ms_e2e, e2e_returns = timed(run_e2e, args.steps, wu, whole_run=True), "caption strings (the reference's return type)"
    clocks_e2e = sampler.stop() if rank == 0 else None
    last = e2e_marks[-args.steps:]
    e2e_gaps = [round((b - a) * 1e3, 2) for a, b in zip(last[:-1], last[1:])]  # host-side gaps between consecutive results

    regions = captions_per_step(args) * world
    value = regions * args.steps / (ms_total / 1e3)
    e2e_value = regions * args.steps / (ms_e2e / 1e3)

    extra = {}
    if rank == 0:
        extra = stage_breakdown(model, ops, resident[0], kw, args, stream)
    parity = None
    if rank == 0 and args.precision == "bf16" and not args.no_parity_sample and model.viecap is None:
        try:
            parity = ids_vs_fp32_mode(model, args, kw, key, resident[0], synth, dev)
        except Exception as e:  # reported next to the number, never required for it
            parity = {"error": f"{type(e).__name__}: {e}"}
    line = None
    if rank == 0:
        pk = peaks()
        gm = extra.pop("_gemm")
        roofs = extra.pop("_rooflines")
        roof = {"kernel": gm["kernel"], "bound": "tensor", "achieved": gm["tflops"], "peak": pk["bf16_tflops"] if args.precision == "bf16" else None,
                "unit": "TFLOP/s", "frac": (gm["tflops"] / pk["bf16_tflops"]) if args.precision == "bf16" else None,
                "traffic": gm.get("traffic"), "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture of this shape: profiles/dominant_kernel.json)",
                "peak_source": pk["source"] + " (burst figure: kernel timed alone, back to back)",
                "shape": gm["shape"], "avg_launch_ms": gm["ms"], "algorithmic_flops_per_launch": gm["flops"]}
        # the whole step against the sustained tensor peak: algorithmic FLOPs of ViT + projection + decode / step time
        step_flops = extra.get("_step_flops", 0.0)
        extra.pop("_step_flops", None)
        step_tflops = step_flops / (ms_total / args.steps) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic", "config": workload_config(args), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "returns": e2e_returns, "clocks": clocks_e2e,
                        "host_gaps_ms": e2e_gaps,
                        "h2d_bytes_per_step": int(h2d_bytes) * world,
                        "d2h_bytes_per_step": int(captions_per_step(args) * decode_steps(args) * 4) * world},
                "e2e_ids": {"value": regions * args.steps / (ms_e2e_ids / 1e3), "unit": UNIT, "ms_per_step": ms_e2e_ids / args.steps,
                            "returns": "int32 ids (return_ids=True)"},
                "gpu_launches": int(launches), "roofline": roof, "rooflines": roofs,
                "step_frac": {"achieved_tflops": step_tflops, "peak": pk["bf16_tflops_sustained"],
                              "frac": step_tflops / pk["bf16_tflops_sustained"] if args.precision == "bf16" else None,
                              "what": "algorithmic FLOPs of one step (ViT + projection + KV-cached decode) / ms_per_step, against the "
                                      "sustained bf16 peak (" + pk["source"] + ")"},
                "ids_vs_fp32_mode": parity, "stages": extra,
                "vit_images_per_s": extra.get("vit_images_per_s")}
        if not args.no_cpu_baseline and world == 1:
            try:
                v, sec, sample = cpu_sample(args, 1, 1)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample,
                                        "seconds_per_sample_step": sec}
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": f"failed: {e}"}
        if not args.no_gpu_eager and world == 1 and args.workload == "dense":
            try:
                model = None
                torch.cuda.empty_cache()
                line["gpu_eager_baseline"] = gpu_eager_baseline(args, synth, dev, host[0])
                best = max(v for k, v in line["gpu_eager_baseline"]["captions_per_s"].items() if v)
                line["vs_gpu_eager"] = {"vs_best_eager_variant": value / best,
                                        "vs_reference_algorithm_fp32": value / line["gpu_eager_baseline"]["captions_per_s"]["fp32_reference_algorithm"]}
            except Exception as e:
                line["gpu_eager_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    if world > 1 and not args.no_sharded_bank:
        # BASELINE configs[4]: the one path with a collective (bank rows sharded over the ranks, max / sum-exp all-reduce),
        # after the timed legs so that it cannot disturb them; rank 0 attaches the record to the line
        try:
            model = None
            resident = None
            torch.cuda.empty_cache()
            from tools.sharded_bank_nccl import run_sharded_bank

            rec = run_sharded_bank(dev, rank, world, 1_000_000, 4096, 5)
            if rank == 0:
                line["sharded_bank"] = rec
        except Exception as e:
            if rank == 0:
                line["sharded_bank"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def stage_breakdown(model, ops, batch, kw, args, stream):
    """One extra instrumented step (not the timed one): CUDA-event time per stage and for the dominant kernel."""
    from patchioner_b200 import _lib as L

    def flops_per_image(size):  # 12*(24*N*D^2 + 4*N^2*D) + 2*P*588*D
        p_ = (size // 14) ** 2
        n_ = p_ + 5
        return 12 * (24 * n_ * 768 ** 2 + 4 * n_ * n_ * 768) + 2 * p_ * 588 * 768

    def ev():
        return torch.cuda.Event(enable_timing=True)

    B, R, S = args.batch, args.boxes, args.size
    imgs = batch["imgs"]
    P = (S // 14) ** 2
    marks = [ev() for _ in range(5)]
    torch.cuda.synchronize()
    marks[0].record(stream)
    tokens, attn, _ = model.dino.forward(imgs, want_attn=True)
    marks[1].record(stream)
    patch = tokens[:, 5:]
    if args.workload == "traces":
        w = ops.trace_bins(batch["traces"], S // 14, patch.device, attn)
        feats = ops.pool_grid(patch, w.reshape(B, 1, P), 1.0 / P)[:, 0]
        n_out = 1
    elif args.workload.startswith("regionset"):
        feats = ops.pool_boxes(patch, batch["bboxes"], 14, True, 1.0, None, get_single_embedding_per_image=True)
        n_out = 1
    else:
        feats = ops.pool_boxes(patch, batch["bboxes"], 14, kw["gaussian_avg"], kw["gaussian_bbox_variance"],
                               attn if kw["use_attn_map_for_bboxes"] else None)
        n_out = R
    marks[2].record(stream)
    if model.viecap is not None:
        pre = model.viecap.prompt_embeddings(feats.reshape(-1, 768).clone())  # mapping network + entity prompt ("project" stage)
        marks[3].record(stream)
        model.viecap.gpt.decode(pre, 64)
    else:
        pre = model.embed_tokens(feats.reshape(-1, 768))
        marks[3].record(stream)
        model.decoder.decode(pre, 30)
    marks[4].record(stream)
    torch.cuda.synchronize()
    t = [marks[i].elapsed_time(marks[i + 1]) for i in range(4)]
    pk = peaks()
    vit_flops = flops_per_image(S) * B
    pool_bytes = B * (P * 768 * 4 + n_out * 768 * 4 + R * 16)
    proj_flops = 4.0 * model.im_proj.M * 768 * B * n_out if model.im_proj is not None else 0.0
    dec_flops = 2 * (4 * 12 * 768 * 768 + 50257 * 768) * 30 * B * n_out
    if model.viecap is not None:  # GPT-2 small over prompt + 63 positions, lm-head on 64 of them; mapping network counted in "project"
        proj_flops = (2 * 20 * 8 * 8 * 768 * 768 + 2 * 768 * 7680) * B * n_out
        dec_flops = (2 * 12 * 12 * 768 * 768 * (pre.shape[1] + 63) + 2 * 50257 * 768 * 64) * B * n_out
    out = {"vit_ms": t[0], "pool_ms": t[1], "project_ms": t[2], "decode_ms": t[3],
           "vit_images_per_s": B / (t[0] / 1e3), "vit_tflops": vit_flops / t[0] / 1e9,
           "pool_gbs": pool_bytes / t[1] / 1e6, "pool_frac_of_hbm": pool_bytes / t[1] / 1e6 / pk["hbm_gbs"],
           "project_tflops": proj_flops / t[2] / 1e9 if t[2] > 0 else None, "decode_tflops": dec_flops / t[3] / 1e9}
    # dominant kernel: the dense layers (tcgen05 GEMM in bf16 mode).  Time the ViT fc1 shape alone.
    N = 5 + P
    M = 64 * N  # the BASELINE configs[1] batch, whatever the workload: the committed ncu capture is of this shape
    dt = torch.bfloat16 if args.precision == "bf16" else torch.float32
    A = torch.randn(M, 768, device=imgs.device).to(dt)
    W = torch.randn(3072, 768, device=imgs.device).to(dt) / 28
    bias = torch.zeros(3072, device=imgs.device)
    Cbuf = torch.empty(M, 3072, device=imgs.device, dtype=dt)
    for _ in range(3):
        ops.linear(A, W, args.precision, bias=bias, act=L.ACT_GELU_ERF, out=Cbuf)
    e0, e1 = ev(), ev()
    reps = 10
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        ops.linear(A, W, args.precision, bias=bias, act=L.ACT_GELU_ERF, out=Cbuf)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * M * 3072 * 768
    traffic = None
    try:  # per-launch DRAM traffic of this very shape from the committed ncu capture (null for any other shape)
        dk = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel.json")))
        if dk.get("shape") == [M, 3072, 768] and args.precision == "bf16":
            traffic = dk["traffic_bytes_per_launch"]
    except Exception:
        pass
    out["_gemm"] = {"kernel": "gemm_tc2_kernel: 2-CTA tcgen05 GEMM, 256x256 pair tiles (ViT fc1 + GELU)" if args.precision == "bf16"
                    else "sgemm_tn_kernel (ViT fc1 + GELU)",
                    "shape": [M, 3072, 768], "ms": ms, "flops": flops, "tflops": flops / ms / 1e9, "traffic": traffic}
    # ---- the other kernels the north star names, each against its own roofline (SURVEY.md 8d work-per-unit figures)
    roofs = [{"kernel": out["_gemm"]["kernel"], "bound": "tensor", "achieved": out["_gemm"]["tflops"], "peak": pk["bf16_tflops"],
              "unit": "TFLOP/s", "frac": out["_gemm"]["tflops"] / pk["bf16_tflops"], "timed": "alone, 10 launches back to back"}]
    if args.precision == "bf16":
        qkv = torch.randn(B, N, 2304, device=imgs.device).to(dt)
        for _ in range(2):
            ops.vit_attention(qkv)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record(stream)
        for _ in range(5):
            ops.vit_attention(qkv)
        e1.record(stream)
        torch.cuda.synchronize()
        ms_a = e0.elapsed_time(e1) / 5
        fl_a = 4.0 * N * N * 64 * 12 * B
        roofs.append({"kernel": "vit_attention_tc_kernel (one ViT layer, all heads)", "bound": "tensor", "achieved": fl_a / ms_a / 1e9,
                      "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": fl_a / ms_a / 1e9 / pk["bf16_tflops"], "avg_launch_ms": ms_a,
                      "shape": [B, N, 12, 64], "timed": "alone, 5 launches back to back"})
        del qkv
    if proj_flops > 0 and t[2] > 0:
        roofs.append({"kernel": "caption-memory projection (similarity GEMM with exp epilogue + recombination GEMM per bank chunk)",
                      "bound": "tensor", "achieved": proj_flops / t[2] / 1e9, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                      "frac": proj_flops / t[2] / 1e9 / pk["bf16_tflops"], "ms": t[2], "timed": "stage of one instrumented step"})
    # decode: per step max(weights + KV over HBM, R * F over the tensor pipe)  
    Rr = B * n_out
    steps_d = 64 if model.viecap is not None else 30
    n_layer = 12 if model.viecap is not None else 4
    w_bytes = (n_layer * 12 * 768 * 768 + 50257 * 768) * 2
    kv_bytes = sum(2 * n_layer * 768 * 2 * tt for tt in range(1, steps_d + 1)) / steps_d * Rr  # mean over the steps
    f_tok = 2.0 * (n_layer * 12 * 768 * 768 + 50257 * 768)
    hbm_ms = (w_bytes + kv_bytes) / (pk["hbm_gbs"] * 1e6)
    tc_ms = Rr * f_tok / (pk["bf16_tflops_sustained"] * 1e9)
    floor_ms = max(hbm_ms, tc_ms) * steps_d
    roofs.append({"kernel": "greedy decode (all kernels of the %d steps)" % steps_d, "bound": "hbm" if hbm_ms >= tc_ms else "tensor",
                  "achieved": (w_bytes + kv_bytes) * steps_d / t[3] / 1e6 if hbm_ms >= tc_ms else dec_flops / t[3] / 1e9,
                  "peak": pk["hbm_gbs"] if hbm_ms >= tc_ms else pk["bf16_tflops_sustained"], "unit": "GB/s" if hbm_ms >= tc_ms else "TFLOP/s",
                  "frac": floor_ms / t[3], "ms": t[3], "floor_ms": floor_ms, "regions": Rr,
                  "what": "per step max(weight + KV bytes / HBM peak, regions x FLOPs per token / sustained tensor peak)",
                  "timed": "stage of one instrumented step"})
    roofs.append({"kernel": "region pooling (pool_box_kernel / pool_slab_kernel)", "bound": "hbm", "achieved": out["pool_gbs"],
                  "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": out["pool_frac_of_hbm"], "ms": t[1], "timed": "stage of one instrumented step"})
    out["_rooflines"] = roofs
    out["_step_flops"] = vit_flops + proj_flops + dec_flops
    if model.viecap is not None and args.precision == "bf16":
        # the kernel with the largest share of this workload's launches: the decode-step GEMM at M = regions rows (qkv shape).
        # Weight-streaming bound at this size: its algorithmic bytes are the bf16 weight panel + the activations.
        Md = B * n_out
        A2 = torch.randn(Md, 768, device=imgs.device).to(dt)
        W2 = torch.randn(2304, 768, device=imgs.device).to(dt) / 28
        b2 = torch.zeros(2304, device=imgs.device)
        C2 = torch.empty(Md, 2304, device=imgs.device, dtype=dt)
        for _ in range(3):
            ops.linear(A2, W2, args.precision, bias=b2, out=C2)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record(stream)
        for _ in range(50):
            ops.linear(A2, W2, args.precision, bias=b2, out=C2)
        e1.record(stream)
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 50
        bytes2 = (2304 * 768 + Md * 768 + Md * 2304) * 2
        out["decode_gemm"] = {"kernel": "gemm_tc_kernel<64> (decode-step qkv GEMM)", "shape": [Md, 2304, 768], "avg_launch_ms": ms2,
                              "bound": "hbm", "achieved": bytes2 / ms2 / 1e6, "peak": pk["hbm_gbs"], "unit": "GB/s",
                              "frac": bytes2 / ms2 / 1e6 / pk["hbm_gbs"], "tflops": 2.0 * Md * 2304 * 768 / ms2 / 1e9,
                              "note": "back-to-back launches of one shape (weights L2-resident); launch-latency-bound at this M"}
    return out


def ids_vs_fp32_mode(model16, args, kw, key, batch, synth, dev):
    """The bf16 number always travels with its id agreement: captions of the first >= 256 regions of the bench batch from the
    bf16 tensor-core mode vs the fp32 parity mode (the mode whose ids match the reference's on >= 99 % of regions)."""
    from patchioner_b200 import Patchioner

    per_img = args.boxes if args.workload == "dense" else 1
    n_img = min(args.batch, max(4, -(-256 // per_img)), 64)
    sub = {k: (v[:n_img] if torch.is_tensor(v) else v[:n_img]) for k, v in batch.items()}
    bank = synth.synth_bank(args.bank_rows, 768, seed=7) if args.bank_rows > 0 else None
    cfg = {"decap_weights": synth.make_decoder_weights(1234), "prefix_size": 768, "support_memory_size": args.bank_rows,
           "dino_model": "dinov2_vitb14_reg", "normalize": True, "resize_dim": args.size, "crop_dim": args.size,
           "dino_weights": synth.make_vit_weights(1234), "memory_bank": bank, "precision": "fp32"}
    m32 = Patchioner.from_config(cfg, device=dev)
    del cfg, bank
    a = model16(**sub, get_cls_capt=False, return_ids=True, **kw)[key]
    b = m32(**sub, get_cls_capt=False, return_ids=True, **kw)[key]
    a, b = a.reshape(-1, a.shape[-1]).cpu(), b.reshape(-1, b.shape[-1]).cpu()
    same = (a == b)
    prefix = same.long().cumprod(dim=1).sum(dim=1).float()
    del m32
    torch.cuda.empty_cache()
    return {"regions": int(a.shape[0]), "identical_caption_rate": float(same.all(dim=1).float().mean()),
            "mean_common_prefix": float(prefix.mean()), "steps": int(a.shape[1]),
            "first_token_agreement": float(same[:, 0].float().mean()),
            "what": "bf16 tensor-core mode vs fp32 parity mode (SIMT fp32), same inputs and random-init weights"}


def gpu_eager_baseline(args, synth, dev, host_batch):
    """SURVEY 8d / BASELINE.md 4 'also timed': the same step in eager PyTorch on this GPU (cuBLAS + SDPA, tools/gpu_eager.py),
    fp32 and bf16 autocast, with the reference's decode algorithm (no KV cache) and with a KV cache."""
    from tools.gpu_eager import EagerPipeline

    vit_w, dec_w = synth.make_vit_weights(1234), synth.make_decoder_weights(1234)
    bank = synth.synth_bank(args.bank_rows, 768, seed=7) if args.bank_rows > 0 else None
    imgs, boxes = host_batch["imgs"].to(dev), host_batch["bboxes"].to(dev)
    res, n = {}, imgs.shape[0] * boxes.shape[1]
    for name, bf16 in (("fp32", False), ("bf16_autocast", True)):
        ep = EagerPipeline(vit_w, dec_w, bank, dev, autocast_bf16=bf16)
        for algo, cache in (("reference_algorithm", False), ("kv_cache", True)):
            ep.dense_step(imgs[:8], boxes[:8], use_cache=cache)  # warm-up (cuBLAS handles, autotuning)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 1 if (not bf16 and not cache) else 2
            e0.record()
            for _ in range(reps):
                ep.dense_step(imgs, boxes, gaussian=kw_gauss(args), use_cache=cache)
            e1.record()
            torch.cuda.synchronize()
            res[f"{name}_{algo}"] = n * reps / (e0.elapsed_time(e1) / 1e3)
        del ep
        torch.cuda.empty_cache()
    return {"captions_per_s": res, "unit": UNIT,
            "sample": f"the full step: {imgs.shape[0]} x {args.size}px images, {boxes.shape[1]} boxes each = {n} captions, bank M={args.bank_rows}, "
                      "inputs resident, CUDA events, 1 small warm-up + 1-2 timed steps per variant",
            "what": "eager PyTorch (cuBLAS GEMMs, SDPA attention, stock elementwise kernels); boxes pooled by one bmm instead of the "
                    "reference's per-box Python loop; 'reference_algorithm' re-runs the growing sequence every decode step "
                    "(decap.py:130-155), 'kv_cache' is the same model with a cache"}


def kw_gauss(args):
    return args.pool == "gauss"


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
