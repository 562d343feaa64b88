#!/usr/bin/env python
"""Benchmark of the hot path: stage-2 Performer training (BASELINE.json configs[1] / configs[3]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One "step" = one pass of the hot path over one synthetic batch: embedding -> 12 x (QKV GEMM, FAVOR+
causal linear attention, out-proj, LN, FFN, LN) -> logits -> CE -> full backward -> (NCCL all-reduce
of the flat gradient when N > 1) -> global-norm clip + Adam.  Prints ONE JSON line (rank 0).

  value : tokens/s, whole job, inputs already resident in HBM, K steps between CUDA events,
          max over ranks.
  e2e   : same metric through the reference-facing module API with HOST (pinned) token buffers:
          H2D of tokens/segments/targets and D2H of the loss statistics inside the timed region.
  roofline     : the dominant kernel class, timed live with CUDA events around every launch of
                 that class inside the timed region.
  cpu_baseline : the oracle port of the reference train step timed on this box's host cores
                 (bounded sample), rank 0, N = 1 only.
  --impl reference : the CPU arm alone, same config/metric/unit (the reference has no GPU-free
                 path we can install offline: fast_transformers is absent -> oracle port).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

V_FUNCTIONAL_STAGE2 = 329      # functional-representation stage-2 vocabulary (+1 PAD)
T_SEQ = 2048
CFG = dict(n_layer=12, n_head=8, d_model=512, d_ff=2048)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "tf_burst": float(d.get("bf16_tflops", 1590.0)),
                    "tf_sustained": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PF burst / ~1.4 PF sustained)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w": statistics.median(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_traffic(kernel_substr):
    """mean DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel, from the ncu
    pass over this same command that is committed under profiles/ (profiles/r01_traffic.json, written by
    scripts/ncu_traffic.py); None when no capture for this batch size has been committed."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        return d
    except Exception:
        return None


def make_batches(n, B, T, V, seed0, pinned):
    from emo_disentanger_b200.synth import synthetic_batch
    out = []
    for i in range(n):
        tok, seg, tgt = synthetic_batch(V, B, T, seed0 + i)
        if pinned:
            tok, seg, tgt = tok.pin_memory(), seg.pin_memory(), tgt.pin_memory()
        out.append((tok, seg, tgt))
    return out


def run_reference(args):
    """CPU arm: the reference train step (oracle port) on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.cpu_train import time_cpu_train
    r = time_cpu_train(V=V_FUNCTIONAL_STAGE2, B=1, T=T_SEQ, steps=args.steps, warmup=max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": "stage2_performer_train_tokens_per_sec", "value": r["value"],
            "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": 1,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.batch, 1, "n/a (CPU arm: each step is a bounded sample of the workload -- ONE "
                                      "2048-token sequence through the same train step, see cpu_baseline.sample)"),
            "cpu_baseline": {"value": r["value"], "unit": "tokens/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference stage2 train step restated in torch fp32 (oracle port): fast_transformers (un-pinned "
                    "3rd-party dep of the reference) is not installable offline and /root/reference does not travel "
                    "to the GPU box"}
    print(json.dumps(line), flush=True)
    return 0


def bench_decode(model, V, n_tokens=384, prompt=64, cpu=True):
    """BASELINE.json's second metric: 1-GPU autoregressive decode tokens/s of the stage-2 Performer (bf16), through
    the decode engine the inference script uses: incremental FAVOR+ state (CUDA-graph step), fused temperature +
    nucleus sampler on the device, ONE int64 per sequence per step back to the host.  batch 1 = one quadrant,
    batch 4 = the four emotion quadrants of a lead-sheet pair decoded together."""
    import numpy as np
    import torch
    from emo_disentanger_b200.decode import Stage2Decoder
    from emo_disentanger_b200.generate import DeviceSampler
    model.eval()
    out = {"metric": "stage2_performer_decode_tokens_per_sec", "unit": "tokens/s", "temperature": 1.2, "top_p": 0.9,
           "prompt_tokens": prompt, "generated_tokens_per_sequence": n_tokens, "dtype": "bf16",
           "api": "Stage2Decoder.step_sample (model step + temperature/top-p sampler in one CUDA graph; host loop, one H2D of "
                  "tokens/uniforms and one D2H of the sampled ids every step)"}
    np.random.seed(0)
    for B in (1, 4):
        dec = Stage2Decoder(model, batch=B, max_len=2048)
        smp = DeviceSampler(dec.dev, rows=B)
        rng = np.random.RandomState(B)
        for b in range(B):
            dec.append(b, rng.randint(0, V - 1, size=prompt).tolist(), [0] * prompt)
        toks = [5] * B
        for phase, n in (("warm", 16), ("timed", n_tokens)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                toks, _st = dec.step_sample(toks, [1] * B, rng.random_sample(B), 1.2, 0.9)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out["batch%d" % B] = {"value": B * n_tokens / dt, "us_per_step": 1e6 * dt / n_tokens}
        del dec
    if cpu:
        # the reference loop (inference.py:252-272) re-runs the model over the WHOLE prefix for every token; timed on
        # the oracle port at one representative prefix length (cost grows linearly with the prefix)
        from oracle import performer_oracle as PO
        import os
        torch.set_num_threads(os.cpu_count() or 1)
        shapes = PO.performer_state_shapes(V, CFG["n_layer"])
        sd = PO.seeded_state(shapes, 0)
        sd["pe.pe"] = PO.sinusoid_pe(12000, 512)
        plen, reps = 512, 3
        tok = torch.randint(0, V - 1, (1, plen))
        seg = torch.ones(1, plen, dtype=torch.long)
        with torch.no_grad():
            om = [PO.draw_omega(64, 64) for _ in range(CFG["n_layer"])]
            PO.performer_forward(sd, tok, seg, om, CFG["n_layer"], 8, 512)
            t0 = time.perf_counter()
            for _ in range(reps):
                om = [PO.draw_omega(64, 64) for _ in range(CFG["n_layer"])]
                PO.performer_forward(sd, tok, seg, om, CFG["n_layer"], 8, 512)
            dt = (time.perf_counter() - t0) / reps
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "%d full-prefix forwards at prefix length %d (one per generated token, as the "
                                         "reference loop does), fp32 oracle port" % (reps, plen)}
    model.train()
    return out


def workload_config(B, world, l2note):
    return {"workload": "stage2 Performer train step: 12L d512 8h ff2048 FAVOR+ M=128, functional repr V=%d, "
                        "seq=%d, per-GPU batch %d, dropout 0.1, clip 0.5 + Adam, Omega redrawn per forward"
                        % (V_FUNCTIONAL_STAGE2, T_SEQ, B),
            "global_batch": B * world, "seq_len": T_SEQ, "per_gpu_batch": B,
            "parallelism": "dp%d" % world, "l2": l2note}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--batch", type=int, default=74,
                    help="per-GPU batch (sequences of 2048 tokens); 74 = 2 per SM pair: every GEMM and FAVOR+ launch is a "
                         "whole number of waves on 148 SMs (swept 4..74 in profiles/)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="skip the 1-GPU autoregressive decode measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps or 3
        args.warmup = args.warmup or 1
        return run_reference(args)
    args.steps = args.steps or 20
    args.warmup = max(3, args.warmup if args.warmup is not None else 5)

    # the contract is ONE JSON line on stdout: libraries that write banners to fd 1 (NCCL prints its version there)
    # are sent to stderr for the whole run; the line itself goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from emo_disentanger_b200 import _lib, ops, dp
    from emo_disentanger_b200.stage2 import MusicPerformer
    from emo_disentanger_b200.optim import FusedAdam

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    rank, local, world = dp.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, T, V = args.batch, T_SEQ, V_FUNCTIONAL_STAGE2
    torch.manual_seed(1234)                         # same init on every rank
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # the reference's "[info] model init completed" print
        model = MusicPerformer(V, CFG["n_layer"], CFG["n_head"], CFG["d_model"], CFG["d_ff"], CFG["d_model"],
                               dropout=0.1, use_segment_emb=True, n_segment_types=2, favor_feature_dims=128,
                               compute_dtype=torch.bfloat16)
    model = model.cuda(local).train()
    sync = dp.GradSync(model)
    sync.broadcast_params()
    opt = FusedAdam(model, lr=1e-4, max_grad_norm=0.5)

    nb = 4
    host = make_batches(nb, B, T, V, 1000 * (rank + 1), pinned=True)
    devb = [tuple(t.to(dev) for t in b) for b in host]

    def step_resident(i):
        tok, seg, tgt = devb[i % nb]
        acc = model.train_step(tok, seg, tgt, count_allreduce=sync.count_allreduce)
        sync.allreduce_grads()
        opt.step()
        return acc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also: find the dominant kernel class) -------------------------------------------
    ops.TIMER.enable(["gemm", "favor_fwd", "favor_bwd", "emo_ln_fwd", "emo_ln_bwd", "emo_colsum", "emo_ce_fwd_bwd",
                      "emo_embed_fwd", "emo_embed_bwd", "emo_adam_step", "emo_sumsq"])
    for i in range(args.warmup):
        if i == args.warmup - 1:
            ops.TIMER.enable(list(ops.TIMER.classes))       # keep only the last warm-up step's events
        step_resident(i)
    barrier()
    breakdown = ops.TIMER.summary()
    tot_ms = sum(v[1] for v in breakdown.values()) or 1.0
    shares = {k: round(v[1] / tot_ms, 4) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1][1])}
    classes = {"gemm": breakdown["gemm"][1], "favor": breakdown["favor_fwd"][1] + breakdown["favor_bwd"][1]}
    dominant = max(classes, key=classes.get)
    timed_classes = ["gemm"] if dominant == "gemm" else ["favor_fwd", "favor_bwd"]

    # ---- timed region: value ---------------------------------------------------------------------
    ops.TIMER.enable(timed_classes)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        acc = step_resident(i)
    e1.record()
    barrier()
    ck = clocks.stop() if rank == 0 else None
    launches = _lib.LAUNCHES - launches0
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms[0])
    ksum = ops.TIMER.summary()
    ops.TIMER.disable()
    loss_last = float(acc[1] / acc[0]) if world == 1 else None
    value = B * T * world * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel class -----------------------------------------------------
    pk = peaks()
    if dominant == "gemm":
        n, kms, flops = ksum["gemm"]
        ach = flops / (kms * 1e-3) / 1e12
        roof = {"kernel": "gemm_tc_kernel (tcgen05 bf16 GEMM: QKV/out/FFN/logits fwd, dgrad, wgrad)",
                "bound": "tensor", "achieved": round(ach, 2), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / pk["tf_sustained"], 4), "traffic": None, "launches": n,
                "algorithmic_flops_per_launch": round(flops / max(n, 1)),
                "avg_launch_us": round(1e3 * kms / max(n, 1), 2), "share_of_step": round(kms / (ms or 1), 4),
                "work_per_launch": "2*M*N*K flops of each launch, summed (%.3f GFLOP/token)" % (flops / (B * T * args.steps) / 1e9),
                "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)"}
    else:
        n = ksum["favor_fwd"][0] + ksum["favor_bwd"][0]
        kms = ksum["favor_fwd"][1] + ksum["favor_bwd"][1]
        byts = ksum["favor_fwd"][2] + ksum["favor_bwd"][2]
        ach = byts / (kms * 1e-3) / 1e9
        roof = {"kernel": "favor_fwd_kernel + favor_bwd_kernel (FAVOR+ feature map + causal prefix-sum)",
                "bound": "hbm", "achieved": round(ach, 2), "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(ach / pk["hbm_gbs"], 4), "traffic": None, "launches": n,
                "avg_launch_us": round(1e3 * kms / max(n, 1), 2), "share_of_step": round(kms / (ms or 1), 4),
                "work_per_launch": "bf16 bytes: fwd 4*512*2 B/token/layer (q,k,v in, out), bwd 8*512*2 B/token/layer",
                "peak_source": pk["source"]}

    tr = measured_traffic("gemm_tc_kernel")
    if tr and tr.get("per_gpu_batch") == B and dominant == "gemm":
        roof["traffic"] = tr["gemm_mean_dram_bytes_per_launch"]
        roof["traffic_source"] = tr.get("source")

    # ---- e2e: host buffers, H2D + D2H inside the timed region -----------------------------------------
    def step_e2e(i):
        tok, seg, tgt = (t.to(dev, non_blocking=True) for t in host[i % nb])
        acc = model.train_step(tok, seg, tgt, count_allreduce=sync.count_allreduce)
        sync.allreduce_grads()
        opt.step()
        return acc.cpu()                   # D2H read of [count, loss_sum, n_correct] (synchronises the step)

    for i in range(2):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        stats = step_e2e(i)
    e1.record()
    barrier()
    ems = e0.elapsed_time(e1)
    tms = torch.tensor([ems], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ems = float(tms[0])
    e2e = {"value": B * T * world * args.steps / (ems * 1e-3), "unit": "tokens/s",
           "h2d_bytes_per_step": 3 * B * T * 8 * world, "d2h_bytes_per_step": 12 * world,
           "ms_per_step": ems / args.steps,
           "api": "MusicPerformer.train_step(tok, seg, tgt) + FusedAdam.step() on pinned-host int64 batches"}

    line = {"metric": "stage2_performer_train_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(B, world, "per-step working set (bf16 weights 76 MB + ~%.1f GB saved activations) "
                                      "exceeds the 126 MB L2; no explicit flush" % (B * T * 148e3 / 1e9)),
            "e2e": e2e, "gpu_launches": launches, "roofline": roof, "kernel_time_shares": shares,
            "clocks": ck, "loss_last_step": loss_last, "impl": "b200"}

    if rank == 0 and world == 1 and not args.no_decode:
        del devb, opt, sync
        model.zero_grad()
        torch.cuda.empty_cache()
        line["decode"] = bench_decode(model, V, cpu=not args.no_cpu_baseline)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_train import time_cpu_train
        r = time_cpu_train(V=V, B=1, T=T, steps=3, warmup=1)
        line["cpu_baseline"] = {"value": r["value"], "unit": "tokens/s", "cores": r["cores"], "kind": "port",
                                "sample": r["sample"]}
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    return 0


if __name__ == "__main__":
    sys.exit(main())
