#!/usr/bin/env python
"""bench.py -- ELBO steps/sec of the HetMOGP hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA engine through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...  CPU arm: the reference's algorithm on the host cores
                                                           (diag-only numpy port; the literal reference is O(N^2)
                                                           in memory and cannot run this workload)

Workload (config.workload): BASELINE.json configs[2] "cfg3" -- N=1e6 rows per output, M=500, Q=3 RBF latents,
T=5 outputs [HetGaussian, Bernoulli, Categorical(K=4), Gamma, Beta] (J=10 output functions), seeded synthetic data.
One step = one evaluation equivalent to SVMOGP.parameters_changed(): ELBO and ALL gradients (q(U), Z, kernel and
coregionalisation hyper-parameters).  Total N is fixed; under torchrun the rows are sharded over the ranks and the
packed sufficient statistics are summed with one NCCL all-reduce per step ("strong" scaling).

value   steps/s with X, Y and the parameters resident in HBM (device pointers through the C-ABI).
e2e     steps/s through the reference-facing call SVMOGPInf.inference(...) with HOST numpy buffers (pinned): every
        step uploads X, Y and the parameters and downloads ELBO + gradients inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--rows", type=int, default=None, help="rows per task (default: the config's N)")
    ap.add_argument("--precision", default=os.environ.get("HMOGP_PRECISION", "auto"))
    ap.add_argument("--what", default="full", choices=["full", "ve", "elbo"])
    ap.add_argument("--cpu-rows", type=int, default=5000, help="rows per task of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc=d.get("bf16_tflops_sustained", d["bf16_tflops"]), tc_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tc=1400.0, tc_burst=1590.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def algorithmic_work(N_tasks, M, Q, Xdim, what="full"):
    """U = (sum_t Q N_t) M^2 multiply-adds.  Forward quadratic form 2U flops; the weighted Gram H^1 of
    the backward pass is symmetric-aware U flops; the hyper-parameter column statistics of a full step need the
    projection again (transposed): 2U  =>  ELBO only 2U, VE step 3U, full step 5U.  Irreducible HBM bytes:
    sum_t N_t (Xdim + 1) 8."""
    P = sum(Q * n for n in N_tasks)
    U = float(P) * M * M
    fl = {"elbo": 2.0, "ve": 3.0, "full": 5.0}[what]
    return dict(U=U, flops_full=fl * U, flops_fwd=2 * U, flops_bwd_proj=2 * U, flops_gram=U,
                bytes=float(sum(N_tasks)) * (Xdim + 1) * 8)


def cpu_baseline(cfg_name, n_rows, steps=1):
    """The CPU path beside the GPU number: diag-only fp64 numpy/OpenBLAS port of the reference's algorithm
    (oracle/diag_oracle.py; arithmetic-identical to hetmogp/svmogp_inf.py for ELBO and gradients) on
    a bounded row sample of the same workload; cost is linear in N, so steps/s is extrapolated to the full N."""
    from hetmogp_b200 import synth
    from oracle import diag_oracle
    c = synth.CONFIGS[cfg_name]
    prob = synth.make_config(cfg_name, N=n_rows)
    ts = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        diag_oracle.elbo_and_grads(prob, chunk=8192)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    full_t = t * (c["N"] / float(n_rows))
    return {"value": 1.0 / full_t, "unit": "ELBO steps/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d of %d rows per task (all %d tasks), %.2f s per sample step, extrapolated linearly in N" % (n_rows, c["N"], len(c["liks"]), t),
            "sample_seconds": t}


def run_reference(args, rank, world):
    if rank != 0:
        return
    c_rows = args.cpu_rows
    base = None
    ts = []
    for i in range(args.warmup + args.steps):
        b = cpu_baseline(args.config, c_rows, steps=1)
        if i >= args.warmup:
            ts.append(b["sample_seconds"])
        base = b
    from hetmogp_b200 import synth
    c = synth.CONFIGS[args.config]
    t = float(np.mean(ts)) * (c["N"] / float(c_rows))
    base["value"] = 1.0 / t
    base["sample_seconds"] = float(np.mean(ts))
    line = {"impl": "reference", "metric": "ELBO steps/sec (ELBO + all gradients)", "value": 1.0 / t, "unit": "ELBO steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, c), "cpu_baseline": base,
            "e2e": {"value": 1.0 / t, "unit": "ELBO steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, c):
    N = args.rows or c["N"]
    return {"workload": "%s: N=%d rows/output, M=%d, Q=%d, T=%d outputs %s, Xdim=%d; step = ELBO + all gradients (%s)" % (
        args.config, N, c["M"], c["Q"], len(c["liks"]), [s[0] + (str(s[1]) if s[0] == "Categorical" else "") for s in c["liks"]], c["Xdim"], args.what),
        "l2": "working set per step (X, Y, per-row a/c and row weights, Gram partials: >0.5 GB) exceeds the 126 MB L2; "
              "a 256 MB buffer is also written between timed iterations", "seed": 1234 + int(args.config[3:])}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from hetmogp_b200 import Engine, shard_rows, synth, _lib
    from hetmogp_b200.svmogp_inf import SVMOGPInf
    from hetmogp_b200 import likelihoods as L
    from hetmogp_b200.het_likelihood import HetLikelihood
    from hetmogp_b200.gpy_shim import RBF, Coregionalize

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    prec = args.precision
    if prec == "auto":
        prec = "tc" if _lib.lib.hmogp_tc_built() else "fp32"

    c = synth.CONFIGS[args.config]
    N = args.rows or c["N"]
    prob = synth.make_config(args.config, N=N)            # same seed on every rank -> same data, each keeps its shard
    T, Q, M, Xdim = prob["T"], prob["Q"], prob["M"], prob["Xdim"]
    Ns = [x.shape[0] for x in prob["X"]]
    begin, count = shard_rows(Ns, rank, world)
    Xs = [prob["X"][t][begin[t]:begin[t] + count[t]] for t in range(T)]
    Ys = [prob["Y"][t][begin[t]:begin[t] + count[t]] for t in range(T)]
    bscale = [1.0] * T

    # ---------------------------------------------------------------- resident arm ("value")
    eng = Engine(prob["lik_specs"], M, Q, Xdim, precision=prec, device=local, group=group)
    eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    eng.set_data([torch.as_tensor(x, device=dev) for x in Xs], [torch.as_tensor(y, device=dev) for y in Ys])
    pkeys = ("Z", "m_u", "L_u", "rbf_var", "rbf_ls", "W", "kappa")
    params_dev = {k: torch.as_tensor(np.ascontiguousarray(prob[k]), device=dev) for k in pkeys}
    params_dev["batch_scale"] = torch.as_tensor(np.asarray(bscale), device=dev)
    out_dev, _ = eng._alloc_out({"elbo": 0, "ve": 1, "full": 2}[args.what], True, False)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    eng.enable_timing(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        barrier()
        if sampler:
            sampler.start()
        per = []
        phase = []
        t_wall = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1)                                   # evict L2 between timed iterations (not timed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            per.append(a.elapsed_time(b))
            phase.append(eng.last_timing())
        barrier()
        wall = time.perf_counter() - t_wall
        tot = torch.tensor([sum(per)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()) / steps, phase, wall

    sampler = ClockSampler(local) if rank == 0 else None
    ms_step, phases, wall = timed(lambda: eng.evaluate(params_dev, what=args.what, out=out_dev), args.steps, max(3, args.warmup), sampler)
    clocks = sampler.stop() if sampler else None
    elbo_resident = float(out_dev["log_marginal"].cpu()[0, 0])
    launches = int(np.sum([p["launches"] for p in phases]))

    # ---------------------------------------------------------------- end-to-end arm (host buffers through the plugin API)
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory().numpy()
        Xh, Yh = [pin(x) for x in Xs], [pin(y) for y in Ys]
        liks = HetLikelihood([L.from_spec(s) for s in prob["lik_specs"]])
        meta = liks.generate_metadata()
        kern_list = [RBF(Xdim, variance=prob["rbf_var"][q], lengthscale=prob["rbf_ls"][q]) for q in range(Q)]
        B_list = [Coregionalize(Xdim, prob["J"], 1, W=prob["W"][:, q:q + 1], kappa=prob["kappa"][:, q]) for q in range(Q)]
        m_u, L_u, Z = pin(prob["m_u"]), pin(prob["L_u"]), pin(prob["Z"])
        inf = SVMOGPInf(precision=prec, device=local, group=group)
        inf._eng, inf._key = eng, (tuple(tuple(l.spec) for l in liks.likelihoods_list), M, Q, Xdim, prec, local)
        res = {}

        def e2e_step():
            lm, grads, _, _ = inf.inference(m_u, L_u, Xh, Yh, Z, kern_list, liks, B_list, meta, batch_scale=bscale, what=args.what)
            res["lm"] = float(lm[0, 0])
        ms_e2e, _, _ = timed(e2e_step, args.steps, 3)
        h2d = sum(x.nbytes + y.nbytes for x, y in zip(Xh, Yh)) + sum(np.asarray(prob[k]).nbytes for k in pkeys) + 8 * T
        d2h = 8 * (2 + T + M * Q + (M * (M + 1) // 2) * Q + Q * M * M + 2 * Q + 2 * prob["J"] * Q + M * Q * Xdim)
        e2e = {"value": 1e3 / ms_e2e, "unit": "ELBO steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_e2e, "elbo": res.get("lm")}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- roofline of the dominant kernel (live CUDA-event times)
    peaks = load_peaks()
    work = algorithmic_work(count, M, Q, Xdim, args.what)   # this rank's shard
    med = {k: float(np.median([p[k] for p in phases])) for k in phases[0] if k.endswith("_ms")}
    # (kernel name, algorithmic flops per launch, launches per step) of the N-sized kernels behind each phase timer
    if prec == "tc":
        kern = {"forward_ms": ("tc_fwd_kernel (K_fu build + K_fu C_q on tcgen05 cta_group::2 + row reductions)", work["flops_fwd"], 1),
                "bwd_gram_ms": ("tc_gram2_kernel (H^1 = K_fu^T diag(omega) K_fu on tcgen05 cta_group::2, three-level accumulation)", work["flops_gram"], 1)}
        if args.what == "full":
            kern["bwd_proj_ms"] = ("tc_bwd_kernel (transposed projection C_q K_fu^T on tcgen05 cta_group::2 + column sums)", work["flops_bwd_proj"], 1)
    else:
        kern = {"forward_ms": ("proj_fwd (K_fu build + projection)", work["flops_fwd"], 1),
                "bwd_proj_ms": ("proj_bwd (K_fu rebuild + projection + hyper column stats)", work["flops_bwd_proj"], 1),
                "bwd_gram_ms": ("gram (K_fu^T diag(w) K_fu)", work["flops_gram"], 1)}
    per_launch = {k: med.get(k, 0.0) / kern[k][2] for k in kern}
    dom = max(kern, key=lambda k: per_launch[k])
    ach = kern[dom][1] / (per_launch[dom] * 1e-3) / 1e12 if per_launch[dom] > 0 else 0.0
    n_kernel_ms = sum(med.get(k, 0.0) for k in kern)
    Mc = -(-M // 256) * 256
    issued = {"forward_ms": 3 * 2.0 * work["U"] / (M * M) * Mc * Mc,                       # 3 split-fp16 products, padded M
              "bwd_proj_ms": 3 * 2.0 * work["U"] / (M * M) * Mc * Mc,
              "bwd_gram_ms": 3 * 2.0 * work["U"] / (M * M) * (Mc * Mc) * (0.75 if Mc % 256 == 0 else 0.625)}   # lower block-triangle of 256x256 (pair kernel) / 128x256 tiles
    traffic = {"bwd_gram_ms": 3.11e8, "forward_ms": 3.08e8, "bwd_proj_ms": 4.97e8}   # dram read+write per launch, ncu --set full (profiles/r1_tc3_ncu_summary.txt)
    roofline = {"bound": "tensor", "kernel": kern[dom][0], "achieved": ach, "peak": peaks["tc"], "unit": "TFLOP/s",
                "frac": ach / peaks["tc"],
                "traffic": traffic.get(dom) if (prec == "tc" and args.config == "cfg3" and world == 1 and not args.rows) else None,
                "peak_source": peaks["src"] + " bf16 sustained (kernel timed inside a long step)",
                "operand_format": {"tc": "split-fp16 hi/lo, 3 tcgen05.mma products per algorithmic product (fp32-class accuracy); "
                                         "the algorithmic fraction is therefore bounded by 1/3 of the bf16 peak",
                                   "fp32": "fp32 FFMA (CUDA cores)", "fp64": "fp64 DFMA"}[prec],
                "algorithmic_flops_per_launch": kern[dom][1], "ms_per_launch": per_launch[dom], "launches_per_step": kern[dom][2],
                "issued_mma_tflops": ({k: issued[k] / (med[k] * 1e-3) / 1e12 for k in kern if med.get(k, 0) > 0} if prec == "tc" else None),
                "all_kernels": {kern[k][0].split(" ")[0]: {"ms_per_launch": per_launch[k], "launches": kern[k][2],
                                                            "achieved_TFLOPs": kern[k][1] / (per_launch[k] * 1e-3) / 1e12 if per_launch[k] > 0 else 0.0}
                                for k in kern},
                "step_algorithmic_tflops": work["flops_full"] / (ms_step * 1e-3) / 1e12,
                "hbm_view": {"achieved_GBps": work["bytes"] / (n_kernel_ms * 1e-3) / 1e9, "peak_GBps": peaks["hbm"],
                             "frac": work["bytes"] / (n_kernel_ms * 1e-3) / 1e9 / peaks["hbm"],
                             "note": "path is a dense contraction (2 M^2 flops per 16 B row): not HBM-bound for M >~ 3 (SURVEY 8d)"},
                "phase_ms_median": med}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(args.config, args.cpu_rows, steps=1)
    line = {"metric": "ELBO steps/sec (ELBO + all gradients)", "value": 1e3 / ms_step, "unit": "ELBO steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": {"tc": "f16x3 (split fp16 on tcgen05, fp32 accumulate; fp64 M x M algebra)", "fp32": "f32", "fp64": "f64"}[prec],
            "data": "synthetic", "config": workload_config(args, c), "elbo": elbo_resident, "clocks": clocks,
            "gpu_launches": launches, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "wall_s_timed_region": wall}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
