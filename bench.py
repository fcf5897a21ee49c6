#!/usr/bin/env python
"""bench.py -- TLR GEMM tile-GEMMs/s on B200 (BASELINE.json metric), one process per GPU.

A "step" is one full pass of the hot path over the workload: C = A*B on compressed tiles, i.e. for every C tile the
sequential k-sum of HCore::Gemm calls with recompression (examples/matrix_multiplication/omp_main.cpp:112-126).
Workload at N=1: BASELINE.json configs[2] -- 16384 x 16384 double, tile 1024, accuracy 1e-8, compressed A, B, C
(16^3 = 4096 tile-GEMMs per step).  N>1: weak scaling, C tiles 2D block-cyclic over a P x Q grid (16P x 16Q C tiles,
k = 16), A row-panels / B column-panels moved by NCCL broadcast, no other data-path collective.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched through torch.distributed.run)
  python bench.py --impl reference ...                      times the reference's own CPU path (oracle/_ref) instead

Prints ONE JSON line (rank 0).  The oracle / compiled reference is used here only as checker and as the reported CPU
baseline -- never inside the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tlr_gemm_tile_gemms_per_s"
UNIT = "tile-GEMM/s"


# --------------------------------------------------------------------------------------------------------------------
_JSON_OUT = None  # private copy of the process's stdout, see main()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles", type=int, default=16, help="tiles per matrix dimension per GPU-grid unit")
    ap.add_argument("--nb", type=int, default=1024)
    ap.add_argument("--acc", type=float, default=1e-8)
    ap.add_argument("--rank", type=int, default=0, help="rank of the synthetic A/B tiles (0: from the spectrum law)")
    ap.add_argument("--kc-bound", type=int, default=0,
                    help="rank bound used to size scratch for C tiles (0: calibrate with one untimed pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    ap.add_argument("--compress-tiles", type=int, default=64,
                    help="initial-compression leg (SURVEY.md 8d: reported separately): dense tiles compressed (0: skip)")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0)
    return ap.parse_args()


def spectrum(nb, dtype=np.float64):
    """Reference generator law (src/helpers/generators/LatmsGenerator.cpp:36-53)."""
    eps = float(np.finfo(dtype).eps)
    sep = eps * 10
    i = np.arange(nb, dtype=np.float64)
    s = (sep ** (1.0 / 80.0)) ** i
    if nb > 80:
        b2 = (eps / sep) ** (1.0 / (nb - 1 - 80))
        s = np.where(i < 80, s, sep * b2 ** (i - 80))
    return s


def rank_for_accuracy(nb, acc):
    """What CalculateNewRank (omp/kernels.cpp:82-104) keeps of the law: first i >= 1 with sigma_i < acc."""
    s = spectrum(nb)
    for i in range(1, nb):
        if s[i] < acc:
            return i
    return nb


def grid_shape(n):
    p = int(math.sqrt(n))
    while n % p:
        p -= 1
    return p, n // p  # P x Q, P <= Q  (1x1, 1x2, 2x2, 2x4)


def flops_ccc(nb, ka, kb, kc, rk):
    """SURVEY.md 8d closed forms (implicit apply-Q): (contraction flops, recompression flops) for one tile-GEMM."""
    r = kc + ka
    f_contr = 2 * ka * nb * kb + 2 * ka * kb * nb + 2 * r ** 3
    f_recomp = 2 * (2 * nb * r * r - (2 * r ** 3) / 3) + 22 * r ** 3 + 2 * (4 * nb * r * rk - 2 * r * r * rk)
    return f_contr, f_recomp


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak():
    """The path computes in f64: no tcgen05 kind exists for it, the tensor instruction is DMMA.  MEASURED_PEAKS.json
    only holds bf16, so the yardstick is cuBLAS DGEMM measured on this pool's B200 by scripts/fp64_peak.py."""
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_yardstick.json")))
        return float(p["dgemm_tflops"]), "measured cuBLAS DGEMM 8192^3 f64 (profiles/r01_fp64_yardstick.json; MEASURED_PEAKS.json has no f64 entry)"
    except Exception:
        return 35.5, "fallback: cuBLAS DGEMM measured in round 1 (35.5 TFLOP/s)"


def ncu_traffic(kernel):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (or None)."""
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        return p.get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------------------------
# reference CPU arm / baseline (oracle/_ref = the reference's own sources compiled unmodified)
# --------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(args, Uc_A, Vc_A, Uc_B, Vc_B, rank, budget_s, want_outputs=False):
    """Times the reference's tile loop (omp_main.cpp:112-126 via oracle/ref.py) on a bounded sample of the SAME
    workload: the first `cols` block-columns of C (all k), all host threads, serial BLAS inside (SURVEY.md 8c iii).
    Uc_*/Vc_*: host arrays (ntiles, rank, nb) / (ntiles, nb, rank), tile order lin = row + col*T."""
    from oracle import ref as R
    T, nb = args.tiles, args.nb
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, R.lib().hcref_max_threads()))
    # ~60 ms per tile-GEMM per core at nb = 1024 with race-free inputs (C ranks grow to ~313; measured: 4096 tile-GEMMs in
    # 15.5 s on 16 cores): size the sample for the budget, whole block-columns
    est_per_gemm = 60e-3 * (nb / 1024.0) ** 1.2
    cols = int(max(1, min(T, budget_s * threads / (est_per_gemm * T * T))))
    p = R.Params(args.acc)
    tileA = lambda j, k: R.RefTile.from_uv(Uc_A[j + k * T].T, Vc_A[j + k * T].T)
    tileB = lambda k, i: R.RefTile.from_uv(Uc_B[k + i * T].T, Vc_B[k + i * T].T)
    A = [[tileA(j, k) for k in range(T)] for j in range(T)]
    B = [[tileB(k, i) for i in range(cols)] for k in range(T)]
    z_u, z_v = np.zeros((nb, 1)), np.zeros((1, nb))
    Cg = [[R.RefTile.from_uv_cap(z_u, z_v, max(nb // 3, 1)) for _ in range(cols)] for _ in range(T)]
    sec, flops = R.matmul(A, B, Cg, 1.0, 1.0, p, nthreads=threads)
    n_gemms = T * cols * T
    out = {"seconds": sec, "tile_gemms": n_gemms, "value": n_gemms / sec, "cores": threads, "cols": cols,
           "ref_flop_counter": flops}
    if want_outputs:
        out["ranks"] = np.array([[t.info()["rank"] for t in r] for r in Cg])
        out["tiles"] = Cg
    return out


# --------------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    # stdout carries exactly ONE line (the JSON): everything any library writes to fd 1 (NCCL prints its version banner
    # there when NCCL_DEBUG is set in the environment) is sent to stderr; the JSON goes to a private copy of fd 1.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank_env = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    krank = args.rank or rank_for_accuracy(args.nb, args.acc)
    T, nb = args.tiles, args.nb

    if args.impl == "reference":
        if rank_env != 0:
            return 0
        run_reference_arm(args, krank)
        return 0

    import torch
    import torch.distributed as dist
    import hcorepp_b200 as hc
    from hcorepp_b200 import _capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hcorepp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner / debug lines go to stderr: stdout is the ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, Q = grid_shape(world)
    pr, pc = rank_env // Q, rank_env % Q
    ctx = hc.RunContext(local)
    dt = torch.float64
    prm = hc.CompressionParameters(args.acc)
    dev = ctx.device

    # ---- synthetic inputs, generated directly in compressed form with the reference spectrum law:
    #      tile = Q_u diag(sigma_0..k-1) Q_v^T, Haar-like Q from QR of Gaussians, deterministic per (matrix, tile).
    sig = torch.from_numpy(spectrum(nb)[:krank].copy()).to(dev)

    def synth(n_tiles, seed):
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        U = torch.empty(n_tiles, krank, nb, dtype=dt, device=dev)   # column-major nb x k per tile
        V = torch.empty(n_tiles, nb, krank, dtype=dt, device=dev)   # column-major k x nb per tile
        for c0 in range(0, n_tiles, 64):
            c1 = min(n_tiles, c0 + 64)
            qu, _ = torch.linalg.qr(torch.randn(c1 - c0, nb, krank, generator=g, dtype=dt, device=dev))
            qv, _ = torch.linalg.qr(torch.randn(c1 - c0, nb, krank, generator=g, dtype=dt, device=dev))
            U[c0:c1] = qu.transpose(1, 2)
            V[c0:c1] = qv * sig[None, None, :]                       # V = diag(sigma) Qv^T, stored (n, k) row-major
        return U, V

    cap_in = krank  # A/B tiles never grow: tight capacity (the reference's from-U/V constructor does the same)
    if world == 1:
        mt = nt = kt = T
        Ua, Va = synth(mt * kt, 1)
        Ub, Vb = synth(kt * nt, 2)
        A = hc.TileMatrix(mt, kt, nb, nb, dt, ctx, compressed=True, max_rank=cap_in, rank_bound=krank)
        B = hc.TileMatrix(kt, nt, nb, nb, dt, ctx, compressed=True, max_rank=cap_in, rank_bound=krank)
        A.load_factors(Ua, Va, krank)
        B.load_factors(Ub, Vb, krank)
        Cm = hc.TileMatrix.zeros_compressed(mt, nt, nb, nb, dt, ctx, rank_bound=args.kc_bound)
        n_local_gemms = mt * nt * kt
        info = torch.zeros(mt * nt, dtype=torch.int32, device=dev)
        if args.kc_bound == 0:
            # untimed calibration pass with the safe bound (max_rank): learn how far the C ranks grow, then size the
            # scratch / grids of the timed passes from that (+ margin). A violated bound is detected on the device.
            hc.tile_matrix_multiplication(A, B, Cm, 1.0, 1.0, ctx, prm, info=info)
            ctx.Sync()
            args.kc_bound = int(min(Cm.max_rank, (int(Cm.ranks.max().item()) + 8 + 7) // 8 * 8))
            Cm.set_rank_bound(args.kc_bound)

        def one_pass():
            Cm.reset_to_zero()
            hc.tile_matrix_multiplication(A, B, Cm, 1.0, 1.0, ctx, prm, info=info)
    else:
        one_pass, Cm, info, n_local_gemms, A, B, local_factors = setup_distributed(args, hc, ctx, dist, torch, synth, krank, cap_in, P, Q,
                                                                    pr, pc, prm)
        verify = lambda: verify_distributed(args, torch, hc, synth, Cm, krank, P, Q, pr, pc)
        if args.kc_bound == 0:  # untimed calibration pass, bound agreed across ranks
            one_pass()
            ctx.Sync()
            mx = Cm.ranks.max().to(torch.int64)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            args.kc_bound = int(min(Cm.max_rank, (int(mx.item()) + 8 + 7) // 8 * 8))
            Cm.set_rank_bound(args.kc_bound)
    total_gemms = n_local_gemms * world

    # ---- warm-up (also grows the scratch arena once), then the timed region
    for _ in range(max(args.warmup, 3)):
        one_pass()
    ctx.Sync()
    _capi.lib.hcb_launch_count_reset()
    _capi.check(_capi.lib.hcb_ctx_phase_timing(ctx.h, 1))
    sampler = ClockSampler(local)
    if rank_env == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank_env == 0 else None
    launches = int(_capi.lib.hcb_launch_count())
    import ctypes as C
    ph_ms = (C.c_double * _capi.N_PHASES)()
    ph_n = (C.c_uint64 * _capi.N_PHASES)()
    _capi.check(_capi.lib.hcb_ctx_phase_times(ctx.h, ph_ms, ph_n))
    _capi.check(_capi.lib.hcb_ctx_phase_timing(ctx.h, 0))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    bad = int((info & 5).max().item())  # 1: Jacobi not converged, 4: rank bound exceeded (2 = clipped at maxRank is legal)
    sweeps = int((info >> 8).max().item())
    ms_per_step = ms / args.steps
    value = total_gemms / (ms_per_step * 1e-3)

    # ---- rank trace (untimed): kc before / rk after / Jacobi sweeps of every k-step, for the flop and byte accounting
    if world == 1:
        Cm.reset_to_zero()
        kc_hist, rk_hist, sw_hist = [], [], []
        for k in range(T):
            kc_hist.append(Cm.ranks.clone())
            hc.tile_matrix_multiplication(A, B, Cm, 1.0, 1.0, ctx, prm, k_range=(k, k + 1), info=info)
            rk_hist.append(Cm.ranks.clone())
            sw_hist.append(((info >> 8) & 0xff).clone())
        ctx.Sync()
        kc_all = torch.stack(kc_hist).cpu().numpy().astype(np.float64)
        rk_all = torch.stack(rk_hist).cpu().numpy().astype(np.float64)
        sw_all = torch.stack(sw_hist).cpu().numpy().astype(np.float64)
    else:
        kc_all = rk_all = sw_all = None

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (compressed tiles U diag(sigma) V^T, reference LATMS spectrum law, rank %d)" % krank,
        "config": {"workload": "TLR GEMM %dx%d f64, tile %d, acc %.0e, compressed A,B,C, %d tile-GEMMs/step%s" % (
            T * nb * (P if world > 1 else 1), T * nb * (Q if world > 1 else 1), nb, args.acc, total_gemms,
            "" if world == 1 else ", 2D block-cyclic %dx%d + NCCL panel broadcast" % (P, Q)),
            "baseline_config": "BASELINE.json configs[2]" if (T, nb, world) == (16, 1024, 1) else "custom",
            "l2": "inputs larger than L2 (A+B live factors %.0f MB per GPU, C scratch re-written every k)" % (
                2 * T * T * 2 * nb * krank * 8 / 1e6)},
        "library": _capi.lib.hcb_version().decode(),
        "gpu_launches": launches, "jacobi_or_bound_flags": bad, "jacobi_sweeps_last_step_max": sweeps,
        "c_rank_bound": args.kc_bound,
    }
    if world > 1 and not args.no_e2e:
        # end to end at N GPUs: every rank uploads ITS A / B tiles from pinned host memory, runs the pass (panel
        # broadcasts included) and reads ITS C tiles (ranks + live factors) back; barrier on both sides, max over ranks
        e2e_ms, h2d, d2h = run_e2e_distributed(args, torch, dist, ctx, A, B, Cm, local_factors, krank, one_pass)
        if rank_env == 0:
            result["e2e"] = {"value": total_gemms / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                             "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world}
    if world > 1 and rank_env == 0:
        result["parity"] = verify()
    if rank_env == 0:
        result["clocks"] = clocks
        phases = {}
        for i in range(_capi.N_PHASES):
            nm = _capi.lib.hcb_phase_name(i).decode()
            phases[nm] = {"ms_per_step": ph_ms[i] / args.steps, "launches_per_step": ph_n[i] / args.steps}
        result["phases"] = phases
        hbm, src = peaks()
        if kc_all is not None:
            # algorithmic traffic / flops per step from the true ranks
            ka = kb = krank
            r = kc_all + ka
            fc, fr = flops_ccc(nb, ka, kb, kc_all, rk_all)
            b_qr = 8 * (2 * 2 * nb * r)                 # read both stacks + write both reflector panels, per tile-GEMM
            b_recomp = 8 * (2 * nb * (3 * r + rk_all) + 6 * r * r)
            dom = max(("panel_qr", "jacobi_svd", "core_lq", "apply_q", "contraction", "stack"), key=lambda n: phases[n]["ms_per_step"])
            t_dom = phases[dom]["ms_per_step"] * 1e-3
            launch_ms = phases[dom]["ms_per_step"] / max(phases[dom]["launches_per_step"], 1)
            # algorithmic flops of the FP64-bound phases (DESIGN.md "rooflines"): Jacobi = sweeps * r(r-1)/2 rotations
            # of 6r flops; QR = 2 stacks of 2 m r^2 - 2/3 r^3; rebuild = 2 sides of 4 m r rk - 2 r^2 rk
            alg_flops = {"jacobi_svd": (sw_all * (r * (r - 1) / 2) * 6 * r).sum(),
                         "panel_qr": (2 * (2 * nb * r * r - 2.0 / 3.0 * r ** 3)).sum(),
                         "apply_q": (2 * (4 * nb * r * rk_all - 2 * r * r * rk_all)).sum()}
            if dom in alg_flops:
                pk, pk_src = fp64_peak()
                ach = float(alg_flops[dom]) / t_dom / 1e12
                tr = ncu_traffic(dom)
                result["roofline"] = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk, "unit": "TFLOP/s",
                                      "frac": ach / pk, "traffic": tr, "peak_source": pk_src, "launch_ms": launch_ms,
                                      "note": "f64 path: flop-bound (C ranks grow to ~313, r = kc + 44); the FP64 pipe "
                                              "/ DMMA rate is the ceiling, HBM traffic is <1% of peak"}
            else:
                alg = {"core_lq": (8 * 4 * r * r).sum(), "contraction": (8 * (2 * nb * (ka + kb) + nb * r)).sum(),
                       "stack": (8 * 2 * 2 * nb * r).sum()}[dom]
                ach = alg / t_dom / 1e9
                result["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s",
                                      "frac": ach / hbm, "traffic": ncu_traffic(dom), "peak_source": src, "launch_ms": launch_ms}
            t_rec = sum(phases[n]["ms_per_step"] for n in ("stack", "panel_qr", "core_lq", "jacobi_svd", "vsigma_truncate",
                                                           "apply_q", "finalize")) * 1e-3
            t_con = phases["contraction"]["ms_per_step"] * 1e-3
            result["effective"] = {
                "dense_equivalent_gflops": 2.0 * (T * nb) ** 3 / (ms_per_step * 1e-3) / 1e9,
                "lowrank_gflops": float((fc + fr).sum()) / (ms_per_step * 1e-3) / 1e9,
                "recompression_gbs_of_hbm": float(b_recomp.sum()) / t_rec / 1e9 / hbm,
                "recompression_gflops": float(fr.sum()) / t_rec / 1e9,
                "contraction_gflops": float(fc.sum()) / max(t_con, 1e-9) / 1e9,
                "c_rank_final_mean": float(rk_all[-1].mean()), "c_rank_max": float(rk_all.max()),
                # the three FP64-bound recompression phases against the measured FP64 peak (algorithmic flops, true ranks)
                "phase_fp64_tflops": {n: float(alg_flops[n]) / (phases[n]["ms_per_step"] * 1e-3) / 1e12 for n in alg_flops},
                "phase_fp64_frac_of_peak": {n: float(alg_flops[n]) / (phases[n]["ms_per_step"] * 1e-3) / 1e12 / fp64_peak()[0]
                                            for n in alg_flops},
            }
        # ---- end-to-end: same pass through the public API with HOST buffers (pinned), H2D + D2H inside the timing
        if world == 1 and not args.no_e2e:
            result["e2e"] = run_e2e(args, torch, hc, ctx, A, B, Cm, Ua, Va, Ub, Vb, krank, prm, info, total_gemms)
            if args.compress_tiles > 0:
                try:
                    result["initial_compression"] = run_compression(args, torch, hc, ctx, prm, not args.no_cpu_baseline)
                except Exception as e:  # a separate, reported-only leg
                    result["initial_compression"] = {"error": repr(e)}
            if not args.no_cpu_baseline:
                try:
                    result["cpu_baseline"], result["parity"] = run_cpu_baseline(args, torch, hc, ctx, A, B, Cm, Ua, Va, Ub, Vb,
                                                                                krank, prm)
                except Exception as e:  # the baseline is reported, never required for the GPU number
                    result["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(result), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(args, torch, hc, ctx, A, B, Cm, Ua, Va, Ub, Vb, krank, prm, info, total_gemms):
    """Inputs start in pinned HOST memory each step; the step copies them to the device, runs the product and reads
    the result (ranks + live factors of C) back to pinned host memory."""
    T, nb = args.tiles, args.nb
    hUa, hVa, hUb, hVb = (x.cpu().pin_memory() for x in (Ua, Va, Ub, Vb))
    cap = Cm.max_rank
    h_ranks = torch.empty(T * T, dtype=torch.int32).pin_memory()
    kcb = args.kc_bound
    hU = torch.empty(T * T, nb * kcb, dtype=torch.float64).pin_memory()
    hV = torch.empty(T * T, nb * kcb, dtype=torch.float64).pin_memory()
    h2d = sum(x.numel() * x.element_size() for x in (hUa, hVa, hUb, hVb))

    def step():
        A.load_factors(hUa, hVa, krank)
        B.load_factors(hUb, hVb, krank)
        Cm.reset_to_zero()
        hc.tile_matrix_multiplication(A, B, Cm, 1.0, 1.0, ctx, prm, info=info)
        h_ranks.copy_(Cm.ranks, non_blocking=True)
        v = Cm.buf.view(T * T, Cm.tile_elems)
        hU.copy_(v[:, : nb * kcb], non_blocking=True)
        hV.copy_(v[:, nb * cap: nb * cap + nb * kcb], non_blocking=True)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    d2h = h_ranks.numel() * 4 + (hU.numel() + hV.numel()) * 8
    return {"value": total_gemms / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h)}


def run_e2e_distributed(args, torch, dist, ctx, A, B, Cm, factors, krank, one_pass):
    """N > 1 end-to-end leg: per step H2D of the rank's own A / B factor stacks (pinned), the distributed pass, D2H of the
    rank's C tiles.  Returns (ms per step as max over ranks, H2D bytes, D2H bytes per rank and step)."""
    T, nb = args.tiles, args.nb
    hUa, hVa, hUb, hVb = (x.cpu().pin_memory() for x in factors)
    cap, kcb = Cm.max_rank, args.kc_bound
    nC = Cm.mt * Cm.nt
    h_ranks = torch.empty(nC, dtype=torch.int32).pin_memory()
    hU = torch.empty(nC, nb * kcb, dtype=torch.float64).pin_memory()
    hV = torch.empty(nC, nb * kcb, dtype=torch.float64).pin_memory()

    def step():
        A.load_factors(hUa, hVa, krank)
        B.load_factors(hUb, hVb, krank)
        one_pass()
        h_ranks.copy_(Cm.ranks, non_blocking=True)
        v = Cm.buf.view(nC, Cm.tile_elems)
        hU.copy_(v[:, : nb * kcb], non_blocking=True)
        hV.copy_(v[:, nb * cap: nb * cap + nb * kcb], non_blocking=True)
    step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=Cm.buf.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h2d = sum(x.numel() * x.element_size() for x in (hUa, hVa, hUb, hVb))
    d2h = h_ranks.numel() * 4 + (hU.numel() + hV.numel()) * 8
    return float(t.item()), h2d, d2h


def run_cpu_baseline(args, torch, hc, ctx, A, B, Cm, Ua, Va, Ub, Vb, krank, prm):
    """Reference CPU path (oracle/_ref, kind 'reference') on a bounded sample of the same workload + parity of the GPU
    result against it on that sample."""
    T, nb = args.tiles, args.nb
    hUa, hVa, hUb, hVb = (x.cpu().numpy() for x in (Ua, Va, Ub, Vb))
    ref = cpu_reference_sample(args, hUa, hVa, hUb, hVb, krank, args.cpu_budget_s, want_outputs=True)
    cols = ref["cols"]
    Cm.reset_to_zero()
    hc.tile_matrix_multiplication(A, B, Cm, 1.0, 1.0, ctx, prm)
    ctx.Sync()
    g_ranks = Cm.rank_table()[:, :cols]
    num = den = 0.0
    for i in range(min(cols, 2)):          # dense reconstruction of up to 2 block-columns
        for j in range(T):
            d_ref = ref["tiles"][j][i].to_dense()
            d_gpu = Cm.GetTile(j, i).to_dense()
            num += np.linalg.norm(d_gpu - d_ref) ** 2
            den += np.linalg.norm(d_ref) ** 2
    parity = {"rel_fro_err_vs_reference": math.sqrt(num / den), "tolerance": 10 * args.acc,
              "max_rank_diff": int(np.max(np.abs(g_ranks - ref["ranks"]))), "c_tiles_compared": int(T * min(cols, 2)),
              "ranks_compared": int(T * cols)}
    parity["pass"] = bool(parity["rel_fro_err_vs_reference"] <= parity["tolerance"] and parity["max_rank_diff"] <= 1)
    base = {"value": ref["value"], "unit": UNIT, "cores": ref["cores"], "kind": "reference",
            "sample": "first %d of %d block-columns of C, all k (%d of %d tile-GEMMs), %.1f s" % (
                cols, T, ref["tile_gemms"], T ** 3, ref["seconds"])}
    return base, parity


def run_compression(args, torch, hc, ctx, prm, with_reference):
    """Initial compression (the compressing constructor, Compressed.cpp:75-146; TileMatrix.cpp:150-171) of dense nb x nb
    tiles that follow the reference generator's full spectrum law: ONE batched device call (sketched range finder + small
    Jacobi SVD + rank rule on the device; full SVD fallback for flat spectra).  Reported separately from the GEMM metric; the reference's own
    constructor is timed on a 2-tile sample of the same tiles and its ranks / reconstruction compared."""
    nb, n = args.nb, args.compress_tiles
    dev = ctx.device
    g = torch.Generator(device=dev)
    g.manual_seed(4242)
    sig = torch.from_numpy(spectrum(nb)).to(dev)
    qu, _ = torch.linalg.qr(torch.randn(n, nb, nb, generator=g, dtype=torch.float64, device=dev))
    qv, _ = torch.linalg.qr(torch.randn(n, nb, nb, generator=g, dtype=torch.float64, device=dev))
    tiles = (qu * sig[None, None, :]) @ qv.transpose(1, 2)            # (n, nb, nb)
    raw = tiles.permute(1, 0, 2).reshape(nb, n * nb).contiguous()     # one block-row of n tiles
    hc.TileMatrix.from_dense(raw, nb, nb, ctx, prm)                    # warm-up (scratch arena, module load)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tm = hc.TileMatrix.from_dense(raw, nb, nb, ctx, prm)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ranks = tm.rank_table().reshape(-1)
    out = {"tiles": n, "nb": nb, "ms_total": ms, "tiles_per_s": n / (ms * 1e-3), "ranks": [int(r) for r in ranks],
           "algorithmic_bytes_per_tile": int(8 * (nb * nb + 2 * nb * int(ranks.max())))}
    if with_reference:
        from oracle import ref as R
        p = R.Params(args.acc)
        t0 = time.time()
        errs, rdiff, k = [], [], min(2, n)
        for t in range(k):
            a = np.asfortranarray(tiles[t].cpu().numpy())
            rt = R.RefTile.compress(a, p)
            d_ref = rt.to_dense()
            d_gpu = tm.GetTile(0, t).to_dense()
            errs.append(float(np.linalg.norm(d_gpu - d_ref) / np.linalg.norm(d_ref)))
            rdiff.append(abs(int(rt.info()["rank"]) - int(ranks[t])))
        out["reference_cpu"] = {"tiles": k, "s_per_tile": (time.time() - t0) / k, "cores": 1,
                                "rel_fro_err_vs_reference": max(errs), "max_rank_diff": max(rdiff),
                                "pass": bool(max(errs) <= 10 * args.acc and max(rdiff) <= 1)}
    return out


def run_reference_arm(args, krank):
    """--impl reference: the reference's own CPU implementation (oracle/_ref) on the host cores, same metric/config."""
    from oracle import tlr_oracle as O
    T, nb = args.tiles, args.nb
    rng_tiles = {}

    def stack(n_tiles, seed):
        U = np.empty((n_tiles, krank, nb))
        V = np.empty((n_tiles, nb, krank))
        for t in range(n_tiles):
            tile = O.synth_compressed_tile(nb, krank, seed * 100003 + t)
            U[t], V[t] = tile.U.T, tile.V.T
        return U, V
    Ua, Va = stack(T * T, 1)
    Ub, Vb = stack(T * T, 2)
    steps = max(1, args.steps)
    budget = max(5.0, min(args.cpu_budget_s, 150.0 / (steps + max(args.warmup, 1))))
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_sample(args, Ua, Va, Ub, Vb, krank, budget)
    t = []
    last = None
    for _ in range(steps):
        last = cpu_reference_sample(args, Ua, Va, Ub, Vb, krank, budget)
        t.append(last["seconds"])
    sec = float(np.mean(t))
    value = last["tile_gemms"] / sec
    sample = "first %d of %d block-columns of C per step (%d of %d tile-GEMMs)" % (last["cols"], T, last["tile_gemms"], T ** 3)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (same law, rank %d)" % krank,
        "config": {"workload": "TLR GEMM %dx%d f64, tile %d, acc %.0e, compressed A,B,C (reference CPU path, bounded sample)" % (
            T * nb, T * nb, nb, args.acc)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), file=_JSON_OUT, flush=True)


def setup_distributed(args, hc, ctx, dist, torch, synth, krank, cap_in, P, Q, pr, pc, prm):
    """2D block-cyclic TLR GEMM over a P x Q grid: C(j,i) on (j mod P, i mod Q); A(j,k) lives on
    (j mod P, k mod Q), B(k,i) on (k mod P, i mod Q).  Per k: the owner column broadcasts its A(:,k) row-panel along
    the grid row, the owner row broadcasts its B(k,:) column-panel along the grid column (NCCL, comm stream,
    double-buffered so that the transfer of panel k+1 overlaps the recompression of step k)."""
    from hcorepp_b200 import partition as part
    T, nb = args.tiles, args.nb
    dt = torch.float64
    dev = ctx.device
    rank = pr * Q + pc
    mt_l, nt_l, kt = T, T, T                      # per-GPU C tiles: T x T ; global grid (T*P) x (T*Q), k = T
    row_groups = [dist.new_group([r * Q + c for c in range(Q)]) for r in range(P)]
    col_groups = [dist.new_group([r * Q + c for r in range(P)]) for c in range(Q)]
    # local A tiles: rows j_l (global j = j_l*P + pr), columns k with k mod Q == pc  -> column-major (mt_l x kA_l)
    kA = part.owned_indices(kt, Q, pc)
    kB = part.owned_indices(kt, P, pr)
    Ua, Va = synth(mt_l * max(len(kA), 1), 1000 + rank)
    Ub, Vb = synth(nt_l * max(len(kB), 1), 2000 + rank)
    A_loc = hc.TileMatrix(mt_l, max(len(kA), 1), nb, nb, dt, ctx, compressed=True, max_rank=cap_in, rank_bound=krank)
    B_loc = hc.TileMatrix(nt_l, max(len(kB), 1), nb, nb, dt, ctx, compressed=True, max_rank=cap_in, rank_bound=krank)  # B^T grid
    A_loc.load_factors(Ua, Va, krank)
    B_loc.load_factors(Ub, Vb, krank)
    panA = [hc.TileMatrix(mt_l, 1, nb, nb, dt, ctx, compressed=True, max_rank=cap_in, rank_bound=krank) for _ in range(2)]
    panB = [hc.TileMatrix(nt_l, 1, nb, nb, dt, ctx, compressed=True, max_rank=cap_in, rank_bound=krank) for _ in range(2)]
    Cm = hc.TileMatrix.zeros_compressed(mt_l, nt_l, nb, nb, dt, ctx, rank_bound=args.kc_bound)
    info = torch.zeros(mt_l * nt_l, dtype=torch.int32, device=dev)
    comm = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    # descriptor lists for the batched call at buffer b: C(j_l, i_l) += panA[b](j_l) * panB[b](i_l)
    import ctypes as C
    from hcorepp_b200._capi import hcb_tile
    n = mt_l * nt_l
    descs = []
    for b in range(2):
        da, db = (hcb_tile * n)(), (hcb_tile * n)()
        for i in range(nt_l):
            for j in range(mt_l):
                da[j + i * mt_l] = panA[b].descs[j]
                db[j + i * mt_l] = panB[b].descs[i]
        descs.append((da, db))
    fn = getattr(_lib(), "hcb_dtlr_gemm_batched")
    cprm = prm.c()
    slab_elems = A_loc.tile_elems * mt_l

    def fetch(k, b):
        with torch.cuda.stream(comm):
            comm.wait_event(done[b])
            src_c, src_r = k % Q, k % P
            if src_c == pc:
                kl = k // Q
                panA[b].buf.copy_(A_loc.buf[kl * slab_elems:(kl + 1) * slab_elems], non_blocking=True)
                panA[b].ranks.copy_(A_loc.ranks[kl * mt_l:(kl + 1) * mt_l], non_blocking=True)
            if Q > 1:
                dist.broadcast(panA[b].buf, src=pr * Q + src_c, group=row_groups[pr])
                dist.broadcast(panA[b].ranks, src=pr * Q + src_c, group=row_groups[pr])
            if src_r == pr:
                kl = k // P
                panB[b].buf.copy_(B_loc.buf[kl * slab_elems:(kl + 1) * slab_elems], non_blocking=True)
                panB[b].ranks.copy_(B_loc.ranks[kl * nt_l:(kl + 1) * nt_l], non_blocking=True)
            if P > 1:
                dist.broadcast(panB[b].buf, src=src_r * Q + pc, group=col_groups[pc])
                dist.broadcast(panB[b].ranks, src=src_r * Q + pc, group=col_groups[pc])
            ready[b].record(comm)

    def one_pass():
        Cm.reset_to_zero()
        done[0].record(main)
        done[1].record(main)
        fetch(0, 0)
        for k in range(kt):
            b = k & 1
            if k + 1 < kt:
                fetch(k + 1, b ^ 1)
            main.wait_event(ready[b])
            da, db = descs[b]
            from hcorepp_b200._capi import check
            check(fn(ctx.h, n, da, 0, db, 0, Cm.descs, C.c_double(1.0), C.c_double(1.0), C.byref(cprm), info.data_ptr()))
            done[b].record(main)
    return one_pass, Cm, info, mt_l * nt_l * kt, A_loc, B_loc, (Ua, Va, Ub, Vb)


def verify_distributed(args, torch, hc, synth, Cm, krank, P, Q, pr, pc):
    """Rank-local plumbing check for N > 1: this rank's C(0,0) tile against the dense sum over k of A(j,k) B(k,i), with
    the A / B tiles REGENERATED here from their owners' deterministic generator streams (so a panel that was broadcast
    from the wrong owner, or landed in the wrong slot, shows up as an O(1) error).  The TLR arithmetic itself is
    checked against the reference in the single-GPU run."""
    from hcorepp_b200 import partition as part
    T, nb = args.tiles, args.nb
    kt = T
    dense = torch.zeros(nb, nb, dtype=torch.float64, device=Cm.buf.device)
    cache = {}

    def owner_tiles(kind, owner):
        if (kind, owner) not in cache:
            opr, opc = part.grid_pos(owner, P, Q)
            cnt = len(part.owned_indices(kt, Q, opc)) if kind == "A" else len(part.owned_indices(kt, P, opr))
            cache[(kind, owner)] = synth(T * max(cnt, 1), (1000 if kind == "A" else 2000) + owner)
        return cache[(kind, owner)]
    for k in range(kt):
        oa, ob = part.owner_of_a(pr, k, P, Q), part.owner_of_b(k, pc, P, Q)
        Ua, Va = owner_tiles("A", oa)
        Ub, Vb = owner_tiles("B", ob)
        la, lb = 0 + (k // Q) * T, 0 + (k // P) * T      # local linear index of A(jl=0, kl) / B(il=0, kl) at the owner
        Am = Ua[la].t() @ Va[la].t()                      # (nb x k)(k x nb)
        Bm = Ub[lb].t() @ Vb[lb].t()
        dense += Am @ Bm
    U, V = Cm.GetTile(0, 0).factors()
    err = (torch.linalg.norm(U @ V - dense) / torch.linalg.norm(dense)).item()
    return {"rel_fro_err_tile00_vs_dense": err, "tolerance": 1e-5, "pass": bool(err <= 1e-5),
            "note": "distributed plumbing check; TLR-vs-reference parity is the N=1 block"}


def _lib():
    from hcorepp_b200 import _capi
    return _capi.lib


if __name__ == "__main__":
    sys.exit(main())
