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
    ap.add_argument("--profile", action="store_true",
                    help="profiling runs under ncu only (its numbers are never bench values): honour --warmup below 3 and skip "
                         "the untimed rank-trace pass, so that the launch list holds just warm-up + timed passes")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (BASELINE configs[3])")
    ap.add_argument("--no-cholesky", action="store_true", help="skip the TLR Cholesky leg (BASELINE configs[4], N=1 only)")
    ap.add_argument("--chol-tiles", type=int, default=32)
    ap.add_argument("--chol-nb", type=int, default=1024)
    ap.add_argument("--chol-acc", type=float, default=1e-8)
    ap.add_argument("--strong-tiles", type=int, default=32)
    ap.add_argument("--strong-nb", type=int, default=2048)
    ap.add_argument("--strong-acc", type=float, default=1e-6)
    ap.add_argument("--strong-steps", type=int, default=2)
    ap.add_argument("--strong-parity-tiles", type=int, default=64)
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


def workload_string(T, nb, acc, world):
    """config.workload of the headline (weak-scaling) line -- printed identically by both arms"""
    P, Q = grid_shape(world)
    return "TLR GEMM %dx%d f64, tile %d, acc %.0e, compressed A,B,C, %d tile-GEMMs/step%s" % (
        T * nb * P, T * nb * Q, nb, acc, T ** 3 * world,
        "" if world == 1 else ", 2D block-cyclic %dx%d + NCCL panel broadcast" % (P, Q))


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


_FP64_PEAK = None


def fp64_peak(torch=None):
    """The path computes in f64: no tcgen05 kind exists for it, the tensor instruction is DMMA.  MEASURED_PEAKS.json only
    holds bf16, so the yardstick is cuBLAS DGEMM (torch.matmul, 8192^3, best of 6 after 2 warm-ups, CUDA events) measured
    IN THIS RUN on this GPU (the denominator should be driver-visible, not a committed constant); the
    committed round-1 measurement is only the fallback when the run cannot measure (reference arm, no torch)."""
    global _FP64_PEAK
    if _FP64_PEAK is not None:
        return _FP64_PEAK
    if torch is not None:
        try:
            n = 8192
            a = torch.randn(n, n, dtype=torch.float64, device="cuda")
            b = torch.randn(n, n, dtype=torch.float64, device="cuda")
            best = 1e9
            for i in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                c = a @ b
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    best = min(best, e0.elapsed_time(e1))
            del a, b, c
            torch.cuda.empty_cache()
            _FP64_PEAK = (2 * n ** 3 / (best * 1e-3) / 1e12,
                          "measured in this run: cuBLAS DGEMM 8192^3 f64, best of 6 (MEASURED_PEAKS.json has no f64 entry; "
                          "register-resident DMMA issue peak 37.1 TFLOP/s, profiles/r01_dmma_rate.txt)")
            return _FP64_PEAK
        except Exception:
            pass
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_yardstick.json")))
        return float(p["dgemm_tflops"]), "committed round-1 measurement of cuBLAS DGEMM 8192^3 f64 (profiles/r01_fp64_yardstick.json)"
    except Exception:
        return 35.5, "fallback: cuBLAS DGEMM measured in round 1 (35.5 TFLOP/s)"


def ncu_traffic(kernel):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (or None)."""
    try:
        name = "r02_ncu_traffic.json" if os.path.exists(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) else "r01_ncu_traffic.json"
        p = json.load(open(os.path.join(ROOT, "profiles", name)))
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
    # all the host cores this process may run on -- NOT omp_get_max_threads(): torch.distributed.run exports
    # OMP_NUM_THREADS=1 to its workers, which made the round-1 reference arm run on one core at N > 1.  The tile loop
    # passes an explicit num_threads() clause (oracle/ref_capi.cpp tile_matmul), so the environment does not cap it.
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    threads = max(1, cores)
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
        # weak scaling through the LIBRARY's multi-GPU driver: (T*P) x (T*Q) C tiles, k = T, 2D block-cyclic, NCCL panels
        from hcorepp_b200 import distributed as D
        grid = D.Grid2D(P, Q)
        leg = DistLeg(torch, hc, ctx, grid, T * P, T * Q, T, nb, args.acc, krank, kc_bound=args.kc_bound)
        Cm, info, n_local_gemms, A, B = leg.C.local, leg.info, leg.n_local * T, leg.A.local, leg.B.local
        one_pass = leg.one_pass
        if args.kc_bound == 0:  # untimed calibration pass, bound agreed across ranks
            args.kc_bound = leg.calibrate()
    total_gemms = n_local_gemms * world

    # ---- warm-up (also grows the scratch arena once), then the timed region
    for _ in range(args.warmup if args.profile else max(args.warmup, 3)):
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
    fast_paths = ctx.stats(reset=True)  # adaptive fast paths taken over warm-up + timed passes (CholeskyQR2 panels, one-pass Gram-Schmidt)
    bad = int((info & 5).max().item())  # 1: Jacobi not converged, 4: rank bound exceeded (2 = clipped at maxRank is legal)
    sweeps = int((info >> 8).max().item())
    ms_per_step = ms / args.steps
    value = total_gemms / (ms_per_step * 1e-3)

    # ---- rank trace (untimed): kc before / rk after / Jacobi sweeps of every k-step, for the flop and byte accounting
    if world == 1 and not args.profile:
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
        "config": {"workload": workload_string(T, nb, args.acc, world),
            "baseline_config": "BASELINE.json configs[2]" if (T, nb, world) == (16, 1024, 1) else "custom",
            "l2": "inputs larger than L2 (A+B live factors %.0f MB per GPU, C scratch re-written every k)" % (
                2 * T * T * 2 * nb * krank * 8 / 1e6)},
        "library": _capi.lib.hcb_version().decode(),
        "gpu_launches": launches, "jacobi_or_bound_flags": bad, "jacobi_sweeps_last_step_max": sweeps,
        "c_rank_bound": args.kc_bound, "fast_paths": fast_paths,
    }
    if world > 1 and not args.no_e2e:
        # end to end at N GPUs: every rank uploads ITS A / B tiles from pinned host memory, runs the pass (panel
        # broadcasts included) and reads ITS C tiles (ranks + live factors) back; barrier on both sides, max over ranks
        leg.kc_bound = args.kc_bound
        e2e_ms, h2d, d2h = run_e2e_dist(leg, args.steps)
        if rank_env == 0:
            result["e2e"] = {"value": total_gemms / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                             "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                             "bytes_note": "bytes are summed over the ranks; the time is the maximum over the ranks"}
    if world > 1:
        # weak leg: every rank checks a few of its own C tiles against the reference CPU path (all ranks call: collective)
        par = None
        if not args.no_cpu_baseline:
            try:
                leg.one_pass()
                ctx.Sync()
                par = leg.parity_vs_reference(max(8, 2 * world))
            except Exception as e:
                par = {"error": repr(e)}
        if rank_env == 0 and par is not None:
            result["parity"] = par
        del leg, Cm, A, B, one_pass
        torch.cuda.empty_cache()
    # ---- strong-scaling leg on BASELINE configs[3] (all ranks take part; reported next to the headline line)
    strong = None
    if not args.no_strong:
        try:
            from hcorepp_b200 import distributed as D
            sgrid = grid if world > 1 else D.Grid2D(1, 1)
            strong = run_strong_leg(args, torch, hc, ctx, sgrid, rank_env)
        except Exception as e:  # reported, never fatal for the headline number
            strong = {"error": repr(e)}
    if rank_env == 0 and strong is not None:
        result["strong_scaling"] = strong
        result["config"]["strong_workload"] = strong.get("workload")
    if world == 1 and not args.no_cholesky:
        try:
            result["cholesky"] = run_cholesky_leg(args, torch, hc, ctx)
        except Exception as e:
            result["cholesky"] = {"error": repr(e)}
        torch.cuda.empty_cache()
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
                pk, pk_src = fp64_peak(torch)
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
                # Jacobi sweeps per (k-step, tile): histogram over the whole pass and over the last k-step
                "jacobi_sweep_hist": {str(int(v)): int(c) for v, c in zip(*np.unique(sw_all, return_counts=True))},
                "jacobi_sweep_hist_last_k": {str(int(v)): int(c) for v, c in zip(*np.unique(sw_all[-1], return_counts=True))},
                # the three FP64-bound recompression phases against the measured FP64 peak (algorithmic flops, true ranks)
                "phase_fp64_tflops": {n: float(alg_flops[n]) / (phases[n]["ms_per_step"] * 1e-3) / 1e12 for n in alg_flops},
                "phase_fp64_frac_of_peak": {n: float(alg_flops[n]) / (phases[n]["ms_per_step"] * 1e-3) / 1e12 / fp64_peak(torch)[0]
                                            for n in alg_flops},
                "phase_note": "panel_qr / apply_q rates count the ALGORITHMIC flops of the task as SURVEY.md 8d defines them "
                              "(Householder QR of both n x r stacks, implicit apply-Q); since round 2 the incremental path "
                              "executes fewer: block Gram-Schmidt of the kp new columns (16 nb kc kp), two kp-column panels, "
                              "an R-only QR of the r x r matrix RV*Pi, and GEMM rebuilds (4 nb r rk)",
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
    out = {"value": total_gemms / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h),
           "mode": "serial: H2D of A and B, the product, D2H of C, one after the other on one stream"}
    # ---- the same steps as a double-buffered pipeline (what a throughput-oriented caller does): step i+1's uploads and
    #      step i-1's download run on the copy engines while step i computes.  Every step still moves all of its inputs and
    #      its whole result inside the timed region; reported next to the serial number, not instead of it.
    try:
        main = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        A2 = hc.TileMatrix(A.mt, A.nt, nb, nb, torch.float64, ctx, compressed=True, max_rank=A.max_rank, rank_bound=krank)
        B2 = hc.TileMatrix(B.mt, B.nt, nb, nb, torch.float64, ctx, compressed=True, max_rank=B.max_rank, rank_bound=krank)
        C2 = hc.TileMatrix.zeros_compressed(T, T, nb, nb, torch.float64, ctx, rank_bound=kcb)
        bufs = [(A, B, Cm), (A2, B2, C2)]
        hR = [h_ranks, torch.empty_like(h_ranks).pin_memory()]
        hUo = [hU, torch.empty_like(hU).pin_memory()]
        hVo = [hV, torch.empty_like(hV).pin_memory()]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]

        def pipelined(n_steps):
            for i in range(n_steps):
                b = i % 2
                Ai, Bi, Ci = bufs[b]
                s_in.wait_event(ev_done[b])              # the compute that last read this A / B pair is over
                with torch.cuda.stream(s_in):
                    Ai.load_factors(hUa, hVa, krank)
                    Bi.load_factors(hUb, hVb, krank)
                    ev_in[b].record(s_in)
                main.wait_event(ev_in[b])
                main.wait_event(ev_out[b])               # the download that last read this C is over
                Ci.reset_to_zero()
                hc.tile_matrix_multiplication(Ai, Bi, Ci, 1.0, 1.0, ctx, prm, info=info)
                ev_done[b].record(main)
                s_out.wait_event(ev_done[b])
                with torch.cuda.stream(s_out):
                    hR[b].copy_(Ci.ranks, non_blocking=True)
                    v = Ci.buf.view(T * T, Ci.tile_elems)
                    hUo[b].copy_(v[:, : nb * kcb], non_blocking=True)
                    hVo[b].copy_(v[:, nb * cap: nb * cap + nb * kcb], non_blocking=True)
                    ev_out[b].record(s_out)
            main.wait_stream(s_out)
        pipelined(2)
        torch.cuda.synchronize()
        n_p = max(args.steps, 4)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        pipelined(n_p)
        p1.record()
        torch.cuda.synchronize()
        pms = p0.elapsed_time(p1) / n_p
        same = bool(torch.equal(hR[0], hR[1]))
        out["pipelined"] = {"value": total_gemms / (pms * 1e-3), "unit": UNIT, "ms_per_step": pms, "steps": n_p,
                            "ranks_equal_across_buffers": same,
                            "mode": "double-buffered A/B/C: uploads of step i+1 and the download of step i-1 overlap the "
                                    "product of step i (two copy streams); same bytes per step as the serial mode"}
        del A2, B2, C2
    except Exception as e:  # reported, never fatal
        out["pipelined"] = {"error": repr(e)}
    return out


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
        # same workload as the GPU arm at this N (its tiles are i.i.d. draws of one law, so the per-tile-GEMM cost of the
        # bounded sample named in cpu_baseline.sample is the workload's)
        "config": {"workload": workload_string(T, nb, args.acc, max(1, args.gpus))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), file=_JSON_OUT, flush=True)


def tile_synth(torch, dev, nb, krank):
    """Deterministic PER GLOBAL TILE generator of compressed inputs with the reference spectrum law:
    tile (matrix id, r, c) = Q_u diag(sigma_0..k-1) Q_v^T, Haar-like Q from the QR of Gaussians seeded by the tile's global
    coordinates -- any rank can regenerate any tile (the parity legs rely on it).  Returns (U, V) stacks in the layout
    TileMatrix.load_factors takes: U (n, krank, nb) [column-major nb x k], V (n, nb, krank) [column-major k x nb]."""
    sig = torch.from_numpy(spectrum(nb)[:krank].copy()).to(dev)
    dt = torch.float64
    g = torch.Generator(device=dev)

    def synth(coords, mid):
        n = len(coords)
        U = torch.empty(max(n, 1), krank, nb, dtype=dt, device=dev)
        V = torch.empty(max(n, 1), nb, krank, dtype=dt, device=dev)
        for c0 in range(0, n, 32):
            c1 = min(n, c0 + 32)
            R = torch.empty(2 * (c1 - c0), nb, krank, dtype=dt, device=dev)
            for t in range(c0, c1):
                r, c = coords[t]
                g.manual_seed(1_000_003 * (mid + 1) + 4099 * r + c)
                R[2 * (t - c0): 2 * (t - c0) + 2] = torch.randn(2, nb, krank, generator=g, dtype=dt, device=dev)
            q, _ = torch.linalg.qr(R)
            U[c0:c1] = q[0::2].transpose(1, 2)
            V[c0:c1] = q[1::2] * sig[None, None, :]
        return U, V
    return synth


class DistLeg:
    """One distributed workload (weak or strong leg): A (mt x kt), B (kt x nt), C (mt x nt) tiles of nb x nb on the
    process grid, driven through the LIBRARY's multi-GPU driver (hcorepp_b200.distributed.tlr_matmul_distributed)."""

    def __init__(self, torch, hc, ctx, grid, mt, nt, kt, nb, acc, krank, kc_bound=0):
        from hcorepp_b200 import distributed as D
        self.torch, self.hc, self.D, self.ctx, self.grid = torch, hc, D, ctx, grid
        self.mt, self.nt, self.kt, self.nb, self.acc, self.krank = mt, nt, kt, nb, acc, krank
        dt = torch.float64
        self.prm = hc.CompressionParameters(acc)
        self.synth = tile_synth(torch, ctx.device, nb, krank)
        self.A = D.DistTileMatrix(mt, kt, nb, nb, dt, ctx, grid, panel_rows=False, max_rank=krank, rank_bound=krank)
        self.B = D.DistTileMatrix(kt, nt, nb, nb, dt, ctx, grid, panel_rows=True, max_rank=krank, rank_bound=krank)
        self.C = D.DistTileMatrix(mt, nt, nb, nb, dt, ctx, grid, panel_rows=False, rank_bound=kc_bound)
        self.fa = self.synth(self.A.global_coords(), 0)
        self.fb = self.synth(self.B.global_coords(), 1)
        self.A.local.load_factors(self.fa[0], self.fa[1], krank)
        self.B.local.load_factors(self.fb[0], self.fb[1], krank)
        self.n_local = len(self.C.rows) * len(self.C.cols)
        self.info = torch.zeros(max(self.n_local, 1), dtype=torch.int32, device=ctx.device)
        self.kc_bound = kc_bound

    def one_pass(self):
        self.C.local.reset_to_zero()
        self.D.tlr_matmul_distributed(self.A, self.B, self.C, 1.0, 1.0, self.ctx, self.prm, info=self.info)

    def calibrate(self):
        """untimed pass with the safe bound (max_rank): learn how far the C ranks grow, agree on the bound across ranks"""
        torch, dist = self.torch, self.grid.dist
        self.one_pass()
        self.ctx.Sync()
        mx = self.C.local.ranks.max().to(torch.int64) if self.n_local else torch.zeros((), dtype=torch.int64, device=self.ctx.device)
        if dist is not None and self.grid.world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        self.kc_bound = int(min(self.C.max_rank, (int(mx.item()) + 8 + 7) // 8 * 8))
        self.C.local.set_rank_bound(self.kc_bound)
        return self.kc_bound

    def timed(self, steps, warmup):
        """`warmup` untimed passes, then `steps` passes between barriers; returns ms per pass (max over ranks)."""
        torch, dist = self.torch, self.grid.dist
        multi = dist is not None and self.grid.world > 1
        for _ in range(warmup):
            self.one_pass()
        self.ctx.Sync()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.one_pass()
        e1.record()
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=self.ctx.device)
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def flags(self):
        torch, dist = self.torch, self.grid.dist
        f = torch.stack([(self.info & 5).max(), ((self.info >> 8) & 0xff).max()]).to(torch.int64)
        if dist is not None and self.grid.world > 1:
            dist.all_reduce(f, op=dist.ReduceOp.MAX)
        return int(f[0].item()), int(f[1].item())

    def parity_vs_reference(self, n_tiles_total):
        """Every rank checks a sub-grid of ITS OWN C tiles (about n_tiles_total / world of them) against the reference's
        CPU path (oracle/_ref, compiled unmodified): the A row-panels and B column-panels those tiles need are
        regenerated from the global per-tile generator, the reference runs the full k-sum on them, and the GPU result is
        compared as a dense reconstruction (rel. Frobenius <= 10 * acc) and by rank (+/-1).  Results are reduced over the
        ranks: the worst error / rank difference anywhere, and the number of tiles compared."""
        import numpy as np
        from oracle import ref as R
        torch, dist, g = self.torch, self.grid.dist, self.grid
        want = max(1, n_tiles_total // g.world)
        na = max(1, min(len(self.C.rows), int(round(math.sqrt(want)))))
        nbc = max(1, min(len(self.C.cols), (want + na - 1) // na))
        rows = [self.C.rows[(i * max(1, len(self.C.rows) // na)) % len(self.C.rows)] for i in range(na)]
        cols = [self.C.cols[(i * max(1, len(self.C.cols) // nbc)) % len(self.C.cols)] for i in range(nbc)]
        rows, cols = sorted(set(rows)), sorted(set(cols))
        p = R.Params(self.acc)
        nb, kt = self.nb, self.kt
        ua, va = self.synth([(j, k) for j in rows for k in range(kt)], 0)
        ub, vb = self.synth([(k, i) for i in cols for k in range(kt)], 1)
        ua, va, ub, vb = (x.cpu().numpy() for x in (ua, va, ub, vb))
        A = [[R.RefTile.from_uv(ua[a * kt + k].T, va[a * kt + k].T) for k in range(kt)] for a in range(len(rows))]
        B = [[R.RefTile.from_uv(ub[b * kt + k].T, vb[b * kt + k].T) for b in range(len(cols))] for k in range(kt)]
        z_u, z_v = np.zeros((nb, 1)), np.zeros((1, nb))
        Cg = [[R.RefTile.from_uv_cap(z_u, z_v, max(nb // 3, 1)) for _ in cols] for _ in rows]
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        threads = max(1, cores // g.world)
        sec, _ = R.matmul(A, B, Cg, 1.0, 1.0, p, nthreads=threads)
        num = den = 0.0
        rdiff = 0
        for a, j in enumerate(rows):
            for b, i in enumerate(cols):
                d_ref = Cg[a][b].to_dense()
                t = self.C.GetTile(j, i)
                d_gpu = t.to_dense()
                num += float(np.linalg.norm(d_gpu - d_ref) ** 2)
                den += float(np.linalg.norm(d_ref) ** 2)
                rdiff = max(rdiff, abs(int(Cg[a][b].info()["rank"]) - t.GetTileRank()))
        v = torch.tensor([num, den, float(len(rows) * len(cols)), sec], dtype=torch.float64, device=self.ctx.device)
        m = torch.tensor([rdiff], dtype=torch.int64, device=self.ctx.device)
        if dist is not None and g.world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
        err = math.sqrt(v[0].item() / max(v[1].item(), 1e-300))
        out = {"rel_fro_err_vs_reference": err, "tolerance": 10 * self.acc, "max_rank_diff": int(m.item()),
               "c_tiles_compared": int(v[2].item()), "ranks_checking": g.world,
               "reference": "oracle/_ref (reference CPU path compiled unmodified), full k-sum of every compared tile",
               "reference_cpu_s_sum_over_ranks": v[3].item()}
        out["pass"] = bool(err <= out["tolerance"] and out["max_rank_diff"] <= 1)
        return out


def run_e2e_dist(leg, steps):
    """N > 1 end-to-end: per step H2D of the rank's own A / B factor stacks (pinned), the distributed pass, D2H of the
    rank's C tiles.  Returns (ms per step as max over ranks, H2D bytes, D2H bytes per rank and step)."""
    torch, dist = leg.torch, leg.grid.dist
    hUa, hVa, hUb, hVb = (x.cpu().pin_memory() for x in (*leg.fa, *leg.fb))
    Cl = leg.C.local
    cap, kcb, nb = Cl.max_rank, leg.kc_bound, leg.nb
    nC = Cl.mt * Cl.nt
    h_ranks = torch.empty(nC, dtype=torch.int32).pin_memory()
    hU = torch.empty(nC, nb * kcb, dtype=torch.float64).pin_memory()
    hV = torch.empty(nC, nb * kcb, dtype=torch.float64).pin_memory()

    def step():
        leg.A.local.load_factors(hUa, hVa, leg.krank)
        leg.B.local.load_factors(hUb, hVb, leg.krank)
        leg.one_pass()
        h_ranks.copy_(Cl.ranks, non_blocking=True)
        v = Cl.buf.view(nC, Cl.tile_elems)
        hU.copy_(v[:, : nb * kcb], non_blocking=True)
        hV.copy_(v[:, nb * cap: nb * cap + nb * kcb], non_blocking=True)
    step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=Cl.buf.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h2d = sum(x.numel() * x.element_size() for x in (hUa, hVa, hUb, hVb))
    d2h = h_ranks.numel() * 4 + (hU.numel() + hV.numel()) * 8
    return float(t.item()), h2d, d2h


def run_cholesky_leg(args, torch, hc, ctx):
    """BASELINE.json configs[4]: tile Cholesky of a synthetic covariance matrix (32768 points in the unit square along a
    Morton curve, Matern-5/2 kernel, nugget 1e-2), tile 1024, accuracy 1e-8: dense diagonal tiles, compressed tiles below
    (compressed on the device, untimed), factorised by hcb_dtlr_potrf (potrf / panel trsm / syrk / one batched recompressing
    GEMM per step).  Timed with CUDA events; checked through the size-independent property A = L L^T on sampled tiles."""
    nt, nb, acc = args.chol_tiles, args.chol_nb, args.chol_acc
    dev, dt = ctx.device, torch.float64
    g = torch.Generator(device=dev)
    g.manual_seed(77)
    pts = torch.rand(nt * nb, 2, generator=g, dtype=dt, device=dev)
    q = (pts * 65535).to(torch.int64)

    def spread(v):
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        return (v | (v << 1)) & 0x55555555
    pts = pts[torch.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << 1), stable=True)]
    ell, nugget = 0.1, 1e-2

    def tile(i, j):
        a, b = pts[i * nb:(i + 1) * nb], pts[j * nb:(j + 1) * nb]
        z = math.sqrt(5.0) * torch.cdist(a, b) / ell
        t = (1 + z + z * z / 3.0) * torch.exp(-z)
        if i == j:
            t = t + nugget * torch.eye(nb, dtype=dt, device=dev)
        return t
    prm = hc.CompressionParameters(acc)
    t0 = time.time()
    S = hc.SymTileMatrix.from_tiles(tile, nt, nb, dt, ctx, prm)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    ranks0 = S.low.rank_table()
    low_mask = np.tril(np.ones((nt, nt), dtype=bool), -1)
    keep = (S.diag.clone(), S.low.buf.clone(), S.low.ranks.clone(), S.low.state.clone())
    pinfo = torch.zeros(nt, dtype=torch.int32, device=dev)

    def restore():
        S.diag.copy_(keep[0]); S.low.buf.copy_(keep[1]); S.low.ranks.copy_(keep[2]); S.low.state.copy_(keep[3])
    hc.tlr_cholesky(S, ctx, prm, potrf_info=pinfo)   # warm-up (scratch arena)
    ctx.Sync()
    ms = []
    for _ in range(2):
        restore()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hc.tlr_cholesky(S, ctx, prm, potrf_info=pinfo)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = min(ms)
    ranks1 = S.low.rank_table()
    # property check on sampled tiles: (L L^T)(i, j) = sum_{k <= j} L(i, k) L(j, k)^T against the generated block
    def Lt(i, k):
        if i == k:
            return torch.tril(S.diag[k].view(nb, nb).t())
        U, V = S.low.GetTile(i, k).factors()
        return U @ V
    num = den = 0.0
    samples = [(nt - 1, nt - 1), (nt - 1, 0), (nt // 2, nt // 2 - 1), (nt - 1, nt - 2), (nt // 3, nt // 3)]
    for (i, j) in samples:
        acc_t = torch.zeros(nb, nb, dtype=dt, device=dev)
        for k in range(j + 1):
            acc_t += Lt(i, k) @ Lt(j, k).t()
        ref = tile(i, j)
        num += float(torch.linalg.norm(acc_t - ref) ** 2)
        den += float(torch.linalg.norm(ref) ** 2)
    err = math.sqrt(num / den)
    n_updates = sum((nt - k - 1) * (nt - k - 2) // 2 for k in range(nt))
    return {"workload": "TLR Cholesky %dx%d f64, tile %d, acc %.0e, Matern-5/2 covariance (ell %.2f, nugget %.0e), dense diagonal + "
                        "compressed lower tiles" % (nt * nb, nt * nb, nb, acc, ell, nugget),
            "baseline_config": "BASELINE.json configs[4]" if (nt, nb, acc) == (32, 1024, 1e-8) else "custom",
            "ms": ms, "tile_updates": n_updates, "tile_updates_per_s": n_updates / (ms * 1e-3),
            "dense_equivalent_tflops": (nt * nb) ** 3 / 3.0 / (ms * 1e-3) / 1e12,
            "build_and_compress_s": t_build, "rank_in_mean": float(ranks0[low_mask].mean()), "rank_in_max": int(ranks0[low_mask].max()),
            "rank_out_mean": float(ranks1[low_mask].mean()), "rank_out_max": int(ranks1[low_mask].max()),
            "potrf_info_max": int(pinfo.abs().max().item()),
            "check": {"rel_fro_err_LLt_vs_A_sampled_tiles": err, "tiles": len(samples), "tolerance": 10 * acc,
                      "pass": bool(err <= 10 * acc and int(pinfo.abs().max().item()) == 0)},
            "driver": "hcorepp_b200.tlr_cholesky (hcb_dtlr_potrf)"}


def run_strong_leg(args, torch, hc, ctx, grid, rank_env):
    """BASELINE.json configs[3]: 65536 x 65536 f64, tile 2048, accuracy 1e-6, C tiles 2D block-cyclic over the grid, NCCL
    panel broadcast -- FIXED total work (32^3 tile-GEMMs) at every N: strong scaling.  Timed like the headline (barrier +
    synchronize on both sides, CUDA events, max over ranks), parity of sampled C tiles of every rank against oracle/_ref."""
    T, nb, acc = args.strong_tiles, args.strong_nb, args.strong_acc
    krank = rank_for_accuracy(nb, acc)
    leg = DistLeg(torch, hc, ctx, grid, T, T, T, nb, acc, krank)
    bound = leg.calibrate()
    import ctypes as C
    from hcorepp_b200 import _capi
    ph_ms, ph_n = (C.c_double * _capi.N_PHASES)(), (C.c_uint64 * _capi.N_PHASES)()
    _capi.check(_capi.lib.hcb_ctx_phase_times(ctx.h, ph_ms, ph_n))   # flush
    ph_ms, ph_n = (C.c_double * _capi.N_PHASES)(), (C.c_uint64 * _capi.N_PHASES)()
    _capi.check(_capi.lib.hcb_ctx_phase_timing(ctx.h, 1))
    ms = leg.timed(args.strong_steps, 1)
    _capi.check(_capi.lib.hcb_ctx_phase_times(ctx.h, ph_ms, ph_n))
    _capi.check(_capi.lib.hcb_ctx_phase_timing(ctx.h, 0))
    bad, sweeps = leg.flags()
    total = T ** 3
    out = {"workload": "TLR GEMM %dx%d f64, tile %d, acc %.0e, compressed A,B,C, %d tile-GEMMs/step, 2D block-cyclic %dx%d%s" % (
               T * nb, T * nb, nb, acc, total, grid.P, grid.Q, " + NCCL panel broadcast" if grid.world > 1 else ""),
           "baseline_config": "BASELINE.json configs[3]" if (T, nb, acc) == (32, 2048, 1e-6) else "custom",
           "scaling": "strong", "n_gpus": grid.world, "value": total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
           "steps": args.strong_steps, "warmup": 1, "c_rank_bound": bound, "jacobi_or_bound_flags": bad,
           "jacobi_sweeps_max": sweeps, "input_rank": krank,
           "dense_equivalent_tflops": 2.0 * (T * nb) ** 3 / (ms * 1e-3) / 1e12,
           "driver": "hcorepp_b200.distributed.tlr_matmul_distributed (hcb_dtlr_matmul_panel_step per k)",
           # rank 0's phases over the warm-up + timed steps, per step
           "phases_ms_per_step": {_capi.lib.hcb_phase_name(i).decode(): ph_ms[i] / (args.strong_steps + 1)
                                  for i in range(_capi.N_PHASES)}}
    try:
        ref1 = json.load(open(os.path.join(ROOT, "profiles", "r02_strong_n1.json")))
        out["speedup_vs_committed_n1"] = ref1["ms_per_step"] / ms
        out["committed_n1_ms_per_step"] = ref1["ms_per_step"]
    except Exception:
        pass
    if args.strong_parity_tiles > 0:
        try:
            out["parity"] = leg.parity_vs_reference(args.strong_parity_tiles)
        except Exception as e:
            out["parity"] = {"error": repr(e)}
    del leg
    torch.cuda.empty_cache()
    return out


def _lib():
    from hcorepp_b200 import _capi
    return _capi.lib


if __name__ == "__main__":
    sys.exit(main())
