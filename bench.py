#!/usr/bin/env python3
"""bench.py -- MLUPS of the D3Q19 cavity hot path on N B200s (one JSON line on rank 0).

A bench "step" is one residual block of the reference's time loop (main.c:104-124): `--block`
(default 500 = the reference's `skip`, main.c:43) iterations of collide+stream+walls followed by the
convergence residual.  Workload (BASELINE.json configs[2], the north_star's target grid):
    512^3, Re=1000, TRT, fp64 -- on ONE GPU at --gpus 1 and as z-slabs over N GPUs (NCCL halo) at
    --gpus N, so the 1 -> 8 curve is strong scaling on one grid.
    --grid G : override (256 = configs[1]; 1024 on eight GPUs = configs[3])
The N=1 line also carries `config5_d2q9`: BASELINE configs[4] (8192^2 D2Q9 variant, --workload d2q9).

value  = cells * iterations / device time (CUDA events on the launching stream, max over ranks);
e2e    = the same block driven through the C ABI with pinned HOST buffers: every step uploads its
         distributions (lbm_set_f_async / lbm_set_f_commit), iterates, reads the residual and
         downloads the five macroscopic fields (lbm_macrovars_async); a host that feeds a sequence of
         states overlaps the upload of state k+1 and the download of state k-1 with the iterations of
         state k (what the timed loop does); `e2e.serial` is the same without any overlap.
roofline: algorithmic bytes = 304 per cell update (19 loads + 19 stores of fp64) against the
         measured copy bandwidth of MEASURED_PEAKS.json.  The two-iterations-per-pass kernel
         (csrc/lbm3d_pair.cuh) moves 304 B per cell per TWO updates, so `frac` can exceed 1.
cpu_baseline / --impl reference: the reference's own CPU code timed on this box's host cores --
         oracle/_ref (main.c compiled by oracle/build_ref.py, kind "reference") where it can run; its
         `int` indexing overflows at N >= 484, so 512^3 uses the 64-bit port (kind "port").
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = 304.0     # bytes per lattice update, D3Q19 fp64: 2 * 19 * 8
B_ALG_2D = 144.0  # D2Q9
NVSMI_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
           "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
           "clocks_event_reasons.sw_power_cap")


def workload_label(N, Re, world):
    """The SAME string in both arms (the driver compares it)."""
    return "3D lid-driven cavity D3Q19 %d^3 fp64 Re=%d TRT%s" % (N, Re, "" if world == 1 else ", z-slabs over %d GPUs" % world)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh).get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi sampled every 200 ms DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.path = os.path.join(tempfile.gettempdir(), "lbm_bench_clocks_%d_%d.csv" % (os.getpid(), index))
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + NVSMI_Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, power, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---- the reference's CPU implementation (the ONLY place bench.py touches oracle/) -------------------
def cpu_reference(N, Re, iters, warm, steps, threads, seconds=None):
    """Time the reference's own CPU implementation of the path (oracle/_ref if it was compiled for
    this grid, else the oracle port with 64-bit indexing).  `seconds`: run whole iterations until that
    much wall time has passed instead of a fixed count.  Returns (mlups, kind, sample, seconds)."""
    from oracle import oracle as O
    cfg = dict(N=N, Re=Re)
    if O.ref_available(**cfg):
        exe = O.ref_binary(**cfg)
        if seconds:
            cmd = [exe, "sample", str(seconds), str(warm), str(threads)]
        else:
            cmd = [exe, "bench", str(iters), str(warm), str(steps), str(threads)]
        r = subprocess.run(cmd, capture_output=True, text=True, check=True)
        j = json.loads(r.stdout.strip().splitlines()[-1])
        return j["mlups"], "reference", "%d+%d iterations of the %d^3 grid (%.1f s), oracle/_ref (reference main.c)" % (
            j["warmup"] * j["iters_per_step"], j["steps"] * j["iters_per_step"], N, j["seconds"]), j["seconds"]
    o = O.Oracle3D(N=N, Re=Re, threads=threads)
    if warm:
        o.step(warm * iters, want_diff=False)
    if seconds:
        n, sec = 0, 0.0
        while n < 3 or sec < seconds:
            sec += o.time_steps(1)
            n += 1
        o.close()
        return N ** 3 * n / sec / 1e6, "port", ("%d+%d iterations of the %d^3 grid (%.1f s), 64-bit oracle port (the reference's int "
                                                "indexing cannot run this grid)" % (warm, n, N, sec)), sec
    sec = o.time_steps(steps * iters)
    o.close()
    return N ** 3 * steps * iters / sec / 1e6, "port", "%d+%d steps x %d iterations of the %d^3 grid, 64-bit oracle port" % (
        warm, steps, iters, N), sec


def cpu_reference_2d(N, threads):
    """The REAL 2D program (oracle/_ref, built over oracle/mpi_shim by build_ref.py) on `threads` ranks:
    11 iterations of N x N, MLUPS as the program itself prints it after its last check (2D:469-472)."""
    from oracle import oracle as O
    cfg = dict(N=N, M=N, Re="10000", MAXITER="11", prntInt="5", THRESH="0")
    exe = O.ref2d_binary(**cfg)
    if not os.path.exists(exe):
        return None
    ranks = max(3, min(threads, N // 8, 128))   # oracle/mpi_shim: SHIM_MAXP
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        r = subprocess.run([exe], capture_output=True, text=True, cwd=tmp, env=dict(os.environ, SHIM_NP=str(ranks)), check=True)
        wall = time.perf_counter() - t0
    rows = [l.split() for l in r.stdout.splitlines() if l[:1].isdigit() and len(l.split()) == 3]
    mlups = float(rows[-1][2])
    return {"value": mlups, "unit": "MLUPS", "cores": ranks, "kind": "reference",
            "sample": "11 iterations of the %d^2 grid on %d ranks of the reference's main.cpp over oracle/mpi_shim (%.1f s wall incl. "
                      "initialisation); MLUPS as the program prints it (2D:472)" % (N, ranks, wall)}


def run_reference_arm(a, world, rank):
    if rank != 0:
        return 0
    N = a.grid or 512
    threads = host_threads()
    iters = a.ref_iters
    mlups, kind, sample, sec = cpu_reference(N, a.Re, iters, a.warmup, a.steps, threads)
    line = {
        "impl": "reference", "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec / a.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_label(N, a.Re, a.gpus), "grid": N, "Re": a.Re, "method": "TRT",
                   "iterations_per_step": iters, "cpu_code": kind,
                   "note": "the reference's CPU implementation of the path on the host cores; MLUPS is per cell update, so a "
                           "bounded sample (%d iteration(s) per step) measures the same metric" % iters},
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- parity probes: every bench line carries evidence that the timed library computes the reference's bits
def parity_probe(lbm, device, opts):
    """12^3 TRT cavity, 10 iterations on this rank's GPU, sha256 of the raw distributions against the
    hash recorded from the REAL reference (tests/golden/ref_hashes.json <- oracle/_ref)."""
    case = "12^3 Re=100 TRT, 10 iterations"
    try:
        with open(os.path.join(ROOT, "tests", "golden", "ref_hashes.json")) as fh:
            want = json.load(fh)["N12_Re100_m1_g1.0_d0_u0.1"]["10"]["sha256"]
        with lbm.Cavity3D(N=12, Re=100, device=device) as sim:
            for k, v in opts.items():
                sim.set_option(k, v)
            sim.step(10)
            got = hashlib.sha256(sim.f().tobytes()).hexdigest()
        return {"case": case, "multi": False, "sha256_equals_reference": got == want}
    except Exception as e:
        return {"case": case, "multi": False, "error": repr(e)}


def parity_probe_multi(lbm, dist, torch, dev, rank, world, local_rank, nccl_id, halo, opts):
    """The same over ALL ranks: 64^3 Re=253 TRT as z-slabs with the bench's own halo transport and
    stepping options, 30 iterations, gathered, sha256 against the hash recorded from the real reference."""
    case = "64^3 Re=253 TRT over %d z-slabs (halo=%s), 30 iterations" % (world, halo)
    out = {"case": case, "multi": True, "ranks": world}
    try:
        with open(os.path.join(ROOT, "tests", "golden", "ref_hashes.json")) as fh:
            want = json.load(fh)["N64_Re253_m1_g1.0_d0_u0.1"]["30"]["sha256"]
        N = 64
        sim = lbm.Cavity3D(N=N, Re=253, rank=rank, nranks=world, device=local_rank, nccl_id=nccl_id)
        sim.set_option("host_local_layout", 1)
        if halo == "p2p":
            import cfd_codes_b200.distributed as D
            D.connect_direct_halo_via(sim, dist, dev)
        for k, v in opts.items():
            sim.set_option(k, v)
        sim.step(30)
        nloc = sim.info()["cells_local"]
        h = torch.empty(19 * nloc, dtype=torch.float64, pin_memory=True)
        sim._L.lbm_get_f(sim._h, h.data_ptr())
        sim.close()
        mine = h.to(dev)
        if rank == 0:
            parts = [mine.cpu().numpy().reshape(19, -1)]
            for r in range(1, world):
                k0, k1 = lbm.slab_range(N, r, world)
                buf = torch.empty(19 * (k1 - k0) * N * N, dtype=torch.float64, device=dev)
                dist.recv(buf, src=r)
                parts.append(buf.cpu().numpy().reshape(19, -1))
            import numpy as np
            full = np.concatenate(parts, axis=1)      # [q][k][j][i] = LU4 order (main.c:48)
            out["sha256_equals_reference"] = hashlib.sha256(np.ascontiguousarray(full).tobytes()).hexdigest() == want
        else:
            dist.send(mine, dst=0)
    except Exception as e:
        out["error"] = repr(e)
    return out


# ---- BASELINE config 5: D2Q9 8192^2 ------------------------------------------------------------------
def bench_d2q9(a, lbm, torch, rank=0, world=1, local_rank=0, nccl_id=None, dist=None, dev=None):
    """2D lid-driven cavity, D2Q9, fp64, N x N on `world` GPUs (row slabs); 144 B per cell update."""
    N = a.grid2d or 8192
    Re = 10000.0  # tau = 0.746 at N = 8192 (the 2D program's default Re = 100 would give tau = 25)
    block = a.block2d
    steps, warm = max(1, min(a.steps, 4)), max(3, min(a.warmup, 3))
    multi = world > 1

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sim = lbm.Cavity2D(N=N, M=N, Re=Re, rank=rank, nranks=world, device=local_rank, nccl_id=nccl_id)
    for _ in range(warm):
        sim.step(block)
        sim.residual()
    barrier()
    l0 = sim.info()["kernel_launches"]
    import ctypes
    sim._L.lbm_timer_start(sim._h)
    res = None
    for _ in range(steps):
        sim.step(block)
        res = sim.residual()
    ms = ctypes.c_float()
    sim._L.lbm_timer_stop(sim._h, ctypes.byref(ms))
    barrier()
    ms = allmax(ms.value)
    launches = sim.info()["kernel_launches"] - l0
    ms_k = allmax(sim.step_timed(200) / 200)
    peak, peak_src = peaks()
    cells = float(N) * N
    value = cells * block * steps / (ms * 1e-3) / 1e6
    achieved = B_ALG_2D * cells / world / (ms_k * 1e-3) / 1e9
    # e2e: host dist in, u / v / rho out, every step (serial: the 2D path has no asynchronous transfer API)
    e2e = None
    if not a.no_e2e:
        sim.set_option("host_local_layout", 1)
        nloc = sim.info()["cells_local"]
        h_d = torch.empty(9 * nloc, dtype=torch.float64, pin_memory=True)
        h_o = [torch.empty(nloc, dtype=torch.float64, pin_memory=True) for _ in range(3)]
        sim._L.lbm_get_f(sim._h, h_d.data_ptr())

        def e2e_step():
            sim._L.lbm_set_f(sim._h, h_d.data_ptr(), None)
            sim.step(block)
            r = sim.residual()
            sim._L.lbm_macrovars(sim._h, h_o[0].data_ptr(), h_o[1].data_ptr(), None, h_o[2].data_ptr(), None)
            return r

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        barrier()
        dt = allmax(time.perf_counter() - t0)
        e2e = {"value": cells * block * steps / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": int(9 * 8 * cells),
               "d2h_bytes_per_step": int(8 + 3 * 8 * cells), "ms_per_step": dt / steps * 1e3,
               "call": "lbm_set_f(host dist) + lbm_step(%d) + lbm_residual + lbm_macrovars(host u, v, rho), serial" % block}
    sim.close()
    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        try:
            cpu = cpu_reference_2d(N, host_threads())
        except Exception as e:
            cpu = {"value": None, "unit": "MLUPS", "cores": host_threads(), "kind": "reference", "sample": "failed: %r" % (e,)}
    return {
        "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "2D lid-driven cavity D2Q9 %d^2 fp64 Re=%g TRT (two-lattice pull variant of the same kernel)" % (N, Re),
                   "grid": N, "Re": Re, "iterations_per_step": block,
                   "l2": "inputs larger than L2: %.2f GB lattice vs 126 MB" % (9 * 8 * cells / world / 1e9), "residual_last": res,
                   "parity": "per-iteration parity vs the reference is unpinned by the reference itself (in-place, rank-dependent sweep, "
                             "SURVEY 8c); bit-equal to the synchronous oracle, same steady state (DESIGN.md 7)"},
        "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("d2q9/%d/%d" % (N, world)), "kernel": "lbm2d::k2_step", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": B_ALG_2D * cells / world, "launch_ms": ms_k, "frac_of_nominal_8TBs": achieved / 8000.0},
        "cpu_baseline": cpu}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=0, help="grid points per side (default 512)")
    ap.add_argument("--Re", type=int, default=1000)
    ap.add_argument("--block", type=int, default=500, help="iterations per bench step (the reference's skip)")
    ap.add_argument("--ref-iters", type=int, default=1, help="iterations per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-d2q9", action="store_true", help="N=1: do not append the config-5 (D2Q9 8192^2) sub-record")
    ap.add_argument("--e2e-steps", type=int, default=6, help="upper bound on the timed e2e steps")
    ap.add_argument("--block-x", type=int, default=0)
    ap.add_argument("--block-y", type=int, default=0)
    ap.add_argument("--pair", type=int, default=-1, help="two-iterations-per-pass kernel: -1 library default, 0 off, 1 on")
    ap.add_argument("--pair-variant", type=int, default=-1)
    ap.add_argument("--pair-lz", type=int, default=0)
    ap.add_argument("--halo", default="auto", choices=["auto", "nccl", "p2p"],
                    help="multi-GPU halo transport: NCCL send/recv, the fused direct-NVLink kernel (p2p), or auto = "
                         "p2p below 2^21 cells per GPU (latency-bound slabs), NCCL above (profiles/r01_halo_p2p_vs_nccl.md)")
    ap.add_argument("--workload", default="cavity3d", choices=["cavity3d", "d2q9"],
                    help="d2q9: BASELINE config 5 alone (8192^2 D2Q9 variant of the kernel)")
    ap.add_argument("--grid2d", type=int, default=0)
    ap.add_argument("--block2d", type=int, default=1000, help="iterations per step of the D2Q9 workload (the 2D program's prntInt)")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        return run_reference_arm(a, world, rank)
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            print("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % a.gpus, file=sys.stderr)
            return 2
        a.gpus = world

    import torch
    import torch.distributed as dist
    from lbmpkg import lbm

    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the hot path has no CPU fallback", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    multi = world > 1

    def new_nccl_id():
        if not multi:
            return None
        idt = torch.zeros(lbm.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.tensor(list(lbm.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    if multi:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    if a.workload == "d2q9":
        rec = bench_d2q9(a, lbm, torch, rank, world, local_rank, new_nccl_id(), dist, dev)
        if rank == 0:
            rec["clocks"] = None
            print(json.dumps(rec), flush=True)
        if multi:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    N = a.grid or 512
    sim = lbm.Cavity3D(N=N, Re=a.Re, UMAX=0.1, METHOD=1, rank=rank, nranks=world, device=local_rank,
                       track_residual=True, nccl_id=new_nccl_id())
    if a.halo == "auto":
        a.halo = "p2p" if (multi and float(N) ** 3 / world < 2 ** 21) else "nccl"
    if multi and a.halo == "p2p":
        import cfd_codes_b200.distributed as D
        D.connect_direct_halo_via(sim, dist, dev)
    opts = {}
    if a.block_x:
        opts["block_x"] = a.block_x
    if a.block_y:
        opts["block_y"] = a.block_y
    if a.pair >= 0:
        opts["pair"] = a.pair
    if a.pair_variant >= 0:
        opts["pair_variant"] = a.pair_variant
    if a.pair_lz > 0:
        opts["pair_lz"] = a.pair_lz
    for k, v in opts.items():
        sim.set_option(k, v)
    info0 = sim.info()
    cells = float(N) ** 3
    block = a.block

    def one_step():
        sim.step(block)
        return sim.residual()

    # ---- device-resident measurement ------------------------------------------------------------
    for _ in range(a.warmup):
        one_step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    i0 = sim.info()
    l0 = i0["kernel_launches"]
    t_wall = time.perf_counter()
    sim.timer_start()
    res = None
    for _ in range(a.steps):
        res = one_step()
    ms = sim.timer_stop()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    i1 = sim.info()
    launches = allsum(i1["kernel_launches"] - l0)
    ms = allmax(ms)
    value = cells * block * a.steps / (ms * 1e-3) / 1e6

    # ---- the dominant kernel: average duration of the plain iterations INSIDE the timed region (the
    # library brackets every batch of plain iterations of lbm_step with CUDA events on its launching
    # stream); slowest rank.  With the two-iterations-per-pass kernel one launch performs TWO updates
    # of every cell of the slab.
    n_k = i1["plain_step_iterations"] - i0["plain_step_iterations"]
    ms_k = allmax(i1["plain_step_ms"] - i0["plain_step_ms"])
    pair_on = bool(i1.get("pair_enabled", 0))
    pair_tma = pair_on and i1.get("pair_variant", -1) in (8, 9, 10)
    upd = 2 if pair_on else 1
    peak, peak_src = peaks()
    cells_per_launch = cells / world  # per GPU: its slab
    achieved = B_ALG * cells_per_launch / (ms_k * 1e-3 / n_k) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic("%d/%d%s" % (N, world, ("/pair_tma" if pair_tma else "/pair") if pair_on else "")),
                "kernel": ("lbm3d::k_step_pair<TRT> (two iterations per pass, variant %d: %s)" % (
                    i1.get("pair_variant", -1), "TMA-staged pulls, UTMALDG.4D" if pair_tma else "direct pulls")) if pair_on
                else "lbm3d::k_step<TRT, MODE_STEP>",
                "peak_source": peak_src, "updates_per_launch": upd,
                "algorithmic_bytes_per_launch": B_ALG * cells_per_launch * upd, "launch_ms": ms_k / n_k * upd,
                "launches_timed": int(n_k // upd), "timed_with": "CUDA events around the plain-iteration batches inside the timed region",
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "whole_step_GBs_per_gpu": B_ALG * value * 1e6 / 1e9 / world,
                # the same throughput with the residual cadence excluded (plain iterations only)
                "mlups_plain_iterations": cells / (ms_k * 1e-3 / n_k) / 1e6}

    # ---- end to end through the C ABI with pinned host buffers -----------------------------------
    e2e = None
    if not a.no_e2e:
        sim.set_option("host_local_layout", 1)
        nloc = info0["cells_local"]
        # several ranks: allocate the pinned buffers on the GPU's own NUMA node (PCIe traffic of 8 ranks
        # otherwise crosses the socket interconnect); undone before the CPU baseline uses all cores
        old_affinity = None
        if multi:
            import cfd_codes_b200.distributed as D
            old_affinity = D.bind_to_gpu_numa(local_rank)
        h_f = torch.empty(19 * nloc, dtype=torch.float64, pin_memory=True)
        h_mv = [[torch.empty(nloc, dtype=torch.float64, pin_memory=True) for _ in range(5)] for _ in range(2)]
        sim._L.lbm_get_f(sim._h, h_f.data_ptr())       # a valid state to restart from
        h2d = h_f.numel() * 8
        d2h = 8 + sum(t.numel() * 8 for t in h_mv[0])
        n_e2e = max(1, min(a.steps, a.e2e_steps))

        def serial_step():
            sim.raw_set_f(h_f.data_ptr(), None)
            sim.step(block)
            r = sim.residual()
            sim.raw_macrovars([t.data_ptr() for t in h_mv[0]])
            return r

        serial_step()                                   # warm-up (allocates the staging buffers)
        barrier()
        t0 = time.perf_counter()
        serial_step()
        barrier()
        dt_serial = allmax(time.perf_counter() - t0)

        # pipelined: the upload of state k+1 and the download of state k-1 overlap the iterations of state k
        sim.raw_set_f_async(h_f.data_ptr(), None)
        barrier()
        t0 = time.perf_counter()
        for k in range(n_e2e):
            sim.set_f_commit()                          # state k becomes current (stream-ordered after its upload)
            if k + 1 < n_e2e:
                sim.raw_set_f_async(h_f.data_ptr(), None)
            sim.step(block)
            sim.residual()
            sim.raw_macrovars_async([t.data_ptr() for t in h_mv[k & 1]])
        sim.io_wait()
        barrier()
        dt = allmax(time.perf_counter() - t0)
        e2e = {"value": cells * block * n_e2e / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": int(allsum(h2d)),
               "d2h_bytes_per_step": int(allsum(d2h)), "ms_per_step": dt / n_e2e * 1e3, "steps": n_e2e,
               "call": "per step: lbm_set_f_async(pinned host f) -> lbm_set_f_commit, lbm_step(%d), lbm_residual, "
                       "lbm_macrovars_async(pinned host u1,u2,u3,rho,vmag); lbm_io_wait at the end" % block,
               "overlap": "upload of step k+1 and download of step k-1 run on the copy stream during the iterations of step k",
               "serial": {"value": cells * block / dt_serial / 1e6, "unit": "MLUPS", "ms_per_step": dt_serial * 1e3,
                          "call": "lbm_set_f + lbm_step(%d) + lbm_residual + lbm_macrovars, nothing overlapped" % block}}
        del h_f, h_mv
        if old_affinity is not None:
            os.sched_setaffinity(0, old_affinity)
        e2e["numa_bound"] = old_affinity is not None
    sim.close()

    if multi:
        parity = parity_probe_multi(lbm, dist, torch, dev, rank, world, local_rank, new_nccl_id(), a.halo, opts)
    else:
        parity = parity_probe(lbm, local_rank, opts)

    # ---- CPU baseline (rank 0): the reference's own CPU code on this box's cores, a bounded sample ----
    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        try:
            thr = host_threads()
            m, kind, sample, sec = cpu_reference(N, a.Re, 1, 1, 3, thr, seconds=12.0)
            cpu = {"value": m, "unit": "MLUPS", "cores": thr, "kind": kind, "sample": sample, "seconds": sec}
        except Exception as e:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "MLUPS", "cores": host_threads(), "kind": "port", "sample": "failed: %r" % (e,)}

    d2q9 = None
    if not multi and not a.no_d2q9 and a.grid in (0, 512, 256):
        try:
            d2q9 = bench_d2q9(a, lbm, torch, dev=dev)
        except Exception as e:
            d2q9 = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_label(N, a.Re, world),
                       "grid": N, "Re": a.Re, "UMAX": 0.1, "method": "TRT", "iterations_per_step": block,
                       "step": "one residual block of the reference loop: %d iterations + diff" % block,
                       "l2": "inputs larger than L2: %.2f GB lattice per GPU vs 126 MB" % (19 * 8 * cells / world / 1e9),
                       "parallelism": "z-slab x%d" % world, "halo": a.halo if world > 1 else None, "residual_last": res,
                       "two_iterations_per_pass": pair_on, "options": opts,
                       "scaling_note": "every N runs the same %d^3 grid (strong scaling)" % N},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": t_wall, "parity_probe": parity,
        }
        if d2q9 is not None:
            line["config5_d2q9"] = d2q9
        print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
