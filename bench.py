#!/usr/bin/env python3
"""bench.py -- MLUPS of the D3Q19 cavity hot path on N B200s (one JSON line on rank 0).

A bench "step" is one residual block of the reference's time loop (main.c:104-124): `--block`
(default 500 = the reference's `skip`, main.c:43) iterations of collide+stream+walls followed by the
convergence residual.  Workloads (BASELINE.json configs):
    --gpus 1 : 256^3, Re=1000, TRT, fp64          (configs[1]; the lattice, 2.55 GB, is >> the 126 MB L2)
    --gpus N : 512^3, Re=1000, TRT, z-slabs on N GPUs, NCCL halo exchange (configs[2], strong scaling)
    --grid G : override (e.g. 512 on one GPU, 1024 on eight)

value  = cells * iterations / device time (CUDA events on the launching stream, max over ranks);
e2e    = the same block driven through the C ABI with HOST buffers: upload the distributions
         (lbm_set_f, pinned host memory), iterate, read back the residual and the macroscopic fields.
roofline: 304 algorithmic bytes per cell update (19 loads + 19 stores of fp64) against the measured
         copy bandwidth of MEASURED_PEAKS.json.
cpu_baseline / --impl reference: the reference's own main.c (oracle/_ref, built from
         /root/reference by oracle/build_ref.py) timed on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = 304.0  # bytes per lattice update, D3Q19 fp64: 2 * 19 * 8
NVSMI_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
           "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
           "clocks_event_reasons.sw_power_cap")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled every 200 ms DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.path = "/tmp/lbm_bench_clocks_%d_%d.csv" % (os.getpid(), index)
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
        return N ** 3 * n / sec / 1e6, "port", "%d+%d iterations of the %d^3 grid (%.1f s), oracle port" % (warm, n, N, sec), sec
    sec = o.time_steps(steps * iters)
    o.close()
    return N ** 3 * steps * iters / sec / 1e6, "port", "%d+%d steps x %d iterations of the %d^3 grid, oracle port" % (
        warm, steps, iters, N), sec


def run_reference_arm(a, world, rank):
    if rank != 0:
        return 0
    N = a.grid or (256 if a.gpus == 1 else 512)
    threads = host_threads()
    iters = a.ref_iters
    mlups, kind, sample, sec = cpu_reference(N, a.Re, iters, a.warmup, a.steps, threads)
    line = {
        "impl": "reference", "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec / a.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D lid-driven cavity D3Q19 %d^3 fp64 Re=%d TRT (reference CPU implementation)" % (N, a.Re),
                   "grid": N, "Re": a.Re, "method": "TRT", "iterations_per_step": iters},
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def parity_probe(lbm, device):
    """Not a measurement: a 2-second sanity probe so that every bench line carries its own evidence
    that the timed library computes the reference's bits -- 12^3 TRT cavity, 10 iterations, sha256
    of the raw distributions against the hash recorded from the REAL reference
    (tests/golden/ref_hashes.json, made by tests/golden/make_golden.py from oracle/_ref)."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "tests", "golden", "ref_hashes.json")) as fh:
            want = json.load(fh)["N12_Re100_m1_g1.0_d0_u0.1"]["10"]["sha256"]
        with lbm.Cavity3D(N=12, Re=100, device=device) as sim:
            sim.step(10)
            got = hashlib.sha256(sim.f().tobytes()).hexdigest()
        return {"case": "12^3 Re=100 TRT, 10 iterations", "sha256_equals_reference": got == want}
    except Exception as e:
        return {"case": "12^3 Re=100 TRT, 10 iterations", "error": repr(e)}


def run_d2q9(a):
    """BASELINE config 5: 2D lid-driven cavity, D2Q9, 8192^2 fp64 on one B200 (144 B per cell update)."""
    from lbmpkg import lbm
    N = a.grid or 8192
    Re = 10000.0  # tau = 0.746 at N = 8192 (the 2D program's default Re = 100 would give tau = 25)
    block = a.block
    with lbm.Cavity2D(N=N, M=N, Re=Re) as sim:
        for _ in range(a.warmup):
            sim.step(block)
            sim.residual()
        sampler = ClockSampler(0)
        sampler.start()
        l0 = sim.info()["kernel_launches"]
        sim._L.lbm_timer_start(sim._h)
        for _ in range(a.steps):
            sim.step(block)
            res = sim.residual()
        import ctypes
        ms = ctypes.c_float()
        sim._L.lbm_timer_stop(sim._h, ctypes.byref(ms))
        ms = ms.value
        clocks = sampler.stop()
        launches = sim.info()["kernel_launches"] - l0
        ms_k = sim.step_timed(200) / 200
    peak, peak_src = peaks()
    cells = float(N) * N
    value = cells * block * a.steps / (ms * 1e-3) / 1e6
    achieved = 144.0 * cells / (ms_k * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            traffic = json.load(fh).get("d2q9/%d/1" % N, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    print(json.dumps({
        "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "2D lid-driven cavity D2Q9 %d^2 fp64 Re=%g TRT (two-lattice pull variant of the same kernel)" % (N, Re),
                   "grid": N, "Re": Re, "iterations_per_step": block, "l2": "inputs larger than L2: %.2f GB lattice vs 126 MB" % (9 * 8 * cells / 1e9),
                   "residual_last": res},
        "clocks": clocks, "e2e": None, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "lbm2d::k2_step", "peak_source": peak_src, "algorithmic_bytes_per_launch": 144.0 * cells,
                     "launch_ms": ms_k, "frac_of_nominal_8TBs": achieved / 8000.0},
        "cpu_baseline": None}), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=0, help="grid points per side (default 256 on 1 GPU, 512 on >1)")
    ap.add_argument("--Re", type=int, default=1000)
    ap.add_argument("--block", type=int, default=500, help="iterations per bench step (the reference's skip)")
    ap.add_argument("--ref-iters", type=int, default=1, help="iterations per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--block-x", type=int, default=0)
    ap.add_argument("--block-y", type=int, default=0)
    ap.add_argument("--halo", default="auto", choices=["auto", "nccl", "p2p"],
                    help="multi-GPU halo transport: NCCL send/recv, the fused direct-NVLink kernel (p2p), or auto = "
                         "p2p below 2^21 cells per GPU (latency-bound slabs), NCCL above (profiles/r01_halo_p2p_vs_nccl.md)")
    ap.add_argument("--workload", default="cavity3d", choices=["cavity3d", "d2q9"],
                    help="d2q9: BASELINE config 5 (8192^2 D2Q9 variant of the kernel, 1 GPU; extra line, not the headline)")
    a = ap.parse_args()
    if a.workload == "d2q9":
        return run_d2q9(a)

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
    nccl_id = None
    if multi:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        idt = torch.zeros(lbm.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.tensor(list(lbm.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())

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

    N = a.grid or (256 if world == 1 else 512)
    sim = lbm.Cavity3D(N=N, Re=a.Re, UMAX=0.1, METHOD=1, rank=rank, nranks=world, device=local_rank,
                       track_residual=True, nccl_id=nccl_id)
    if a.halo == "auto":
        a.halo = "p2p" if (multi and float(N) ** 3 / world < 2 ** 21) else "nccl"
    if multi and a.halo == "p2p":
        import cfd_codes_b200.distributed as D
        if not dist.is_initialized():
            raise RuntimeError("process group missing")
        # the blobs travel as CPU tensors: needs a CPU-capable backend
        D.connect_direct_halo_via(sim, dist, dev)
    if a.block_x:
        sim.set_option("block_x", a.block_x)
    if a.block_y:
        sim.set_option("block_y", a.block_y)
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

    # ---- the dominant kernel: average duration of the plain step iterations INSIDE the timed region
    # (the library brackets every batch of plain iterations of lbm_step with CUDA events on its
    # launching stream); slowest rank.
    n_k = i1["plain_step_iterations"] - i0["plain_step_iterations"]
    ms_k = allmax(i1["plain_step_ms"] - i0["plain_step_ms"])
    peak, peak_src = peaks()
    cells_per_launch = cells / world  # per GPU: its slab
    achieved = B_ALG * cells_per_launch / (ms_k * 1e-3 / n_k) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            traffic = json.load(fh).get("%d/%d" % (N, world), {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "lbm3d::k_step<TRT, MODE_STEP>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": B_ALG * cells_per_launch, "launch_ms": ms_k / n_k,
                "launches_timed": int(n_k), "timed_with": "CUDA events around the plain-step batches inside the timed region",
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "whole_step_GBs_per_gpu": B_ALG * value * 1e6 / 1e9 / world,
                # the same throughput with the residual cadence excluded (plain iterations only)
                "mlups_plain_iterations": cells / (ms_k * 1e-3 / n_k) / 1e6}

    # ---- end to end through the C ABI with host buffers -----------------------------------------
    e2e = None
    if not a.no_e2e:
        sim.set_option("host_local_layout", 1)
        nloc = info0["cells_local"]
        h_f = torch.empty(19 * nloc, dtype=torch.float64, pin_memory=True)
        h_prev = torch.empty(19 * nloc, dtype=torch.float64, pin_memory=True)
        h_mv = [torch.empty(nloc, dtype=torch.float64, pin_memory=True) for _ in range(5)]
        sim._L.lbm_get_f(sim._h, h_f.data_ptr())       # a valid state to restart from
        h_prev.copy_(h_f)
        h2d = 2 * 0 + h_f.numel() * 8                   # lbm_set_f uploads f (f_prev: only N*nz doubles are read)
        d2h = 8 + sum(t.numel() * 8 for t in h_mv)

        def e2e_step():
            sim.raw_set_f(h_f.data_ptr(), h_prev.data_ptr())
            sim.step(block)
            r = sim.residual()
            sim.raw_macrovars([t.data_ptr() for t in h_mv])
            return r

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        barrier()
        dt = allmax(time.perf_counter() - t0)
        e2e = {"value": cells * block * a.steps / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": int(allsum(h2d)),
               "d2h_bytes_per_step": int(allsum(d2h)), "ms_per_step": dt / a.steps * 1e3,
               "call": "lbm_set_f(host) + lbm_step(%d) + lbm_residual + lbm_macrovars(host)" % block}
    sim.close()

    parity = parity_probe(lbm, local_rank) if rank == 0 else None

    # ---- CPU baseline (rank 0, N=1 only): the reference's own main.c on this box's cores ---------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            thr = host_threads()
            m, kind, sample, sec = cpu_reference(N, a.Re, 1, 1, 3, thr, seconds=12.0)   # a bounded ~12 s sample
            cpu = {"value": m, "unit": "MLUPS", "cores": thr, "kind": kind, "sample": sample, "seconds": sec}
        except Exception as e:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "MLUPS", "cores": host_threads(), "kind": "port", "sample": "failed: %r" % (e,)}

    if rank == 0:
        line = {
            "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3D lid-driven cavity D3Q19 %d^3 fp64 Re=%d TRT%s" % (
                           N, a.Re, "" if world == 1 else ", z-slabs over %d GPUs" % world),
                       "grid": N, "Re": a.Re, "UMAX": 0.1, "method": "TRT", "iterations_per_step": block,
                       "step": "one residual block of the reference loop: %d iterations + diff" % block,
                       "l2": "inputs larger than L2: %.2f GB lattice per GPU vs 126 MB" % (19 * 8 * cells / world / 1e9),
                       "parallelism": "z-slab x%d" % world, "halo": a.halo if world > 1 else None, "residual_last": res,
                       "scaling_note": "N=1 runs BASELINE configs[1] (256^3); N>1 strong-scale configs[2] (512^3) over z-slabs; "
                                       "MLUPS is normalised per cell update (one GPU: 21.1 GLUPS at 256^3, 21.3 at 512^3, "
                                       "profiles/r01_summary.md); use --grid 512 at N=1 for the strict strong-scaling base"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": t_wall, "parity_probe": parity,
        }
        print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
