#!/usr/bin/env python
"""bench.py -- H*v matvec throughput of the engine on the BASELINE.json headline workloads.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

Headline workload: XXZ chain L=32, Sz=0 sector (601,080,390 states, 192 terms) -- the configuration BASELINE.json's metric
is quoted on; one "step" = one matrix-free matvec y = H x over the whole basis.  N>1 (one process per GPU under torchrun):
the same problem row-sharded over N ranks (strong scaling) through the library's multi-GPU context -- partition, halo
copies over NVLink, NCCL fences and the Lanczos all-reduces all run inside libedcuda.so (torch.distributed only hands out
the NCCL unique id).  The line also carries, as sub-objects, BASELINE's other named workloads at the same N: `lanczos`
(config 5: 100 device-resident Lanczos steps), `tri6x6` (config 4: 6x6 triangular, k=0 A1, reduced matvec matrix-free
and through the cached CSR) and, at N=1, `sparse` (configs 1-2 assembly times) and `parity_sample` (max relative error
of the timed kernel's output against the oracle's C twin on sampled rows of the same x).
`--impl reference` times the reference algorithm's CPU restatement (oracle/ed_oracle_c.c, OpenMP, every host core) on a
bounded row sample of the same workload; it never loads libedcuda.so.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

WORKLOADS = {
    "xxz_chain_L32_sz0": dict(n=32, kind="xxz", desc="XXZ chain L=32 (Delta=1), Sz=0, periodic, Pauli normalisation"),
    "j1j2_chain_L28_sz0": dict(n=28, kind="j1j2", desc="J1-J2 chain L=28 (J2=0.5), Sz=0"),
    "xxz_chain_L24_sz0": dict(n=24, kind="xxz", desc="XXZ chain L=24, Sz=0 (small, for quick checks)"),
    "xxz_chain_L28_sz0": dict(n=28, kind="xxz", desc="XXZ chain L=28, Sz=0"),
    "xxz_chain_L30_sz0": dict(n=30, kind="xxz", desc="XXZ chain L=30, Sz=0"),
    "xxz_chain_L16_sz0": dict(n=16, kind="xxz", desc="Heisenberg chain L=16, Sz=0 (reference CPU-runnable case)"),
    "tri6x6_k0A1_sz0": dict(n=36, kind="tri", desc="6x6 triangular Heisenberg, T x| C6v k=0 A1, Sz=0: reduced matvec (ComplexF64), |G|=432"),
}

_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner at N > 1), so the real
    stdout is kept aside for emit() and file descriptor 1 is pointed at stderr for everything else."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def host_cores() -> int:
    """cores this process may run on -- NOT OMP_NUM_THREADS, which torchrun sets to 1 for its workers"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def nnz_eff(n_sites: int, n_bonds: int) -> int:
    """D + n_bonds * 2 * C(N-2, N/2-1)."""
    return math.comb(n_sites, n_sites // 2) + n_bonds * 2 * math.comb(n_sites - 2, n_sites // 2 - 1)


def n_bonds_of(w) -> int:
    return w["n"] if w["kind"] == "xxz" else 2 * w["n"]


def base_config(name, w):
    """identical in both arms (the driver compares the dicts): the workload, nothing about how it was run"""
    n = w["n"]
    if w["kind"] == "tri":
        return {"workload": name, "description": w["desc"], "n_sites": n, "dim": 21029820, "n_terms": 648}
    return {"workload": name, "description": w["desc"], "n_sites": n, "dim": math.comb(n, n // 2), "n_terms": 6 * n_bonds_of(w)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm (oracle only)
def oracle_terms(w):
    """the workload's term list from the ORACLE's operator algebra (simplify order): the CPU arm never imports the product"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ed_oracle as O
    n = w["n"]
    hs, pauli = O.spin_half_system(n)

    def bonds_op(bonds, j):
        h = None
        for (a, b) in bonds:
            t = (2.0 * j) * (pauli(a, "+") * pauli(b, "-")) + (2.0 * j) * (pauli(a, "-") * pauli(b, "+")) + float(j) * (pauli(a, "z") * pauli(b, "z"))
            h = t if h is None else h + t
        return h

    h = bonds_op([(i, (i + 1) % n) for i in range(n)], 1.0)
    if w["kind"] == "j1j2":
        h = h + bonds_op([(i, (i + 2) % n) for i in range(n)], 0.5)
    return O.term_arrays(O.simplify(h))


class CpuReference:
    """Reference-algorithm restatement (oracle C twin: term walk in order + binary search per hit + static row
    partition = the reference's apply_parallel!) on a bounded contiguous row sample, every core of the host."""

    def __init__(self, w, x=None):
        os.environ["ED_ORACLE_NATIVE"] = "1"      # rebuilt on this machine with -march=native (falls back to the portable .so)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import numpy as np
        import ed_oracle_c as OC
        self.np, self.OC = np, OC
        self.threads = host_cores()
        OC.set_num_threads(self.threads)          # explicit: torchrun exports OMP_NUM_THREADS=1
        self.n = w["n"]
        self.n_bonds = n_bonds_of(w)
        self.terms = oracle_terms(w)
        self.basis = OC.basis_fixed_popcount(self.n, self.n // 2)
        self.dim = len(self.basis)
        self.x = x if x is not None else np.random.default_rng(20260717 + 5).standard_normal(self.dim)
        self.n0 = min(self.dim, 20000 * self.threads)
        self.rate = self.n0 / self._time(self.n0)[0]          # rows per second (calibration slice)

    def _time(self, rows, lo=None):
        lo = max(0, self.dim // 2 - rows // 2) if lo is None else lo
        out = self.np.zeros(rows)
        t0 = time.perf_counter()
        self.OC.apply(self.basis, self.terms, self.x, out, lo, lo + rows)
        return time.perf_counter() - t0, lo, out

    def sample(self, target_seconds):
        rows = int(min(self.dim, max(self.n0, self.rate * target_seconds)))
        dt, lo, _ = self._time(rows)
        matvec_s = dt * self.dim / rows
        return {"value": 1.0 / matvec_s, "unit": "matvec/s", "cores": self.threads, "kind": "port",
                "sample": f"rows [{lo},{lo + rows}) of {self.dim} ({rows / self.dim:.4%}) timed {dt:.2f}s on {self.threads} OpenMP threads "
                          f"(all cores of the host; OMP_NUM_THREADS ignored), extrapolated to the full matvec; reference-algorithm "
                          f"restatement oracle/ed_oracle_c.c [{self.OC.BUILD}], not Julia (not installed)",
                "seconds_per_matvec_extrapolated": matvec_s, "gnnz_per_s": nnz_eff(self.n, self.n_bonds) / matvec_s / 1e9}

    def parity(self, y_rows, blocks):
        """max |y - oracle| / max |oracle| over the given (lo, hi) row blocks; y_rows(lo, hi) -> numpy"""
        worst, scale, total = 0.0, 0.0, 0
        for lo, hi in blocks:
            _, _, exp = self._time(hi - lo, lo)
            worst = max(worst, float(self.np.max(self.np.abs(y_rows(lo, hi) - exp))))
            scale = max(scale, float(self.np.max(self.np.abs(exp))))
            total += hi - lo
        return {"max_rel_err": worst / scale if scale > 0 else None, "rows": total, "blocks": len(blocks),
                "against": "oracle/ed_oracle_c.c apply_parallel! restatement on the same x"}


def run_reference(args, w, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if w["kind"] == "tri":
        # the reference keeps 32 B per PARENT state (9.08e9 states -> 290 GB of maps): not runnable on this host
        emit({"impl": "reference", "unavailable": "reference algorithm needs ~360 GB of host memory for the 6x6 triangular parent space (SURVEY 8d); no CPU arm for this workload"})
        return
    per_step = max(1.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    ref = CpuReference(w)
    vals, cb = [], None
    for i in range(args.warmup + args.steps):
        cb = ref.sample(per_step)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = sum(vals) / len(vals)
    cb["value"] = v
    emit({"impl": "reference", "metric": "H*v matvecs/sec", "value": v, "unit": "matvec/s", "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True,
          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "gnnz_per_s": nnz_eff(w["n"], n_bonds_of(w)) * v / 1e9, "config": base_config(name, w), "cpu_baseline": cb,
          "e2e": {"value": v, "unit": "matvec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


# ------------------------------------------------------------------------------------------ engine arm
def build_model(ed, w):
    n = w["n"]
    if w["kind"] == "xxz":
        hs, h = ed.models.xxz_chain(n, 1.0)
    else:
        hs, h = ed.models.j1j2_chain(n, 0.5)
    return hs, h


def time_sharded(ctx, fn, warmup, steps):
    """fn() enqueues one step on the context's streams; returns ms per step (device time, max over ranks)"""
    for _ in range(warmup):
        fn()
    ctx.barrier()
    ctx.timer_record(0)
    for _ in range(steps):
        fn()
    ctx.timer_record(1)
    return ctx.timer_elapsed(0, 1) / steps


def bench_tri6x6(ed, np, ctx, steps, peak):
    """BASELINE config 4: reduced matvec of the 6x6 triangular k=0 A1 sector, rows sharded over the ranks of ctx
    (every rank enumerates the reduced basis itself; x all-gathered over NCCL per matvec): matrix-free (K6), then with
    the owned rows cached as CSR."""
    from edcuda.distributed import ShardedOperator
    t0 = time.perf_counter()
    hs, h = ed.models.heisenberg_triangular(6)
    symops = ed.lattices.triangular_space_group_irrep(6, "A1")

    def make():
        return ed.represent(ed.symmetry_reduce(ed.represent(ed.HilbertSpaceSector(hs, 0)), symops), h)

    sh = ShardedOperator(ctx, make)
    ctx.sync()
    t_setup = time.perf_counter() - t0
    d = sh.dimension
    x, y = sh.vector(), sh.vector()
    x.randn(20260717 + 4, 1.0 / math.sqrt(d))
    l0 = ed.kernel_launch_count()
    ms_free = time_sharded(ctx, lambda: sh.apply(y, x), 1, max(2, min(steps, 3)))
    launches_free = (ed.kernel_launch_count() - l0) // (1 + max(2, min(steps, 3)))
    dot_free = sh.apply(y, x, dot=True)
    t0 = time.perf_counter()
    nnz_local = sum(o.cache_matrix() for o in sh.oprs)
    ctx.sync()
    t_cache = time.perf_counter() - t0
    nnz = int(ctx.allreduce([float(nnz_local)])[0])
    ms_csr = time_sharded(ctx, lambda: sh.apply(y, x), 3, max(steps, 10))
    dot_csr = sh.apply(y, x, dot=True)
    n_local = sh.info(0)["n_local"]
    alg = 40.0 * n_local                       # 8 B word + 16 B x + 16 B y per owned row
    out = {"workload": "tri6x6_k0A1_sz0", "dim": d, "parent_dim": 9075135300, "n_terms": len(h.terms), "group_order": len(symops),
           "dtype": "c128", "rows_per_gpu": n_local, "exchange": "none" if ctx.world == 1 else "NCCL all-gather of x per matvec (in-library), window by window on a side stream; the cached SpMV's column-block pass b starts when window b is there",
           "setup_seconds": t_setup,
           "matrix_free": {"ms_per_matvec": ms_free, "matvec_per_s": 1e3 / ms_free, "gnnz_per_s": nnz / (ms_free * 1e-3) / 1e9,
                           "roofline": {"bound": "hbm", "achieved": alg / (ms_free * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                        "frac": alg / (ms_free * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg},
                           "gpu_launches": int(launches_free), "kernel": "K6 staged (emit / canonicalize / combine)",
                           "note": "bound by shared-memory table look-ups of the orbit search per off-diagonal hit, not by HBM (SURVEY H1)"},
           "cached_csr": {"nnz": nnz, "assemble_seconds": t_cache, "ms_per_matvec": ms_csr, "matvec_per_s": 1e3 / ms_csr,
                          "gnnz_per_s": nnz / (ms_csr * 1e-3) / 1e9,
                          "roofline": {"bound": "hbm", "achieved": alg / (ms_csr * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": alg / (ms_csr * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg}},
           "checksum_x_dot_Hx": [dot_free.real, dot_csr.real]}
    x.close(); y.close(); sh.close()
    return out


def bench_sparse(ed, np):
    """sparse() assembly times of BASELINE configs 1 and 2 (and L=24 for scale): seconds and nnz/s, device work only"""
    out = {}

    def timed(opr):
        opr.sparse_csc()                     # warm-up (allocations, term upload)
        t0 = time.perf_counter()
        cp, rv, nz = opr.sparse_csc()
        return time.perf_counter() - t0, int(cp[-1] - 1)

    hs, h = ed.models.heisenberg_chain(16)
    dt, nnz = timed(ed.represent(ed.represent(ed.HilbertSpaceSector(hs, 0)), h))
    out["config1_L16"] = {"seconds": dt, "nnz": nnz, "nnz_per_s": nnz / dt}
    hs, h = ed.models.heisenberg_square(4, 4)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    t_all, nnz_all = 0.0, 0
    for k2 in range(4):
        for k1 in range(4):
            rhsr = ed.symmetry_reduce(hsr, ed.lattices.torus_translation_irrep(4, 4, k1, k2))
            dt, nnz = timed(ed.represent(rhsr, h))
            t_all += dt; nnz_all += nnz
    out["config2_sq4x4_16_sectors"] = {"seconds": t_all, "nnz": nnz_all, "nnz_per_s": nnz_all / t_all}
    hs, h = ed.models.xxz_chain(24, 1.0)
    dt, nnz = timed(ed.represent(ed.represent(ed.HilbertSpaceSector(hs, 0)), h))
    out["xxz_L24"] = {"seconds": dt, "nnz": nnz, "nnz_per_s": nnz / dt, "includes": "device assembly + 0.7 GB D2H of the CSC arrays"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="edcuda")
    ap.add_argument("--workload", default="xxz_chain_L32_sz0")
    ap.add_argument("--kernel", type=int, default=0, help="0 = automatic (fast path), 1 = generic term-walk kernel")
    ap.add_argument("--exchange", default="auto", choices=["auto", "halo", "push", "pull", "cepush", "nccl", "allgather"],
                    help="N>1: how remote rows of x reach a rank: halo copies over NVLink (packed by the owner, pulled by copy engines) or an NCCL all-gather per matvec")
    ap.add_argument("--chunks", type=int, default=0, help="N>1: launch chunks per matvec (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the lanczos / tri6x6 / sparse / parity_sample sub-objects")
    ap.add_argument("--lanczos", type=int, default=100, help="device-resident Lanczos steps for the `lanczos` sub-object (0 = skip)")
    args = ap.parse_args()
    claim_stdout()
    name = args.workload
    w = WORKLOADS[name]
    if args.impl == "reference":
        run_reference(args, w, name)
        return
    args.warmup = max(args.warmup, 3)

    sys.path.insert(0, os.path.join(ROOT, "exactdiagonalization.jl_b200"))
    import numpy as np
    import torch
    import edcuda as ed
    from edcuda.distributed import Context, ShardedOperator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if ed.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libedcuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    ctx = Context.from_env()                  # world 1: a trivial context; world > 1: NCCL communicator inside the library
    peak, peak_src = peaks()

    if w["kind"] == "tri":
        tri = bench_tri6x6(ed, np, ctx, args.steps, peak)
        if rank == 0:
            mf = tri["matrix_free"]
            emit({"metric": "H*v matvecs/sec", "value": mf["matvec_per_s"], "unit": "matvec/s", "n_gpus": world, "steps": args.steps,
                  "warmup": args.warmup, "ms_per_step": mf["ms_per_matvec"], "higher_is_better": True, "scaling": "strong",
                  "vs_baseline": None, "dtype": "c128", "data": "synthetic", "gnnz_per_s": mf["gnnz_per_s"],
                  "config": base_config(name, w), "roofline": dict(mf["roofline"], traffic=None, peak_source=peak_src),
                  "gpu_launches": mf["gpu_launches"], "tri6x6": tri})
        ctx.close()
        return

    n = w["n"]
    n_bonds = n_bonds_of(w)
    hs, h = build_model(ed, w)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    dim = hsr.dimension
    seed = 20260717 + 5
    details = {}
    dev = torch.device("cuda", local_rank)

    if world == 1:
        # ---- one GPU: the plain C-ABI call (ed_apply_async) on torch's current stream, CUDA events around it
        import ctypes as C
        from edcuda._lib import lib, check
        opr = ed.represent(hsr, h).set_kernel(args.kernel)
        x = torch.empty(dim, dtype=torch.float64, device=dev)
        y = torch.zeros(dim, dtype=torch.float64, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib.ed_set_stream(stream, 1))
        check(lib.ed_vector_randn_async(x.data_ptr(), dim, ed.ED_F64, seed, 0))
        check(lib.ed_vector_scale_async(x.data_ptr(), dim, ed.ED_F64, 1.0 / math.sqrt(dim)))

        def step():
            check(lib.ed_apply_async(opr._handle, y.data_ptr(), x.data_ptr(), ed.ED_F64, 0, 0, None))

        sampler = ClockSampler(local_rank)
        sampler.start()                       # nvidia-smi needs ~0.2 s to deliver its first sample: it runs from the warm-up on
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        t_w = time.perf_counter()
        while not sampler.samples and time.perf_counter() - t_w < 3.0:     # same load as the timed steps
            step()
            torch.cuda.synchronize()
        launches0 = ed.kernel_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        torch.cuda.synchronize()
        lib.ed_set_stream(None, 0)
        launches = ed.kernel_launch_count() - launches0
        clocks = sampler.stop()
        ms_per_step = ev0.elapsed_time(ev1) / args.steps
        kern_ms = ms_per_step                  # one launch per step: the step IS the kernel
        n_local = dim
        checksum = float(torch.dot(x, y))
        details.update({"rows_per_gpu": dim, "sharding": "none", "exchange": "none",
                        "kernel": "generic term-walk (k2_apply_generic)" if args.kernel == 1 else "tiled U(1) kernel k2_apply_u1"})
    else:
        # ---- N GPUs: the library's multi-GPU context end to end
        sh = ShardedOperator(ctx, lambda: ed.represent(hsr, h).set_kernel(args.kernel), exchange=args.exchange, n_chunks=args.chunks)
        info = sh.info(0)
        xv, yv = sh.vector(), sh.vector()
        xv.randn(seed, 1.0 / math.sqrt(dim))
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            sh.apply(yv, xv)
        ctx.barrier()
        for _ in range(40):                   # keep the load up until nvidia-smi has delivered a sample (all ranks in step)
            flag = ctx.allreduce([1.0 if (rank != 0 or sampler.samples) else 0.0], "sum")[0]
            if flag >= world:
                break
            for _ in range(5):
                sh.apply(yv, xv)
            ctx.barrier()
        launches0 = ed.kernel_launch_count()
        ctx.timer_record(0)
        for _ in range(args.steps):
            sh.apply(yv, xv)
        ctx.timer_record(1)
        ms_per_step = ctx.timer_elapsed(0, 1) / args.steps
        ctx.barrier()
        launches = ed.kernel_launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        kern_ms = ms_per_step
        n_local = info["n_local"]
        checksum = sh.apply(yv, xv, dot=True).real
        halo_max = int(ctx.allreduce([float(info["n_halo"])], "max")[0])
        phases = sh.profile(yv, xv) if info["exchange"] == "halo" else None
        details.update({"rows_per_gpu": n_local, "sharding": "tiles assigned by the library's planner (ed_u1_shard_layout)" if info["exchange"] == "halo" else "contiguous row ranges",
                        "exchange": ("halo exchange over NVLink (%s), interior tiles first, %d launch chunks, one NCCL fence per matvec" % (
                                     {"owner pushes": "owners write the tiles their peers read into the peers' halo buffers with remote stores and bump a per-chunk arrival counter",
                                      "owner copy-engine pushes": "owner-side pack, the owner's copy engines write the pieces into the peers' halo buffers, per-chunk arrival counters",
                                      "nccl send/recv": "owner-side pack, one grouped ncclSend/ncclRecv per launch chunk",
                                      "copy-engine pulls": "owner-side pack + reader-side copy-engine pulls of the packed tiles"}[info["halo_transport"]], info["n_chunks"]))
                        if info["exchange"] == "halo" else "NCCL all-gather of x per matvec", "halo_rows_max": halo_max,
                        "nvlink_bytes_per_rank_per_matvec_max": halo_max * 8, "pulls_per_matvec": info["n_pulls"],
                        "phases_run_back_to_back_ms": phases,
                        "halo_GBps_per_rank": (halo_max * 8 / (phases["pull_ms"] * 1e-3) / 1e9) if phases and phases["pull_ms"] > 0 else None,
                        "nccl_version": ctx.nccl_version, "collectives": "in-library NCCL (no torch.distributed on the data path)",
                        "kernel": "generic term-walk (k2_apply_generic)" if args.kernel == 1 else "tiled U(1) kernel k2_apply_u1"})
    value = 1e3 / ms_per_step
    details["checksum_x_dot_Hx"] = checksum

    # ---- e2e: HOST vectors in and out, H2D / D2H inside the timed region ----------------------------------------
    e2e = None
    if not args.no_e2e:
        e_steps = max(2, min(args.steps, 5))
        if world == 1:
            xh = torch.empty(dim, dtype=torch.float64).pin_memory()
            yh = torch.empty(dim, dtype=torch.float64).pin_memory()
            xh.copy_(x)
            xn, yn = xh.numpy(), yh.numpy()
            ed.mul_b(yn, opr, xn)             # warm-up (staging allocations)
            t0 = time.perf_counter()
            for _ in range(e_steps):
                ed.mul_b(yn, opr, xn)
            dt = (time.perf_counter() - t0) / e_steps
            assert abs(float(np.dot(xn, yn)) - checksum) <= 1e-9 * max(1.0, abs(checksum))
            e2e = {"value": 1.0 / dt, "unit": "matvec/s", "h2d_bytes_per_step": int(dim * 8), "d2h_bytes_per_step": int(dim * 8),
                   "ms_per_step": dt * 1e3, "api": "edcuda.mul_b(out, opr, x) = mul!(out, opr, x) with pinned host vectors -> ed_apply"}
        else:
            # every rank keeps its rows of x and y in pinned host memory; a step = H2D of its x rows, the sharded matvec, D2H of its y rows
            st = torch.cuda.ExternalStream(ctx.stream(0), device=dev)
            xt, yt = xv.tensor(0), yv.tensor(0)
            xh = torch.empty(n_local, dtype=torch.float64).pin_memory()
            yh = torch.empty(n_local, dtype=torch.float64).pin_memory()
            with torch.cuda.stream(st):
                xh.copy_(xt)
            ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                with torch.cuda.stream(st):
                    xt.copy_(xh, non_blocking=True)
                sh.apply(yv, xv)
                with torch.cuda.stream(st):
                    yh.copy_(yt, non_blocking=True)
                ctx.sync()
            ctx.barrier()
            dt = float(ctx.allreduce([(time.perf_counter() - t0) / e_steps], "max")[0])
            e2e = {"value": 1.0 / dt, "unit": "matvec/s", "h2d_bytes_per_step": int(dim * 8), "d2h_bytes_per_step": int(dim * 8),
                   "ms_per_step": dt * 1e3, "api": "ed_apply_sharded with per-rank pinned host shards (H2D of the owned rows + pack/pull exchange + kernel + D2H)"}
            del xh, yh

    extras = {}
    if not args.no_extras:
        # ---- parity of the timed output against the oracle's C twin (N=1: sampled rows of the same x) ------------
        cpu = None
        if world == 1 and w["kind"] in ("xxz", "j1j2"):
            xh_np = xh.numpy() if not args.no_e2e else x.cpu().numpy()
            cpu = CpuReference(w, x=xh_np)
            half = math.comb(n - 1, n // 2)
            blocks = sorted({(max(0, lo), min(dim, lo + 20000)) for lo in (0, dim - 20000, half - 10000, dim // 2, dim // 3, (1 << 29) - 10000 if dim > (1 << 29) + 10000 else dim // 5)})
            extras["parity_sample"] = cpu.parity(lambda lo, hi: y[lo:hi].cpu().numpy(), blocks)
            if not args.no_cpu_baseline:
                extras["cpu_baseline"] = cpu.sample(12.0)
            del cpu
        # ---- config 5: device-resident Lanczos on the same representation -------------------------------------
        if args.lanczos > 0:
            if world == 1:
                del x, y
                if e2e is not None:
                    del xh, yh, xn, yn
                ed._lib.lib.ed_release_staging()
                torch.cuda.empty_cache()
                sh = ShardedOperator(ctx, lambda: opr)
            else:
                xv.close(); yv.close()
            sh.lanczos(3, seed=1)            # warm-up
            res = sh.lanczos(args.lanczos, seed=20260717)
            extras["lanczos"] = {"steps": int(res.steps), "ms_per_step": res.ms_per_step, "lowest_ritz": float(res.ritz[0]) if len(res.ritz) else None,
                                 "e0_per_site_over_4": (float(res.ritz[0]) / (4.0 * n)) if len(res.ritz) else None,
                                 "api": "ed_lanczos_sharded (in-library loop; scalars all-reduced over NCCL at N>1)",
                                 "note": "three-term Lanczos, unnormalised device-resident Krylov vectors, fused <u,Hu> and update+norm kernels"}
            sh.close()
        torch.cuda.empty_cache()
        # ---- config 4 (6x6 triangular) and sparse() assembly -------------------------------------------------------
        if name == "xxz_chain_L32_sz0":
            hsr = None
            extras["tri6x6"] = bench_tri6x6(ed, np, ctx, args.steps, peak)
            if world == 1:
                extras["sparse"] = bench_sparse(ed, np)

    if rank == 0:
        alg_bytes = 24.0 * n_local     # 8 B basis word + 8 B x + 8 B y per owned row
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if world == 1 and args.kernel == 0 and os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(name, {}).get("dram_bytes_per_launch")
        line = {
            "metric": "H*v matvecs/sec", "value": value, "unit": "matvec/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "gnnz_per_s": nnz_eff(n, n_bonds) * value / 1e9,
            "config": dict(base_config(name, w)),
            "details": dict(details, l2="inputs larger than L2 (x and y are %.2f GB each per GPU); no flush needed" % (n_local * 8 / 1e9)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "ncu --set full capture committed under profiles/ (bytes per launch)" if traffic else None,
                         "peak_source": peak_src, "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "bytes_definition": "SURVEY 8(d): (8 B basis word + 8 B x + 8 B y) per owned row",
                         "frac_of_vector_bytes_only": 16.0 * n_local / (kern_ms * 1e-3) / 1e9 / peak,
                         "vector_bytes_note": "the tiled kernel reads no basis word (combinatorial ranks): 16 B/row is what it must move",
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = e2e
        for k, v in extras.items():
            line[k] = v
        emit(line)
    ctx.close()


if __name__ == "__main__":
    main()
