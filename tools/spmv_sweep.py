"""Cached-CSR SpMV of the 6x6 triangular k=0 A1 sector: column-block width x lanes-per-row sweep (one process)."""
import math, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "exactdiagonalization.jl_b200"))
import torch
import edcuda as ed

hs, h = ed.models.heisenberg_triangular(6)
hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
rhsr = ed.symmetry_reduce(hsr, ed.lattices.triangular_space_group_irrep(6, "A1"))
d = rhsr.dimension
ropr = ed.represent(rhsr, h)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(d, dtype=torch.complex128, device="cuda", generator=g) / math.sqrt(d)
y = torch.empty_like(x)
ref = None
args = [a for a in sys.argv[1:] if not a.startswith("--lanes=")]
lanes_arg = [tuple(int(v) for v in a.split("=")[1].split(",")) for a in sys.argv[1:] if a.startswith("--lanes=")]
for cols in args or ["0", "7100000", "5300000", "3540000"]:
    os.environ.pop("EDCUDA_CSR_NOBLOCK", None)
    os.environ.pop("EDCUDA_CSR_BLOCK_COLS", None)
    if cols == "0":
        os.environ["EDCUDA_CSR_NOBLOCK"] = "1"
    else:
        os.environ["EDCUDA_CSR_BLOCK_COLS"] = cols
    ropr.drop_cache()
    t0 = time.perf_counter()
    ropr.cache_matrix()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    for lanes in ((8,) if cols == "0" else (lanes_arg[0] if lanes_arg else (1, 2, 4, 8))):
        os.environ["EDCUDA_CSR_LANES"] = str(lanes)
        for _ in range(3):
            ed.mul_b(y, ropr, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ed.mul_b(y, ropr, x)
        e1.record()
        torch.cuda.synchronize()
        if ref is None:
            ref = y.clone()
        err = float((y - ref).abs().max() / ref.abs().max())
        print(f"block_cols {cols:>8} lanes {lanes}: {e0.elapsed_time(e1) / 10:.3f} ms/matvec  build {t_build:.2f}s  rel.diff vs first {err:.2e}", flush=True)
