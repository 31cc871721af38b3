"""Peer-copy bandwidth between the per-GPU processes of one node: every rank pulls `mb` MB in total from `fanin` peers
at once (cudaMemcpyAsync from IPC-mapped peer buffers, one stream per peer).  Prints GB/s per rank (min over ranks)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "exactdiagonalization.jl_b200"))
import numpy as np, torch, torch.distributed as dist
from edcuda._lib import lib, check
from edcuda.lanczos import DeviceBuffer
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); check(lib.ed_set_device(lr))
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
for fanin in sorted(set([1, min(2, world - 1), min(4, world - 1), world - 1])):
    n = mb * (1 << 20) // 8 // fanin
    src = DeviceBuffer(n * fanin, np.float64); src.tensor().fill_(rank)
    dst = torch.empty(n * fanin, dtype=torch.float64, device="cuda")
    hs = [None] * world
    dist.all_gather_object(hs, src.ipc_handle())
    peers = [(rank + 1 + i) % world for i in range(fanin)]
    views = []
    for p in peers:
        ptr = C.c_void_p(); check(lib.ed_ipc_open_handle((C.c_uint8 * 64).from_buffer_copy(hs[p]), C.byref(ptr)))
        class V: pass
        v = V(); v.__cuda_array_interface__ = {"shape": (n * fanin,), "typestr": "<f8", "data": (ptr.value, False), "version": 2, "strides": None}
        views.append((torch.as_tensor(v, device="cuda"), ptr))
    streams = [torch.cuda.Stream() for _ in peers]
    def run():
        for i, ((pv, _), st) in enumerate(zip(views, streams)):
            with torch.cuda.stream(st):
                dst[i * n:(i + 1) * n].copy_(pv[i * n:(i + 1) * n], non_blocking=True)
        for st in streams: st.synchronize()
    for _ in range(2): run()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): run()
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / 5
    t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = bool((dst[:n] == peers[0]).all())
    if rank == 0: print(f"fanin {fanin}: {n * fanin * 8 / float(t[0]) / 1e9:.0f} GB/s per rank inbound ({mb} MB, ok={ok})", flush=True)
    for _, ptr in views: lib.ed_ipc_close_handle(ptr)
    del src
dist.destroy_process_group()
