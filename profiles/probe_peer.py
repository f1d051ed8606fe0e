"""2-GPU probe: does CUDA IPC (own C-ABI) and torch symmetric memory work between the ranks on this box?"""
import ctypes as C, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smnngp_b200 as sm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
lib = sm._lib.load()
try:
    ptr = C.c_void_p(); h = (C.c_ubyte * 64)()
    rc = lib.smnngp_peer_alloc(1 << 20, C.byref(ptr), h)
    print(rank, "peer_alloc rc", rc, hex(ptr.value or 0), flush=True)
    ht = torch.tensor(list(h), dtype=torch.uint8, device="cuda")
    allh = [torch.empty_like(ht) for _ in range(world)]
    dist.all_gather(allh, ht)
    peers = []
    for r in range(world):
        if r == rank:
            peers.append(ptr.value); continue
        q = C.c_void_p(); hh = (C.c_ubyte * 64)(*allh[r].cpu().tolist())
        rc = lib.smnngp_peer_open(hh, C.byref(q))
        print(rank, "peer_open", r, "rc", rc, hex(q.value or 0), sm._lib.load().smnngp_last_error(), flush=True)
        peers.append(q.value)
    # write my rank id into every peer's buffer at slot rank via cudaMemcpy
    class Arr:
        def __init__(self, p, n): self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(p, False), version=3)
    views = [torch.as_tensor(Arr(p, 16), device="cuda") for p in peers]
    for r in range(world):
        views[r][rank] = float(rank + 1)
    torch.cuda.synchronize(); dist.barrier()
    print(rank, "my buffer after peer writes:", views[rank][:world].tolist(), flush=True)
except Exception as e:
    print(rank, "IPC EXC", repr(e), flush=True)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1024, dtype=torch.float64, device="cuda")
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "symm_mem ok", [hex(p) for p in hdl.buffer_ptrs], flush=True)
except Exception as e:
    print(rank, "SYMM EXC", repr(e), flush=True)
dist.barrier()
dist.destroy_process_group()
