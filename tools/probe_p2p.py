"""Which peer-memory transports work between the ranks of one box?  (torchrun, >= 2 ranks)

  1. torch.distributed._symmetric_memory: empty() + rendezvous() -> peer buffer pointers
  2. CUDA IPC: cudaMalloc + cudaIpcGetMemHandle -> all_gather -> cudaIpcOpenMemHandle
Each is exercised with a write into the right neighbour's buffer and a read-back there.
"""
import ctypes
import os
import sys
import traceback

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
right = (rank + 1) % world
left = (rank - 1) % world


def report(name, ok, msg=""):
  flags = [None] * world
  dist.all_gather_object(flags, (bool(ok), msg))
  if rank == 0:
    print(f"{name}: {'OK' if all(f[0] for f in flags) else 'FAILED'} {[f[1] for f in flags if f[1]]}", flush=True)


try:
  import torch.distributed._symmetric_memory as symm
  t = symm.empty(1 << 20, dtype=torch.uint8, device=torch.device("cuda", local))
  hdl = symm.rendezvous(t, dist.group.WORLD)
  ptrs = [int(p) for p in hdl.buffer_ptrs]
  t.fill_(0)
  dist.barrier()
  peer = hdl.get_buffer(right, (1 << 20,), torch.uint8)
  peer.fill_(rank + 1)
  torch.cuda.synchronize()
  dist.barrier()
  ok = int(t[0].item()) == left + 1 and int(t[-1].item()) == left + 1
  report("symmetric_memory", ok, f"r{rank} ptrs={[hex(p) for p in ptrs]}")
except Exception as e:                                       # noqa: BLE001
  traceback.print_exc()
  report("symmetric_memory", False, repr(e)[:200])

try:
  rt = ctypes.CDLL("libcudart.so.12")
  buf = ctypes.c_void_p()
  assert rt.cudaMalloc(ctypes.byref(buf), ctypes.c_size_t(1 << 20)) == 0
  assert rt.cudaMemset(buf, 0, ctypes.c_size_t(1 << 20)) == 0
  handle = (ctypes.c_ubyte * 64)()
  rc = rt.cudaIpcGetMemHandle(handle, buf)
  assert rc == 0, f"cudaIpcGetMemHandle rc={rc}"
  mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
  allh = [torch.empty_like(mine) for _ in range(world)]
  dist.all_gather(allh, mine)
  hb = (ctypes.c_ubyte * 64)(*allh[right].cpu().tolist())
  peer = ctypes.c_void_p()

  class H(ctypes.Structure):
    _fields_ = [("b", ctypes.c_ubyte * 64)]
  rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), H, ctypes.c_uint]
  rc = rt.cudaIpcOpenMemHandle(ctypes.byref(peer), H(hb), 1)
  assert rc == 0, f"cudaIpcOpenMemHandle rc={rc}"
  assert rt.cudaMemset(peer, rank + 1, ctypes.c_size_t(1 << 20)) == 0
  torch.cuda.synchronize()
  dist.barrier()
  host = (ctypes.c_ubyte * 16)()
  assert rt.cudaMemcpy(host, buf, ctypes.c_size_t(16), 2) == 0
  report("cuda_ipc", host[0] == left + 1, f"r{rank} got {host[0]}")
except Exception as e:                                       # noqa: BLE001
  traceback.print_exc()
  report("cuda_ipc", False, repr(e)[:200])

dist.barrier()
dist.destroy_process_group()
