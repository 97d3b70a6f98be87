"""bore_allreduce_maxloc on N GPUs of one box without torch.distributed: one process, one raw NCCL
communicator per device (ncclCommInitAll), the calls grouped.  Each device holds the packed key of
its shard; afterwards every device holds the maximum = the global first-minimum winner.
usage (gpurun --gpus 2): python tools/nccl_maxloc_2gpu.py [N]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bore_b200 import _lib
from bore_b200.engine import NativeMLP
lib = _lib.require_cuda()
N = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
nccl = C.CDLL("libnccl.so.2")
comms = (C.c_void_p * N)()
assert nccl.ncclCommInitAll(comms, N, (C.c_int * N)(*range(N))) == 0
rs = np.random.RandomState(0)
S = 1000
fun_all = rs.normal(size=(N, S))
status_all = rs.choice([0, 0, 0, 1, 2], size=(N, S)).astype(np.int32)
keys = []
for r in range(N):
    torch.cuda.set_device(r)
    net = NativeMLP([2, 4, 1], ["relu", "sigmoid"], device=r)
    keys.append(net.select_best(torch.from_numpy(fun_all[r]).cuda(r), torch.from_numpy(status_all[r]).cuda(r), idx_offset=r * S))
assert nccl.ncclGroupStart() == 0
for r in range(N):
    torch.cuda.set_device(r)
    _lib.check(lib.bore_allreduce_maxloc(C.c_void_p(comms[r]), C.c_void_p(keys[r].data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream(r).cuda_stream)))
assert nccl.ncclGroupEnd() == 0
for r in range(N):
    torch.cuda.synchronize(r)
got = [0x7fffffff - (int(k.item()) & 0x7fffffff) for k in keys]
ok = (status_all != 2)
f = np.where(ok, fun_all.astype(np.float32), np.inf).reshape(-1)
want = int(np.argmin(f))
print("winner on every device:", got, "expected", want)
assert all(g == want for g in got)
for r in range(N):
    nccl.ncclCommDestroy(C.c_void_p(comms[r]))
print("ok")
