"""Latency of the hot path's only collective (all-gather of one int64 per rank) under torchrun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from flacarray_b200.mpi import TorchComm
comm = TorchComm()
for it in range(25):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = comm.allgather_int64(1000 + rank)
    t1 = time.perf_counter()
    if rank == 0 and (it < 3 or it % 8 == 0): print(f"it{it}: allgather_int64 {1e3*(t1-t0):.3f} ms -> {r}", flush=True)
dist.barrier(); dist.destroy_process_group()
