import os
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 1 << 22
t = symm.empty(n, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; multicast", hdl.multicast_ptr != 0, "ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
t.fill_(float(rank))
hdl.barrier(0)
peer = (rank + 1) % world
pb = hdl.get_buffer(peer, (n,), torch.float32, 0)
src = torch.full((1024,), 100.0 + rank, device=dev)
pb[:1024].copy_(src)          # P2P store into the peer
hdl.barrier(1)
torch.cuda.synchronize()
print(rank, "local head after peer write:", t[:2].tolist(), "tail", t[2048:2050].tolist(), flush=True)
# barrier latency
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100):
    hdl.barrier(i % 2)
e1.record(); torch.cuda.synchronize()
print(rank, "symm barrier us:", e0.elapsed_time(e1) * 10, flush=True)
# P2P copy bandwidth (16 MB to next peer)
big = torch.empty(n, dtype=torch.float32, device=dev)
for _ in range(3): pb.copy_(big)
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(10): pb.copy_(big)
e1.record(); torch.cuda.synchronize()
print(rank, "p2p copy GB/s:", 10 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9, flush=True)
# NCCL small all_reduce latency for comparison
x = torch.zeros(8, device=dev)
for _ in range(5): dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(100): dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
print(rank, "nccl tiny all_reduce us:", e0.elapsed_time(e1) * 10, flush=True)
dist.barrier(); dist.destroy_process_group()
