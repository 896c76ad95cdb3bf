"""torchrun -n W tools/peer_allreduce_check.py : team_peer_allreduce_f32 against the NCCL all-reduce (values and time)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from team_b200 import capi, parallel    # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
capi.require_device()
import itertools
for n, use_mc in itertools.product((4, 1000, 1848832, 1 << 24), (False, True)):
    ar = parallel.PeerAllReduce(n, dev, multicast=use_mc)
    if use_mc and not ar.multicast_ptr:
        if rank == 0:
            print("no multicast mapping on this box"); 
        continue
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    x = torch.randn(n, generator=g).to(dev)
    ref = x.clone()
    dist.all_reduce(ref)
    # exact expectation: rank-ordered sum
    parts = [torch.randn(n, generator=torch.Generator(device="cpu").manual_seed(1234 + r)) for r in range(world)]
    want = parts[0].clone()
    for r in range(1, world):
        want += parts[r]
    for it in range(3):
        ar.buffer.copy_(x)
        ar()
        torch.cuda.synchronize()
        if not use_mc:
            assert torch.equal(ar.buffer.cpu(), want), (n, it, float((ar.buffer.cpu() - want).abs().max()))
    assert torch.allclose(ar.buffer, ref, rtol=1e-5, atol=1e-5)
    # timing: graph of 20 all-reduces (buffer keeps growing; values irrelevant)
    ar.buffer.zero_()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        ar(); st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(20):
                ar()
        gr.replay(); st.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); gr.replay(); e1.record(st); st.synchronize()
        t_peer = e0.elapsed_time(e1) / 20 * 1e3
        y = torch.zeros(n, device=dev)
        for _ in range(3):
            dist.all_reduce(y)
        torch.cuda.synchronize(); dist.barrier()
        e0.record(st)
        for _ in range(20):
            dist.all_reduce(y)
        e1.record(st); st.synchronize()
        t_nccl = e0.elapsed_time(e1) / 20 * 1e3
    if rank == 0:
        print(f"n={n:9d} ({n * 4 / 2**20:7.2f} MiB) world={world} multicast={bool(ar.multicast_ptr)}: peer {t_peer:7.1f} us   nccl {t_nccl:7.1f} us   exact rank-ordered sum OK", flush=True)
    del ar
dist.destroy_process_group()
