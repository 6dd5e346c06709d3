"""Host<->device copy probe for the e2e path (run alone or under torchrun): D2H/H2D rate of a
pinned buffer per rank, all ranks at once, with and without NUMA binding; host memcpy rate
with 1..8 threads.  Prints one line per measurement."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custom_envs_b200.sharding import bind_to_gpu_numa, gpu_numa_node   # noqa: E402

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
from custom_envs_b200.sharding import device_pci_bus_id   # noqa: E402
bus = device_pci_bus_id(local)
print('rank', rank, 'bus', bus, 'numa', gpu_numa_node(bus), 'cpus', len(os.sched_getaffinity(0)),
      'OMP', os.environ.get('OMP_NUM_THREADS'), 'torch threads', torch.get_num_threads(), flush=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rate(tag, nbytes=4 << 30):
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    devbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for direction in ('d2h', 'h2d'):
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            if direction == 'd2h':
                host.copy_(devbuf, non_blocking=True)
            else:
                devbuf.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print('rank %d %s %s %.1f GB/s' % (rank, tag, direction, 3 * nbytes / dt / 1e9), flush=True)
    # host memcpy pageable -> pinned with n threads
    src = np.ones(1 << 30, np.uint8)
    dst = host.numpy()[:1 << 30]
    for threads in (1, 4, 8):
        pool = ThreadPoolExecutor(threads)
        cuts = np.linspace(0, src.size, threads + 1).astype(np.int64)
        barrier()
        t0 = time.perf_counter()
        list(pool.map(lambda i: np.copyto(dst[cuts[i]:cuts[i + 1]], src[cuts[i]:cuts[i + 1]]),
                      range(threads)))
        dt = time.perf_counter() - t0
        print('rank %d %s memcpy %d threads %.1f GB/s' % (rank, tag, threads, src.size / dt / 1e9),
              flush=True)
        pool.shutdown()
    del host, devbuf


rate('unbound')
node = bind_to_gpu_numa(bus)
print('rank', rank, 'bound to node', node, 'cpus', len(os.sched_getaffinity(0)), flush=True)
if node >= 0:
    rate('bound')
if world > 1:
    dist.destroy_process_group()
