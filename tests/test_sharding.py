"""CPU tests of the multi-GPU plumbing with the gloo backend, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from custom_envs_b200.sharding import gather_env_stats, shard_range, shard_seeds


def test_shard_ranges_cover_every_env_once():
    for num_envs in (1, 7, 8, 4096, 4099):
        for world in (1, 2, 4, 8):
            owned = []
            for rank in range(world):
                first, count = shard_range(num_envs, world, rank)
                owned += list(range(first, first + count))
            assert owned == list(range(num_envs))
    assert shard_seeds(list(range(10, 20)), 4, 1) == [13, 14, 15]


def _free_port():
    with socket.socket() as sock:
        sock.bind(('127.0.0.1', 0))
        return sock.getsockname()[1]


def _worker(rank, world, port, num_envs, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    first, count = shard_range(num_envs, world, rank)
    # per-env "episode statistics": [global env index, episode return]
    local = torch.stack([torch.arange(first, first + count, dtype=torch.float32),
                         torch.arange(first, first + count, dtype=torch.float32) * 0.5], dim=1)
    full = gather_env_stats(local)
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank), full.numpy())
    dist.destroy_process_group()


def test_gather_env_stats_gloo_world2(tmp_path):
    num_envs, world = 7, 2                      # uneven shards: 4 + 3
    mp.spawn(_worker, args=(world, _free_port(), num_envs, str(tmp_path)), nprocs=world, join=True)
    want = np.stack([np.arange(num_envs), np.arange(num_envs) * 0.5], axis=1).astype(np.float32)
    for rank in range(world):
        assert np.array_equal(np.load(tmp_path / ('rank%d.npy' % rank)), want)


def test_gather_is_identity_without_process_group():
    stats = torch.ones(3, 2)
    assert gather_env_stats(stats) is stats


def test_numa_binding_reads_the_gpu_node_from_sysfs(tmp_path):
    """bind_to_gpu_numa against a fake sysfs tree: GPU on node 1 whose CPUs include the ones this
    process may run on; a platform that reports no node (-1, VMs) changes nothing."""
    from custom_envs_b200.sharding import _parse_cpulist, bind_to_gpu_numa, gpu_numa_node
    assert _parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    allowed = sorted(os.sched_getaffinity(0))
    keep = allowed[:max(1, len(allowed) // 2)]
    pci = tmp_path / 'bus/pci/devices/0000:1b:00.0'
    pci.mkdir(parents=True)
    (pci / 'numa_node').write_text('1\n')
    node = tmp_path / 'devices/system/node/node1'
    node.mkdir(parents=True)
    (node / 'cpulist').write_text(','.join(str(c) for c in keep) + ',4093-4095\n')
    assert gpu_numa_node('00000000:1B:00.0', str(tmp_path)) == 1          # CUDA's spelling of the id
    try:
        assert bind_to_gpu_numa('0000:1b:00.0', str(tmp_path)) == 1
        assert sorted(os.sched_getaffinity(0)) == keep
    finally:
        os.sched_setaffinity(0, allowed)
    (pci / 'numa_node').write_text('-1\n')
    assert bind_to_gpu_numa('0000:1b:00.0', str(tmp_path)) == -1
    assert bind_to_gpu_numa('0000:ff:00.0', str(tmp_path)) == -1          # unknown device
    assert sorted(os.sched_getaffinity(0)) == allowed
