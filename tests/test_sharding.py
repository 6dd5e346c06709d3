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


def _host_episode_table(steps, first_env):
    """What the reference's per-env Monitor would log (utils_logging.py:96-113): reward sum,
    length, last reward, episode counter per finished episode."""
    rows, total, episode = [], {}, {}
    for reward, done, info in steps:
        for e in range(len(reward)):
            total[e] = total.get(e, 0.0) + float(reward[e])
            if done[e]:
                rows.append(dict(env=first_env + e, r=round(total[e], 6), l=int(info[e, 15]),
                                 current_reward=float(reward[e]), episode=episode.get(e, 1),
                                 batch_loss=float(info[e, 1])))
                total[e] = 0.0
                episode[e] = episode.get(e, 1) + 1
    return rows


def _fake_steps(num_envs, count, seed):
    rng = np.random.RandomState(seed)
    length = np.zeros(num_envs)
    steps = []
    for _ in range(count):
        length += 1
        reward = rng.normal(size=num_envs).astype(np.float32)
        done = rng.uniform(size=num_envs) < 0.3
        info = np.zeros((num_envs, 16))
        info[:, 15], info[:, 1] = length, rng.uniform(size=num_envs)
        steps.append((reward, done.astype(np.uint8), info))
        length[done] = 0
    return steps


def test_device_episode_monitor_matches_per_env_monitors(tmp_path):
    import pandas as pd
    from custom_envs_b200.wrappers.device_monitor import DeviceEpisodeMonitor
    steps = _fake_steps(6, 25, seed=1)
    monitor = DeviceEpisodeMonitor(6, str(tmp_path / 'job'), info_keywords=('batch_loss',), first_env=12,
                                   device='cpu', capacity=64)
    split = DeviceEpisodeMonitor(6, str(tmp_path / 'monitor'), first_env=0, device='cpu', capacity=64,
                                 split_by_env=True)
    got = []
    for t, (reward, done, info) in enumerate(steps):
        args = (t, None, torch.from_numpy(reward), torch.from_numpy(done), torch.from_numpy(info))
        monitor.on_step(*args)
        split.on_step(*args)
        if t % 10 == 9:
            got += monitor.flush()                       # flushes in the middle keep the running sums
    got += monitor.close()
    split.close()
    want = _host_episode_table(steps, 12)
    assert len(got) == len(want) > 20
    key = lambda row: (row['env'], row['episode'])       # noqa: E731
    for a, b in zip(sorted(got, key=key), sorted(want, key=key)):
        assert all(a[name] == b[name] for name in ('env', 'l', 'episode', 'current_reward', 'batch_loss'))
        assert abs(a['r'] - b['r']) < 1e-5
    frame = pd.read_csv(tmp_path / 'job.mon.csv')
    assert list(frame.columns) == sorted(frame.columns) and len(frame) == len(want)       # utils_logging.py:64
    files = sorted(tmp_path.glob('monitor_*.mon.csv'))                                     # compile_exp.py:15-16
    assert len(files) == 6 and 'env' not in pd.read_csv(files[0]).columns
    assert sum(len(pd.read_csv(name)) for name in files) == len(want)
    small = DeviceEpisodeMonitor(6, None, device='cpu', capacity=2)
    for t, (reward, done, info) in enumerate(steps):
        small.on_step(t, None, torch.from_numpy(reward), torch.from_numpy(done), torch.from_numpy(info))
    with pytest.raises(RuntimeError):
        small.flush()                                    # dropped episodes are an error, not silence


def _monitor_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from custom_envs_b200.wrappers.device_monitor import DeviceEpisodeMonitor
    first, count = shard_range(7, world, rank)
    monitor = DeviceEpisodeMonitor(count, os.path.join(out_dir, 'job'), first_env=first, device='cpu', capacity=64)
    for t, (reward, done, info) in enumerate(_fake_steps(count, 12, seed=10 + rank)):
        monitor.on_step(t, None, torch.from_numpy(reward), torch.from_numpy(done), torch.from_numpy(info))
    rows = monitor.close()
    np.save(os.path.join(out_dir, 'rows%d.npy' % rank), np.array([[r['env'], r['l'], r['episode']] for r in rows]))
    dist.destroy_process_group()


def test_device_episode_monitor_gathers_shards_gloo_world2(tmp_path):
    import pandas as pd
    mp.spawn(_monitor_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    want = []
    for rank in range(2):
        first, count = shard_range(7, 2, rank)
        want += [[r['env'], r['l'], r['episode']] for r in _host_episode_table(_fake_steps(count, 12, seed=10 + rank), first)]
    for rank in range(2):                                # every rank sees the whole job's episodes
        assert np.load(tmp_path / ('rows%d.npy' % rank)).tolist() == want
    frame = pd.read_csv(tmp_path / 'job.mon.csv')       # written once, by rank 0
    assert frame[['env', 'l', 'episode']].values.tolist() == want


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """bench.py --impl reference launched like the driver launches it for N > 1: rank 0 alone times
    the oracle port and prints ONE JSON line with the contract's keys; the other rank exits 0."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), os.path.join(root, 'bench.py'),
           '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=280, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['n_gpus'] == 2 and line['unit'] == 'env-steps/s'
    assert line['value'] > 0 and line['higher_is_better'] is True and line['vs_baseline'] is None
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0,
                           'd2h_bytes_per_step': 0}
    assert line['config']['workload'].startswith('MultiOptLRs MLP 784-64-10')
