"""Episode statistics sink for a fused env batch, accumulated on the device.

The reference wraps every env in a ``Monitor`` (utils/utils_logging.py:15-156) that sums the
rewards of an episode on the host and appends ``r, l, t, current_reward, episode`` plus chosen
``info`` keys to ``<path>.mon.csv`` (sorted columns, :57-68).  With thousands of envs per GPU
that bookkeeping stays in HBM here: per-env running sums are device tensors, finished episodes
are appended to a device table without a host synchronisation, and ``flush()`` brings the table
to the host, all-gathers it over the process group (NCCL on GPUs, gloo in the CPU tests) and lets
rank 0 append the rows in the reference's CSV format.  ``on_step`` has the signature of
``device_rollout``'s hook."""
import time

import numpy as np
import torch

from custom_envs_b200.sharding import gather_env_stats
from custom_envs_b200.vectorize.optvecenv import INFO_KEYS
from custom_envs_b200.wrappers.monitor import _EpisodeCsv

_FIXED = ('env', 'r', 'l', 'current_reward', 'episode', 't')


class DeviceEpisodeMonitor:
    def __init__(self, num_envs, file_path=None, info_keywords=(), first_env=0, capacity=None,
                 chunk_size=1, device='cuda:0', group=None, split_by_env=False):
        unknown = [key for key in info_keywords if key not in INFO_KEYS]
        if unknown:
            raise KeyError('info keys the env does not report: %s' % unknown)
        self.num_envs, self.first_env = int(num_envs), int(first_env)
        self.info_keywords = tuple(info_keywords)
        self._info_cols = [INFO_KEYS.index(key) for key in self.info_keywords]
        self.columns = _FIXED + self.info_keywords
        self.capacity = int(capacity or 4 * self.num_envs)
        self.device, self.group, self.split_by_env = torch.device(device), group, split_by_env
        self.file_path, self.chunk_size = file_path, chunk_size
        self._sinks = {}
        dev = self.device
        self._ret = torch.zeros(self.num_envs, dtype=torch.float64, device=dev)
        self._episode = torch.ones(self.num_envs, dtype=torch.float64, device=dev)   # Monitor counts from 1
        self._env_ids = torch.arange(self.first_env, self.first_env + self.num_envs, dtype=torch.float64, device=dev)
        # finished episodes; row `capacity` is a dump slot for envs that did not finish this step
        self._table = torch.zeros((self.capacity + 1, len(self.columns)), dtype=torch.float64, device=dev)
        self._count = torch.zeros((), dtype=torch.int64, device=dev)
        self._dropped = torch.zeros((), dtype=torch.int64, device=dev)
        self.t_start = time.time()
        self.rows = []                                   # host copy of everything flushed so far (rank 0)

    @torch.no_grad()
    def on_step(self, t, obs, reward, done, info):
        """Account one batched step: device tensors reward [E], done [E], info [E,16]."""
        finished = done.to(torch.bool)
        self._ret += reward.to(torch.float64)
        elapsed = time.time() - self.t_start
        rows = torch.stack([self._env_ids, self._ret, info[:, 15], reward.to(torch.float64), self._episode,
                            torch.full_like(self._ret, elapsed)]
                           + [info[:, col] for col in self._info_cols], dim=1)
        slot = self._count + torch.cumsum(finished, 0) - 1
        fits = finished & (slot < self.capacity)
        self._table.index_copy_(0, torch.where(fits, slot, torch.full_like(slot, self.capacity)), rows)
        self._count += fits.sum()
        self._dropped += (finished & ~fits).sum()
        self._ret.masked_fill_(finished, 0.0)
        self._episode += finished.to(torch.float64)

    def flush(self):
        """Device table -> host -> all ranks' rows on every rank; rank 0 appends them to the CSV.
        Returns this call's rows (all ranks, ordered by rank then by time) as a list of dicts."""
        count = int(self._count.item())
        if int(self._dropped.item()):
            raise RuntimeError('DeviceEpisodeMonitor: %d finished episodes did not fit the table of %d rows; '
                               'flush more often or raise capacity' % (int(self._dropped.item()), self.capacity))
        local = self._table[:count].clone()
        self._count.zero_()
        rank = 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank = torch.distributed.get_rank(self.group)
            local = gather_env_stats(local, self.group)
        table = local.cpu().numpy()
        rows = [self._row(line) for line in table]
        if rank == 0:
            self.rows += rows
            if self.file_path is not None:
                for row in rows:
                    key = row['env'] if self.split_by_env else None
                    if key not in self._sinks:
                        path = self.file_path if key is None else '%s_%d' % (self.file_path, key)
                        self._sinks[key] = _EpisodeCsv(path, self.chunk_size)
                    out = dict(row)
                    if self.split_by_env:
                        out.pop('env')                   # one file per env, the reference's layout
                    self._sinks[key].add(out)
                for sink in self._sinks.values():
                    sink.flush()
        return rows

    def _row(self, line):
        row = {name: float(value) for name, value in zip(self.columns, line)}
        for name in ('env', 'l', 'episode'):
            row[name] = int(row[name])
        row['r'], row['t'] = round(row['r'], 6), round(row['t'], 6)
        if 'loss' in row and np.isnan(row['loss']):
            row['loss'] = None
        return row

    def close(self):
        return self.flush()
