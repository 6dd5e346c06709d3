"""Episode monitors: per-episode reward / length / time (+ chosen info keys) appended to
``<path>.mon.csv`` with sorted columns, one row per episode.

Two flavours exist in the reference and both are kept: the strict one
(wrappers/monitor.py:11-163: resets only after ``done`` unless ``allow_early_resets``,
stepping a finished env raises) and the lenient callback one the training scripts use
(utils/utils_logging.py:15-156)."""
import time
from pathlib import Path

import pandas as pd

from custom_envs_b200.compat import Wrapper


class _EpisodeCsv:
    """Buffered CSV sink: header on first write, appended afterwards."""
    EXT = '.mon.csv'

    def __init__(self, file_path, chunk_size):
        self.file_path = None if file_path is None else Path(file_path).resolve().with_suffix(self.EXT)
        self.chunk_size = chunk_size
        self.rows = []

    def add(self, row):
        self.rows.append(row)
        if len(self.rows) >= self.chunk_size:
            self.flush()

    def flush(self):
        if self.file_path is not None and self.rows:
            first = not self.file_path.is_file()
            frame = pd.DataFrame(self.rows)
            frame = frame.reindex(sorted(frame.columns), axis=1)
            frame.to_csv(self.file_path, header=first, index=False, mode='w' if first else 'a')
        self.rows = []


class Monitor(Wrapper):
    """Strict monitor (reference wrappers/monitor.py)."""
    EXT = _EpisodeCsv.EXT

    def __init__(self, env, file_path, allow_early_resets=False, reset_keywords=(),
                 info_keywords=(), chunk_size=1):
        Wrapper.__init__(self, env=env)
        self.t_start = time.time()
        self._sink = _EpisodeCsv(file_path, chunk_size)
        self.file_path = self._sink.file_path
        self.chunk_size = chunk_size
        self.reset_keywords = reset_keywords
        self.info_keywords = info_keywords
        self.allow_early_resets = allow_early_resets
        self.rewards = None
        self.needs_reset = True
        self.episode_rewards, self.episode_lengths, self.episode_times = [], [], []
        self.total_steps = 0
        self.current_reset_info = {}

    @property
    def data(self):
        return self._sink.rows

    def save(self):
        self._sink.flush()

    def reset(self, **kwargs):
        if not self.allow_early_resets and not self.needs_reset:
            raise RuntimeError('Tried to reset an environment before done. If you want to allow '
                               'early resets, wrap your env with Monitor(env, path, '
                               'allow_early_resets=True)')
        self.rewards = []
        self.needs_reset = False
        for key in self.reset_keywords:
            if kwargs.get(key) is None:
                raise ValueError('Expected you to pass kwarg %s into reset' % key)
            self.current_reset_info[key] = kwargs[key]
        return self.env.reset(**kwargs)

    def step(self, action):
        if self.needs_reset:
            raise RuntimeError('Tried to step environment that needs reset')
        observation, reward, done, info = self.env.step(action)
        self.rewards.append(reward)
        if done:
            self.needs_reset = True
            elapsed = time.time() - self.t_start
            episode = {'r': round(sum(self.rewards), 6), 'l': len(self.rewards), 't': round(elapsed, 6)}
            episode.update({key: info[key] for key in self.info_keywords})
            self.episode_rewards.append(sum(self.rewards))
            self.episode_lengths.append(len(self.rewards))
            self.episode_times.append(elapsed)
            episode.update(self.current_reset_info)
            self._sink.add(episode)
            info['episode'] = episode
        self.total_steps += 1
        return observation, reward, done, info

    def close(self):
        self._sink.flush()
        super().close()

    def get_total_steps(self):
        return self.total_steps

    def get_episode_rewards(self):
        return self.episode_rewards

    def get_episode_lengths(self):
        return self.episode_lengths

    def get_episode_times(self):
        return self.episode_times
