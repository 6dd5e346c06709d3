"""The lenient ``Monitor`` the training scripts use and ``create_env`` (reference
utils/utils_logging.py:15-175): accepts an env factory, never refuses a reset, calls
``callbacks`` with every step's result, logs ``r, l, t, current_reward, episode`` plus
``info_keywords`` per episode."""
import time
from collections import defaultdict

from custom_envs_b200.compat import Wrapper, make
from custom_envs_b200.wrappers.monitor import _EpisodeCsv


class Monitor(Wrapper):
    EXT = _EpisodeCsv.EXT

    def __init__(self, env, file_path, info_keywords=(), chunk_size=1, callbacks=None):
        if callable(env):
            env = env()
        Wrapper.__init__(self, env=env)
        self.t_start = time.time()
        self._sink = _EpisodeCsv(file_path, chunk_size)
        self.file_path = self._sink.file_path
        self.chunk_size = chunk_size
        self.info_keywords = info_keywords
        self.last_info = {}
        self.rewards = None
        self.metric_history = defaultdict(list)
        self.current_episode = 0
        self.callbacks = [] if callbacks is None else callbacks

    @property
    def data(self):
        return self._sink.rows

    def save(self):
        self._sink.flush()

    def reset(self, **kwargs):
        self.rewards = []
        self.current_episode += 1
        return self.env.reset(**kwargs)

    def step(self, action):
        observation, reward, done, info = self.env.step(action)
        for callback in self.callbacks:
            callback({'observation': observation, 'reward': reward, 'done': done, 'info': info,
                      'episode': self.current_episode})
        self.rewards.append(reward)
        if done:
            self._finish_episode(sum(self.rewards), len(self.rewards), reward, info)
        return observation, reward, done, info

    def _finish_episode(self, total, length, reward, info):
        """The episode-end branch of ``step`` (reference utils_logging.py:104-113).  The fused
        ``OptVecEnv`` calls it directly with the episode's reward sum and length, which it keeps
        vectorised over the envs."""
        elapsed = time.time() - self.t_start
        episode = {'r': round(total, 6), 'l': length, 't': round(elapsed, 6),
                   'current_reward': reward, 'episode': self.current_episode}
        self.last_info = info
        episode.update({key: info[key] for key in self.info_keywords})
        self._sink.add(episode)
        info['episode'] = episode
        self.metric_history['rewards'].append(total)
        self.metric_history['lengths'].append(length)
        self.metric_history['times'].append(elapsed)

    def close(self):
        self._sink.flush()
        super().close()

    def get_episode_rewards(self):
        return self.metric_history.get('rewards', [])

    def get_episode_lengths(self):
        return self.metric_history.get('lengths', [])

    def get_episode_times(self):
        return self.metric_history.get('times', [])


def create_env(env_name, log_dir=None, num_of_envs=1, **kwarg):
    """Build ``num_of_envs`` registered envs, optionally each behind a Monitor."""
    from pathlib import Path
    envs = [make(env_name, **kwarg) for _ in range(num_of_envs)]
    if log_dir is not None:
        envs = [Monitor(env, str(Path(log_dir) / str(i)), chunk_size=10) for i, env in enumerate(envs)]
    return envs
