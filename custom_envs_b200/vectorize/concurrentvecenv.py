"""Generic host-side vectoriser: one worker (thread or process) per env behind a pipe.

Same surface and semantics as the reference's ``ConcurrentVecEnv`` family
(vectorize/concurrentvecenv.py:27-61 worker loop, :64-198 VecEnv methods, :230-268
Thread/Subproc variants): auto-reset when ``any(done)`` with the RESET observation
returned, spaces taken from env 0, idempotent ``close`` that drains a pending step.
It is the fallback for arbitrary gym envs (and the reference's own StubEnv tests); the
device-backed optimise envs never go through it -- ``OptVecEnv`` fuses them into one
``BatchedOptEnv`` instead.
"""
import multiprocessing as mp
from collections import OrderedDict
from threading import Thread

import numpy as np

from custom_envs_b200.compat import CloudpickleWrapper, VecEnv, spaces, tile_images


def _any_done(done):
    return bool(np.any(done))


def _serve(pipe, env_factory):
    """Worker loop: answer commands until 'close' or a dropped pipe."""
    env = env_factory.var()

    def do_step(action):
        observation, reward, done, info = env.step(action)
        if _any_done(done):
            observation = env.reset()            # terminal observation is dropped
        return observation, reward, done, info

    handlers = {
        'step': do_step,
        'reset': lambda _: env.reset(),
        'render': lambda data: env.render(*data[0], **data[1]),
        'get_spaces': lambda _: (env.observation_space, env.action_space),
        'env_method': lambda data: getattr(env, data[0])(*data[1], **data[2]),
        'get_attr': lambda name: getattr(env, name),
        'set_attr': lambda data: setattr(env, data[0], data[1]),
    }
    try:
        while True:
            command, payload = pipe.recv()
            if command == 'close':
                pipe.close()
                break
            if command not in handlers:
                raise NotImplementedError(command)
            pipe.send(handlers[command](payload))
    except EOFError:
        pass
    finally:
        env.close()


def _stack_observations(observations, space):
    """Per-env observations -> batch with the env index first (dict / tuple / array)."""
    assert isinstance(observations, (list, tuple)) and len(observations) > 0
    if isinstance(space, spaces.Dict):
        return OrderedDict((key, np.stack([obs[key] for obs in observations]))
                           for key in space.spaces.keys())
    if isinstance(space, spaces.Tuple):
        return tuple(np.stack([obs[i] for obs in observations])
                     for i in range(len(space.spaces)))
    return np.stack(observations)


class ConcurrentVecEnv(VecEnv):
    """:param env_fns: callables building one gym env each
    :param create_method: Thread-like factory (target=, args=, daemon=)"""

    def __init__(self, env_fns, create_method):
        self.waiting = False
        self.closed = False
        pipes = [mp.Pipe(True) for _ in env_fns]
        self.remotes = tuple(p[0] for p in pipes)
        self.work_remotes = tuple(p[1] for p in pipes)
        self.processes = []
        for work_remote, env_fn in zip(self.work_remotes, env_fns):
            worker = create_method(target=_serve, args=(work_remote, CloudpickleWrapper(env_fn)),
                                   daemon=True)      # a crashed parent must not hang
            worker.start()
            self.processes.append(worker)
        observation_space, action_space = self._ask(self.remotes[:1], 'get_spaces', None)[0]
        VecEnv.__init__(self, len(env_fns), observation_space, action_space)

    def _ask(self, remotes, command, payload):
        for remote in remotes:
            remote.send((command, payload))
        return [remote.recv() for remote in remotes]

    def step_async(self, actions):
        for remote, action in zip(self.remotes, actions):
            remote.send(('step', action))
        self.waiting = True

    def step_wait(self):
        results = [remote.recv() for remote in self.remotes]
        self.waiting = False
        observations, rewards, dones, infos = zip(*results)
        return (_stack_observations(observations, self.observation_space), np.stack(rewards),
                np.stack(dones), infos)

    def reset(self):
        return _stack_observations(self._ask(self.remotes, 'reset', None), self.observation_space)

    def close(self):
        if self.closed:
            return
        if self.waiting:
            for remote in self.remotes:
                remote.recv()
        for remote in self.remotes:
            remote.send(('close', None))
        for worker in self.processes:
            worker.join()
        self.closed = True

    def get_images(self):
        return self._ask(self.remotes, 'render', ((), {'mode': 'rgb_array'}))

    def render(self, *args, **kwargs):
        mode = kwargs.get('mode', 'human')
        kwargs['mode'] = 'rgb_array'
        image = tile_images(self._ask(self.remotes, 'render', (args, kwargs)))
        if mode == 'rgb_array':
            return image
        if mode == 'human':
            import cv2
            cv2.imshow('vecenv', image[:, :, ::-1])
            cv2.waitKey(1)
            return None
        raise NotImplementedError

    def env_method(self, method_name, *method_args, **method_kwargs):
        return self._ask(self.remotes, 'env_method', (method_name, method_args, method_kwargs))

    def get_attr(self, attr_name, indices=None):
        return self._ask(self._pick(indices), 'get_attr', attr_name)

    def set_attr(self, attr_name, value, indices=None):
        return self._ask(self._pick(indices), 'set_attr', (attr_name, value))

    def _pick(self, indices):
        if indices is None:
            return self.remotes
        if isinstance(indices, int):
            indices = [indices]
        return [self.remotes[i] for i in indices]


class SubprocVecEnv(ConcurrentVecEnv):
    """One process per env ('forkserver' when available, else 'spawn')."""

    def __init__(self, env_fns, start_method=None):
        if start_method is None:
            start_method = 'forkserver' if 'forkserver' in mp.get_all_start_methods() else 'spawn'
        super().__init__(env_fns, mp.get_context(start_method).Process)


class ThreadVecEnv(ConcurrentVecEnv):
    """One thread per env."""

    def __init__(self, env_fns):
        super().__init__(env_fns, Thread)
