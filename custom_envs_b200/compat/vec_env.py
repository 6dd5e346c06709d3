"""Stand-in for the stable-baselines 2 ``VecEnv`` surface the reference subclasses
(reference vectorize/concurrentvecenv.py:9-10,64-198).  Used when stable_baselines is
not importable (it is not installable in the build image)."""
from abc import ABC, abstractmethod

import numpy as np


class VecEnv(ABC):
    metadata = {'render.modes': ['human', 'rgb_array']}

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def step_async(self, actions):
        pass

    @abstractmethod
    def step_wait(self):
        pass

    @abstractmethod
    def close(self):
        pass

    @abstractmethod
    def get_attr(self, attr_name, indices=None):
        pass

    @abstractmethod
    def set_attr(self, attr_name, value, indices=None):
        pass

    @abstractmethod
    def env_method(self, method_name, *method_args, **method_kwargs):
        pass

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def get_images(self):
        raise NotImplementedError

    def render(self, *args, **kwargs):
        return None

    @property
    def unwrapped(self):
        return self


class CloudpickleWrapper:
    def __init__(self, var):
        self.var = var

    def __getstate__(self):
        import cloudpickle
        return cloudpickle.dumps(self.var)

    def __setstate__(self, obs):
        import pickle
        self.var = pickle.loads(obs)


def tile_images(img_nhwc):
    img_nhwc = np.asarray(img_nhwc)
    n_images, height, width, n_channels = img_nhwc.shape
    new_height = int(np.ceil(np.sqrt(n_images)))
    new_width = int(np.ceil(float(n_images) / new_height))
    img_nhwc = np.array(list(img_nhwc) + [img_nhwc[0] * 0
                                          for _ in range(n_images, new_height * new_width)])
    out = img_nhwc.reshape(new_height, new_width, height, width, n_channels)
    out = out.transpose(0, 2, 1, 3, 4)
    return out.reshape(new_height * height, new_width * width, n_channels)
