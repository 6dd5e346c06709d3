"""Third-party surfaces the reference boundary needs: real packages when importable,
in-repo stand-ins otherwise (``gym`` and ``stable_baselines`` cannot be installed in
the build image)."""
try:                                                     # pragma: no cover
    import gym                                           # noqa: F401
    from gym import spaces
    from gym.core import Env, Wrapper
    from gym.envs.registration import register
    from gym import make
    from gym.utils.seeding import np_random
    HAVE_GYM = not hasattr(gym, '__standin__')
except ImportError:
    from custom_envs_b200.compat import gym_standin as _standin
    gym = _standin.install_as_gym()
    spaces = gym.spaces
    Env, Wrapper = _standin.Env, _standin.Wrapper
    register, make = _standin.register, _standin.make
    np_random = _standin.np_random
    HAVE_GYM = False

try:                                                     # pragma: no cover
    from stable_baselines.common.vec_env import VecEnv, CloudpickleWrapper
    from stable_baselines.common.tile_images import tile_images
    HAVE_SB = True
except ImportError:
    from custom_envs_b200.compat.vec_env import VecEnv, CloudpickleWrapper, tile_images
    HAVE_SB = False

__all__ = ['gym', 'spaces', 'Env', 'Wrapper', 'register', 'make', 'np_random',
           'VecEnv', 'CloudpickleWrapper', 'tile_images', 'HAVE_GYM', 'HAVE_SB']
