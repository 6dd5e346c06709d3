"""Vectorised environments (reference custom_envs/vectorize/)."""
from custom_envs_b200.vectorize.concurrentvecenv import (ConcurrentVecEnv, SubprocVecEnv,
                                                         ThreadVecEnv)
from custom_envs_b200.vectorize.optvecenv import DeviceOptVecEnv, OptVecEnv

__all__ = ['ConcurrentVecEnv', 'SubprocVecEnv', 'ThreadVecEnv', 'OptVecEnv', 'DeviceOptVecEnv']
