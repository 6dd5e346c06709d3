from custom_envs_b200.vectorize.concurrentvecenv import ConcurrentVecEnv, SubprocVecEnv, ThreadVecEnv  # noqa: F401
