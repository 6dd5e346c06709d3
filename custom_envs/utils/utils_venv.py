from custom_envs_b200.vectorize import SubprocVecEnv, ThreadVecEnv  # noqa: F401
