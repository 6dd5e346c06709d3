from custom_envs_b200.vectorize.optvecenv import OptVecEnv, OptEnvRunner, flatten_dictionary  # noqa: F401
