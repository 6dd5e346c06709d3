from custom_envs_b200.wrappers.monitor import Monitor  # noqa: F401
