from custom_envs_b200.data import load_data  # noqa: F401
