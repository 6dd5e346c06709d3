from custom_envs_b200.utils.utils_logging import Monitor, create_env  # noqa: F401
