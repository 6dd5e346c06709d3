from custom_envs_b200.wrappers.optimizewrappers import HistoryWrapper, SubSetWrapper  # noqa: F401
