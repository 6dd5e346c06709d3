from custom_envs_b200.envs.baseenvironment import BaseEnvironment, BaseMultiEnvironment  # noqa: F401
