"""Environments (reference custom_envs/envs/)."""
from custom_envs_b200.envs.baseenvironment import BaseEnvironment, BaseMultiEnvironment
from custom_envs_b200.envs.multioptlrs import MultiOptLRs

SINGLE_AGENT_ENVIRONMENTS = (MultiOptLRs,)

__all__ = ['BaseEnvironment', 'BaseMultiEnvironment', 'MultiOptLRs', 'SINGLE_AGENT_ENVIRONMENTS']
