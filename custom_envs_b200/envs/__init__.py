"""Environments (reference custom_envs/envs/)."""
from custom_envs_b200.envs.baseenvironment import BaseEnvironment, BaseMultiEnvironment
from custom_envs_b200.envs.multioptimize import MultiOptimize
from custom_envs_b200.envs.multioptlrs import MultiOptLRs

SINGLE_AGENT_ENVIRONMENTS = (MultiOptimize, MultiOptLRs)

__all__ = ['BaseEnvironment', 'BaseMultiEnvironment', 'MultiOptLRs', 'MultiOptimize', 'SINGLE_AGENT_ENVIRONMENTS']
