"""custom_envs_b200: B200-native batched step of custom_envs' optimisation environments.

Importing the package registers the reference's gym ids (custom_envs/__init__.py:12-40)
that are built on the device path."""
from custom_envs_b200.compat import register

register(id='MultiOptLRs-v0', entry_point='custom_envs_b200.envs.multioptlrs:MultiOptLRs')
register(id='MultiOptimize-v0', entry_point='custom_envs_b200.envs.multioptimize:MultiOptimize')

__version__ = '0.1.0'
