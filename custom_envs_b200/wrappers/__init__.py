"""Wrappers (reference custom_envs/wrappers/)."""
from custom_envs_b200.wrappers.optimizewrappers import HistoryWrapper, SubSetWrapper
from custom_envs_b200.wrappers.monitor import Monitor

__all__ = ['HistoryWrapper', 'SubSetWrapper', 'Monitor']
