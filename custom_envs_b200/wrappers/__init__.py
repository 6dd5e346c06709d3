"""Wrappers (reference custom_envs/wrappers/) and their device-side counterparts."""
from custom_envs_b200.wrappers.optimizewrappers import (DeviceHistoryWrapper, DeviceSubSetWrapper,
                                                        HistoryWrapper, SubSetWrapper)
from custom_envs_b200.wrappers.monitor import Monitor

__all__ = ['HistoryWrapper', 'SubSetWrapper', 'DeviceHistoryWrapper', 'DeviceSubSetWrapper', 'Monitor']
