"""Host-side data set containers (reference custom_envs/dataset/)."""
from custom_envs_b200.dataset.inmemorydataset import BatchType, DataSet, DeviceDataSet, InMemoryDataSet

__all__ = ['BatchType', 'DataSet', 'DeviceDataSet', 'InMemoryDataSet']
