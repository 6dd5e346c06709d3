from custom_envs_b200.data.load_data import load_data

__all__ = ['load_data']
