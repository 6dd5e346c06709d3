"""Drop-in alias: ``import custom_envs`` resolves the reference's module paths to the
B200 implementation in ``custom_envs_b200`` (reference custom_envs/__init__.py)."""
import custom_envs_b200  # noqa: F401  (registers the gym ids)
from custom_envs_b200.data import load_data  # noqa: F401
