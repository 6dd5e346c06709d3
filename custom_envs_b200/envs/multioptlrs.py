"""MultiOptLRs: one agent per parameter chooses that parameter's learning rate
(reference envs/multioptlrs.py:19-138); the step itself is the fused CUDA kernel."""
from custom_envs_b200.envs.device_env import DeviceEnvFront, VersionType
from custom_envs_b200.problems import get_problem
from custom_envs_b200.utils import utils_env

BOUNDS = 1e2


class MultiOptLRs(DeviceEnvFront):
    """``MultiOptLRs(problem='func', max_batches=400, max_history=5)`` as in the reference;
    ``problem_kwargs`` (e.g. ``dict(layers=(64,), data_set=...)``) is forwarded to
    ``get_problem`` and ``device`` picks the GPU."""

    def __init__(self, problem='func', max_batches=400, max_history=5, problem_kwargs=None,
                 device='cuda:0'):
        super().__init__()
        model = problem if hasattr(problem, 'spec') else get_problem(problem, **(problem_kwargs or {}))
        model.device = device
        obs_space, _ = utils_env.get_obs_version((model.size,), max_history, 3)
        act_space = utils_env.get_action_space_optlrs(2)
        self._setup(model, obs_space, act_space, max_batches, max_history,
                    VersionType(3, 3, 0, 6), device)

    def __repr__(self):
        return '<MultiOptLRs({})>'.format(self.version)

    def _terminal(self):
        return self.current_step >= self.max_batches
