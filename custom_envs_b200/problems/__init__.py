"""Problems the optimise envs work on (reference custom_envs/problems/)."""
from custom_envs_b200.problems.base_problem import BaseProblem, ProblemTuple
from custom_envs_b200.problems.device_problems import OptimizeFunction, OptimizeNN


def get_problem(name='func', **kwargs):
    """Reference problems/__init__.py:7-16."""
    if name == 'nn':
        return OptimizeNN.create(**kwargs)
    if name == 'func':
        return OptimizeFunction.create(**kwargs)
    raise RuntimeError('Not a name of a problem.')


__all__ = ['BaseProblem', 'ProblemTuple', 'OptimizeNN', 'OptimizeFunction', 'get_problem']
