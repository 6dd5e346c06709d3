"""The problem interface (reference problems/base_problem.py:8-73)."""
from collections import namedtuple

ProblemTuple = namedtuple('ProblemTuple', ['gradient', 'loss', 'parameters'])


class BaseProblem:
    @classmethod
    def create(cls, *args, **kwargs):
        return cls(*args, **kwargs)

    def reset(self):
        raise NotImplementedError

    def get_gradient(self):
        raise NotImplementedError

    def get_loss(self):
        raise NotImplementedError

    def get_parameters(self):
        raise NotImplementedError

    def set_parameters(self, parameters):
        raise NotImplementedError

    def next(self):
        """For problems that need to advance (next minibatch)."""

    def get(self):
        return ProblemTuple(self.get_gradient(), self.get_loss(), self.get_parameters())

    @property
    def size(self):
        return len(self.get_parameters())

    @property
    def parameters(self):
        return self.get_parameters()
