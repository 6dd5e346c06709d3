"""What an optimisation env asks of a problem (reference problems/base_problem.py:8-73): flat
parameter / gradient vectors of length ``size``, a scalar loss, ``get()`` for the three at once,
``next()`` to move to the following minibatch, ``reset()`` for a new episode."""
import collections

ProblemTuple = collections.namedtuple('ProblemTuple', 'gradient loss parameters')


def _abstract(name):
    def method(self, *args, **kwargs):
        raise NotImplementedError('%s.%s' % (type(self).__name__, name))
    method.__name__ = name
    return method


class BaseProblem:
    reset = _abstract('reset')
    get_gradient = _abstract('get_gradient')
    get_loss = _abstract('get_loss')
    get_parameters = _abstract('get_parameters')
    set_parameters = _abstract('set_parameters')

    @classmethod
    def create(cls, *args, **kwargs):
        return cls(*args, **kwargs)

    def next(self):
        """Advance to the next minibatch; problems without data ignore it."""

    def get(self):
        return ProblemTuple(gradient=self.get_gradient(), loss=self.get_loss(), parameters=self.get_parameters())

    parameters = property(lambda self: self.get_parameters())
    size = property(lambda self: len(self.get_parameters()))
