"""RLTrainerMixin: the RL hyper-parameter views every trainer exposes
(reagent/training/rl_trainer_pytorch.py:14-72): `gamma`, `tau`, `rl_temperature` read straight
from `rl_parameters`; `multi_steps`, `maxq_learning` and `use_seq_num_diff_as_time_diff` can be
overridden per trainer instance and otherwise fall back to `rl_parameters`."""
from ..core.parameters import RLParameters


class _Overridable:
    """Attribute that reads `rl_parameters.<name>` until it is assigned on the instance; the
    assigned value is kept in `_<name>` (the attribute name the reference uses)."""

    def __set_name__(self, owner, name):
        self.name = name
        self.slot = "_" + name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        own = obj.__dict__.get(self.slot, getattr(type(obj), self.slot, None))
        return getattr(obj.rl_parameters, self.name) if own is None else own

    def __set__(self, obj, value):
        # nn.Module.__setattr__ is bypassed on purpose: these are plain Python values
        obj.__dict__[self.slot] = value


class _FromRl:
    """Read-only view of one `rl_parameters` field."""

    def __init__(self, field):
        self.field = field

    def __get__(self, obj, objtype=None):
        return self if obj is None else getattr(obj.rl_parameters, self.field)


class RLTrainerMixin:
    # score given to an action that is not possible: worse than any legitimate one
    ACTION_NOT_POSSIBLE_VAL = -1e9

    rl_parameters: RLParameters
    _multi_steps = None
    _maxq_learning = None
    _use_seq_num_diff_as_time_diff = None

    gamma = _FromRl("gamma")
    tau = _FromRl("target_update_rate")
    rl_temperature = _FromRl("temperature")
    multi_steps = _Overridable()
    maxq_learning = _Overridable()
    use_seq_num_diff_as_time_diff = _Overridable()
