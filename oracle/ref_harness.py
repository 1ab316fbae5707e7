"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Harness that imports the UNMODIFIED reference (facebookresearch/ReAgent) from
/root/reference in the build container and drives its hot path, so that

  * oracle/make_golden.py can dump golden vectors into tests/golden/, and
  * the restatement in oracle/*.py can be pinned against the reference itself.

/root/reference does not exist on the GPU box: nothing imported by `-m gpu`
tests, smoke() or bench.py may import this module at run time there (they use
the committed fixtures and the restatement instead).

Two glue stubs are needed:
  * `torchrec` -- imported at reagent/core/types.py:22-23, never used on this path;
  * `pytorch_lightning` -- base class only (reagent/training/reagent_lightning_module.py:8,18);
and `reagent.training` is pre-registered as a bare package so that
reagent/training/__init__.py:5-24 (imports every trainer) is skipped; likewise `reagent.gym`
and `reagent.gym.policies` (their __init__ import the gym environments; `gym` is absent).

All arithmetic that runs is reference + torch code.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("REAGENT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "reagent"))


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    """Install torchrec / pytorch_lightning stubs and put the reference on sys.path."""
    import torch

    if "reagent" in sys.modules and getattr(sys.modules["reagent"], "_rb200_harness", False):
        return
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # ---- torchrec (unused on this path) ----
    class _Dummy:
        def __init__(self, *a, **k):
            pass

    class PoolingType:
        MEAN = "MEAN"
        SUM = "SUM"
        NONE = "NONE"

    _stub_module("torchrec", PoolingType=PoolingType, EmbeddingBagConfig=_Dummy,
                 EmbeddingBagCollection=_Dummy)
    _stub_module("torchrec.sparse")
    _stub_module("torchrec.sparse.jagged_tensor", KeyedJaggedTensor=_Dummy, JaggedTensor=_Dummy)
    _stub_module("torchrec.modules")
    _stub_module("torchrec.modules.embedding_configs", EmbeddingBagConfig=_Dummy, PoolingType=PoolingType)
    _stub_module("torchrec.modules.embedding_modules", EmbeddingBagCollection=_Dummy)
    _stub_module("torchrec.models")
    _stub_module("torchrec.models.dlrm", SparseArch=_Dummy, InteractionArch=_Dummy)

    # ---- pytorch_lightning (base class only) ----
    class DummyExperiment:
        def nop(self, *a, **k):
            return None

        def __getattr__(self, _):
            return self.nop

        def __getitem__(self, idx):
            return self

    class _Logger:
        def __init__(self):
            self.metrics = []

        def log_metrics(self, metrics, step=None):
            self.metrics.append((step, metrics))

    class _TrainerStub:
        log_every_n_steps = 10 ** 9

    class LightningModule(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.logger = _Logger()
            self.trainer = _TrainerStub()
            self._logged = {}

        def log(self, name, value, **kwargs):
            self._logged[name] = value

        @property
        def on_gpu(self):
            return False

        def optimizers(self, use_pl_optimizer=True):
            return [o["optimizer"] for o in self.configure_optimizers()]

    class Callback:
        pass

    pl = _stub_module("pytorch_lightning", LightningModule=LightningModule, Callback=Callback)
    loggers = _stub_module("pytorch_lightning.loggers")
    base = _stub_module("pytorch_lightning.loggers.base", DummyExperiment=DummyExperiment,
                        LoggerCollection=type("LoggerCollection", (), {}))
    tb = _stub_module("pytorch_lightning.loggers.tensorboard",
                      TensorBoardLogger=type("TensorBoardLogger", (), {}))
    loggers.base = base
    loggers.tensorboard = tb
    pl.loggers = loggers
    _stub_module("pytorch_lightning.utilities")
    _stub_module("pytorch_lightning.utilities.distributed", sync_ddp_if_available=lambda x, *a, **k: x)

    import reagent  # noqa: F401  (reference package root)

    sys.modules["reagent"]._rb200_harness = True
    # Skip reagent/training/__init__.py (it imports every trainer, some need pl.loops).
    pkg = types.ModuleType("reagent.training")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "reagent", "training")]
    sys.modules["reagent.training"] = pkg
    # `gym` itself: only named in type annotations / isinstance checks of code not run here
    if "gym" not in sys.modules:
        spaces = _stub_module("gym.spaces", Discrete=type("Discrete", (), {}),
                              Box=type("Box", (), {}), MultiDiscrete=type("MultiDiscrete", (), {}))
        _stub_module("gym", Env=type("Env", (), {}), spaces=spaces)
    # Same for reagent/gym/__init__.py (imports the gym environments; `gym` is not installed):
    # only the act-time samplers / scorers below it are used.
    for sub in ("gym", "gym/policies", "gym/preprocessors", "gym/policies/scorers",
                "gym/policies/samplers"):
        name = "reagent." + sub.replace("/", ".")
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "reagent", *sub.split("/"))]
        sys.modules[name] = pkg


def ref(name):
    """import a reference module, e.g. ref('reagent.training.dqn_trainer')."""
    install_stubs()
    return importlib.import_module(name)


# ---------------------------------------------------------------------------
# Lightning 1.6 automatic-optimisation loop with several optimizers, emulated
# (SURVEY.md 8c "Oracle driver loop").  Assumption on record:
# a `None` yield skips zero_grad/backward/step for that optimizer
# (intent stated at reagent/training/td3_trainer.py:197).
# ---------------------------------------------------------------------------
def _toggle(opts, cur):
    saved = {}
    for o in opts:
        for g in o.param_groups:
            for p in g["params"]:
                if p not in saved:
                    saved[p] = p.requires_grad
                p.requires_grad = False
    for g in cur.param_groups:
        for p in g["params"]:
            p.requires_grad = saved[p]
    return saved


def _untoggle(saved):
    for p, rg in saved.items():
        p.requires_grad = rg


def run_update(trainer, batch, batch_idx, opts=None, capture=None):
    """One full update = exhaust train_step_gen under the toggle loop.
    Returns the list of yielded losses (detached floats / None).
    `capture`, if a dict, receives per-optimizer grads: capture[i] = [grad clones]."""
    if opts is None:
        opts = [o["optimizer"] for o in trainer.configure_optimizers()]
    gen = trainer.train_step_gen(batch, batch_idx)
    losses = []
    for i, opt in enumerate(opts):
        saved = _toggle(opts, opt)
        try:
            loss = next(gen)
            if loss is not None:
                opt.zero_grad()
                loss.backward()
                if capture is not None:
                    capture[i] = [
                        None if p.grad is None else p.grad.detach().clone()
                        for g in opt.param_groups for p in g["params"]
                    ]
                opt.step()
                losses.append(float(loss.detach()))
            else:
                losses.append(None)
        finally:
            _untoggle(saved)
    try:
        next(gen)
        raise RuntimeError("train_step_gen yielded more times than there are optimizers")
    except StopIteration:
        pass
    return losses
