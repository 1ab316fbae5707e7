"""Lightning-1.6-style automatic optimisation with several optimizers, as the reference is
driven (reagent/training/reagent_lightning_module.py:108-133; pytorch_lightning==1.6.0 is
not vendored, so the order is restated: for each optimizer in configure_optimizers() order
-> next(train_step_gen) -> zero_grad -> backward -> step; a `None` yield skips that
optimizer, the intent stated at reagent/training/td3_trainer.py:197)."""


def run_update(trainer, batch, batch_idx: int, optimizers=None):
    """One full update (all sub-optimizer steps incl. the soft update).  Returns the list
    of yielded losses (tensors / None)."""
    opts = trainer.optimizers() if optimizers is None else optimizers
    losses = []
    for i, opt in enumerate(opts):
        loss = trainer.training_step(batch, batch_idx, i)
        if loss is not None:
            opt.zero_grad()
            loss.backward()
            opt.step()
        losses.append(loss)
    return losses
