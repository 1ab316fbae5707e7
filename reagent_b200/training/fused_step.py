"""Whole-update CUDA graphs: replay sample -> TD step -> weight gradients -> Adam + Polyak.

`FusedDqnStep` is the public fast path for the reference workflow "rb.sample_transition_batch
-> trainer_preprocessor -> Lightning optimizer loop" (reagent/gym/datasets/
replay_buffer_dataset.py:122-133 + reagent/training/reagent_lightning_module.py:108-133):
per update the host only draws the random numbers (Python's `random` stream for the
prioritized buffer, torch.randint for the uniform one -- bit-exact index parity with the
reference), writes them to pinned memory and replays one captured graph that does the
host->device copy, the kernels and the device->host copy of the loss.

`prefetch=True` pipelines the sampler one update ahead, the way the reference's DataLoader
over ReplayBufferDataset prefetches batches: the graph of update k trains on the batch drawn
during update k-1 while the replay-sample kernel for update k+1 runs on a second stream (it
does not depend on the parameters).  The host's random stream is consumed in the same order;
the only observable difference is that a transition added between two `step()` calls can be
sampled one update later.
"""
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..replay_memory.prioritized_replay_buffer import PrioritizedReplayBuffer


class FusedDqnStep:
    def __init__(self, trainer, replay_buffer, batch_size: int, process_group=None,
                 slots: int = 2, prefetch: bool = False, shard=None, rng: str = "host",
                 online: bool = False):
        """`shard = (rank, world)`: data-parallel strong scaling.  `batch_size`
        is the GLOBAL minibatch; the replay buffer is replicated and every rank consumes the
        identical random stream, so all ranks select the same global indices, and this
        rank gathers and trains on rows [rank*B/world, (rank+1)*B/world) only.

        `rng="device"` (prioritized buffer): the buffer goes device-resident
        (replay_memory/device_replay.py) -- Python's `random` state is uploaded once and the
        stratified draws, tree descents and retries of sample_index_batch run in a kernel
        inside the captured graph: no host random numbers, no per-step query upload.
        `online=True` (needs rng="device"): `step(transition)` also ADDS one transition before
        drawing, the reference's online loop (reagent/gym/runners/gymrunner.py: one env step
        -> replay_buffer.add -> one update); the transition is the step's only host->device
        traffic, staged in pinned memory and copied + inserted by the same graph replay."""
        if rng not in ("host", "device"):
            raise ValueError("rng must be 'host' or 'device'")
        if online and rng != "device":
            raise ValueError("online=True needs rng='device' (device-resident replay)")
        self.rng, self.online = rng, bool(online)
        self.trainer = trainer
        self.rb = replay_buffer
        self.B_global = batch_size
        self.row0 = 0
        if shard is not None and shard[1] > 1:
            from .data_parallel import shard_rows

            self.row0, hi = shard_rows(batch_size, shard[0], shard[1])
            batch_size = hi - self.row0
        self.B = batch_size
        self.pg = process_group
        self.prioritized = isinstance(replay_buffer, PrioritizedReplayBuffer)
        self.dev = replay_buffer._dev()
        self.A = trainer.num_actions
        self.slots = []
        self.k = 0
        self.h2d_bytes = batch_size * 8
        self._side = torch.cuda.Stream(device=self.dev)
        self._side2 = torch.cuda.Stream(device=self.dev)
        self.d2h_bytes = 4
        self.prefetch = bool(prefetch)
        replay_buffer._flush()
        self.dr = None
        if rng == "device":
            from ..replay_memory.device_replay import DeviceReplay

            if not self.prioritized:
                raise NotImplementedError("rng='device' covers the prioritized buffer")
            self.dr = getattr(replay_buffer, "_device_resident", None) or DeviceReplay(
                replay_buffer, stage_rows=1, stage_slots=2)
            if self.dr.stage_slots < 2:
                self.dr._alloc_stage(self.dr.stage_rows, 2)
            self._idx_buf = [torch.zeros(self.B_global, dtype=torch.int64, device=self.dev)
                             for _ in range(2)]
            self._status_host = torch.zeros(2, dtype=torch.int32).pin_memory()
            self._status_np = self._status_host.numpy()
            self.h2d_bytes = self.dr.h2d_bytes_per_add if self.online else 0
            self.d2h_bytes = 4 + 8
        # warm-up outside capture (lazy allocations, cudaFuncSetAttribute, optimizer state)
        self._one_update(None)
        torch.cuda.synchronize()
        if self.prefetch:
            # two fixed sets of batch tensors: update k trains on set k%2 while the sampler
            # fills set (k+1)%2.  Set 0 gets the first real draw now; set 1 is only allocated
            # (given indices: no random numbers consumed).
            self._pools = [{}, {}]
            self._batches = [None, None]
            with self.rb.output_buffers(self._pools[0]):
                self._batches[0] = self._sample(None)
            with self.rb.output_buffers(self._pools[1]):
                self._batches[1] = self.rb.sample_discrete_dqn_batch(
                    self.B, self.A, indices=self._batches[0].indices.reshape(-1))
            torch.cuda.synchronize()
            slots = 2
        for i in range(slots):
            self.slots.append(self._capture(i))
        self._param_versions = self._versions()

    # -- parameters changed from outside (load_state_dict, manual edits) ----------------------
    def _versions(self):
        t = self.trainer
        return tuple(p._version for p in t.q_network.parameters()) + tuple(
            p._version for p in t.q_network_target.parameters())

    def invalidate_tc_images(self):
        """The captured update keeps the tensor-core weight images of K2 current by itself (the
        Adam kernel rewrites them).  If the parameters of q_network / q_network_target are
        changed OUTSIDE this object, the images must be rebuilt before the next replay: writes
        through torch (load_state_dict, `p.copy_`, ...) are detected by `step()`; call this
        after anything torch's version counters cannot see (a raw kernel writing the arena)."""
        self._param_versions = None

    def _refresh_tc_images(self):
        v = self._versions()
        if v != self._param_versions:
            self.trainer._tc_images_state = None
            self.trainer.tc_prepack()  # eager, on the current stream, before the replay
            self._param_versions = v

    # -- one update on the current stream ---------------------------------------
    def _one_update(self, rnd_dev):
        # the tcgen05 K2 wants hi/lo weight images: they only depend on the parameters, so they
        # are built on a side stream while the replay-sample kernel runs (fork/join is
        # captured into the graph like any other dependency)
        main = torch.cuda.current_stream()
        prepack = getattr(self.trainer, "tc_prepack", None)
        forked = False
        if prepack is not None:
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                forked = prepack()
        batch = self._sample(rnd_dev)
        if forked:
            main.wait_stream(self._side)
        return self.trainer.train_batch(batch, process_group=self.pg)

    def _prefetch_update(self, i, rnd_dev, overrides=None):
        """Update on batch set i; the sampler fills set 1-i concurrently (second stream)."""
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        self._side2.wait_stream(main)
        prepack = getattr(self.trainer, "tc_prepack", None)
        forked = False
        if prepack is not None:
            with torch.cuda.stream(self._side2):
                forked = prepack()
        with torch.cuda.stream(self._side), self.rb.output_buffers(self._pools[1 - i]):
            if self.dr is not None:
                nxt = self._device_sample(1 - i, self.online and rnd_dev is not None, stage_row=i)
            elif overrides is None:
                nxt = self._sample(rnd_dev)
            else:
                nxt = self.rb.sample_discrete_dqn_batch(self.B, self.A, query_dev=rnd_dev,
                                                        overrides=overrides)
        if forked:
            main.wait_stream(self._side2)
        loss = self.trainer.train_batch(self._batches[i], process_group=self.pg)
        main.wait_stream(self._side)
        self._batches[1 - i] = nxt
        return loss

    def _host_draw(self):
        """This rank's rows of one GLOBAL host draw: (values, override positions, indices)."""
        lo, hi = self.row0, self.row0 + self.B
        if self.prioritized:
            q, pos, idxs = self.rb.host_queries(self.B_global)
            keep = [(p - lo, i) for p, i in zip(pos, idxs) if lo <= p < hi]
            return q[lo:hi], [p for p, _ in keep], [i for _, i in keep]
        n_valid = self.rb._num_valid_indices
        if n_valid == 0:
            raise RuntimeError(f"Cannot sample {self.B_global} since there are no valid indices so far.")
        return torch.randint(n_valid, (self.B_global,))[lo:hi].numpy(), [], []

    def _device_sample(self, slot: int, add: bool, stage_row: int = 0):
        """Device-resident draw: (optionally insert the staged transition,) select the global
        indices with the device MT19937 stream, gather this rank's rows."""
        if add:
            self.dr.launch_add(1, slot=stage_row)
        idx = self.dr.draw_indices(self.B_global, out=self._idx_buf[slot])
        return self.rb.sample_discrete_dqn_batch(self.B, self.A,
                                                 indices=idx[self.row0:self.row0 + self.B])

    def _sample(self, rnd_dev):
        if self.dr is not None:
            return self._device_sample(0, False)
        if rnd_dev is None and self.B != self.B_global:
            q, pos, idxs = self._host_draw()
            qd = torch.from_numpy(np.ascontiguousarray(q)).to(self.dev)
            if self.prioritized:
                kw = {"query_dev": qd}
                if pos:
                    kw["overrides"] = (pos, idxs)
            else:
                kw = {"ranks_dev": qd}
            batch = self.rb.sample_discrete_dqn_batch(self.B, self.A, **kw)
        elif rnd_dev is None:
            batch = self.rb.sample_discrete_dqn_batch(self.B, self.A)
        elif self.prioritized:
            batch = self.rb.sample_discrete_dqn_batch(self.B, self.A, query_dev=rnd_dev)
        else:
            batch = self.rb.sample_discrete_dqn_batch(self.B, self.A, ranks_dev=rnd_dev)
        return batch

    def _capture_device(self, i=0):
        loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        g = torch.cuda.CUDAGraph()
        marker = torch.zeros(1, device=self.dev)  # non-None: "inside the captured step"
        with torch.cuda.graph(g):
            if self.prefetch:
                loss = self._prefetch_update(i, marker)
            else:
                if self.online:
                    self.dr.launch_add(1, slot=i)
                loss = self.trainer.train_batch(self._device_sample(0, False), process_group=self.pg)
            loss_host.copy_(loss.reshape(1), non_blocking=True)
            self._status_host.copy_(self.dr.status, non_blocking=True)
        return {"graph": g, "loss_host": loss_host, "done": torch.cuda.Event(), "used": False}

    def _capture(self, i=0):
        if self.dr is not None:
            return self._capture_device(i)
        dt = torch.float64 if self.prioritized else torch.int64
        host = torch.zeros(self.B, dtype=dt).pin_memory()
        devb = torch.zeros(self.B, dtype=dt, device=self.dev)
        loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            devb.copy_(host, non_blocking=True)
            loss = self._prefetch_update(i, devb) if self.prefetch else self._one_update(devb)
            loss_host.copy_(loss.reshape(1), non_blocking=True)
        return {"graph": g, "host": host, "dev": devb, "loss_host": loss_host,
                "done": torch.cuda.Event(), "used": False}

    # -- public --------------------------------------------------------------------
    def step(self, transition=None) -> torch.Tensor:
        """One full update.  Returns the pinned host tensor that will hold the loss once the
        stream reaches the end of this update (call torch.cuda.current_stream().synchronize()
        or keep going: slots are recycled only after their event completed).
        `transition` (online mode): dict of the add() keyword arguments of the new transition."""
        s = self.slots[self.k % len(self.slots)]
        self.k += 1
        if s["used"]:
            s["done"].synchronize()
        self._refresh_tc_images()
        if self.dr is not None:
            if self._status_np[0] != 0:  # sticky device status of an earlier step
                self.dr.raise_if_failed(self._status_host)
            if self.online:
                if transition is None:
                    raise ValueError("online FusedDqnStep.step() needs the new transition")
                # every slot's graph copies from its own pinned staging row; the slot's previous
                # replay (and with it that H2D copy) was waited for above
                self.dr.stage(0, (self.k - 1) % len(self.slots), **transition)
            s["graph"].replay()
            s["done"].record()
            s["used"] = True
            return s["loss_host"]
        # bring device mirrors up to date OUTSIDE the captured graph (adds / set_priority)
        self.rb._flush()
        if self.prioritized:
            self.rb.sum_tree.device_heap(self.dev)
        else:
            self.rb._ensure_valid_index()
        if self.prioritized:
            q, pos, idxs = self._host_draw()
            if pos:  # rare retry path: resolved on the host, run this update un-captured
                qd = torch.from_numpy(q).to(self.dev)
                if self.prefetch:
                    loss = self._prefetch_update((self.k - 1) % 2, qd, overrides=(pos, idxs))
                else:
                    batch = self.rb.sample_discrete_dqn_batch(self.B, self.A, query_dev=qd,
                                                              overrides=(pos, idxs))
                    loss = self.trainer.train_batch(batch, process_group=self.pg)
                s["loss_host"].copy_(loss.reshape(1), non_blocking=True)
                s["done"].record()
                s["used"] = True
                return s["loss_host"]
            s["host"].numpy()[:] = q
        else:
            s["host"].copy_(torch.from_numpy(self._host_draw()[0]))
        s["graph"].replay()
        s["done"].record()
        s["used"] = True
        return s["loss_host"]


def capture_device_only(trainer, rb, batch_size, steps, queries_dev, process_group=None,
                        overlap_sampling=True):
    """`steps` consecutive updates in ONE graph with all random numbers already resident in
    HBM (queries_dev[k] is the k-th update's draw) -- the kernel-only measurement of
    bench.py.  With `overlap_sampling` the replay-sample kernel of update k+1 is captured on a
    second stream and runs concurrently with the TD / weight-gradient / Adam kernels of update
    k (the row-tile kernels leave ~20 SMs and most of HBM idle; sampling does not depend on
    the parameters, and no trainer of the path writes priorities back -- SURVEY.md fact 5)."""
    A = trainer.num_actions
    prioritized = isinstance(rb, PrioritizedReplayBuffer)

    def sample(k):
        if prioritized:
            return rb.sample_discrete_dqn_batch(batch_size, A, query_dev=queries_dev[k])
        return rb.sample_discrete_dqn_batch(batch_size, A, ranks_dev=queries_dev[k])

    g = torch.cuda.CUDAGraph()
    keep = []  # every batch stays alive until the capture ends: no cross-stream block reuse
    side = torch.cuda.Stream()
    with torch.cuda.graph(g):
        main = torch.cuda.current_stream()
        batch = sample(0)
        keep.append(batch)
        for k in range(steps):
            nxt = None
            if overlap_sampling and k + 1 < steps:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    nxt = sample(k + 1)
                keep.append(nxt)
            trainer.train_batch(batch, process_group=process_group)
            if k + 1 < steps:
                if nxt is None:
                    nxt = sample(k + 1)
                    keep.append(nxt)
                else:
                    main.wait_stream(side)
                batch = nxt
    g._rb200_keep = keep
    return g
