"""Host-buffer entry point of the env step: pinned host arrays in, pinned host arrays out.

This is the call a CPU-side caller makes (the reference's samplers hold states, actions and costs as host
arrays: milo/milo/sampler.py:48-66, mjrl/mjrl/algos/batch_reinforce.py:103-169).  Three CUDA streams carry the
host->device copies, the step itself and the device->host copies; a batch is cut into tile-aligned chunks so
that the PCIe transfers of one chunk overlap the tensor-core work of its neighbours, and `depth` independent
batches can be in flight (submit / collect) so that batch i+1's upload and batch i-1's download overlap batch
i's compute — the way a sampler that alternates between two groups of environments would drive it.

    pipe = HostStepPipeline(engine, E)
    out = pipe.step(state_h, action_h, member_h, steps_h, w_dev, lambda_b, threshold)      # synchronous
    t0 = pipe.submit(...); t1 = pipe.submit(...); out0 = pipe.collect(); ...                 # pipelined
"""
import collections

import torch


class _Slot:
    """Device buffers + pinned output buffers of one in-flight batch."""

    def __init__(self, eng, E):
        dev, S, A = eng.device, eng.S, eng.A
        f32 = dict(device=dev, dtype=torch.float32)
        self.d_state = torch.empty((E, S), **f32)
        self.d_action = torch.empty((E, A), **f32)
        self.d_member = torch.empty((E,), device=dev, dtype=torch.int32)
        self.d_steps = torch.empty((E,), device=dev, dtype=torch.int32)
        self.d_next = torch.empty((E, S), **f32)
        self.d_disc = torch.empty((E,), **f32)
        self.d_done = torch.empty((E,), device=dev, dtype=torch.uint8)
        self.d_cost = torch.empty((E,), **f32)
        self.d_ipm = torch.empty((E,), **f32)
        self.d_bonus = torch.empty((E,), **f32)
        pin = dict(pin_memory=True)
        self.h_next = torch.empty((E, S), dtype=torch.float32, **pin)
        self.h_disc = torch.empty((E,), dtype=torch.float32, **pin)
        self.h_done = torch.empty((E,), dtype=torch.uint8, **pin)
        self.h_cost = torch.empty((E,), dtype=torch.float32, **pin)
        self.h_steps = torch.empty((E,), dtype=torch.int32, **pin)
        self.ev_compute_done = None   # last compute event of the batch that used this slot
        self.ev_out_done = None       # its last device->host copy


class HostStepPipeline:
    def __init__(self, engine, num_envs, n_chunks=4, with_cost=True, depth=2):
        self.eng = engine
        self.E = int(num_envs)
        self.with_cost = with_cost
        dev = engine.device
        S, A, E = engine.S, engine.A, self.E
        n_chunks = max(1, min(int(n_chunks), (E + 255) // 256))
        rows = -(-E // n_chunks)
        rows = -(-rows // 256) * 256  # whole 256-row GEMM tiles per chunk (CTA pairs)
        self.bounds = [(r0, min(E, r0 + rows)) for r0 in range(0, E, rows)]
        self.depth = max(1, int(depth))
        self.slots = [_Slot(engine, E) for _ in range(self.depth)]
        self._next_slot = 0
        self._inflight = collections.deque()
        self.s_in = torch.cuda.Stream(dev)
        self.s_compute = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self.h2d_bytes_per_step = E * (S + A) * 4 + E * 4 + E * 4
        self.d2h_bytes_per_step = E * S * 4 + E * 4 + E * 4 + E + E * 4

    @staticmethod
    def pinned_like(shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    def submit(self, state_h, action_h, member_h, steps_h, w_dev=None, lambda_b=0.0, threshold=1.0, c_min=-1.0,
               c_max=0.0, clamp_cost=True):
        """Enqueue one batched env step on host buffers; returns immediately.  At most `depth` batches may be
        in flight: collect() the oldest before submitting another."""
        if len(self._inflight) >= self.depth:
            raise RuntimeError("HostStepPipeline: collect() a batch before submitting more than `depth`")
        eng = self.eng
        slot = self.slots[self._next_slot]
        self._next_slot = (self._next_slot + 1) % self.depth
        cur = torch.cuda.current_stream(eng.device)
        self.s_in.wait_stream(cur)
        if slot.ev_compute_done is not None:      # the slot's previous batch must have consumed its inputs
            self.s_in.wait_event(slot.ev_compute_done)
        if slot.ev_out_done is not None:          # ... and its outputs must have left the device buffers
            self.s_compute.wait_event(slot.ev_out_done)
        ev_in = []
        for (r0, r1) in self.bounds:
            with torch.cuda.stream(self.s_in):
                slot.d_state[r0:r1].copy_(state_h[r0:r1], non_blocking=True)
                slot.d_action[r0:r1].copy_(action_h[r0:r1], non_blocking=True)
                slot.d_member[r0:r1].copy_(member_h[r0:r1], non_blocking=True)
                slot.d_steps[r0:r1].copy_(steps_h[r0:r1], non_blocking=True)
                e = torch.cuda.Event()
                e.record(self.s_in)
                ev_in.append(e)
        for i, (r0, r1) in enumerate(self.bounds):
            with torch.cuda.stream(self.s_compute):
                self.s_compute.wait_event(ev_in[i])
                if self.with_cost:
                    eng.step_cost(slot.d_state[r0:r1], slot.d_action[r0:r1], slot.d_member[r0:r1],
                                  slot.d_steps[r0:r1], w_dev, lambda_b, threshold, c_min, c_max, clamp_cost,
                                  next_state=slot.d_next[r0:r1], disc=slot.d_disc[r0:r1], done=slot.d_done[r0:r1],
                                  cost=slot.d_cost[r0:r1], ipm=slot.d_ipm[r0:r1], bonus=slot.d_bonus[r0:r1])
                else:
                    eng.step(slot.d_state[r0:r1], slot.d_action[r0:r1], slot.d_member[r0:r1], slot.d_steps[r0:r1],
                             next_state=slot.d_next[r0:r1], disc=slot.d_disc[r0:r1], done=slot.d_done[r0:r1])
                ev_c = torch.cuda.Event()
                ev_c.record(self.s_compute)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_c)
                slot.h_next[r0:r1].copy_(slot.d_next[r0:r1], non_blocking=True)
                slot.h_disc[r0:r1].copy_(slot.d_disc[r0:r1], non_blocking=True)
                slot.h_done[r0:r1].copy_(slot.d_done[r0:r1], non_blocking=True)
                slot.h_steps[r0:r1].copy_(slot.d_steps[r0:r1], non_blocking=True)
                if self.with_cost:
                    slot.h_cost[r0:r1].copy_(slot.d_cost[r0:r1], non_blocking=True)
        slot.ev_compute_done = ev_c
        slot.ev_out_done = torch.cuda.Event()
        slot.ev_out_done.record(self.s_out)
        self._inflight.append(slot)
        return slot

    def collect(self):
        """Wait for the oldest in-flight batch.  Returns pinned host tensors
        (next_state [E,S], cost [E] or None, done [E] uint8, disc [E], num_steps [E]), valid until that slot is
        submitted again (i.e. for the next `depth - 1` submits)."""
        slot = self._inflight.popleft()
        slot.ev_out_done.synchronize()  # the caller reads the host buffers next
        return slot.h_next, (slot.h_cost if self.with_cost else None), slot.h_done, slot.h_disc, slot.h_steps

    def step(self, state_h, action_h, member_h, steps_h, w_dev=None, lambda_b=0.0, threshold=1.0, c_min=-1.0,
             c_max=0.0, clamp_cost=True):
        """One batched env step on host buffers, synchronous: submit + collect."""
        self.submit(state_h, action_h, member_h, steps_h, w_dev, lambda_b, threshold, c_min, c_max, clamp_cost)
        return self.collect()
