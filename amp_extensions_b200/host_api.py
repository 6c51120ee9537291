"""Host-buffer entry point of the env step: pinned host arrays in, pinned host arrays out.

This is the call a CPU-side caller makes (the reference's samplers hold states, actions and costs as host
arrays: milo/milo/sampler.py:48-66, mjrl/mjrl/algos/batch_reinforce.py:103-169).  The batch is cut into
chunks; chunk i's host->device copy, its step on the compute stream and its device->host copy run on three
streams so that PCIe transfers overlap the tensor-core work of neighbouring chunks.
"""
import torch


class HostStepPipeline:
    def __init__(self, engine, num_envs, n_chunks=4, with_cost=True):
        self.eng = engine
        self.E = int(num_envs)
        self.with_cost = with_cost
        dev = engine.device
        S, A, E = engine.S, engine.A, self.E
        n_chunks = max(1, min(int(n_chunks), (E + 127) // 128))
        rows = -(-E // n_chunks)
        rows = -(-rows // 128) * 128  # tile-aligned chunks keep every GEMM tile full
        self.bounds = [(r0, min(E, r0 + rows)) for r0 in range(0, E, rows)]
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
        self.s_in = torch.cuda.Stream(dev)
        self.s_compute = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self.h2d_bytes_per_step = E * (S + A) * 4 + E * 4 + E * 4
        self.d2h_bytes_per_step = E * S * 4 + E * 4 + E * 4 + E + E * 4

    @staticmethod
    def pinned_like(shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    def step(self, state_h, action_h, member_h, steps_h, w_dev=None, lambda_b=0.0, threshold=1.0, c_min=-1.0,
             c_max=0.0, clamp_cost=True):
        """One batched env step on host buffers.  Returns pinned host tensors
        (next_state [E,S], cost [E] or None, done [E] uint8, disc [E], num_steps [E]) valid until the next call."""
        eng = self.eng
        cur = torch.cuda.current_stream(eng.device)
        self.s_in.wait_stream(cur)
        ev_in, ev_c = [], []
        for (r0, r1) in self.bounds:
            with torch.cuda.stream(self.s_in):
                self.d_state[r0:r1].copy_(state_h[r0:r1], non_blocking=True)
                self.d_action[r0:r1].copy_(action_h[r0:r1], non_blocking=True)
                self.d_member[r0:r1].copy_(member_h[r0:r1], non_blocking=True)
                self.d_steps[r0:r1].copy_(steps_h[r0:r1], non_blocking=True)
                e = torch.cuda.Event()
                e.record(self.s_in)
                ev_in.append(e)
        for i, (r0, r1) in enumerate(self.bounds):
            with torch.cuda.stream(self.s_compute):
                self.s_compute.wait_event(ev_in[i])
                if self.with_cost:
                    eng.step_cost(self.d_state[r0:r1], self.d_action[r0:r1], self.d_member[r0:r1],
                                  self.d_steps[r0:r1], w_dev, lambda_b, threshold, c_min, c_max, clamp_cost,
                                  next_state=self.d_next[r0:r1], disc=self.d_disc[r0:r1], done=self.d_done[r0:r1],
                                  cost=self.d_cost[r0:r1], ipm=self.d_ipm[r0:r1], bonus=self.d_bonus[r0:r1])
                else:
                    eng.step(self.d_state[r0:r1], self.d_action[r0:r1], self.d_member[r0:r1], self.d_steps[r0:r1],
                             next_state=self.d_next[r0:r1], disc=self.d_disc[r0:r1], done=self.d_done[r0:r1])
                e = torch.cuda.Event()
                e.record(self.s_compute)
                ev_c.append(e)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_c[i])
                self.h_next[r0:r1].copy_(self.d_next[r0:r1], non_blocking=True)
                self.h_disc[r0:r1].copy_(self.d_disc[r0:r1], non_blocking=True)
                self.h_done[r0:r1].copy_(self.d_done[r0:r1], non_blocking=True)
                self.h_steps[r0:r1].copy_(self.d_steps[r0:r1], non_blocking=True)
                if self.with_cost:
                    self.h_cost[r0:r1].copy_(self.d_cost[r0:r1], non_blocking=True)
        self.s_out.synchronize()  # the caller reads the host buffers next
        return self.h_next, (self.h_cost if self.with_cost else None), self.h_done, self.h_disc, self.h_steps
