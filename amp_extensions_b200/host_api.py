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


class _Record:
    """One env group's device state + step outputs as ONE contiguous byte buffer per chunk, mirrored by one pinned
    host buffer: a chunk's results leave the device with a single copy.  Chunk layout (n rows):
    [next_state n x S f32 | cost n f32 | done n u8 (padded) | disc n f32 | num_steps n i32 | pad to 256 B]; the
    optional blocks (disc, num_steps) sit at the end, so a caller that does not want them back still gets its
    results with one copy (of the record's prefix)."""

    def __init__(self, eng, bounds, with_cost, want_disc, want_steps, pinned_mirror, obs_half=False):
        S = eng.S
        self.bounds = bounds
        self.chunks = []
        off = 0
        for (r0, r1) in bounds:
            n = r1 - r0
            lay, o = {}, 0
            lay["next"] = (o, n * S * 4); o += n * S * 4
            core0 = 0
            if obs_half:   # the fp32 state stays on the device; a half-precision copy is what crosses PCIe
                core0 = o
                lay["obs16"] = (o, n * S * 2); o += -(-(n * S * 2) // 16) * 16
            if with_cost:
                lay["cost"] = (o, n * 4); o += n * 4
            lay["done"] = (o, n); o += -(-n // 16) * 16
            core1 = o                          # what every caller wants back: one copy of [core0, core1)
            if want_disc:
                lay["disc"] = (o, n * 4); o += n * 4
            if want_steps:
                lay["steps"] = (o, n * 4); o += n * 4
            lay["_core"] = (core0, core1 - core0)
            lay["_tail"] = (core0, o - core0)
            size = -(-o // 256) * 256
            self.chunks.append((off, size, lay))
            off += size
        self.nbytes = off
        self.dev = torch.empty(off, device=eng.device, dtype=torch.uint8)
        self.host = torch.empty(off, dtype=torch.uint8, pin_memory=True) if pinned_mirror else None

    def view(self, buf, ci, name, dtype, shape):
        off, _, lay = self.chunks[ci]
        if name not in lay:
            return None
        o, sz = lay[name]
        return buf[off + o: off + o + sz].view(dtype).view(shape)


def chunk_bounds(num_envs, n_chunks, gran=256):
    """Row ranges of the chunks a step of `num_envs` envs is processed in: equal chunks of whole 256-row tiles, cut -
    when `gran` (Engine.round_rows) says so - where the forward kernel finishes a whole round of its CTA pairs: 40 000
    envs in two chunks are 18 944 + 21 056 rows (4 + 4.5 rounds of the kernel) instead of 20 224 + 19 776 (4.5 + 4.5).
    The last chunk takes the remainder; there are at most n_chunks of them."""
    E = int(num_envs)
    n_chunks = max(1, min(int(n_chunks), (E + 255) // 256))
    rows = -(-(-(-E // n_chunks)) // 256) * 256
    if gran > 256 and rows >= gran and round(rows / gran) * gran * (n_chunks - 1) < E:
        rows = round(rows / gran) * gran
    bounds = [(r0, min(E, r0 + rows)) for r0 in range(0, E, rows)][:n_chunks]
    bounds[-1] = (bounds[-1][0], E)
    return bounds


class HostEnvPipeline:
    """The reference plugin's own call shape, batched: `step(actions) -> (obs, cost, done, ...)` with the env state
    RESIDENT on the device (gym-simenv/gym_simenv/envs/sim_env.py:140-162 keeps `self.ob` inside the env and takes
    only the action).  Per step only the actions cross PCIe host->device; observations, costs and flags come back
    as one packed record per chunk (one device->host copy each; disc / step counters only on request).

    `groups` independent groups of E envs alternate (submit / collect), so group g+1's action upload and group
    g-1's result download overlap group g's compute — a sampler that steps two sets of environments in turn.

        pipe = HostEnvPipeline(engine, E, groups=2)
        pipe.reset(0, states0_h, member0_h); pipe.reset(1, states1_h, member1_h)
        pipe.submit(0, actions_h, w_dev, lambda_b, threshold); pipe.submit(1, ...)
        obs_h, cost_h, done_h, disc_h, steps_h = pipe.collect()       # group 0's results
    """

    def __init__(self, engine, num_envs, groups=2, n_chunks=2, with_cost=True, want_disc=True, want_steps=True,
                 obs_dtype=torch.float32):
        """obs_dtype=torch.float16: observations come back in half precision (2^-11 relative rounding, inside the
        1e-3 budget; the env state itself stays fp32 on the device) - halves the bytes per step where the host's
        PCIe is the limit (8 GPUs on one host)."""
        self.eng, self.E, self.with_cost = engine, int(num_envs), with_cost
        self.obs_half = obs_dtype in (torch.float16, torch.half)
        dev, S, A, E = engine.device, engine.S, engine.A, self.E
        rows = -(-(-(-E // max(1, min(int(n_chunks), (E + 255) // 256)))) // 256) * 256
        gran = engine.round_rows(rows) if hasattr(engine, "round_rows") else 256
        self.bounds = chunk_bounds(E, n_chunks, gran)
        self.groups = []
        f32 = dict(device=dev, dtype=torch.float32)
        for _ in range(int(groups)):
            g = type("Group", (), {})()
            # two records alternate: the step reads its state from one and writes s' (+ outputs) into the other
            g.rec = [_Record(engine, self.bounds, with_cost, True, True, pinned_mirror=False, obs_half=self.obs_half)
                     for _ in range(2)]
            g.cur = 0
            g.d_action = torch.empty((E, A), **f32)
            g.d_member = torch.zeros((E,), device=dev, dtype=torch.int32)
            g.d_steps = torch.zeros((E,), device=dev, dtype=torch.int32)
            g.d_ipm = torch.empty((E,), **f32)
            g.d_bonus = torch.empty((E,), **f32)
            g.host = torch.empty(g.rec[0].nbytes, dtype=torch.uint8, pin_memory=True)
            g.ev_compute_done = None
            g.ev_out_done = None
            self.groups.append(g)
        self.want_disc, self.want_steps = bool(want_disc), bool(want_steps)
        self._inflight = collections.deque()
        self.s_in, self.s_compute, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.h2d_bytes_per_step = E * A * 4
        # bytes a step actually brings back: the record's payload minus the blocks the caller opted out of
        self.d2h_bytes_per_step = E * (S * (2 if self.obs_half else 4) + 1 + (4 if with_cost else 0) +
                                       (4 if want_disc else 0) + (4 if want_steps else 0))

    def _views(self, g, r, ci, buf):
        (r0, r1) = self.bounds[ci]
        n, S = r1 - r0, self.eng.S
        rec = g.rec[r]
        return dict(next=rec.view(buf, ci, "next", torch.float32, (n, S)),
                    obs16=rec.view(buf, ci, "obs16", torch.float16, (n, S)),
                    cost=rec.view(buf, ci, "cost", torch.float32, (n,)),
                    disc=rec.view(buf, ci, "disc", torch.float32, (n,)),
                    steps=rec.view(buf, ci, "steps", torch.int32, (n,)),
                    done=rec.view(buf, ci, "done", torch.uint8, (n,)))

    def reset(self, group, state_h, member_h=None, steps_h=None):
        """(Re)start every env of a group from host-supplied initial states (the reference draws them from the
        DeepMimic simulator, sim_env.py:270-285); member indices / step counters default to 0."""
        g = self.groups[group]
        cur = torch.cuda.current_stream(self.eng.device)
        for s in (self.s_in, self.s_compute, self.s_out):
            cur.wait_stream(s)
        for ci, (r0, r1) in enumerate(self.bounds):
            self._views(g, g.cur, ci, g.rec[g.cur].dev)["next"].copy_(state_h[r0:r1], non_blocking=True)
        if member_h is None:
            g.d_member.zero_()
        else:
            g.d_member.copy_(member_h, non_blocking=True)
        if steps_h is None:
            g.d_steps.zero_()
        else:
            g.d_steps.copy_(steps_h, non_blocking=True)
        for s in (self.s_in, self.s_compute, self.s_out):
            s.wait_stream(cur)

    def submit(self, group, action_h, w_dev=None, lambda_b=0.0, threshold=1.0, c_min=-1.0, c_max=0.0,
               clamp_cost=True):
        """Enqueue one step of a group's envs on pinned host actions [E, A]; returns immediately."""
        eng, g = self.eng, self.groups[group]
        if g in self._inflight:
            raise RuntimeError("HostEnvPipeline: collect() this group's previous step before submitting the next")
        if g.ev_out_done is not None:          # the previous results must have left the record ...
            self.s_compute.wait_event(g.ev_out_done)
        if g.ev_compute_done is not None:      # ... and the previous step must have consumed its actions
            self.s_in.wait_event(g.ev_compute_done)
        src, dst = g.cur, 1 - g.cur
        ev_c = None
        for ci, (r0, r1) in enumerate(self.bounds):
            with torch.cuda.stream(self.s_in):
                g.d_action[r0:r1].copy_(action_h[r0:r1], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            vin = self._views(g, src, ci, g.rec[src].dev)
            vout = self._views(g, dst, ci, g.rec[dst].dev)
            with torch.cuda.stream(self.s_compute):
                self.s_compute.wait_event(ev_in)
                if self.with_cost:
                    eng.step_cost(vin["next"], g.d_action[r0:r1], g.d_member[r0:r1], g.d_steps[r0:r1], w_dev,
                                  lambda_b, threshold, c_min, c_max, clamp_cost, next_state=vout["next"],
                                  disc=vout["disc"], done=vout["done"], cost=vout["cost"], ipm=g.d_ipm[r0:r1],
                                  bonus=g.d_bonus[r0:r1])
                else:
                    eng.step(vin["next"], g.d_action[r0:r1], g.d_member[r0:r1], g.d_steps[r0:r1],
                             next_state=vout["next"], disc=vout["disc"], done=vout["done"])
                if self.want_steps:
                    vout["steps"].copy_(g.d_steps[r0:r1], non_blocking=True)
                if self.obs_half:
                    vout["obs16"].copy_(vout["next"])   # fp32 -> fp16 on the device
                ev_c = torch.cuda.Event()
                ev_c.record(self.s_compute)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_c)
                off, size, lay = g.rec[dst].chunks[ci]
                if self.want_disc and self.want_steps:   # ONE copy: the whole record (without the fp32 state in
                    o, sz = lay["_tail"]                   # half-precision observation mode)
                    g.host[off + o:off + o + sz].copy_(g.rec[dst].dev[off + o:off + o + sz], non_blocking=True)
                else:  # the optional blocks sit behind the core: one copy of the core, plus the one that is wanted
                    o, sz = lay["_core"]
                    g.host[off + o:off + o + sz].copy_(g.rec[dst].dev[off + o:off + o + sz], non_blocking=True)
                    for name in ("disc", "steps"):
                        if (name == "disc" and self.want_disc) or (name == "steps" and self.want_steps):
                            o, sz = lay[name]
                            g.host[off + o: off + o + sz].copy_(g.rec[dst].dev[off + o: off + o + sz],
                                                               non_blocking=True)
        g.cur = dst   # s' is the next step's state; it never left the device
        g.ev_compute_done = ev_c
        g.ev_out_done = torch.cuda.Event()
        g.ev_out_done.record(self.s_out)
        self._inflight.append(g)

    def collect(self):
        """Wait for the oldest in-flight step.  Returns pinned host tensors (obs, cost or None, done uint8, disc or
        None, num_steps or None), each a list-free view when the batch is one chunk and otherwise per-chunk views
        concatenated lazily: use `collect_chunks()` to avoid the concatenation.  Valid until that group is submitted
        again."""
        chunks = self.collect_chunks()
        if len(chunks) == 1:
            c = chunks[0]
            return c["next"], c["cost"], c["done"], c["disc"], c["steps"]
        cat = lambda k: None if chunks[0][k] is None else torch.cat([c[k] for c in chunks])  # noqa: E731
        return cat("next"), cat("cost"), cat("done"), cat("disc"), cat("steps")

    def collect_chunks(self):
        """Per-chunk views (dicts with next / cost / done / disc / steps) into the group's pinned record: no copy."""
        g = self._inflight.popleft()
        g.ev_out_done.synchronize()
        out = []
        for ci in range(len(self.bounds)):
            v = self._views(g, g.cur, ci, g.host)
            if self.obs_half:
                v["next"] = v["obs16"]   # the observations the caller reads are the half-precision copy
            if not self.want_disc:
                v["disc"] = None
            if not self.want_steps:
                v["steps"] = None
            out.append(v)
        return out
