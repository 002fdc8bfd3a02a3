"""Discrete-event model of the cross-GPU protocol of the NVSwitch collectives
(csrc/coll/collectives.cu) and of the stream ordering the data-parallel engine puts around them
(ddp/naive_ddp.py ``_reduce_bucket`` / ``finalize``).

Every GPU is a set of coroutines -- its compute stream, its communication stream and the CTAs of
the collective kernels -- scheduled in random interleavings:

* ``block_barrier``: one flag word per (block, source rank) in every rank's signal pad; the
  sender flips it 0 -> 1 with a release CAS and spins while it is still 1, the receiver flips it
  1 -> 0 with an acquire CAS.  Two barrier slots alternate (kernel entry / kernel exit) and are
  reused by the next kernel on the stream.
* two-shot all-reduce: barrier, rank r reads slice r of every peer's bucket, writes the sum into
  slice r of every peer's bucket, barrier.
* fused step: barrier, rank r reduces slice r of the gradients, updates its shard, writes the new
  parameters into slice r of every peer's *parameter* buffer, barrier.
* one-shot all-reduce: barrier, every rank reads everything into a private copy, barrier, writes
  its own buffer.
* streams: the compute stream writes a bucket's gradients (and reads that bucket's parameters for
  the input gradients), then records an event; the communication stream waits for the event
  before the kernel; the compute stream joins the communication stream before the optimizer /
  the next forward.

Further down: the MoE dispatch / combine path (ping-pong halves, one barrier per call) and the
fused sequence-parallel GEMMs (ping-pong halves, epoch flags, no barrier at all).

Buffers carry version tags, so the model catches: a reduction that reads gradients which are not
final, a peer's store landing in a buffer that is still being read, a consumer that sees
un-reduced data, lost or duplicated barrier signals, and deadlocks.  The negative tests remove one
ordering edge each -- including the missing wait on the gradient-producing stream that was found on
hardware at 8 GPUs this round -- and must fail."""
import random

import pytest


class Deadlock(AssertionError):
    pass


class World:
    def __init__(self, n, n_buckets, n_blocks, seed):
        self.n, self.n_buckets, self.n_blocks = n, n_buckets, n_blocks
        self.rng = random.Random(seed)
        self.live = []
        # pad[rank][(slot, block, src)] in {0, 1}
        self.pad = [dict() for _ in range(n)]
        cells = [(s, b) for s in range(n) for b in range(n_blocks)]     # (slice, block portion)
        self.cells = cells
        self.grad = [[{c: None for c in cells} for _ in range(n_buckets)] for _ in range(n)]
        self.param = [[{c: 0 for c in cells} for _ in range(n_buckets)] for _ in range(n)]
        self.event = {}            # (rank, step, bucket) -> gradients of that bucket are written
        self.comm_done = {}        # (rank, step, bucket) -> kernel finished on that rank

    # ---- scheduler
    def spawn(self, gen):
        self.live.append(gen)

    def run(self, max_steps=400000):
        idle = 0
        for _ in range(max_steps):
            if not self.live:
                return
            g = self.rng.choice(self.live)
            try:
                progressed = next(g)
            except StopIteration:
                self.live.remove(g)
                idle = 0
                continue
            idle = 0 if progressed else idle + 1
            if idle > 200 * len(self.live):
                raise Deadlock("every coroutine is waiting")
        raise AssertionError("did not finish")


def wait(pred):
    while not pred():
        yield False
    yield True


def block_barrier(w, rank, block, slot, skip=False):
    """collectives.cu block_barrier: thread t < world signals peer t, then waits for peer t's
    signal in its own pad; the threads run independently, a __syncthreads closes the barrier."""
    if skip:
        yield True
        return
    state = [0] * w.n                      # per thread: 0 = put pending, 1 = wait pending, 2 = done
    while any(s != 2 for s in state):
        t = w.rng.choice([i for i, s in enumerate(state) if s != 2])
        if state[t] == 0:
            key = (slot, block, rank)      # my flag in peer t's pad
            if w.pad[t].get(key, 0) == 0:  # CAS 0 -> 1
                w.pad[t][key] = 1
                state[t] = 1
                yield True
            else:
                yield False                # peer has not consumed my previous signal yet
        else:
            key = (slot, block, t)         # peer t's flag in my pad
            if w.pad[rank].get(key, 0) == 1:   # CAS 1 -> 0
                w.pad[rank][key] = 0
                state[t] = 2
                yield True
            else:
                yield False


def two_shot_block(w, rank, step, bucket, block, flaws):
    yield from block_barrier(w, rank, block, 0, skip="no_entry_barrier" in flaws)
    cell = (rank, block)
    for p in w.rng.sample(range(w.n), w.n):                  # shot 1: read slice `rank` everywhere
        got = w.grad[p][bucket][cell]
        assert got == ("local", step, p), f"rank {rank} reduced {got} from rank {p} (step {step})"
        yield True
    for p in w.rng.sample(range(w.n), w.n):                  # shot 2: multicast the sum
        w.grad[p][bucket][cell] = ("avg", step)
        yield True
    yield from block_barrier(w, rank, block, 1, skip="no_exit_barrier" in flaws)


def fused_block(w, rank, step, bucket, block, flaws):
    yield from block_barrier(w, rank, block, 0, skip="no_entry_barrier" in flaws)
    cell = (rank, block)
    for p in w.rng.sample(range(w.n), w.n):                  # reduce-scatter of my slice
        got = w.grad[p][bucket][cell]
        assert got == ("local", step, p), f"rank {rank} reduced {got} from rank {p} (step {step})"
        yield True
    for p in w.rng.sample(range(w.n), w.n):                  # AdamW on my shard, all-gather
        assert w.param[p][bucket][cell] == step, "parameter slice updated twice"
        w.param[p][bucket][cell] = step + 1
        yield True
    yield from block_barrier(w, rank, block, 1, skip="no_exit_barrier" in flaws)


def one_shot_block(w, rank, step, bucket, block, flaws):
    yield from block_barrier(w, rank, block, 0, skip="no_entry_barrier" in flaws)
    mine = [(s, block) for s in range(w.n)]                  # this block's share of the whole buffer
    for c in mine:
        for p in range(w.n):
            got = w.grad[p][bucket][c]
            assert got == ("local", step, p), f"rank {rank} reduced {got} from rank {p}"
            yield True
    yield from block_barrier(w, rank, block, 1, skip="no_exit_barrier" in flaws)   # peers have read me
    for c in mine:
        w.grad[rank][bucket][c] = ("avg", step)
        yield True


KERNELS = {"two_shot": two_shot_block, "fused": fused_block, "one_shot": one_shot_block}


def comm_stream(w, rank, steps, kind, flaws):
    for step in range(steps):
        for bucket in range(w.n_buckets):
            if "no_producer_wait" not in flaws:
                yield from wait(lambda: w.event.get((rank, step, bucket), False))
            else:
                # the bug found at 8 GPUs: the stream waited for was not the one the gradient
                # kernels ran on -- the host had issued them, the device had not finished
                yield from wait(lambda: w.event.get((rank, step, bucket, "issued"), False))
            done = []
            for block in range(w.n_blocks):
                def cta(block=block):
                    yield from KERNELS[kind](w, rank, step, bucket, block, flaws)
                    done.append(block)
                w.spawn(cta())
            yield from wait(lambda: len(done) == w.n_blocks)     # kernel complete (stream order)
            w.comm_done[(rank, step, bucket)] = True


def compute_stream(w, rank, steps, kind, flaws):
    for step in range(steps):
        # forward: reads every parameter
        for bucket in range(w.n_buckets):
            for c in w.cells:
                assert w.param[rank][bucket][c] == step, \
                    f"rank {rank} forward of step {step} saw parameters v{w.param[rank][bucket][c]}"
            yield True
        # backward, bucket by bucket: input gradients read the parameters, weight gradients
        # are written into the bucket, then the reducer's hook records the event
        for bucket in range(w.n_buckets):
            w.event[(rank, step, bucket, "issued")] = True        # host side: kernels enqueued
            for c in w.cells:
                assert w.param[rank][bucket][c] == step, \
                    f"rank {rank}: parameters of bucket {bucket} changed under its backward"
                prev = w.grad[rank][bucket][c]
                assert prev is None or prev == "consumed", \
                    f"rank {rank} overwrote gradient {prev} that nobody consumed"
                w.grad[rank][bucket][c] = ("local", step, rank)
                yield True
            w.event[(rank, step, bucket)] = True
        # reduce_gradients(): the compute stream joins the communication stream
        if "no_join" not in flaws:
            yield from wait(lambda: all(w.comm_done.get((rank, step, b), False)
                                        for b in range(w.n_buckets)))
        # consumer: optimizer (plain modes) -- in the fused mode the kernel was the optimizer
        for bucket in range(w.n_buckets):
            for c in w.cells:
                if kind == "fused":
                    w.grad[rank][bucket][c] = "consumed"
                else:
                    got = w.grad[rank][bucket][c]
                    assert got == ("avg", step), f"rank {rank} optimizer read {got} (step {step})"
                    w.grad[rank][bucket][c] = "consumed"
                    w.param[rank][bucket][c] = step + 1
            yield True


def simulate(kind, n, seed, steps=3, n_buckets=2, n_blocks=2, flaws=()):
    w = World(n, n_buckets, n_blocks, seed)
    for r in range(n):
        w.spawn(compute_stream(w, r, steps, kind, flaws))
        w.spawn(comm_stream(w, r, steps, kind, flaws))
    w.run()
    for r in range(n):
        assert all(v == 0 for v in w.pad[r].values()), "a barrier signal was left behind"
        for b in range(n_buckets):
            assert all(v == steps for v in w.param[r][b].values()), "replicas diverged"


@pytest.mark.parametrize("kind", ["two_shot", "fused", "one_shot"])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_collective_protocol_is_race_and_deadlock_free(kind, n):
    for seed in range(12 if n < 8 else 5):
        simulate(kind, n, seed)


@pytest.mark.parametrize("kind,flaw", [
    ("two_shot", "no_entry_barrier"),     # a rank reduces before its peers' gradients are final
    ("two_shot", "no_exit_barrier"),      # the optimizer reads slices its peers have not written yet
    ("fused", "no_entry_barrier"),        # new parameters land while a peer's backward still reads them
    ("fused", "no_exit_barrier"),         # the next forward runs on a mix of old and new parameters
    ("one_shot", "no_entry_barrier"),
    ("one_shot", "no_exit_barrier"),      # a rank overwrites its buffer while peers still read it
    ("two_shot", "no_producer_wait"),     # the ordering bug found on hardware (see module docstring)
    ("fused", "no_producer_wait"),
    ("two_shot", "no_join"),              # optimizer not ordered behind the communication stream
])
def test_collective_protocol_model_detects_missing_ordering(kind, flaw):
    failures = 0
    for seed in range(40):
        try:
            simulate(kind, 4, seed, flaws=(flaw,))
        except AssertionError:
            failures += 1
    assert failures > 0, f"the model did not notice the missing edge {flaw!r}"


# ------------------------------------------------------------------------------------------------
# MoE dispatch / combine over peer memory (moe/layer.py _scatter / _gather): ping-pong halves and
# ONE cross-GPU barrier per call.
#   scatter, call c: clear half (c+1)&1 locally; store my routed rows into half c&1 of their owners;
#                    barrier; read my half c&1.
#   gather,  call c: write my slot rows into my half c&1; barrier; read the rows I need from the
#                    peers' half c&1.
# The claim to check: one barrier per call is enough because a half is only touched again two
# calls later, with the barrier of the call in between separating the two uses.
# ------------------------------------------------------------------------------------------------
class A2AWorld(World):
    def __init__(self, n, seed, halves=2):
        super().__init__(n, 1, 1, seed)
        self.halves = halves
        # slots[rank][half][src] : "zero" | ("rows", call, src) | ("slot", call)
        self.slots = [[{s: "zero" for s in range(n)} for _ in range(halves)] for _ in range(n)]
        self.unread = [[set() for _ in range(halves)] for _ in range(n)]    # who still has to read


def a2a_barrier(w, rank, slot, skip):
    yield from block_barrier(w, rank, 0, ("a2a", slot), skip=skip)


def scatter_rank(w, rank, calls, flaws):
    H = w.halves
    for c in range(calls):
        use, other = c % H, (c + 1) % H
        if H > 1:
            for s in range(w.n):                               # clear the half of the NEXT call
                assert s not in w.unread[rank][other], "cleared rows that were not read yet"
                w.slots[rank][other][s] = "zero"
                yield True
        for p in w.rng.sample(range(w.n), w.n):                # my rows -> their owners
            if H == 1 and w.slots[p][use][rank] != "zero":
                # single buffer: the owner clears after reading; a store on top of unread rows
                # or a clear on top of fresh rows is the race the ping-pong avoids
                raise AssertionError("store into a slot that still holds the previous call's rows")
            assert w.slots[p][use][rank] == "zero", \
                f"rank {rank} call {c}: slot on rank {p} holds {w.slots[p][use][rank]}"
            w.slots[p][use][rank] = ("rows", c, rank)
            w.unread[p][use].add(rank)
            yield True
        yield from a2a_barrier(w, rank, c & 1, "no_barrier" in flaws)
        for s in range(w.n):                                   # read what landed in my half
            got = w.slots[rank][use][s]
            assert got == ("rows", c, s), f"rank {rank} call {c}: read {got} from source {s}"
            w.unread[rank][use].discard(s)
            if H == 1:
                w.slots[rank][use][s] = "zero"
            yield True


def gather_rank(w, rank, calls, flaws):
    H = w.halves
    for c in range(calls):
        use = c % H
        assert not w.unread[rank][use], \
            f"rank {rank} call {c}: rewrote slot rows that ranks {sorted(w.unread[rank][use])} still need"
        for s in range(w.n):
            w.slots[rank][use][s] = ("slot", c)
        w.unread[rank][use] = set(range(w.n))
        yield True
        yield from a2a_barrier(w, rank, c & 1, "no_barrier" in flaws)
        for p in w.rng.sample(range(w.n), w.n):                # pull my rows from their owners
            got = w.slots[p][use][rank]
            assert got == ("slot", c), f"rank {rank} call {c}: read {got} from rank {p}"
            w.unread[p][use].discard(rank)
            yield True


def simulate_a2a(role, n, seed, calls=6, halves=2, flaws=()):
    w = A2AWorld(n, seed, halves)
    for r in range(n):
        w.spawn(role(w, r, calls, flaws))
    w.run()
    for r in range(n):
        assert all(v == 0 for v in w.pad[r].values()), "a barrier signal was left behind"


@pytest.mark.parametrize("role", [scatter_rank, gather_rank], ids=["dispatch", "combine"])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_moe_all_to_all_one_barrier_per_call(role, n):
    for seed in range(10 if n < 8 else 4):
        simulate_a2a(role, n, seed)


@pytest.mark.parametrize("role,kw", [
    (scatter_rank, dict(flaws=("no_barrier",))),     # reads rows that have not landed
    (gather_rank, dict(flaws=("no_barrier",))),      # pulls rows the owner has not written
    (scatter_rank, dict(halves=1)),                  # one buffer + one barrier: store races the clear
    (gather_rank, dict(halves=1)),                   # one buffer + one barrier: rewrite under readers
], ids=["dispatch-no-barrier", "combine-no-barrier", "dispatch-single-buffer", "combine-single-buffer"])
def test_moe_all_to_all_model_detects_missing_ordering(role, kw):
    failures = 0
    for seed in range(40):
        try:
            simulate_a2a(role, 4, seed, **kw)
        except AssertionError:
            failures += 1
    assert failures > 0


# ------------------------------------------------------------------------------------------------
# Fused sequence-parallel GEMMs (parallel/tensor_parallel/tp_fused.py _Region): all-gather -> GEMM
# and GEMM -> reduce-scatter use ping-pong halves with epoch flags / counters and NO cross-GPU
# barrier at all.  The claim: a rank can be at most one use ahead of its peers, because every use
# ends in a wait on every peer's data of that use -- so when half k%2 is written again in use
# k+2, all its readers of use k are done.
# ------------------------------------------------------------------------------------------------
class TpWorld(World):
    def __init__(self, n, seed, halves):
        super().__init__(n, 1, 1, seed)
        self.halves = halves
        self.buf = [[{s: None for s in range(n)} for _ in range(halves)] for _ in range(n)]
        self.flag = [{s: 0 for s in range(n)} for _ in range(n)]      # AG: epoch per source chunk
        self.counter = [0] * n                                        # RS: tiles landed at the owner


def ag_gemm_rank(w, rank, uses):
    for k in range(uses):
        half = k % w.halves
        for p in w.rng.sample(range(w.n), w.n):          # push my shard, then raise my chunk flag
            w.buf[p][half][rank] = ("x", k, rank)
            yield True
            w.flag[p][rank] = k + 1                      # release store of the epoch
            yield True
        for i in range(w.n):                             # TMA producer: local chunk first
            src = (rank + i) % w.n
            yield from wait(lambda: w.flag[rank][src] >= k + 1)
            for _ in range(2):                           # the chunk is read over many k-blocks
                got = w.buf[rank][half][src]
                assert got == ("x", k, src), f"rank {rank} use {k}: chunk of rank {src} holds {got}"
                yield True


def gemm_rs_rank(w, rank, uses):
    for k in range(uses):
        half = k % w.halves
        for i in range(1, w.n + 1):                      # remote chunks first, own chunk last
            owner = (rank + i) % w.n
            w.buf[owner][half][rank] = ("part", k, rank)     # epilogue TMA store into the owner
            yield True
            w.counter[owner] += 1                            # release increment
            yield True
        yield from wait(lambda: w.counter[rank] >= (k + 1) * w.n)      # rs_reduce
        for src in range(w.n):
            got = w.buf[rank][half][src]
            assert got == ("part", k, src), f"rank {rank} use {k}: partial of rank {src} is {got}"
            yield True


def simulate_tp(role, n, seed, uses=6, halves=2):
    w = TpWorld(n, seed, halves)
    for r in range(n):
        w.spawn(role(w, r, uses))
    w.run()


@pytest.mark.parametrize("role", [ag_gemm_rank, gemm_rs_rank], ids=["ag_gemm", "gemm_rs"])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_fused_tp_ping_pong_needs_no_barrier(role, n):
    for seed in range(10 if n < 8 else 4):
        simulate_tp(role, n, seed)


@pytest.mark.parametrize("role", [ag_gemm_rank, gemm_rs_rank], ids=["ag_gemm", "gemm_rs"])
def test_fused_tp_single_buffer_would_race(role):
    failures = 0
    for seed in range(40):
        try:
            simulate_tp(role, 4, seed, halves=1)
        except AssertionError:
            failures += 1
    assert failures > 0
