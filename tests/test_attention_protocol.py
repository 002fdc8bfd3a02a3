"""Discrete-event model of the synchronisation protocol of the tcgen05 attention kernels
(csrc/attn/*.cu): every role (TMA producer, MMA issuer, softmax / drain warp-groups) is a
generator that performs the same mbarrier waits / arrivals / tcgen05.commit in the same order as
the CUDA code, scheduled in random interleavings.  The model checks that the protocol never
deadlocks, that mbarrier phases are consumed in order, and that no buffer (smem stage, TMEM
region, P tile) is overwritten while a consumer still has it -- the class of bug that cannot be
seen by compiling and is expensive to find on hardware."""
import random

import pytest


class MBar:
    """mbarrier with an arrival count; wait(parity) passes once the phase with that parity has
    completed (parity of the number of completed phases - 1), like mbarrier.try_wait.parity."""

    def __init__(self, count):
        self.count, self.pending, self.completed = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending = self.count
            self.completed += 1

    def passed(self, parity):
        # phase k (k = 0, 1, ...) has parity k & 1; waiting on `parity` succeeds iff the most
        # recently completed phase has that parity (the waiter is never more than one phase behind)
        return self.completed > 0 and ((self.completed - 1) & 1) == parity

    def passed_fresh(self, parity):
        # "wait(parity ^ 1)" idiom on a fresh barrier: passes immediately the first time
        return ((self.completed - 1) & 1) == parity if self.completed > 0 else parity == 1


class Resource:
    """A buffer with one writer and readers; flags overlapping write / read."""

    def __init__(self, name):
        self.name, self.readers, self.version, self.writing = name, 0, -1, False

    def write(self, version):
        assert self.readers == 0, f"{self.name}: overwritten (v{version}) while being read"
        self.version = version

    def begin_read(self, version):
        assert self.version == version, f"{self.name}: read v{version} but holds v{self.version}"
        self.readers += 1

    def end_read(self):
        self.readers -= 1


def run(roles, seed, max_steps=200000):
    rng = random.Random(seed)
    live = list(roles)
    blocked_rounds = 0
    for _ in range(max_steps):
        if not live:
            return
        g = rng.choice(live)
        try:
            progressed = next(g)
        except StopIteration:
            live.remove(g)
            blocked_rounds = 0
            continue
        blocked_rounds = 0 if progressed else blocked_rounds + 1
        assert blocked_rounds < 50 * len(roles), "deadlock: every role is waiting"
    raise AssertionError("did not finish")


def wait(pred):
    while not pred():
        yield False
    yield True


# ------------------------------------------------------------------------------------------------
# forward: mirrors attn_fwd_sm100.cu
# ------------------------------------------------------------------------------------------------
def forward_roles(n_kv, stages=3):
    n_max = max(n_kv)
    k_full = [MBar(1) for _ in range(stages)]
    k_empty = [MBar(1) for _ in range(stages)]
    v_full = [MBar(1) for _ in range(stages)]
    v_empty = [MBar(1) for _ in range(stages)]
    s_full = [MBar(1), MBar(1)]
    p_ready = [MBar(4), MBar(4)]
    o_full = [MBar(1), MBar(1)]
    K = [Resource(f"K{s}") for s in range(stages)]
    V = [Resource(f"V{s}") for s in range(stages)]
    S = [Resource("S0"), Resource("S1")]
    P = [Resource("P0"), Resource("P1")]
    O = [Resource("O0'"), Resource("O1'")]
    pending = []        # (bars to arrive, resources to release) of MMAs not yet "completed"

    def commit(bars, releases=()):
        # tcgen05.commit: arrives once every MMA issued so far has completed; the model completes
        # MMAs lazily (in issue order) to expose missing waits
        pending.append((list(bars), list(releases)))

    def mma_engine():
        while True:
            if pending:
                bars, rel = pending.pop(0)
                for r in rel:
                    r.end_read()
                for b in bars:
                    b.arrive()
                yield True
            else:
                yield False
                if done["mma"] and not pending:
                    return

    done = {"mma": False}

    def producer():
        for j in range(n_max):
            st, ph = j % stages, (j // stages) & 1
            yield from wait(lambda: k_empty[st].passed_fresh(ph ^ 1))
            K[st].write(j)
            k_full[st].arrive()
            yield from wait(lambda: v_empty[st].passed_fresh(ph ^ 1))
            V[st].write(j)
            v_full[st].arrive()
            yield True

    def issue_s(w, j):
        st = j % stages
        K[st].begin_read(j)
        S[w].write(j)
        commit([s_full[w]], [K[st]])

    def mma():
        if n_max > 0:
            yield from wait(lambda: k_full[0].passed(0))
            for w in range(2):
                if n_kv[w] > 0:
                    issue_s(w, 0)
            commit([k_empty[0]])
        for j in range(n_max):
            st, st1 = j % stages, (j + 1) % stages
            ph, ph1 = (j // stages) & 1, ((j + 1) // stages) & 1
            yield from wait(lambda: v_full[st].passed(ph))
            if j + 1 < n_max:
                yield from wait(lambda: k_full[st1].passed(ph1))
            for w in range(2):
                if j >= n_kv[w]:
                    continue
                yield from wait(lambda: p_ready[w].passed(j & 1))
                P[w].begin_read(j)
                V[st].begin_read(j)
                O[w].write(j)
                commit([o_full[w]], [P[w], V[st]])
                if j + 1 < n_kv[w]:
                    issue_s(w, j + 1)
            commit([v_empty[st]])
            if j + 1 < n_max:
                commit([k_empty[st1]])
            yield True
        done["mma"] = True

    def softmax(w):
        for j in range(n_kv[w]):
            yield from wait(lambda: s_full[w].passed(j & 1))
            S[w].begin_read(j)
            yield True                                   # pass A
            if j > 0:
                yield from wait(lambda: o_full[w].passed((j - 1) & 1))
                O[w].begin_read(j - 1)
                O[w].end_read()
            yield True                                   # pass B
            P[w].write(j)
            S[w].end_read()
            for _ in range(4):
                p_ready[w].arrive()
            yield True
        if n_kv[w] > 0:
            yield from wait(lambda: o_full[w].passed((n_kv[w] - 1) & 1))
            O[w].begin_read(n_kv[w] - 1)
            O[w].end_read()
            P[w].write(-2)                               # staging for the output tile

    return [producer(), mma(), mma_engine(), softmax(0), softmax(1)]


@pytest.mark.parametrize("n_kv", [(1, 0), (1, 2), (2, 3), (4, 5), (8, 8), (3, 3), (7, 8)])
def test_forward_protocol(n_kv):
    for seed in range(40):
        run(forward_roles(list(n_kv)), seed)


# ------------------------------------------------------------------------------------------------
# backward: mirrors attn_bwd_sm100.cu
# ------------------------------------------------------------------------------------------------
def backward_roles(n_iter, q_stages=2):
    q_full = [MBar(1) for _ in range(q_stages)]
    q_empty = [MBar(1) for _ in range(q_stages)]
    s_full = MBar(1)
    p_ready = [MBar(4), MBar(4)]
    pds_free = [MBar(1), MBar(1)]
    dq_full, dq_free, dkv_full = MBar(1), MBar(4), MBar(1)
    Q = [Resource(f"Q/dO{s}") for s in range(q_stages)]
    SDP = Resource("S,dP")
    PDS = [Resource("P,dS[0]"), Resource("P,dS[1]")]
    DQ = Resource("dQ tile")
    pending = []
    done = {"mma": False}

    def commit(bars, releases=()):
        pending.append((list(bars), list(releases)))

    def mma_engine():
        while True:
            if pending:
                bars, rel = pending.pop(0)
                for r in rel:
                    r.end_read()
                for b in bars:
                    b.arrive()
                yield True
            else:
                yield False
                if done["mma"] and not pending:
                    return

    def producer():
        for it in range(n_iter):
            s, ph = it % q_stages, (it // q_stages) & 1
            yield from wait(lambda: q_empty[s].passed_fresh(ph ^ 1))
            Q[s].write(it)
            q_full[s].arrive()
            yield True

    def mma():
        if n_iter > 0:
            yield from wait(lambda: q_full[0].passed(0))
            Q[0].begin_read(0)
            SDP.write(0)
            commit([s_full], [Q[0]])
        for it in range(n_iter):
            s, u = it % q_stages, it & 1
            yield from wait(lambda: p_ready[u].passed((it >> 1) & 1))
            if it + 1 < n_iter:
                s1 = (it + 1) % q_stages
                yield from wait(lambda: q_full[s1].passed(((it + 1) // q_stages) & 1))
                Q[s1].begin_read(it + 1)
                SDP.write(it + 1)
                commit([s_full], [Q[s1]])
            if it > 0:
                yield from wait(lambda: dq_free.passed((it - 1) & 1))
            PDS[u].begin_read(it)
            Q[s].begin_read(it)
            DQ.write(it)
            bars = [dq_full, pds_free[u], q_empty[s]] + ([dkv_full] if it == n_iter - 1 else [])
            commit(bars, [PDS[u], Q[s]])
            yield True
        done["mma"] = True

    def softmax():
        for it in range(n_iter):
            u = it & 1
            if it >= 2:
                yield from wait(lambda: pds_free[u].passed(((it - 2) >> 1) & 1))
            yield from wait(lambda: s_full.passed(it & 1))
            SDP.begin_read(it)
            yield True
            PDS[u].write(it)
            SDP.end_read()
            for _ in range(4):
                p_ready[u].arrive()
            yield True
        if n_iter > 0:
            yield from wait(lambda: dkv_full.passed(0))
            for s in range(q_stages):
                Q[s].write(-2)                           # staging for dK / dV

    def drain():
        for it in range(n_iter):
            yield from wait(lambda: dq_full.passed(it & 1))
            DQ.begin_read(it)
            yield True
            DQ.end_read()
            for _ in range(4):
                dq_free.arrive()
            yield True

    return [producer(), mma(), mma_engine(), softmax(), drain()]


@pytest.mark.parametrize("n_iter", [1, 2, 3, 4, 5, 8])
def test_backward_protocol(n_iter):
    for seed in range(40):
        run(backward_roles(n_iter), seed)


# ------------------------------------------------------------------------------------------------
# GEMM with the split epilogue (gemm_sm100.cuh, TDP_GEMM_EPI=split): two epilogue groups, one per
# TMEM accumulator stage; the MMA warp alternates stages per tile
# ------------------------------------------------------------------------------------------------
def split_epilogue_roles(n_tiles):
    tmem_full = [MBar(1), MBar(1)]
    tmem_empty = [MBar(4), MBar(4)]          # four warps of the owning group arrive
    ACC = [Resource("acc0"), Resource("acc1")]
    pending = []
    done = {"mma": False}

    def mma_engine():
        while True:
            if pending:
                for b in pending.pop(0):
                    b.arrive()
                yield True
            else:
                yield False
                if done["mma"] and not pending:
                    return

    def mma():
        acc, acc_phase = 0, 0
        for t in range(n_tiles):
            yield from wait(lambda: tmem_empty[acc].passed_fresh(acc_phase ^ 1))
            ACC[acc].write(t)
            pending.append([tmem_full[acc]])
            yield True
            acc += 1
            if acc == 2:
                acc, acc_phase = 0, acc_phase ^ 1
        done["mma"] = True

    def group(g):
        for t in range(n_tiles):
            if (t & 1) != g:
                continue
            yield from wait(lambda: tmem_full[g].passed((t >> 1) & 1))
            ACC[g].begin_read(t)
            yield True                                   # four sub-tiles of TMEM reads
            yield True
            ACC[g].end_read()
            for _ in range(4):
                tmem_empty[g].arrive()
            yield True                                   # stores drain after the stage is freed

    return [mma(), mma_engine(), group(0), group(1)]


@pytest.mark.parametrize("n_tiles", [1, 2, 3, 7, 10, 11])
def test_gemm_split_epilogue_protocol(n_tiles):
    for seed in range(40):
        run(split_epilogue_roles(n_tiles), seed)
