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
# forward: mirrors attn_fwd_sm100.cu (persistent CTA walking several work items; double-buffered
# Q; K / V rings that run across items; S released to the tensor core as soon as it is in
# registers; the MMA thread serves the two groups' events in arrival order with non-blocking tests)
# ------------------------------------------------------------------------------------------------
class Machine:
    """Barriers, buffers and the in-order tensor-core completion queue shared by the roles."""

    def __init__(self):
        self.pending = []
        self.mma_done = False

    def commit(self, bars, releases=()):
        # tcgen05.commit: arrives once every MMA issued so far has completed; MMAs complete lazily
        # (in issue order) to expose missing waits
        self.pending.append((list(bars), list(releases)))

    def engine(self):
        while True:
            if self.pending:
                bars, rel = self.pending.pop(0)
                for r in rel:
                    r.end_read()
                for b in bars:
                    b.arrive()
                yield True
            else:
                yield False
                if self.mma_done and not self.pending:
                    return


def forward_roles(items, stages=3, rescale_prob=0.3, seed=0):
    """items: list of (n_kv0, n_kv1) per work item of this CTA."""
    rng = random.Random(1000 + seed)
    M = Machine()
    q_full, q_empty = [MBar(1), MBar(1)], [MBar(1), MBar(1)]
    k_full = [MBar(1) for _ in range(stages)]
    k_empty = [MBar(1) for _ in range(stages)]
    v_full = [MBar(1) for _ in range(stages)]
    v_empty = [MBar(1) for _ in range(stages)]
    s_full, s_free = [MBar(1), MBar(1)], [MBar(4), MBar(4)]
    p_ready, o_full = [MBar(4), MBar(4)], [MBar(1), MBar(1)]
    Q = [Resource("Q buf0"), Resource("Q buf1")]
    K = [Resource(f"K{s}") for s in range(stages)]
    V = [Resource(f"V{s}") for s in range(stages)]
    S = [Resource("S0 (TMEM)"), Resource("S1 (TMEM)")]
    P = [Resource("P0 (smem)"), Resource("P1 (smem)")]
    O = [Resource("O0 (TMEM)"), Resource("O1 (TMEM)")]

    def producer():
        g = 0
        for r, n_kv in enumerate(items):
            qb = r & 1
            if r >= 2:
                yield from wait(lambda: q_empty[qb].passed(((r >> 1) - 1) & 1))
            Q[qb].write(r)
            q_full[qb].arrive()
            for _ in range(max(n_kv)):
                st, ph = g % stages, (g // stages) & 1
                yield from wait(lambda: k_empty[st].passed_fresh(ph ^ 1))
                K[st].write(g)
                k_full[st].arrive()
                yield from wait(lambda: v_empty[st].passed_fresh(ph ^ 1))
                V[st].write(g)
                v_full[st].arrive()
                g += 1
                yield True

    def mma():
        g = 0
        a_cnt, b_cnt = [0, 0], [0, 0]
        for r, n_kv in enumerate(items):
            n_max, qb = max(n_kv), r & 1
            a_loc, b_loc, s_iss = [0, 0], [0, 0], [0, 0]
            rel = {"rk": 0, "rv": 0, "q": False}

            def issue_s(w, j):
                st = (g + j) % stages
                K[st].begin_read(g + j)
                Q[qb].begin_read(r)
                S[w].write((r, j))               # asserts that nobody still reads S_w
                M.commit([s_full[w]], [K[st], Q[qb]])
                s_iss[w] = j + 1

            def release():
                while rel["rk"] < n_max and all(s_iss[w] > rel["rk"] or rel["rk"] >= n_kv[w] for w in (0, 1)):
                    M.commit([k_empty[(g + rel["rk"]) % stages]])
                    rel["rk"] += 1
                while rel["rv"] < n_max and all(b_loc[w] > rel["rv"] or rel["rv"] >= n_kv[w] for w in (0, 1)):
                    M.commit([v_empty[(g + rel["rv"]) % stages]])
                    rel["rv"] += 1
                if not rel["q"] and all(s_iss[w] >= n_kv[w] for w in (0, 1)):
                    M.commit([q_empty[qb]])
                    rel["q"] = True

            yield from wait(lambda: q_full[qb].passed((r >> 1) & 1))
            yield from wait(lambda: k_full[g % stages].passed((g // stages) & 1))
            for w in (0, 1):
                if n_kv[w] > 0:
                    issue_s(w, 0)
            release()
            while any(a_loc[w] < n_kv[w] or b_loc[w] < n_kv[w] for w in (0, 1)):
                progressed = False
                for w in (0, 1):
                    if a_loc[w] < n_kv[w] and s_free[w].passed(a_cnt[w] & 1):
                        j = a_loc[w]
                        gj = g + j + 1
                        if j + 1 >= n_kv[w]:
                            a_cnt[w] += 1; a_loc[w] += 1; progressed = True
                        elif k_full[gj % stages].passed((gj // stages) & 1):
                            a_cnt[w] += 1; a_loc[w] += 1; progressed = True
                            issue_s(w, j + 1)
                            release()
                    if b_loc[w] < n_kv[w] and p_ready[w].passed(b_cnt[w] & 1):
                        j = b_loc[w]
                        gj = g + j
                        if v_full[gj % stages].passed((gj // stages) & 1):
                            b_cnt[w] += 1; progressed = True
                            P[w].begin_read((r, j))
                            V[gj % stages].begin_read(gj)
                            O[w].write((r, j))       # asserts that nobody still reads O_w
                            M.commit([o_full[w]], [P[w], V[gj % stages]])
                            b_loc[w] = j + 1
                            release()
                yield progressed
            g += n_max
        M.mma_done = True

    def softmax(w):
        cnt_s = cnt_o = 0
        store_pending = False
        for r, n_kv in enumerate(items):
            n_mine = n_kv[w]
            for j in range(n_mine):
                yield from wait(lambda: s_full[w].passed(cnt_s & 1))
                cnt_s += 1
                S[w].begin_read((r, j))
                yield True                               # tcgen05.ld of the whole row
                S[w].end_read()
                for _ in range(4):
                    s_free[w].arrive()                   # S_w lives in registers now
                pv_done = j == 0
                if j > 0 and rng.random() < rescale_prob:   # rare path: rescale O in TMEM
                    yield from wait(lambda: o_full[w].passed(cnt_o & 1))
                    cnt_o += 1
                    pv_done = True
                    O[w].begin_read((r, j - 1))
                    yield True
                    O[w].end_read()
                yield True                               # exponentials
                if not pv_done:
                    yield from wait(lambda: o_full[w].passed(cnt_o & 1))
                    cnt_o += 1
                if store_pending:                        # output tile of the previous item left
                    P[w].end_read()
                    store_pending = False
                P[w].write((r, j))                       # asserts that P.V(j-1) finished reading
                for _ in range(4):
                    p_ready[w].arrive()
                yield True
            if n_mine > 0:
                yield from wait(lambda: o_full[w].passed(cnt_o & 1))
                cnt_o += 1
                O[w].begin_read((r, n_mine - 1))
                yield True
                O[w].end_read()
                P[w].write(("out", r))                   # staging for the output tile
                P[w].begin_read(("out", r))              # ... which the TMA store now reads
                store_pending = True
        if store_pending:
            P[w].end_read()

    return [producer(), mma(), M.engine(), softmax(0), softmax(1)]


@pytest.mark.parametrize("items", [
    [(1, 0)], [(1, 2)], [(2, 3), (1, 2)], [(4, 5), (3, 4), (1, 2)], [(8, 8), (8, 8)],
    [(7, 8), (5, 6), (3, 4), (1, 2), (1, 2)], [(3, 3), (1, 0), (2, 2)], [(1, 2)] * 6])
def test_forward_protocol(items):
    for seed in range(40):
        run(forward_roles(list(items), seed=seed), seed)


def test_forward_protocol_detects_a_missing_wait():
    """The model must be able to fail: drop the wait for P.V(j-1) before overwriting the P tile."""
    import types
    src = forward_roles.__code__
    bad = forward_roles.__globals__.copy()
    import inspect
    text = inspect.getsource(forward_roles).replace(
        "                if not pv_done:\n"
        "                    yield from wait(lambda: o_full[w].passed(cnt_o & 1))\n"
        "                    cnt_o += 1\n", "                cnt_o += 0 if pv_done else 1\n")
    assert text != inspect.getsource(forward_roles)
    exec(text.replace("def forward_roles", "def broken_roles"), bad)
    with pytest.raises(AssertionError):
        for seed in range(60):
            run(bad["broken_roles"]([(6, 6), (6, 6)], seed=seed), seed)


# ------------------------------------------------------------------------------------------------
# backward dQ: mirrors attn_bwd_dq_sm100.cu (same machinery; S' / dP' pairs, dQ accumulated in TMEM,
# the dS tile doubles as the landing zone of O and as the staging tile of the dQ store)
# ------------------------------------------------------------------------------------------------
def dq_roles(items, stages=4, seed=0):
    M = Machine()
    q_full, q_empty = MBar(1), MBar(1)
    k_full = [MBar(1) for _ in range(stages)]
    k_empty = [MBar(1) for _ in range(stages)]
    v_full = [MBar(1) for _ in range(stages)]
    v_empty = [MBar(1) for _ in range(stages)]
    s_full, s_free = [MBar(1), MBar(1)], [MBar(4), MBar(4)]
    p_ready, dq_full, stage_free = [MBar(4), MBar(4)], [MBar(1), MBar(1)], [MBar(1), MBar(1)]
    QDO = Resource("Q, dO tiles")
    K = [Resource(f"K'{s}") for s in range(stages)]
    V = [Resource(f"V'{s}") for s in range(stages)]
    SDP = [Resource("S'0, dP'0 (TMEM)"), Resource("S'1, dP'1 (TMEM)")]
    DS = [Resource("dS0 / O0 / dQ0 staging"), Resource("dS1 / O1 / dQ1 staging")]

    def producer():
        g = 0
        n_store = [0, 0]
        for r, n_sub in enumerate(items):
            if r > 0:
                yield from wait(lambda: q_empty.passed((r - 1) & 1))
            n_q = 2 if n_sub[1] > 0 else 1
            for w in range(n_q):
                if n_store[w] > 0:
                    yield from wait(lambda: stage_free[w].passed((n_store[w] - 1) & 1))
            QDO.write(r)
            for w in range(n_q):
                DS[w].write(("O", r))                     # O_w lands in the dS tile
                n_store[w] += 1
            q_full.arrive()
            for _ in range(max(n_sub)):
                st, ph = g % stages, (g // stages) & 1
                yield from wait(lambda: k_empty[st].passed_fresh(ph ^ 1))
                K[st].write(g); k_full[st].arrive()
                yield from wait(lambda: v_empty[st].passed_fresh(ph ^ 1))
                V[st].write(g); v_full[st].arrive()
                g += 1
                yield True

    def mma():
        g = 0
        a_cnt, b_cnt = [0, 0], [0, 0]
        for r, n_sub in enumerate(items):
            n_max = max(n_sub)
            a_loc, b_loc, s_iss = [0, 0], [0, 0], [0, 0]
            rel = {"rk": 0, "rv": 0, "q": False}

            def issue_sdp(w, j):
                st = (g + j) % stages
                K[st].begin_read(g + j); V[st].begin_read(g + j); QDO.begin_read(r)
                SDP[w].write((r, j))
                M.commit([s_full[w]], [K[st], V[st], QDO])
                s_iss[w] = j + 1

            def release():
                while rel["rk"] < n_max and all(b_loc[w] > rel["rk"] or rel["rk"] >= n_sub[w] for w in (0, 1)):
                    M.commit([k_empty[(g + rel["rk"]) % stages]]); rel["rk"] += 1
                while rel["rv"] < n_max and all(s_iss[w] > rel["rv"] or rel["rv"] >= n_sub[w] for w in (0, 1)):
                    M.commit([v_empty[(g + rel["rv"]) % stages]]); rel["rv"] += 1
                if not rel["q"] and all(s_iss[w] >= n_sub[w] for w in (0, 1)):
                    M.commit([q_empty]); rel["q"] = True

            yield from wait(lambda: q_full.passed(r & 1))
            yield from wait(lambda: k_full[g % stages].passed((g // stages) & 1))
            yield from wait(lambda: v_full[g % stages].passed((g // stages) & 1))
            for w in (0, 1):
                if n_sub[w] > 0:
                    issue_sdp(w, 0)
            release()
            while any(a_loc[w] < n_sub[w] or b_loc[w] < n_sub[w] for w in (0, 1)):
                progressed = False
                for w in (0, 1):
                    if a_loc[w] < n_sub[w] and s_free[w].passed(a_cnt[w] & 1):
                        j, gj = a_loc[w], g + a_loc[w] + 1
                        if j + 1 >= n_sub[w]:
                            a_cnt[w] += 1; a_loc[w] += 1; progressed = True
                        elif k_full[gj % stages].passed((gj // stages) & 1) \
                                and v_full[gj % stages].passed((gj // stages) & 1):
                            a_cnt[w] += 1; a_loc[w] += 1; progressed = True
                            issue_sdp(w, j + 1)
                            release()
                    if b_loc[w] < n_sub[w] and p_ready[w].passed(b_cnt[w] & 1):
                        b_cnt[w] += 1; progressed = True
                        j = b_loc[w]
                        st = (g + j) % stages
                        assert K[st].version == g + j, "K'(j) was released before dQ(j) read it"
                        DS[w].begin_read((r, j)); K[st].begin_read(g + j)
                        M.commit([dq_full[w]], [DS[w], K[st]])
                        b_loc[w] = j + 1
                        release()
                yield progressed
            g += n_max
        M.mma_done = True

    def softmax(w):
        cnt_s = cnt_dq = 0
        for r, n_sub in enumerate(items):
            n_mine = n_sub[w]
            if n_mine == 0:
                continue
            yield from wait(lambda: q_full.passed(r & 1))
            DS[w].begin_read(("O", r)); QDO.begin_read(r)  # delta = rowsum(dO . O) from smem
            yield True
            DS[w].end_read(); QDO.end_read()
            for j in range(n_mine):
                yield from wait(lambda: s_full[w].passed(cnt_s & 1))
                cnt_s += 1
                SDP[w].begin_read((r, j))
                yield True
                SDP[w].end_read()
                for _ in range(4):
                    s_free[w].arrive()
                yield True
                if j > 0:
                    yield from wait(lambda: dq_full[w].passed(cnt_dq & 1))
                    cnt_dq += 1
                DS[w].write((r, j))                      # asserts that dQ(j-1) finished reading dS
                for _ in range(4):
                    p_ready[w].arrive()
                yield True
            yield from wait(lambda: dq_full[w].passed(cnt_dq & 1))
            cnt_dq += 1
            DS[w].write(("dQ", r))                       # staging of the output tile
            DS[w].begin_read(("dQ", r))                  # TMA store
            yield True
            DS[w].end_read()                             # cp.async.bulk.wait_group.read
            stage_free[w].arrive()

    return [producer(), mma(), M.engine(), softmax(0), softmax(1)]


@pytest.mark.parametrize("items", [
    [(2, 0)], [(2, 4)], [(4, 6), (2, 4)], [(16, 16), (16, 16)], [(14, 16), (10, 12), (6, 8), (2, 4)],
    [(2, 4)] * 5])
def test_backward_dq_protocol(items):
    for seed in range(40):
        run(dq_roles(list(items), seed=seed), seed)


# ------------------------------------------------------------------------------------------------
# backward dK / dV: mirrors attn_bwd_sm100.cu (one CTA per key tile; S / dP handed back to the
# tensor core as soon as both softmax groups hold them in registers; P / dS double buffered)
# ------------------------------------------------------------------------------------------------
def backward_roles(n_iter, q_stages=2):
    M = Machine()
    q_full = [MBar(1) for _ in range(q_stages)]
    q_empty = [MBar(1) for _ in range(q_stages)]
    s_full, sdp_free = MBar(1), MBar(8)
    p_ready = [MBar(8), MBar(8)]
    pds_free = [MBar(1), MBar(1)]
    dkv_full = MBar(1)
    Q = [Resource(f"Q/dO{s}") for s in range(q_stages)]
    SDP = Resource("S,dP (TMEM)")
    # each softmax group owns one key half of the P / dS buffers
    PDS = [[Resource(f"P,dS[{u}] half {h}") for h in range(2)] for u in range(2)]

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
            M.commit([s_full], [Q[0]])
        for it in range(n_iter):
            s, u = it % q_stages, it & 1
            yield from wait(lambda: sdp_free.passed(it & 1))      # S_i / dP_i are in registers
            if it + 1 < n_iter:
                s1 = (it + 1) % q_stages
                yield from wait(lambda: q_full[s1].passed(((it + 1) // q_stages) & 1))
                Q[s1].begin_read(it + 1)
                SDP.write(it + 1)                                 # asserts nobody still reads S / dP
                M.commit([s_full], [Q[s1]])
            yield from wait(lambda: p_ready[u].passed((it >> 1) & 1))
            for h in range(2):
                PDS[u][h].begin_read(it)
            Q[s].begin_read(it)
            bars = [pds_free[u], q_empty[s]] + ([dkv_full] if it == n_iter - 1 else [])
            M.commit(bars, [PDS[u][0], PDS[u][1], Q[s]])
            yield True
        M.mma_done = True

    def softmax(wg):
        for it in range(n_iter):
            u = it & 1
            yield from wait(lambda: s_full.passed(it & 1))
            SDP.begin_read(it)
            yield True                                            # tcgen05.ld of my column half
            SDP.end_read()
            for _ in range(4):
                sdp_free.arrive()
            yield True                                            # exponentials, dS
            if it >= 2:
                yield from wait(lambda: pds_free[u].passed(((it - 2) >> 1) & 1))
            PDS[u][wg].write(it)                                  # asserts dV / dK(it-2) finished
            for _ in range(4):
                p_ready[u].arrive()
            yield True
        if n_iter > 0:
            yield from wait(lambda: dkv_full.passed(0))
            Q[0].write(-2 - wg) if wg == 0 else None              # staging for dV (group 0) ...
            yield True

    return [producer(), mma(), M.engine(), softmax(0), softmax(1)]


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


# ------------------------------------------------------------------------------------------------
# GEMM with the ring-buffered epilogue (gemm_sm100_2cta.cuh, kEpiBufs = 6): NB staging buffers
# indexed by a running sub-tile number q that keeps counting across tiles.  One thread (the issuer)
# commits one TMA-store group per sub-tile and throttles buffer reuse with
# cp.async.bulk.wait_group.read; 8 epilogue warps meet at ONE named barrier per sub-tile.
#   plain : buffer q % NB          - issuer waits "<= NB-2 groups unread" before the closing barrier
#   aux   : buffer pair q % (NB/2) - two outputs per sub-tile, "<= NB/2-2"
#   in    : a row-wise input of sub-tile q+P (P = NB/2) is TMA-loaded into buffer (q+P) % NB at the
#           start of sub-tile q after "<= NB-P-1 groups unread"; the buffer's own mbarrier hands it
#           to the warps, which overwrite it with the output in place
# The model runs the store engine (reads buffers out, in commit order, lazily) and the load engine
# asynchronously and checks that no buffer is overwritten before its store was read out and that
# every warp reads the input of ITS sub-tile.  `slack` shifts the wait constants: +1 must fail.
# ------------------------------------------------------------------------------------------------
def ring_epilogue_roles(mode, subs_per_tile, NB=6, n_warps=4, slack=0, seed=0, store_rate=0.5):
    rng = random.Random(seed)
    P = NB // 2
    n_q = sum(subs_per_tile)
    groups = []                      # committed store groups: dict(q, bufs, read)
    buf_store_pending = [None] * NB  # q whose (unread) store still owns the buffer
    buf_content = [None] * NB        # ("in", q) | ("out", q)
    in_full = [MBar(1) for _ in range(NB)]
    loads = []                       # pending input loads (buffer, q)
    state = {"done": False, "arrived": 0, "generation": 0}

    def unread():
        return sum(1 for g in groups if not g["read"])

    def store_engine():
        while True:
            nxt = next((g for g in groups if not g["read"]), None)
            if nxt is not None and rng.random() < store_rate:
                for b in nxt["bufs"]:
                    assert buf_content[b] == ("out", nxt["q"]), \
                        f"store of sub-tile {nxt['q']} read buffer {b} holding {buf_content[b]}"
                    buf_store_pending[b] = None
                nxt["read"] = True
                yield True
            elif nxt is not None:
                yield True                       # the engine is busy (time passes), not blocked
            else:
                yield False
                if state["done"]:
                    return

    def load_engine():
        while True:
            if loads and rng.random() < 0.5:
                b, q = loads.pop(0)
                assert buf_store_pending[b] is None, \
                    f"input of sub-tile {q} loaded into buffer {b} before store {buf_store_pending[b]} was read"
                buf_content[b] = ("in", q)
                in_full[b].arrive()
                yield True
            elif loads:
                yield True
            else:
                yield False
                if state["done"]:
                    return

    def barrier(my_gen):
        # bar.sync over the epilogue warps: generation counter
        state["arrived"] += 1
        if state["arrived"] == n_warps:
            state["arrived"] = 0
            state["generation"] += 1
        yield True
        yield from wait(lambda: state["generation"] > my_gen)

    def bufs_of(q):
        if mode == "aux":
            pair = 2 * (q % (NB // 2))
            return [pair, pair + 1]
        return [q % NB]

    def warp(w):
        issuer = w == 0
        pf_q = 0
        gen = 0
        in_phase = [0] * NB
        if issuer and mode == "in":
            for _ in range(P):
                if pf_q < n_q:
                    loads.append((pf_q % NB, pf_q))
                    pf_q += 1
        for q in range(n_q):
            bufs = bufs_of(q)
            if issuer and mode == "in":
                yield from wait(lambda: unread() <= NB - P - 1 + slack)
                if pf_q < n_q:
                    loads.append((pf_q % NB, pf_q))
                    pf_q += 1
            if mode == "in":
                b = bufs[0]
                yield from wait(lambda: in_full[b].passed(in_phase[b]))
                in_phase[b] ^= 1
                assert buf_content[b] == ("in", q), f"warp {w} sub-tile {q}: buffer holds {buf_content[b]}"
                yield True
            for b in bufs:                                   # write the output rows of this warp
                assert buf_store_pending[b] is None, \
                    f"warp {w} wrote buffer {b} for sub-tile {q} before store {buf_store_pending[b]} was read"
            yield True
            if issuer and mode != "in":
                limit = (NB // 2 - 2 if mode == "aux" else NB - 2) + slack
                yield from wait(lambda: unread() <= limit)
            yield from barrier(gen)
            gen += 1
            if issuer:
                for b in bufs:
                    buf_content[b] = ("out", q)
                    buf_store_pending[b] = q
                groups.append(dict(q=q, bufs=bufs, read=False))
                yield True
        if issuer:
            yield from wait(lambda: unread() == 0)
            state["done"] = True

    return [warp(w) for w in range(n_warps)] + [store_engine(), load_engine()]


@pytest.mark.parametrize("mode", ["plain", "aux", "in"])
@pytest.mark.parametrize("subs", [[4], [4, 4, 4], [4, 2, 4, 1, 3], [1] * 9])
def test_gemm_ring_epilogue_protocol(mode, subs):
    for seed in range(25):
        run(ring_epilogue_roles(mode, subs, seed=seed, store_rate=(0.5, 0.1, 0.02)[seed % 3]), seed)


@pytest.mark.parametrize("mode", ["plain", "aux", "in"])
def test_gemm_ring_epilogue_wait_constants_are_tight(mode):
    """One more unread store group than the kernel allows and a buffer is overwritten too early."""
    failures = 0
    for seed in range(60):
        try:
            # (a slow store engine: the reads lag far enough behind for the window to matter)
            run(ring_epilogue_roles(mode, [4, 4, 4, 4], slack=1, seed=seed, store_rate=0.02), seed)
        except AssertionError:
            failures += 1
    assert failures > 0
