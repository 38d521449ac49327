"""Discrete model of the mbarrier / buffer protocols of the backward kernels (one CTA), in the spirit of tools/paged_ring_sim.py.

Actors are Python generators scheduled at random: the MMA / TMA issuer thread, the compute warps, (fused2) the drain warps
and the reducer.  Hardware is modelled as two asynchronous agents: an IN-ORDER tensor pipe (every tcgen05.mma / commit
takes a random time; a commit arrives on its barrier when everything issued before it has executed) and a TMA unit whose
loads land after random, independent delays.  Barriers have real phase/parity semantics: `test(P)` passes when the
current phase parity differs from P, so a waiter that is lapped (two completions before it looks) blocks forever -- a
deadlock the run reports -- and a waiter that looks too early on a re-used barrier would pass on the wrong phase.
Every buffer (shared-memory stage or TMEM region) carries a tag of what it holds and how many readers still need it:
a read asserts the tag, a write asserts that the previous tenant has been fully consumed.  A protocol is accepted when
many random schedules finish without a violated assertion or a deadlock.

Models (each mirrors the barrier code of the kernel it names; the kernels' own comments carry the same rules):
  dkvt(two_s)          csrc/kernels/attn_bwd_sm100.cu  bwd_dkv_t_body: one or two S^T buffers (head_dim <= 64 has two)
  dq2()                csrc/kernels/attn_bwd_sm100.cu  bwd_dq2_body: three issuer threads with separate commit scopes
  fused2(wait_dqfree)  csrc/kernels/attn_bwd_fused2_sm100.cu: 64-query half steps, separate drain warps and reducer;
                       wait_dqfree=False is the version that hung under ncu (the issuer overwrote dQ^T(h-1) before the
                       drain warps had read it and lapped them on bar_dq)
usage: python tools/bwd_protocol_sim.py            (runs both models over a few seeds and prints the verdicts)
"""
import random


class Violation(Exception):
    pass


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):                      # mbarrier.try_wait.parity / test_wait.parity
        return (self.phase & 1) != (parity & 1)


class Buf:
    """A tile-sized buffer: `tag` = what it holds, `left` = reads that must still happen before it may be overwritten."""

    def __init__(self, name):
        self.name, self.tag, self.left = name, None, 0

    def write(self, tag, readers):
        if self.left != 0:
            raise Violation(f"{self.name}: {tag} overwrites {self.tag} with {self.left} read(s) outstanding")
        self.tag, self.left = tag, readers

    def retag(self, old, new, readers):          # in-place transformation by its last reader (S^T -> P^T, dP^T -> dS^T)
        if self.tag != old or self.left != 0:
            raise Violation(f"{self.name}: expected {old} fully read before {new}, holds {self.tag} ({self.left} left)")
        self.tag, self.left = new, readers

    def read(self, tag):
        if self.tag != tag:
            raise Violation(f"{self.name}: read of {tag} finds {self.tag}")
        if self.left <= 0:
            raise Violation(f"{self.name}: extra read of {tag}")
        self.left -= 1


class Machine:
    def __init__(self, seed, max_delay=40):
        self.rng = random.Random(seed)
        self.max_delay = max_delay
        self.pipe = []                           # in-order tensor pipe: [remaining, fn]
        self.pipes = {}                          # further issuing threads: ops are ordered (and commits count) per thread only
        self.tma = []                            # independent loads: [remaining, fn]
        self.actors = []

    def mma(self, fn, thread=None):              # tcgen05.mma group: executes in issue order (of its issuing thread)
        (self.pipe if thread is None else self.pipes.setdefault(thread, [])).append([self.rng.randint(1, self.max_delay), fn])

    def commit(self, bar, thread=None):          # tcgen05.commit: arrives once everything this thread issued before has executed
        (self.pipe if thread is None else self.pipes.setdefault(thread, [])).append([1, bar.arrive])

    def load(self, fn):
        self.tma.append([self.rng.randint(1, 3 * self.max_delay), fn])

    def tick(self):
        for q in [self.pipe] + list(self.pipes.values()):
            if q:
                q[0][0] -= 1
                if q[0][0] <= 0:
                    q.pop(0)[1]()
        for t in list(self.tma):
            t[0] -= 1
            if t[0] <= 0:
                self.tma.remove(t)
                t[1]()

    def run(self, max_ticks=4_000_000):
        live = list(self.actors)
        idle = 0
        for _ in range(max_ticks):
            if not live:
                return True, "ok"
            self.tick()
            progressed = False
            self.rng.shuffle(live)
            for a in list(live):
                if self.rng.random() < 0.35:     # this actor is not scheduled in this tick
                    continue
                try:
                    r = next(a)
                except StopIteration:
                    live.remove(a)
                    progressed = True
                    continue
                if r != "wait":
                    progressed = True
            idle = 0 if (progressed or self.pipe or self.tma or any(self.pipes.values())) else idle + 1
            if idle > 2000:
                return False, "deadlock: every actor waits and nothing is in flight"
        return False, "did not finish"


def wait(bar, parity):
    while not bar.test(parity):
        yield "wait"


# ------------------------------------------------------------------------------------------------ dK/dV kernel (transposed)
def dkvt(nsteps=24, two_s=True, W=4, NQ=3, NDO=2, seed=0, per_buffer_bar=True, per_buffer_pbar=True, slow_warp=0):
    """bwd_dkv_t_body.  W compute warps (the kernel has 16).  per_buffer_bar=False: both S^T buffers announce on ONE barrier
    (parity i & 1) -- what the kernel avoids; per_buffer_pbar=False: the "P^T published" barriers are shared by both buffers --
    the first two-buffer version: with S^T(i+1) available early a fast warp arrives for step i+1 before a slow warp has
    arrived for step i, the barrier completes on mixed arrivals and dV(i) reads a P^T that is not there yet; slow_warp > 0
    delays warp 0 before every S^T wait.  Returns (ok, detail)."""
    m = Machine(seed)
    qfull, qfree = [Bar(1) for _ in range(NQ)], [Bar(1) for _ in range(NQ)]
    dofull, dofree = [Bar(1) for _ in range(NDO)], [Bar(1) for _ in range(NDO)]
    bar_s, bar_s1, bar_dp, bar_ds = Bar(1), Bar(1), Bar(1), Bar(W)
    bar_p, bar_pb = [Bar(W), Bar(W)], [Bar(W), Bar(W)]
    two_p = two_s and per_buffer_pbar
    pbar = (lambda i: bar_p[i & 1] if two_p else bar_p[0])
    pbbar = (lambda i: bar_pb[i & 1] if two_p else bar_pb[0])
    ppar = (lambda i: (i >> 1) & 1) if two_p else (lambda i: i & 1)
    Q, DO = [Buf(f"Q[{i}]") for i in range(NQ)], [Buf(f"dO[{i}]") for i in range(NDO)]
    S = [Buf("S^T/P^T buffer 0"), Buf("S^T/P^T buffer 1")]
    DP = Buf("dP^T/dS^T")
    sbuf = (lambda i: i & 1) if two_s else (lambda i: 0)
    sbar = (lambda i: bar_s1 if (two_s and per_buffer_bar and i & 1) else bar_s)
    spar = (lambda i: (i >> 1) & 1) if (two_s and per_buffer_bar) else (lambda i: i & 1)

    def issuer():
        st = {"ql": 0, "dl": 0}

        def pump():
            ql = st["ql"]
            if ql < nsteps and (ql < NQ or qfree[ql % NQ].test(((ql // NQ) - 1) & 1)):
                m.load(lambda i=ql: (Q[i % NQ].write(("Q", i), 2), qfull[i % NQ].arrive()))     # read by S^T(i) and dK(i)
                st["ql"] += 1
            dl = st["dl"]
            if dl < nsteps and (dl < NDO or dofree[dl % NDO].test(((dl // NDO) - 1) & 1)):
                m.load(lambda i=dl: (DO[i % NDO].write(("dO", i), 2), dofull[i % NDO].arrive()))  # dP^T(i) and dV(i)
                st["dl"] += 1

        def pwait(bar, parity):
            while not bar.test(parity):
                pump()
                yield "wait"

        def issue_s(i):
            b = sbuf(i)
            m.mma(lambda: (Q[i % NQ].read(("Q", i)), S[b].write(("S", i), W)))
            m.commit(sbar(i))

        def issue_dp(i):
            m.mma(lambda: (DO[i % NDO].read(("dO", i)), DP.write(("dP", i), W)))
            m.commit(bar_dp)

        for _ in range(NQ):
            pump()
        yield from pwait(qfull[0], 0)
        issue_s(0)
        yield from pwait(dofull[0], 0)
        issue_dp(0)
        for i in range(nsteps):
            if two_s and i + 1 < nsteps:
                yield from pwait(qfull[(i + 1) % NQ], ((i + 1) // NQ) & 1)
                issue_s(i + 1)
            yield from pwait(pbar(i), ppar(i))                  # first halves of P^T(i)
            yield from pwait(pbbar(i), ppar(i))                 # second halves
            m.mma(lambda i=i: (S[sbuf(i)].read(("P", i)), DO[i % NDO].read(("dO", i))))          # dV(i)
            m.commit(dofree[i % NDO])
            if not two_s and i + 1 < nsteps:
                yield from pwait(qfull[(i + 1) % NQ], ((i + 1) // NQ) & 1)
                issue_s(i + 1)                                   # overwrites P^T(i): behind dV(i) in the pipe
            yield from pwait(bar_ds, i & 1)
            m.mma(lambda i=i: (DP.read(("dS", i)), Q[i % NQ].read(("Q", i))))                   # dK(i)
            m.commit(qfree[i % NQ])
            if i + 1 < nsteps:
                yield from pwait(dofull[(i + 1) % NDO], ((i + 1) // NDO) & 1)
                issue_dp(i + 1)                                  # overwrites dS^T(i): behind dK(i) in the pipe

    done_p, done_ds = [0] * nsteps, [0] * nsteps

    def warp(w):
        for i in range(nsteps):
            for _ in range(slow_warp if w == 0 else 0):
                yield
            yield from wait(sbar(i), spar(i))
            S[sbuf(i)].read(("S", i))
            yield
            done_p[i] += 1
            if done_p[i] == W:
                S[sbuf(i)].retag(("S", i), ("P", i), 1)          # consumed by dV(i)
            pbar(i).arrive()
            yield
            pbbar(i).arrive()
            yield from wait(bar_dp, i & 1)
            DP.read(("dP", i))
            yield
            done_ds[i] += 1
            if done_ds[i] == W:
                DP.retag(("dP", i), ("dS", i), 1)                # consumed by dK(i)
            bar_ds.arrive()

    m.actors = [issuer()] + [warp(w) for w in range(W)]
    try:
        return m.run()
    except Violation as e:
        return False, str(e)


# ------------------------------------------------------------------------------------------------ dQ kernel v2
def dq2(n=24, W=4, NK=4, NV=2, seed=0, slow_warp=0):
    """bwd_dq2_body: three issuer threads (S stream + K loads, dP stream + V loads, dQ stream), each with its own commit
    scope, W compute warps (kernel: 16).  S, dP and dS have one buffer each; dQ accumulates."""
    m = Machine(seed)
    kfull, kfree = [Bar(1) for _ in range(NK)], [Bar(1) for _ in range(NK)]
    vfull, vfree = [Bar(1) for _ in range(NV)], [Bar(1) for _ in range(NV)]
    bar_q, bar_do, bar_s, bar_dp, bar_dsfree = Bar(W), Bar(1), Bar(1), Bar(1), Bar(1)
    bar_sfree, bar_dpfree, bar_ds = Bar(W), Bar(W), Bar(W)
    K, V = [Buf(f"K[{i}]") for i in range(NK)], [Buf(f"V[{i}]") for i in range(NV)]
    S, DP, DS = Buf("S"), Buf("dP"), Buf("dS")

    def s_stream():
        st = {"kl": 0}

        def pump():
            kl = st["kl"]
            if kl < n and (kl < NK or kfree[kl % NK].test(((kl // NK) - 1) & 1)):
                m.load(lambda j=kl: (K[j % NK].write(("K", j), 2), kfull[j % NK].arrive()))      # read by S(j) and dQ(j)
                st["kl"] += 1

        def issue_s(j):
            m.mma(lambda: (K[j % NK].read(("K", j)), S.write(("S", j), W)), "s")
            m.commit(bar_s, "s")

        for _ in range(NK):
            pump()
        yield from wait(bar_q, 0)
        yield from wait(kfull[0], 0)
        issue_s(0)
        for j in range(n - 1):
            pump()
            yield from wait(bar_sfree, j & 1)
            while not kfull[(j + 1) % NK].test(((j + 1) // NK) & 1):
                pump()
                yield "wait"
            issue_s(j + 1)
        while st["kl"] < n:
            pump()
            yield "wait"

    def dp_stream():
        st = {"vl": 0}

        def pump():
            vl = st["vl"]
            if vl < n and (vl < NV or vfree[vl % NV].test(((vl // NV) - 1) & 1)):
                m.load(lambda j=vl: (V[j % NV].write(("V", j), 1), vfull[j % NV].arrive()))
                st["vl"] += 1

        def issue_dp(j):
            m.mma(lambda: (V[j % NV].read(("V", j)), DP.write(("dP", j), W)), "dp")
            m.commit(bar_dp, "dp")
            m.commit(vfree[j % NV], "dp")

        m.load(bar_do.arrive)
        for _ in range(NV):
            pump()
        yield from wait(bar_do, 0)
        yield from wait(vfull[0], 0)
        issue_dp(0)
        for j in range(n - 1):
            pump()
            yield from wait(bar_dpfree, j & 1)
            while not vfull[(j + 1) % NV].test(((j + 1) // NV) & 1):
                pump()
                yield "wait"
            issue_dp(j + 1)

    def dq_stream():
        for j in range(n):
            yield from wait(bar_ds, j & 1)
            m.mma(lambda j=j: (DS.read(("dS", j)), K[j % NK].read(("K", j))), "dq")
            m.commit(kfree[j % NK], "dq")
            m.commit(bar_dsfree, "dq")

    done_ds = [0] * n

    def warp(w):
        bar_q.arrive()
        for j in range(n):
            for _ in range(slow_warp if w == 0 else 0):
                yield
            yield from wait(bar_s, j & 1)
            S.read(("S", j))
            bar_sfree.arrive()
            yield
            yield from wait(bar_dp, j & 1)
            DP.read(("dP", j))
            bar_dpfree.arrive()
            yield
            if j > 0:
                yield from wait(bar_dsfree, (j - 1) & 1)
            done_ds[j] += 1
            if done_ds[j] == 1:
                DS.write(("dS", j), 1)
            yield
            bar_ds.arrive()

    m.actors = [s_stream(), dp_stream(), dq_stream()] + [warp(w) for w in range(W)]
    try:
        return m.run()
    except Violation as e:
        return False, str(e)


# ------------------------------------------------------------------------------------------------ fused backward, half steps
def fused2(nsteps=24, wait_dqfree=True, W=4, WD=2, NQ=3, NDO=2, seed=0, drain_delay=0, slow_warp=0, stg_bufs=2):
    """bwd_fused2_body: W P / dS warps (kernel: 16), WD drain warps (kernel: 4), an issuer and a reducer.
    drain_delay > 0 makes the drain warps slow (what ncu's replay passes did)."""
    m = Machine(seed)
    qfull, qfree = [Bar(1) for _ in range(NQ)], [Bar(1) for _ in range(NQ)]
    dofull, dofree = [Bar(1) for _ in range(NDO)], [Bar(1) for _ in range(NDO)]
    bar_s, bar_dp, bar_dp1, bar_dq, bar_dsfree = Bar(1), Bar(1), Bar(1), Bar(1), Bar(1)
    bar_p, bar_ds, bar_dqfree = Bar(W), Bar(W), Bar(WD)
    stgfull, stgfree = [Bar(WD), Bar(WD)], [Bar(1), Bar(1)]
    Q, DO = [Buf(f"Q[{i}]") for i in range(NQ)], [Buf(f"dO[{i}]") for i in range(NDO)]
    S, DP, DQ, DS = Buf("S^T/P^T"), [Buf("dP^T 0"), Buf("dP^T 1")], Buf("dQ^T"), Buf("dS^T tile")
    STG = [Buf("staging 0"), Buf("staging 1")]

    def issuer():
        st = {"ql": 0, "dl": 0}

        def pump():
            ql = st["ql"]
            if ql < nsteps and (ql < NQ or qfree[ql % NQ].test(((ql // NQ) - 1) & 1)):
                m.load(lambda i=ql: (Q[i % NQ].write(("Q", i), 2), qfull[i % NQ].arrive()))
                st["ql"] += 1
            dl = st["dl"]
            if dl < nsteps and (dl < NDO or dofree[dl % NDO].test(((dl // NDO) - 1) & 1)):
                m.load(lambda i=dl: (DO[i % NDO].write(("dO", i), 2), dofull[i % NDO].arrive()))
                st["dl"] += 1

        def pwait(bar, parity):
            while not bar.test(parity):
                pump()
                yield "wait"

        def issue_s(k):
            m.mma(lambda: (Q[k % NQ].read(("Q", k)), S.write(("S", k), W)))
            m.commit(bar_s)

        def issue_dp(k):
            m.mma(lambda: (DO[k % NDO].read(("dO", k)), DP[k & 1].write(("dP", k), W)))
            m.commit(bar_dp1 if k & 1 else bar_dp)

        def issue_dv(k):
            m.mma(lambda: (S.read(("P", k)), DO[k % NDO].read(("dO", k))))
            m.commit(dofree[k % NDO])

        for _ in range(NQ):
            pump()
        yield from pwait(qfull[0], 0)
        issue_s(0)
        yield from pwait(dofull[0], 0)
        issue_dp(0)
        yield from pwait(bar_p, 0)
        issue_dv(0)
        if nsteps > 1:
            yield from pwait(qfull[1 % NQ], 0)
            issue_s(1)
        for h in range(nsteps):
            if h + 1 < nsteps:
                yield from pwait(dofull[(h + 1) % NDO], ((h + 1) // NDO) & 1)
                issue_dp(h + 1)
            yield from pwait(bar_ds, h & 1)
            if wait_dqfree and h > 0:
                yield from pwait(bar_dqfree, (h - 1) & 1)
            m.mma(lambda h=h: (DS.read(("dS", h)), DQ.write(("dQ", h), WD)))                     # dQ^T(h)
            m.commit(bar_dq)
            m.mma(lambda h=h: (DS.read(("dS", h)), Q[h % NQ].read(("Q", h))))                    # dK(h)
            m.commit(qfree[h % NQ])
            m.commit(bar_dsfree)
            if h + 1 < nsteps:
                yield from pwait(bar_p, (h + 1) & 1)
                issue_dv(h + 1)
                if h + 2 < nsteps:
                    yield from pwait(qfull[(h + 2) % NQ], ((h + 2) // NQ) & 1)
                    issue_s(h + 2)

    done_p, done_ds, done_stg = [0] * nsteps, [0] * nsteps, [0] * nsteps

    def p_phase(k):
        yield from wait(bar_s, k & 1)
        S.read(("S", k))
        yield
        done_p[k] += 1
        if done_p[k] == W:
            S.retag(("S", k), ("P", k), 1)
        bar_p.arrive()

    def pds_warp(w):
        yield from p_phase(0)
        for h in range(nsteps):
            for _ in range(slow_warp if w == 0 else 0):
                yield
            yield from wait(bar_dp1 if h & 1 else bar_dp, (h >> 1) & 1)
            DP[h & 1].read(("dP", h))
            yield
            if h > 0:
                yield from wait(bar_dsfree, (h - 1) & 1)
            done_ds[h] += 1
            if done_ds[h] == 1:
                DS.write(("dS", h), 2)                           # read by dQ^T(h) and dK(h); a violation if h-1 is unread
            bar_ds.arrive()
            if h + 1 < nsteps:
                yield from p_phase(h + 1)

    def drain_warp(w):
        for h in range(nsteps):
            yield from wait(bar_dq, h & 1)
            for _ in range(drain_delay):
                yield
            DQ.read(("dQ", h))
            yield
            bar_dqfree.arrive()
            b = (h & 1) if stg_bufs == 2 else 0
            if stg_bufs == 2 and h >= 2:
                yield from wait(stgfree[b], ((h >> 1) - 1) & 1)
            if stg_bufs == 1 and h >= 1:
                yield from wait(stgfree[0], (h - 1) & 1)
            done_stg[h] += 1
            if done_stg[h] == 1:
                STG[b].write(("stg", h), 1)
            yield
            stgfull[b].arrive()

    def reducer():
        for h in range(nsteps):
            b = (h & 1) if stg_bufs == 2 else 0
            yield from wait(stgfull[b], ((h >> 1) & 1) if stg_bufs == 2 else (h & 1))
            STG[b].read(("stg", h))
            yield
            stgfree[b].arrive()

    m.actors = [issuer(), reducer()] + [pds_warp(w) for w in range(W)] + [drain_warp(w) for w in range(WD)]
    try:
        return m.run()
    except Violation as e:
        return False, str(e)


if __name__ == "__main__":
    for name, fn in (("dkvt, one S^T buffer", lambda s: dkvt(two_s=False, seed=s)), ("dkvt, two S^T buffers", lambda s: dkvt(two_s=True, seed=s)),
                     ("dkvt, one S^T buffer, one slow warp", lambda s: dkvt(two_s=False, seed=s, slow_warp=300)),
                     ("dkvt, two S^T buffers, one slow warp", lambda s: dkvt(two_s=True, seed=s, slow_warp=300)),
                     ("dkvt, two buffers, shared P^T barriers, slow warp", lambda s: dkvt(two_s=True, seed=s, per_buffer_pbar=False, slow_warp=300)),
                     ("dkvt, two buffers, shared S^T barrier, slow warp", lambda s: dkvt(two_s=True, seed=s, per_buffer_bar=False, slow_warp=300)),
                     ("dq2 (three issuer threads)", lambda s: dq2(seed=s)), ("dq2, one slow warp", lambda s: dq2(seed=s, slow_warp=300)),
                     ("fused2, one slow P / dS warp", lambda s: fused2(seed=s, slow_warp=300)),
                     ("fused2", lambda s: fused2(seed=s)), ("fused2, slow drain", lambda s: fused2(seed=s, drain_delay=400)),
                     ("fused2 without the bar_dqfree wait, slow drain", lambda s: fused2(wait_dqfree=False, seed=s, drain_delay=400))):
        res = [fn(s) for s in range(8)]
        bad = [d for ok, d in res if not ok]
        print(f"{name:50s}: {len(res) - len(bad)}/{len(res)} schedules clean" + (f"   e.g. {bad[0]}" if bad else ""))
