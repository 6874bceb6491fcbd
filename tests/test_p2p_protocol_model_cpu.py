"""Executable model of the rendezvous protocol of mvae_allreduce_adam_p2p (csrc/dp_p2p.cu), run on CPU threads with
randomised delays: one thread per rank, shared Python lists play the peer-mapped flag / gradient / parameter memory.

It checks the protocol's safety claims for 2..8 ranks and many launches:
  * a rank only reads a peer's gradients while they are FINAL for the same launch (never half-written, never already
    zeroed for the next step);
  * a rank's parameters are only written by peers while it is inside the exchange kernel of that launch (never during
    its forward / backward);
  * nobody dead-locks, and the flag words only ever grow (launch numbers are reused as epochs);
  * the value that ends a wait always EQUALS the waiter's own launch number (the kernel reports anything else as a
    lockstep violation instead of proceeding: error code 2 of dp_p2p.cu).
This guards the LOGIC (barrier 1 / slice work / barrier 2, epoch reuse); memory-ordering questions are the kernel's."""
import random
import threading
import time

import pytest

COMPUTE, IN_KERNEL = 0, 1


class World:
    def __init__(self, n):
        self.n = n
        self.flag1 = [[0] * n for _ in range(n)]      # flag1[dst][src]: "src's gradients are final for launch e"
        self.flag2 = [[0] * n for _ in range(n)]      # flag2[dst][src]: "src's stores into dst's parameters are done"
        self.grad_version = [0] * n                   # launch for which rank r's gradient bucket is final (-1: being written)
        self.param_version = [[0] * n for _ in range(n)]   # param_version[r][slice]
        self.phase = [COMPUTE] * n
        self.kernel_launch = [0] * n
        self.errors = []
        self.lock = threading.Lock()


def rank_thread(w: World, r: int, launches: int, rng: random.Random):
    n = w.n

    def jitter(scale=1.0):
        if rng.random() < 0.3:
            time.sleep(rng.random() * 0.0004 * scale)

    def wait(flags, e):
        t0 = time.time()
        for p in range(n):
            while flags[r][p] < e:
                time.sleep(0)
                if time.time() - t0 > 20:
                    w.errors.append(f"rank {r}: deadlock waiting for rank {p} at launch {e}")
                    return False
            if flags[r][p] != e:        # the kernel's "exchange-number mismatch" check must never fire in lockstep
                w.errors.append(f"rank {r} launch {e}: rank {p}'s flag reads {flags[r][p]}")
                return False
        return True

    for e in range(1, launches + 1):
        # ---- forward / backward of this step: reads my parameters, rewrites my gradient bucket
        w.phase[r] = COMPUTE
        w.grad_version[r] = -1                      # memset + accumulation in progress
        for s in range(n):                          # every slice of my parameters must be the previous launch's
            if w.param_version[r][s] != e - 1:
                w.errors.append(f"rank {r} step {e}: parameter slice {s} is version {w.param_version[r][s]}")
        jitter(3.0)
        w.grad_version[r] = e
        # ---- the exchange kernel
        w.phase[r] = IN_KERNEL
        w.kernel_launch[r] = e
        for p in range(n):                          # barrier 1: announce
            assert w.flag1[p][r] < e
            w.flag1[p][r] = e
        if not wait(w.flag1, e):
            return
        for p in range(n):                          # slice work: peer gradient loads ...
            jitter()
            if w.grad_version[p] != e:
                w.errors.append(f"rank {r} launch {e}: read rank {p}'s gradients in state {w.grad_version[p]}")
        for p in range(n):                          # ... Adam ... peer parameter stores
            jitter()
            if p != r and not (w.phase[p] == IN_KERNEL and w.kernel_launch[p] == e):
                w.errors.append(f"rank {r} launch {e}: wrote rank {p}'s parameters while it was in phase {w.phase[p]} "
                                f"(launch {w.kernel_launch[p]})")
            w.param_version[p][r] = e
        for p in range(n):                          # barrier 2: my stores are done
            w.flag2[p][r] = e
        if not wait(w.flag2, e):
            return
        jitter()


@pytest.mark.parametrize("n", [2, 3, 4, 8])
def test_protocol_is_safe_and_live(n):
    w = World(n)
    threads = [threading.Thread(target=rank_thread, args=(w, r, 40, random.Random(100 * n + r))) for r in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not any(t.is_alive() for t in threads), "dead-lock"
    assert w.errors == [], w.errors[:5]
    assert all(v == 40 for row in w.param_version for v in row)
