"""Discrete-event model of a chained GEMM launch (mvae_gemm_chain): 148 persistent CTAs take tiles in static round-robin
order; a tile needs its producer row blocks; the TMA/MMA pipeline of a CTA runs tile after tile and the epilogue of a
tile overlaps the main loop of the next.  Calibrated against tools/tile_gantt.py (profiles/r01_gantt_v5.txt); used to
try tile orders / split-K choices on the CPU before spending GPU time.

    python tools/chain_sim.py [batch]        spans of the four chained launches, next to the measured ones
    python tools/chain_sim.py --what-if      the same under the round-2 candidates

Model constants (B200, 3xTF32, measured): T_KB us per 128x128x32 k-block, T_FILL us from the first TMA issue of a tile to
its first MMA when the pipeline is cold, T_EPI us per epilogue (BIAS_SWISH / DSWISH ~4, red.add ~2.7), T_FLAG us for a
completion counter to become visible to a waiting producer."""
import sys
from dataclasses import dataclass, field
from typing import List

T_KB, T_FILL, T_FLAG = 0.62, 2.2, 1.0
N_CTA = 148


@dataclass
class Problem:
    name: str
    M: int
    N: int
    K: int
    split: int = 1
    dep: int = -1
    a_mn: bool = False          # wgrad: dependency through the k range
    epi: float = 4.0
    tiles: List[tuple] = field(default_factory=list)

    def build(self):
        tm, tn = -(-self.M // 128), -(-self.N // 128)
        nkb = -(-self.K // 32)
        s = min(self.split, nkb)
        per = -(-nkb // s)
        s = -(-nkb // per)
        self.tm, self.tn, self.s, self.per, self.nkb = tm, tn, s, per, nkb
        self.tiles = [(sp, m, n) for sp in range(s) for m in range(tm) for n in range(tn)]
        self.done_time = [[0.0, 0] for _ in range(tm)]      # per row block: latest completion, tiles completed
        self.target = tn * s


def simulate(problems: List[Problem], n_cta: int = N_CTA, verbose: bool = True):
    for p in problems:
        p.build()
    order = [(pi, t) for pi, p in enumerate(problems) for t in p.tiles]
    # completion time of every row block must be known before dependents are evaluated: process tiles in global index
    # order (a dependent tile always has a larger index than its producers; a CTA's own tiles are in index order too)
    cta_mainloop_free = [0.0] * n_cta      # when the CTA's TMA/MMA pipeline can start the next tile
    cta_epi_free = [0.0] * n_cta           # when its epilogue warps are free
    spans = {}
    wait_total = 0.0
    for idx, (pi, (sp, m, n)) in enumerate(order):
        p = problems[pi]
        c = idx % n_cta
        ready = 0.0
        if p.dep >= 0:
            q = problems[p.dep]
            if p.a_mn:
                k0 = sp * p.per * 32
                k1 = min((sp + 1) * p.per * 32, p.K)
                rbs = range(k0 // 128, (k1 - 1) // 128 + 1)
            else:
                rbs = [m]
            for rb in rbs:
                assert q.done_time[rb][1] == q.target, "producer tile scheduled after its consumer"
                ready = max(ready, q.done_time[rb][0] + T_FLAG)
        kb = min(p.per, p.nkb - sp * p.per)
        start = max(cta_mainloop_free[c], ready)
        wait_total += max(0.0, ready - cta_mainloop_free[c])
        cold = T_FILL if ready > cta_mainloop_free[c] or cta_mainloop_free[c] == 0.0 else 0.3   # prefetch hides the fill
        # narrow tiles are only a little cheaper per k-block: the A operand (TMA + split) costs the same
        ml_end = start + cold + kb * T_KB * (0.75 + 0.25 * min(p.N, 128) / 128)
        epi_start = max(ml_end, cta_epi_free[c])
        epi_end = epi_start + p.epi
        cta_mainloop_free[c] = ml_end if epi_start == ml_end else epi_start   # 2 accumulators: at most one epilogue behind
        cta_epi_free[c] = epi_end
        rb = p.done_time[m]
        rb[0] = max(rb[0], epi_end); rb[1] += 1
        s = spans.setdefault(pi, [start, epi_end])
        s[0] = min(s[0], start); s[1] = max(s[1], epi_end)
    total = max(cta_epi_free)
    if verbose:
        for pi, p in enumerate(problems):
            print(f"  p{pi}:{p.name:28s} tiles {len(p.tiles):4d}  {spans[pi][0]:6.1f} .. {spans[pi][1]:6.1f} us")
        print(f"  span {total:.1f} us, dependency wait {wait_total / n_cta:.1f} us per CTA")
    return total


def mnist_chains(B: int, L: int = 64):
    P = Problem
    nk_dec = max(1, (2 * B) // 32); sd = max(1, min(nk_dec // 16, 32))
    nk_enc = max(1, B // 32); se = max(1, min(nk_enc // 16, 32))
    enc_f = [P("fc1_i", B, 512, 784), P("fc2_t", B, 512, 512), P("fc2_i", B, 512, 512, dep=0),
             P("heads_t", B, 2 * L, 512, dep=1, epi=3.0), P("heads_i", B, 2 * L, 512, dep=2, epi=3.0)]
    dec_f = [P("d_i1", 2 * B, 512, L), P("d_t1", 2 * B, 512, L)]
    for l in (2, 3):
        dec_f += [P(f"d_i{l}", 2 * B, 512, 512, dep=len(dec_f) - 2), P(f"d_t{l}", 2 * B, 512, 512, dep=len(dec_f) - 1)]
    dec_f += [P("d_i4", 2 * B, 784, 512, dep=4, epi=3.0), P("d_t4", 2 * B, 10, 512, dep=5, epi=3.0)]
    dec_b = []
    for l in (4, 3, 2, 1):
        n_i, n_t, K = (784 if l == 4 else 512), (10 if l == 4 else 512), (L if l == 1 else 512)
        base = len(dec_b)
        di, dt = (-1, -1) if l == 4 else (base - 4, base - 3)
        dec_b += [P(f"dg_i{l}", 2 * B, K, n_i, dep=di, epi=4.5 if l > 1 else 2.7), P(f"dg_t{l}", 2 * B, K, n_t, dep=dt, epi=4.5 if l > 1 else 2.7),
                  P(f"wg_i{l}", n_i, K, 2 * B, split=sd, dep=di, a_mn=True, epi=2.7),
                  P(f"wg_t{l}", n_t, K, 2 * B, split=sd, dep=dt, a_mn=True, epi=2.7)]
    enc_b = [P("dg_hi", B, 512, 2 * L, epi=4.5), P("dg_ht", B, 512, 2 * L, epi=4.5), P("wg_hi", 2 * L, 512, B, split=se, a_mn=True, epi=2.7),
             P("wg_ht", 2 * L, 512, B, split=se, a_mn=True, epi=2.7), P("dg_2i", B, 512, 512, dep=0, epi=4.5), P("dg_2t", B, 512, 512, dep=1, epi=3.0),
             P("wg_2i", 512, 512, B, split=se, dep=0, a_mn=True, epi=2.7), P("wg_2t", 512, 512, B, split=se, dep=1, a_mn=True, epi=2.7),
             P("wg_1i", 512, 784, B, split=se, dep=4, a_mn=True, epi=2.7)]
    return {"encoders fwd": enc_f, "decoders fwd": dec_f, "decoders bwd": dec_b, "encoders bwd": enc_b}


def what_if():
    """Sum of the four chain spans of a step under the round-2 candidates: a main loop at the tensor
    pipe's pace (800 cycles per k-block instead of ~1150) and split-K with a last-arriver epilogue for the layers that
    leave SMs idle."""
    global T_KB, T_FILL

    def total(B, t_kb=T_KB, fix_split=0, t_fill=T_FILL):
        global T_KB, T_FILL
        old = (T_KB, T_FILL)
        T_KB, T_FILL = t_kb, t_fill
        tot = 0.0
        for probs in mnist_chains(B).values():
            if fix_split:
                for p in probs:
                    if not p.a_mn and -(-p.M // 128) * -(-p.N // 128) < 100 and p.K >= 256:
                        p.split, p.epi = fix_split, 2.7 + 3.0 / fix_split    # red.add partials + one fix-up pass per tile
            tot += simulate(probs, verbose=False)
        T_KB, T_FILL = old
        return tot
    print("sum of the chain spans of one MNIST step [us] (model):")
    print(f"{'B/GPU':>6s} {'now':>7s} {'800-cycle main loop':>20s} {'split-K x4 fix-up':>18s} {'both':>7s} {'both + 1 us fill':>17s}")
    for B in (4096, 2048, 1024, 512):
        print(f"{B:6d} {total(B):7.0f} {total(B, 0.42):20.0f} {total(B, fix_split=4):18.0f} {total(B, 0.42, 4):7.0f} "
              f"{total(B, 0.42, 4, 1.0):17.0f}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--what-if":
        what_if()
        sys.exit(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    measured = {4096: (61.0, 149.6, 284.5, 89.7), 512: (51.8, 58.6, 73.1, 43.8)}.get(B)
    tot = 0.0
    for i, (name, probs) in enumerate(mnist_chains(B).items()):
        print(f"=== {name} (B={B})" + (f"   measured span {measured[i]} us" if measured else ""))
        tot += simulate(probs)
    print(f"sum of chain spans {tot:.1f} us" + (f"   measured {sum(measured):.1f} us" if measured else ""))
