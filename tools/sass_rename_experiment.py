"""EXPERIMENT (not part of the build): live-range renaming on top of volatilespace_b200/sass_resched.py.

Status at the end of round 1:
  * Renamer: the 314 values of the <4,8,2,3> pair loop that may change registers (written by a movable
    FP64 instruction, read only by FP64 instructions of the same trip, both halves from that write);
  * relaxed_edges: the dependency edges that remain when those values can be renamed (all RAW; WAR / WAW
    only between fixed values) -- with them the annealing reaches 1 059 modelled cycles (floor 1 056);
  * allocate: interval colouring with the fixed values as pre-coloured intervals (read barriers of STS
    honoured); on ptxas's order it reproduces ptxas's assignment;
  * apply + same_dataflow: patched register fields, and the check that every operand still sees the same
    producer (passes);
  * search: annealing with a bound on how much a value's life may stretch and periodic allocation checks.
    It finds almost nothing allocatable: ptxas's loop keeps all 47 pairs of the pool busy at its peak, so
    the search needs the register pressure as a constraint (not just a veto).  That is the open part.

    python tools/sass_rename_experiment.py volatilespace_b200/_build/vsb_gravity_sym_ptxas.o [slack] [iterations]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200 import sass_resched as R  # noqa: E402

FP64 = R.FP64


def reaching(body, order):
    """For the program `order` (instruction ids at positions), steady state: for every (id, reg read)
    the id of the instruction whose write it sees and whether that write happened in the previous
    trip; None = never written in the loop (live-in)."""
    n = len(order)
    last = {}
    out = {}
    for step in range(2 * n):
        k = order[step % n]
        x = body[k]
        if step >= n:
            for r in x.src:
                if r in last:
                    out[(k, r)] = (last[r][0], last[r][1] < n)
                else:
                    out[(k, r)] = (None, False)
        for r in x.dst:
            last[r] = (k, step)
    return out


class Renamer:
    def __init__(self, body):
        self.body, self.n = body, len(body)
        n = self.n
        ident = list(range(n))
        self.reach0 = reaching(body, ident)
        # readers of every def: def id -> list of (reader id, reg, wrapped)
        self.readers = {}
        for (k, r), (d, wrapped) in self.reach0.items():
            if d is not None:
                self.readers.setdefault(d, []).append((k, r, wrapped))
        # renameable values: written by a movable FP64 instruction, read only by FP64 instructions of
        # the same trip, each of them seeing BOTH halves from this write
        self.renameable = set()
        for i, x in enumerate(body):
            if not x.plain:
                continue
            rs = self.readers.get(i, [])
            if not rs:
                continue
            ok = True
            for (k, r, wrapped) in rs:
                y = body[k]
                if wrapped or y.base not in FP64:
                    ok = False
                    break
                base = r & ~1
                if base not in y.slots.values():
                    ok = False
                    break
                if self.reach0.get((k, base), (None,))[0] != i or self.reach0.get((k, base + 1), (None,))[0] != i:
                    ok = False
                    break
            # nobody else may write half of the pair while the value lives (e.g. MUFU high word):
            # covered by the both-halves test above
            if ok:
                self.renameable.add(i)

    def relaxed_edges(self):
        """Dependency edges for the search: all RAW; WAR / WAW only where neither the value being
        read / overwritten nor the new one can be renamed."""
        body, n = self.body, self.n
        out = set()
        last_w, readers = {}, {}
        for pos in range(2 * n):
            k = pos % n
            x = body[k]
            for r in x.src:
                if r in last_w and last_w[r] != pos:
                    out.add((last_w[r], pos, "RAW"))
                readers.setdefault(r, []).append(pos)
            for r in x.dst:
                new_ren = k in self.renameable
                old = last_w.get(r)
                old_ren = old is not None and (old % n) in self.renameable
                if not new_ren and not old_ren:
                    for q in readers.get(r, []):
                        if q != pos:
                            out.add((q, pos, "WAR"))
                    if old is not None:
                        out.add((old, pos, "WAW"))
                last_w[r] = pos
                readers[r] = []
        return sorted((i, j, k) for i, j, k in out if i < n and j - i < n)


def fixed_intervals(rn, body_pos_of, order, after):
    """Per 32-bit register: intervals (a, b) in positions of `order` during which a FIXED value
    occupies it: written at a (+1/2), last read at b.  Wrapped lives are split; live-in registers
    are busy all the time.  A last reader with a read barrier (STS) holds the register until the
    instruction that waits on that barrier."""
    body, n = rn.body, rn.n
    busy = {}
    reach = rn.reach0          # who reads whose value is a property of the PROGRAM, not of the order
    pos = body_pos_of
    written = set()
    for k, x in enumerate(body):
        written |= x.dst
    # reads
    last_read = {}    # (def id or None, reg) -> list of (reader pos, wrapped)
    for (k, r), (d, wrapped) in reach.items():
        if d is not None and d in rn.renameable:
            continue
        end = pos[k]
        x = body[k]
        if x.rb != 7:
            w = after.get((k, x.rb))
            if w is not None:
                # the waiter may sit in the next trip
                wp = pos[w % n] + (n if w >= n else 0)
                end = max(end, wp)
        last_read.setdefault((d, r), []).append((end, wrapped))
    for (d, r), lst in last_read.items():
        iv = busy.setdefault(r, [])
        if d is None:
            iv.append((-1, n))
            continue
        a = pos[d]
        for end, wrapped in lst:
            if wrapped:
                iv.append((a, n))
                iv.append((-1, end if end < n else n))
            else:
                if end >= n:            # read-barrier wait in the next trip
                    iv.append((a, n))
                    iv.append((-1, end - n))
                else:
                    iv.append((a, end))
    # fixed defs nobody reads still write
    for k, x in enumerate(body):
        if k in rn.renameable:
            continue
        for r in x.dst:
            busy.setdefault(r, []).append((pos[k], pos[k] + 1))
    return busy


def allocate(rn, order, after):
    """Registers for the renameable values in the program `order`.  Returns {def id: new pair base}
    or None if some value finds no free pair."""
    body, n = rn.body, rn.n
    pos = [0] * n
    for p, k in enumerate(order):
        pos[k] = p
    busy = fixed_intervals(rn, pos, order, after)
    pool = sorted({min(body[i].dst) for i in rn.renameable})
    # life of every renameable value
    life = {}
    for i in rn.renameable:
        life[i] = (pos[i], max(pos[k] for k, _, _ in rn.readers[i]))
    held = {}          # pair base -> (def id, last)
    assign = {}

    def free_for(pair, a, b):
        for r in (pair, pair + 1):
            for c, d in busy.get(r, ()):
                if c < b and a < d:
                    return False
        h = held.get(pair)
        return h is None or h[1] <= a

    for p in range(n):
        i = order[p]
        if i not in rn.renameable:
            continue
        a, b = life[i]
        if b <= a:
            return None
        orig = min(body[i].dst)
        if free_for(orig, a, b):
            choice = orig
        else:
            # best fit: the free pair whose next fixed occupation starts soonest after b
            choice, gap = None, None
            for pair in pool:
                if pair != orig and free_for(pair, a, b):
                    nxt = min([c for r in (pair, pair + 1) for c, d in busy.get(r, ()) if c >= b] + [n])
                    if gap is None or nxt < gap:
                        choice, gap = pair, nxt
            if choice is None:
                return None
        assign[i] = choice
        held[choice] = (i, b)
    return assign


def apply(rn, order, assign):
    """Instruction copies with the renamed registers (fields and the .dst/.src/.slots model)."""
    import copy
    body, n = rn.body, rn.n
    new = [copy.copy(x) for x in body]
    for x in new:
        x.dst, x.src, x.slots = set(x.dst), set(x.src), dict(x.slots)
    FIELD = {0: 24, 1: 32, 2: 64}
    for i, pair in assign.items():
        old = min(body[i].dst)
        if pair == old:
            continue
        x = new[i]
        x.word = (x.word & ~(0xFF << 16)) | (pair << 16)
        x.dst = {pair, pair + 1}
        for (k, r, _) in rn.readers[i]:
            if r != old:
                continue
            y = new[k]
            for s_, reg in body[k].slots.items():
                if reg == old and rn.reach0[(k, old)][0] == i:
                    y.word = (y.word & ~(0xFF << FIELD[s_])) | (pair << FIELD[s_])
                    y.slots[s_] = pair
            y.src = set()
            for reg in y.slots.values():
                y.src |= {reg, reg + 1}
    return new


def same_dataflow(rn, order, new):
    """Every operand of the renamed program sees the same producer as in ptxas's program."""
    body, n = rn.body, rn.n
    r1 = reaching(new, order)
    # map operands by (instruction, slot) for FP64, by register for the rest (they keep registers)
    for k, x in enumerate(body):
        y = new[k]
        if x.base in FP64:
            for s_, reg in x.slots.items():
                for h in (0, 1):
                    if rn.reach0[(k, reg + h)] != r1[(k, y.slots[s_] + h)]:
                        return False, (k, s_, reg, y.slots[s_], rn.reach0[(k, reg + h)], r1[(k, y.slots[s_] + h)])
        else:
            for reg in x.src:
                if rn.reach0[(k, reg)] != r1[(k, reg)]:
                    return False, (k, reg, rn.reach0[(k, reg)], r1[(k, reg)])
    return True, None



def search(rn, b, iters=400000, window=12, seed=1, t0=0.35, slack=2, check_every=3000, verbose=True):
    """Body.optimise with two additions: a renameable value may not live more than `slack` positions
    longer than in ptxas's order, and a new best only counts if the registers can be allocated
    (checked at most every `check_every` iterations)."""
    import math, random
    rnd = random.Random(seed)
    body, n = b.body, b.n
    order = list(range(n)); pos = list(range(n))
    PP = [p for p in range(n) if body[p].plain]
    cur = start = sum(b._cpos(p, order) for p in range(n))
    best, best_order, best_assign = cur, list(order), None
    life0 = {i: max(k for k, _, _ in rn.readers[i]) - i for i in rn.renameable}
    # values touched by an instruction: the one it defines and the ones it reads
    touch = [set() for _ in range(n)]
    for i in rn.renameable:
        touch[i].add(i)
        for k, _, _ in rn.readers[i]:
            touch[k].add(i)
    last_check = -check_every
    for it in range(iters):
        temp = t0 * (1.0 - it / iters) + 0.02
        a = rnd.randrange(len(PP) - 1)
        bb = min(len(PP) - 1, a + rnd.randint(1, window))
        seq = PP[a:bb + 1]
        old_vals = [order[p] for p in seq]
        vals = [old_vals[-1]] + old_vals[:-1] if rnd.random() < 0.5 else old_vals[1:] + [old_vals[0]]
        last = b.next_pos[seq[-1]]
        span = range(seq[0], (last if last is not None else seq[-1]) + 1)
        before = sum(b._cpos(q, order) for q in span)
        for p, v in zip(seq, vals):
            order[p] = v; pos[v] = p
        delta = sum(b._cpos(q, order) for q in span) - before
        ok = delta <= 0 or rnd.random() < math.exp(-delta / temp)
        if ok:
            ok = all(b._edge_ok(e, pos) for v in vals for e in b.adj[v])
        if ok:
            for i in set().union(*[touch[v] for v in vals]):
                if max(pos[k] for k, _, _ in rn.readers[i]) - pos[i] > life0[i] + slack:
                    ok = False
                    break
        if ok:
            cur += delta
        else:
            for p, v in zip(seq, old_vals):
                order[p] = v; pos[v] = p
        if cur < best and it - last_check >= check_every:
            last_check = it
            asg = allocate(rn, order, b.after)
            if asg is not None:
                best, best_order, best_assign = cur, list(order), asg
                if verbose:
                    print("  it %d: %d cycles, allocatable" % (it, cur), flush=True)
    return best_order, best_assign, start, best


def run(path, pat, slack=2, iters=400000):
    funcs = R.disassemble(path)
    lat = R.min_latency(funcs)
    name = [k for k in funcs if pat in k and 'sym_kernel' in k][0]
    ins = funcs[name]
    lo, hi = R.find_loop(ins)
    body = ins[lo:hi + 1]
    for x in body:
        if x.base in FP64:
            x.yld, x.word = 1, x.word | (1 << 109)
    rn = Renamer(body)
    orig_edges = R.edges
    R.edges = lambda b_: rn.relaxed_edges()
    b = R.Body(body, lat)
    R.edges = orig_edges
    order, assign, st, best = search(rn, b, iters=iters, slack=slack)
    print("slack", slack, "cycles", st, "->", best, "renamed", None if assign is None else sum(assign[i] != min(body[i].dst) for i in assign))
    if assign is None:
        return
    new = apply(rn, order, assign)
    print("same dataflow:", same_dataflow(rn, order, new))


if __name__ == "__main__":
    run(sys.argv[1], 'ILi4ELi8ELi2ELi3E', slack=int(sys.argv[2]) if len(sys.argv) > 2 else 4,
        iters=int(sys.argv[3]) if len(sys.argv) > 3 else 300000)
