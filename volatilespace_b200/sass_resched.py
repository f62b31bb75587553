"""Post-pass over the SASS of the symmetric gravity kernel's pair loop: regroup the FP64
accumulates so that consecutive instructions share an operand slot, and set the .reuse flags.

Why (tools/fp64_stream.cu): a warp-wide FP64 instruction holds the FP64 pipe
for max(2, R) cycles, R = distinct 64-bit registers fetched from the register file; an operand
kept in the operand-reuse cache (same register, same slot, previous instruction of the warp
flagged .reuse, no yield in between) is free.  The four accumulates of a pair are the only
three-operand instructions; ptxas schedules them wherever they fit and knows nothing of this
cost.  This tool moves FP64 instructions of the loop body between the plain FP64 issue positions
(stall 2, no scoreboard, no wait mask) -- every other instruction and every position's control
bits (stall count, yield) stay where ptxas put them -- and accepts a move only if

  * every register dependency of ptxas's order keeps its direction (RAW, WAR, WAW), also across
    the loop's back edge;
  * every RAW edge from a producer without a scoreboard (FP64, integer moves, and the one LDS
    ptxas covers by a later LDS's scoreboard) keeps at least the smallest issue distance ptxas
    itself uses for that kind of edge -- (producer opcode, consumer opcode, operand slot) --
    anywhere in the file: its own latency knowledge;
  * consumers of scoreboarded results (LDS, MUFU) stay behind the instruction that waits on the
    scoreboard, writers of registers an STS still reads stay behind its read-barrier wait;

and the modelled cycles go down.  Registers, opcodes, immediates: untouched, so the patched
kernel computes the SAME BITS as the unpatched one (tools/sym_ab.py compares them on the GPU).

    python -m volatilespace_b200.sass_resched volatilespace_b200/_build/vsb_gravity_sym.o [--write] [--no-yield] [-v]

Without --write it only reports.  The code bytes are patched in place (the cubin sits
uncompressed in the object file; each loop body must be found exactly once).
"""
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD")
LINE = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
HI = re.compile(r"/\* (0x[0-9a-f]{16}) \*/")
REG = re.compile(r"\bR(\d+)\b")


class Ins:
    __slots__ = ("addr", "text", "word", "op", "base", "dst", "src", "slots", "stall", "yld", "wb",
                 "rb", "wait", "plain", "idx")

    def __init__(self, addr, text, word):
        self.addr, self.text, self.word = addr, text, word
        t = re.sub(r"^@!?U?P\d+\s+", "", text)
        self.op = t.split()[0]
        self.base = self.op.split(".")[0]
        ops = [o.strip() for o in t[len(self.op):].split(",")] if " " in t else []
        ctrl = (word >> 105) & ((1 << 21) - 1)
        self.stall, self.yld = ctrl & 0xF, (ctrl >> 4) & 1
        self.wb, self.rb, self.wait = (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3F
        self.dst, self.src, self.slots = set(), set(), {}
        self._regs(ops, t)
        self.plain = (self.base in FP64 and self.stall == 2 and self.wb == 7 and self.rb == 7 and
                      self.wait == 0 and not text.startswith("@"))

    def _regs(self, ops, t):
        def span(o, n):
            m = REG.search(o)
            return set(range(int(m.group(1)), int(m.group(1)) + n)) if m else set()

        if self.base in FP64:
            self.dst = span(ops[0], 2)
            slot_of = {"DFMA": (0, 1, 2), "DMUL": (0, 1), "DADD": (0, 2)}[self.base]
            for o, s in zip(ops[1:], slot_of):
                r = span(o, 2)
                self.src |= r
                if r:
                    self.slots[s] = min(r)
            return
        width = 4 if ".128" in self.op else 2 if ".64" in self.op else 1
        if self.base == "MUFU":      # MUFU.RSQ64H Rd(hi word), Rs(hi word)
            self.dst, self.src = span(ops[0], 1), span(ops[1], 1)
        elif self.base in ("LDS", "LDG", "LD", "LDL"):
            self.dst = span(ops[0], width)
            for o in ops[1:]:
                self.src |= span(o, 1)
        elif self.base in ("STS", "STG", "ST", "STL"):
            self.src = span(ops[0], 1) | span(ops[1], width)
        elif self.base in ("BRA", "BAR", "NOP", "WARPSYNC", "BSYNC", "BSSY", "EXIT", "UISETP",
                           "UIADD3", "ULEA", "UMOV", "ULOP3"):
            for o in ops:
                self.src |= span(o, 1)
        elif ops:                    # integer / move class: first operand written, all others read
            self.dst = span(ops[0], 2 if ".WIDE" in self.op or ".64" in self.op else 1)
            for o in ops[1:]:
                self.src |= span(o, 2 if ".WIDE" in self.op else 1)


def _cuobjdump():
    import os
    import shutil
    for cand in (os.environ.get("CUOBJDUMP"), shutil.which("cuobjdump"), "/usr/local/cuda/bin/cuobjdump"):
        if cand and os.path.exists(cand):
            return cand
    return "cuobjdump"


def disassemble(path):
    txt = subprocess.run([_cuobjdump(), "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    lines = txt.split("\n")
    for i, ln in enumerate(lines):
        if "Function :" in ln:
            cur = ln.split("Function :")[1].strip()
            funcs[cur] = []
            continue
        m = LINE.search(ln)
        if m and cur is not None and i + 1 < len(lines):
            h = HI.search(lines[i + 1])
            if h:
                funcs[cur].append(Ins(int(m.group(1), 16), m.group(2).strip(),
                                      (int(h.group(1), 16) << 64) | int(m.group(3), 16)))
    return funcs


def find_loop(ins):
    best = None
    for k, i in enumerate(ins):
        if i.base == "BRA":
            m = re.search(r"0x([0-9a-f]+)\s*$", i.text)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt < i.addr:
                lo = next(j for j, x in enumerate(ins) if x.addr == tgt)
                if sum(x.base == "STS" for x in ins[lo:k + 1]) >= 2 and (best is None or k - lo < best[1] - best[0]):
                    best = (lo, k)
    return best


def edges(body):
    """Dependency edges (i, j, kind) of the original order, i before j; kind RAW/WAR/WAW.  The body
    is unrolled twice so that loop-carried edges appear as (i, j + n)."""
    n = len(body)
    last_w, readers, out = {}, {}, []
    for pos in range(2 * n):
        x = body[pos % n]
        for r in x.src:
            if r in last_w and last_w[r] != pos:
                out.append((last_w[r], pos, "RAW"))
            readers.setdefault(r, []).append(pos)
        for r in x.dst:
            for q in readers.get(r, []):
                if q != pos:
                    out.append((q, pos, "WAR"))
            if r in last_w:
                out.append((last_w[r], pos, "WAW"))
            last_w[r] = pos
            readers[r] = []
    # keep edges that start in the first copy; de-duplicate
    return sorted({(i, j, k) for i, j, k in out if i < n and j - i < n})


def edge_class(p, c, reg_overlap_slot):
    return (p.base, c.base, reg_overlap_slot)


def overlap_slot(p, c):
    if c.base not in FP64:
        return -1
    for s, r in c.slots.items():
        if {r, r + 1} & p.dst:
            return s
    return -1


def times(order, body):
    """Issue time of every instruction: the stall counts belong to the POSITIONS."""
    t, acc = {}, 0
    for pos, k in enumerate(order):
        t[k] = acc
        acc += body[pos].stall
    return t, acc


def min_latency(funcs):
    """Smallest issue distance ptxas uses per (producer opcode, consumer opcode, operand slot) RAW
    edge whose producer sets no scoreboard, over the pair loops of all kernels in the file."""
    lat = {}
    for ins in funcs.values():
        rng = find_loop(ins)
        if not rng:
            continue
        body = ins[rng[0]:rng[1] + 1]
        n = len(body)
        t, total = times(list(range(n)), body)
        for i, j, kind in edges(body):
            p, c = body[i], body[j % n]
            if kind != "RAW" or p.wb != 7:
                continue
            d = t[j % n] - t[i] + (total if j >= n else 0)
            key = edge_class(p, c, overlap_slot(p, c))
            lat[key] = min(lat.get(key, 1 << 30), d)
    return lat


class Body:
    """One loop body: the dependency model of ptxas's order and the search over new orders."""

    def __init__(self, body, lat):
        self.body, self.n, self.lat = body, len(body), lat
        n = self.n
        self.edges = edges(body)
        # scoreboard waits: for producer i with barrier b -> position of the first later waiter
        self.after = {}
        for i, x in enumerate(body):
            for b in {x.wb, x.rb} - {7}:
                for d in range(1, n + 1):
                    if (body[(i + d) % n].wait >> b) & 1:
                        self.after[(i, b)] = i + d      # may be >= n: waiter in the next trip
                        break
        # "next FP64 instruction" of a position looks through an LDS / MUFU (they never move)
        skip = ("LDS", "MUFU")
        self.next_pos, self.prev_pos = [None] * n, [None] * n
        self.between = [set()] * n                   # registers written between prev_pos[p] and p
        self.hop_yields = [False] * n                # ... and whether one of those instructions yields
        for p in range(n):
            q = p + 1
            while q < n and body[q].base in skip:
                q += 1
            if q < n:
                self.next_pos[p] = q
                if body[p].base not in skip:
                    self.prev_pos[q] = p
                    self.between[q] = set().union(*[body[m].dst for m in range(p + 1, q)]) if q > p + 1 else set()
                    self.hop_yields[q] = any(body[m].yld == 0 for m in range(p + 1, q))
        self.T = [0] * (n + 1)                       # issue time of a POSITION
        for p in range(n):
            self.T[p + 1] = self.T[p] + body[p].stall
        self.total = self.T[n]
        # per edge: minimum issue distance, or the waiter the consumer has to stay behind
        self.adj = [[] for _ in range(n)]
        self.req = {}
        for e in self.edges:
            i, j, kind = e
            pr, c = body[i], body[j % n]
            need, waiter = 0, None
            if kind == "RAW":
                if pr.wb != 7:
                    waiter = self.after.get((i, pr.wb), -1)
                else:
                    need = self.lat.get(edge_class(pr, c, overlap_slot(pr, c)), 1 << 30)
            elif kind == "WAR" and pr.rb != 7:
                waiter = self.after.get((i, pr.rb), -1)
            self.req[e] = (need, waiter)
            self.adj[i].append(e)
            if j % n != i:
                self.adj[j % n].append(e)

    def _edge_ok(self, e, pos):
        i, j, kind = e
        n = self.n
        need, waiter = self.req[e]
        pi, pj = pos[i], pos[j % n] + (n if j >= n else 0)
        if pi >= pj:
            return False
        if waiter is not None:
            return waiter >= 0 and pj >= pos[waiter % n] + (n if waiter >= n else 0)
        if need:
            return self.T[pj % n] + (self.total if pj >= n else 0) - self.T[pi] >= need
        return True

    def verify(self, order):
        """Every dependency of ptxas's order holds in `order` (order[p] = instruction at position p)."""
        pos = [0] * self.n
        for p, k in enumerate(order):
            pos[k] = p
        return all(self._edge_ok(e, pos) for e in self.edges)

    def _kept(self, p, order):
        """Slots of the instruction at position p whose operand the previous FP64 instruction leaves
        in the reuse cache: same register in the same slot, not overwritten on the way, no yield;
        across an LDS / MUFU only slot B (what ptxas itself does)."""
        body = self.body
        q = self.prev_pos[p]
        x = body[order[p]]
        if q is None or x.base not in FP64 or body[q].yld == 0:
            return []
        prev = body[order[q]]
        if prev.base not in FP64:
            return []
        hop = q != p - 1
        if hop and self.hop_yields[p]:
            return []
        return [s_ for s_, r in x.slots.items()
                if prev.slots.get(s_) == r and not (hop and s_ != 1)
                and not ({r, r + 1} & (prev.dst | self.between[p]))]

    def _cpos(self, p, order):
        """modelled cycles of the instruction at position p: max(2, register-file reads)"""
        x = self.body[order[p]]
        if x.base not in FP64:
            return 0
        vals = list(x.slots.values())
        regs = set(vals)
        for s_ in self._kept(p, order):
            if vals.count(x.slots[s_]) == 1:
                regs.discard(x.slots[s_])
        return max(2, len(regs))

    def reuse_flags(self, order):
        """slot mask per position: the operands to keep for the next FP64 instruction"""
        flags = [0] * self.n
        for p in range(self.n):
            for s_ in self._kept(p, order):
                flags[self.prev_pos[p]] |= 1 << s_
        return flags

    def optimise(self, iters=600000, window=12, seed=1, t0=0.35, verbose=False):
        """Seeded simulated annealing over rotations of up to `window` neighbouring movable
        positions.  Returns (order, modelled cycles before, after)."""
        import math
        import random
        rnd = random.Random(seed)
        body, n = self.body, self.n
        order = list(range(n))
        pos = list(range(n))
        assert self.verify(order), "ptxas's own order fails the checks: the model is wrong"
        PP = [p for p in range(n) if body[p].plain]
        cur = start = sum(self._cpos(p, order) for p in range(n))
        best, best_order = cur, list(order)
        for it in range(iters):
            temp = t0 * (1.0 - it / iters) + 0.02
            rnd.random()                         # (keeps the random stream of the first version)
            a = rnd.randrange(len(PP) - 1)
            b = min(len(PP) - 1, a + rnd.randint(1, window))
            seq = PP[a:b + 1]
            old_vals = [order[p] for p in seq]
            vals = [old_vals[-1]] + old_vals[:-1] if rnd.random() < 0.5 else old_vals[1:] + [old_vals[0]]
            last = self.next_pos[seq[-1]]
            span = range(seq[0], (last if last is not None else seq[-1]) + 1)
            before = sum(self._cpos(q, order) for q in span)
            for p, v in zip(seq, vals):
                order[p] = v
                pos[v] = p
            delta = sum(self._cpos(q, order) for q in span) - before
            ok = delta <= 0 or rnd.random() < math.exp(-delta / temp)
            if ok:
                ok = all(self._edge_ok(e, pos) for v in vals for e in self.adj[v])
            if ok:
                cur += delta
            else:
                for p, v in zip(seq, old_vals):
                    order[p] = v
                    pos[v] = p
            if cur < best:
                best, best_order = cur, list(order)
            if verbose and it % 100000 == 0:
                print("  iteration %d: %d cycles (best %d)" % (it, cur, best))
        assert self.verify(best_order), "final order fails the global check"
        return best_order, start, best


def encode(body, order, flags):
    """New instruction words: an instruction keeps its own bits except the position-bound control
    fields (stall, yield) and the recomputed reuse mask (FP64 instructions only)."""
    out = []
    for p, k in enumerate(order):
        w = body[k].word
        if k != p or body[k].base in FP64:
            keep = 0x1F << 105                               # stall (4) + yield (1) of the position
            w = (w & ~keep) | (body[p].word & keep)
            if body[k].base in FP64:
                w = (w & ~(0xF << 122)) | (flags[p] << 122)
        out.append(w)
    return out


def to_bytes(words):
    return b"".join((w & ((1 << 64) - 1)).to_bytes(8, "little") + (w >> 64).to_bytes(8, "little")
                    for w in words)


def _optimise_one(job):
    name, body, lat, iters = job
    b = Body(body, lat)
    order, before, after = b.optimise(iters=iters)
    return name, order, b.reuse_flags(order), before, after


def patch_object(path, pattern="allpairs_sym_kernel", write=True, iters=600000, jobs=None, verbose=False,
                 no_yield=False):
    """Re-schedules the pair loop of every kernel in `path` whose mangled name contains `pattern`
    and patches the code bytes in place.  Returns a list of report lines."""
    import multiprocessing as mp
    funcs = disassemble(path)
    lat = min_latency(funcs)
    if verbose:
        for k, v in sorted(lat.items()):
            print("  min issue distance %-28s %d cycles" % (k, v))
    work, originals = [], {}
    for name, ins in funcs.items():
        if pattern in name:
            rng = find_loop(ins)
            if rng:
                body = ins[rng[0]:rng[1] + 1]
                originals[name] = to_bytes([x.word for x in body])
                if no_yield:
                    # no yield hints on the loop's FP64 instructions: a yield ends every reuse chain
                    # (118 of 512 positions), and the warps of an SMSP hand over at the MUFU / LDS
                    # scoreboard waits anyway (measured: 528.3 -> 526.2 ms per 1 M-body tick)
                    for x in body:
                        if x.base in FP64:
                            x.yld, x.word = 1, x.word | (1 << 109)
                work.append((name, body, lat, iters))
    if not work:
        raise RuntimeError("no kernel matching %r with a pair loop in %s" % (pattern, path))
    jobs = jobs or min(len(work), max(1, (mp.cpu_count() or 2) - 1))
    if jobs > 1:
        with mp.get_context("fork").Pool(jobs) as pool:
            results = pool.map(_optimise_one, work)
    else:
        results = [_optimise_one(w) for w in work]
    data = bytearray(open(path, "rb").read())
    report = []
    for (name, body, _, _), (_, order, flags, before, after) in zip(work, results):
        nfp = sum(x.base in FP64 for x in body)
        short = re.search(r"ILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)E", name)
        report.append("%s<%s>: loop 0x%x..0x%x, %d FP64 instructions, modelled FP64 cycles %d -> %d "
                      "(ceiling %.3f -> %.3f), %d instructions moved" % (
                          pattern, ",".join(short.groups()) if short else "?", body[0].addr, body[-1].addr,
                          nfp, before, after, 2.0 * nfp / before, 2.0 * nfp / after,
                          sum(k != p for p, k in enumerate(order))))
        old = originals[name]
        new = to_bytes(encode(body, order, flags))
        if new != old:
            if data.count(old) != 1:
                raise RuntimeError("loop body of %s not found exactly once in %s" % (name, path))
            at = data.find(old)
            data[at:at + len(old)] = new
    if write:
        open(path, "wb").write(data)
    return report


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    for line in patch_object(args[0], args[1] if len(args) > 1 else "allpairs_sym_kernel",
                             write="--write" in sys.argv, verbose="-v" in sys.argv,
                             no_yield="--no-yield" in sys.argv):
        print(line)


if __name__ == "__main__":
    main()
