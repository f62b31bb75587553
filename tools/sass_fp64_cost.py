"""Operand-read cost model of an FP64 instruction stream on sm_100a, applied to SASS.

tools/fp64_stream.cu (log: profiles/r1_fp64_stream_b200.txt) measures that a warp-wide FP64
instruction occupies its SMSP for max(2, R) cycles, R = number of DISTINCT 64-bit registers it
has to fetch from the register file; an operand carried by the operand-reuse cache (flagged
`.reuse` in the same slot of the previous instruction) is free.  This script applies that model
to the SASS of a loop:

    cuobjdump -sass -fun <mangled> file.o | python tools/sass_fp64_cost.py [lo_hex hi_hex | --auto]

and prints the instruction mix and the modelled cycles of the address range.
"""
import re
import sys

INSTR = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*([^;]*);")


def parse(lines):
    out = []
    for ln in lines:
        m = INSTR.search(ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2), [o.strip() for o in m.group(3).split(",")]))
    return out


def cost(instrs, lo, hi):
    prev_reuse = {}
    tot = {"fp64": 0, "fp64_cycles": 0, "other": 0}
    mix = {}
    by_reads = {}
    for addr, op, ops in instrs:
        if not (lo <= addr <= hi):
            continue
        base = op.split(".")[0]
        mix[base] = mix.get(base, 0) + 1
        cur_reuse = {}
        if base in ("DFMA", "DMUL", "DADD"):
            srcs = ops[1:]
            regs = set()
            for slot, o in enumerate(srcs):
                r = re.search(r"R(\d+)", o)
                if not r:
                    continue
                reg = int(r.group(1))
                if ".reuse" in o:
                    cur_reuse[slot] = reg
                if prev_reuse.get(slot) == reg:
                    continue
                regs.add(reg)
            c = max(2, len(regs))
            by_reads[(base, len(regs))] = by_reads.get((base, len(regs)), 0) + 1
            tot["fp64"] += 1
            tot["fp64_cycles"] += c
        else:
            tot["other"] += 1
        prev_reuse = cur_reuse
    return tot, mix, by_reads


def auto_range(instrs):
    """Innermost loop (backward branch) that contains several STS: the pair loop of the symmetric kernel."""
    best = None
    for addr, op, ops in instrs:
        if op.startswith("BRA") and ops and ops[-1].startswith("0x"):
            tgt = int(ops[-1], 16)
            if tgt < addr and sum(o.startswith("STS") and tgt <= a <= addr for a, o, _ in instrs) >= 2:
                if best is None or addr - tgt < best[1] - best[0]:
                    best = (tgt, addr)
    return best


if __name__ == "__main__":
    ins = parse(sys.stdin.readlines())
    if len(sys.argv) > 1 and sys.argv[1] == "--auto":
        lo, hi = auto_range(ins)
        print("loop 0x%x..0x%x" % (lo, hi))
    else:
        lo = int(sys.argv[1], 16) if len(sys.argv) > 1 else 0
        hi = int(sys.argv[2], 16) if len(sys.argv) > 2 else 1 << 30
    tot, mix, by_reads = cost(ins, lo, hi)
    print("mix:", dict(sorted(mix.items(), key=lambda kv: -kv[1])))
    print("FP64 by register-file reads:", dict(sorted(by_reads.items())))
    n = tot["fp64"]
    print("FP64 instructions %d, modelled FP64 cycles %d (%.3f per instruction; pipe utilisation "
          "ceiling %.3f), other instructions %d" % (n, tot["fp64_cycles"], tot["fp64_cycles"] / max(n, 1),
                                                   2.0 * n / max(tot["fp64_cycles"], 1), tot["other"]))
