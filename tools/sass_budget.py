"""Static instruction budget of one plane of the z-march, from SASS.

usage: cuobjdump -sass -fun <mangled kernel> lib.so | python tools/sass_budget.py [iteration]

The march body is unrolled; one plane = the instructions between two consecutive BAR.SYNC.  For that
range the script prints the opcode mix, grouped the way DESIGN.md budgets it (fp32 / shared-memory and
shuffle traffic / global / integer+select / control), and an issue-cycle estimate under a register-bank rule:
an instruction occupies the dispatch port for max(1, #distinct source registers in one bank) cycles, four banks
(register number % 4, measured: see NBANKS below); operands served from the reuse cache, RZ, immediates,
constant-bank and uniform-register operands are free.  Packed fp32x2 instructions are not modelled.
"""
import collections
import re
import sys

# Register-file banks.  tools/probe/fp_rate_probe.cu on B200: eight FFMA chains with three distinct source registers,
# four of which have two sources in the same (register number % 4) class, issue at 2.60 warp-instr/clk/SM -- the
# 4-bank prediction is 4/1.5 = 2.67, an even/odd 2-bank rule would give 2.0, "three in one 64-bit bank" 4.0.
NBANKS = 4

FP = {"FFMA", "FMUL", "FADD", "FMNMX", "FSEL", "FSET", "FSETP", "FCHK", "MUFU", "FFMA2", "FMUL2", "FADD2"}
MIO = {"LDS", "STS", "SHFL", "LDSM", "STSM"}
GLB = {"LDG", "STG", "LDL", "STL", "LD", "ST"}
CTL = {"BRA", "BSSY", "BSYNC", "BAR", "SYNCS", "EXIT", "WARPSYNC", "NOP", "CALL", "RET", "ELECT", "UTMALDG", "YIELD", "DEPBAR", "ERRBAR", "BMOV"}


def parse(lines):
    ins = []
    for ln in lines:
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);\s*/\*", ln)
        if not m:
            continue
        text = m.group(2).strip()
        pred = None
        pm = re.match(r"(@!?U?P\d+)\s+(.*)", text)
        if pm:
            pred, text = pm.group(1), pm.group(2)
        parts = text.split(None, 1)
        op = parts[0]
        ops = [o.strip() for o in parts[1].split(",")] if len(parts) > 1 else []
        ins.append((int(m.group(1), 16), pred, op, ops))
    return ins


def bank_cycles(op, ops, reuse_live):
    """Dispatch cycles of one instruction; reuse_live = {slot: register} latched by the previous instruction."""
    base = op.split(".")[0]
    srcs = ops[1:] if base not in ("STS", "STG", "STL", "BAR", "BRA") else ops
    banks = [set() for _ in range(NBANKS)]
    latched = {}
    for slot, o in enumerate(srcs):
        m = re.search(r"\bR(\d+)(\.reuse)?", o)
        if not m or "UR" in o and not re.search(r"(?<!U)R\d+", o):
            continue
        r = int(m.group(1))
        wide = 2 if (".64" in op or base.endswith("2")) else 1
        if m.group(2):
            latched[slot] = r
        if reuse_live.get(slot) == r:
            continue
        for k in range(wide):
            banks[(r + k) % NBANKS].add(r + k)
    return max(1, *(len(b) for b in banks)), latched


def main():
    which = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    ins = parse(sys.stdin.readlines())
    bars = [n for n, (_, _, op, _) in enumerate(ins) if op.startswith("BAR.SYNC")]
    if len(bars) < which + 2:
        print("not enough BAR.SYNC in this function:", len(bars))
        return
    body = ins[bars[which]:bars[which + 1]]
    mix = collections.Counter()
    cyc = collections.Counter()
    reuse = {}
    nreuse = 0
    for _, pred, op, ops in body:
        base = op.split(".")[0]
        c, reuse = bank_cycles(op, ops, reuse)
        grp = "fp32" if base in FP else "smem+shfl" if base in MIO else "global/local" if base in GLB else "control" if base in CTL else "int/select/mov"
        mix[(grp, base)] += 1
        cyc[grp] += c
        nreuse += any(".reuse" in o for o in ops)
    n = len(body)
    print(f"instructions in one plane of the march (static, iteration {which}): {n}")
    groups = collections.Counter()
    for (g, b), k in mix.items():
        groups[g] += k
    for g, k in groups.most_common():
        detail = ", ".join(f"{b} {k2}" for (g2, b), k2 in sorted(mix.items(), key=lambda kv: -kv[1]) if g2 == g)
        print(f"  {g:15s} {k:5d} ({100 * k / n:4.1f} %)  dispatch cycles {cyc[g]:5d}   {detail}")
    tot = sum(cyc.values())
    print(f"dispatch-cycle estimate with the register-bank rule (reg % 4): {tot} ({tot / n:.2f} per instruction; "
          f"{nreuse} instructions latch a .reuse operand)")


if __name__ == "__main__":
    main()
