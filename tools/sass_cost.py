"""sass_cost.py -- static issue-cost model for the hot loop of a kernel (developer tool).

Model fitted on B200 with tools/ubench.cu: an SMSP issues one warp instruction per clock but
can fetch only two *fresh* register operands per clock; operands served by the operand
reuse cache (.reuse on the previous instruction, same slot, same register), uniform
registers, constants and immediates are free.  cost(instr) = max(1, fresh_reads / 2);
MUFU adds ~2.2 cycles of dispatch blocking (ubench G).

usage: python tools/sass_cost.py <binary> <mangled-kernel-substring> [pairs_per_iteration]
Finds the largest backward-branch loop and reports cycles per loop iteration.
"""
import re
import subprocess
import sys


def disasm(binary, fn_sub):
    txt = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
    out, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = fn_sub in line
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def loops(ins):
    addr_to_idx = {a: i for i, (a, _) in enumerate(ins)}
    res = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_to_idx:
                res.append((addr_to_idx[tgt], i))
    return res


def cost(body, mufu_extra=2.2, verbose=False):
    cache = {}  # slot -> register kept by .reuse
    total = 0.0
    n3 = 0
    hist = {}
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op, _, rest = t.partition(" ")
        ops = [o.strip() for o in rest.split(",")] if rest else []
        base = op.split(".")[0]
        srcs = ops[1:] if base not in ("STS", "STG", "BRA", "BAR", "ST") else ops
        fresh = set()
        newcache = {}
        for slot, o in enumerate(srcs):
            m = re.match(r"^[-|~]*R(\d+)(\.reuse)?", o)
            if not m or o.startswith("RZ"):
                continue
            reg = int(m.group(1))
            if cache.get(slot) == reg:
                pass  # served by the reuse cache
            else:
                fresh.add(reg)
            if m.group(2):
                newcache[slot] = reg
        # 64-bit / 128-bit destinations don't matter; only source reads are modelled
        cache = newcache
        c = max(1.0, len(fresh) / 2.0)
        if base == "MUFU":
            c += mufu_extra
        if len(fresh) >= 3:
            n3 += 1
        hist[base] = hist.get(base, 0) + 1
        total += c
        if verbose:
            print(f"{c:4.1f} {len(fresh)} {t}")
    return total, n3, hist


def main():
    binary, fn = sys.argv[1], sys.argv[2]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    ins = disasm(binary, fn)
    if not ins:
        sys.exit("kernel not found")
    lp = sorted(loops(ins), key=lambda ab: ab[1] - ab[0], reverse=True)
    for (a, b) in lp[:int(sys.argv[4]) if len(sys.argv) > 4 else 1]:
        body = [t for _, t in ins[a:b + 1]]
        tot, n3, hist = cost(body, verbose="-v" in sys.argv)
        print(f"loop {ins[a][0]:#x}..{ins[b][0]:#x}: {len(body)} instr, model {tot:.1f} cycles, "
              f"{n3} instr with >=3 fresh reads; per unit: {len(body)/per:.2f} instr, {tot/per:.2f} cycles")
        print("   ", ", ".join(f"{k}:{v}" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main()
