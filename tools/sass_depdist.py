#!/usr/bin/env python
"""Dependency distance of the fp64 instructions of every kernel in an object / cubin (a CPU-side check, no GPU):
for each DADD / DMUL / DFMA, the distance (in instructions) to the nearest earlier instruction that wrote one of its
source registers.  fp64 results take ~8 cycles and a warp issues one fp64 instruction every 2 cycles, so with two
resident warps per scheduler a kernel whose fp64 instructions mostly sit at distance 1-2 runs at the pipe's latency, not
its throughput.  This is how the collapsed leapfrog step of the four-warp HMC kernel was found (profiles/r2_summary.md).
    python tools/sass_depdist.py gpurun_out/klb_build/default/klb_hmc_ws_0.o [name-filter]
"""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, ops, addrs = None, [], []
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            if name:
                yield name, ops, addrs
            name, ops, addrs = m.group(1), [], []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ops.append(m.group(2))
            addrs.append(int(m.group(1), 16))
    if name:
        yield name, ops, addrs


def loops(ops, addrs):
    """innermost loop bodies: [target of a backward branch, the branch]"""
    idx = {a: i for i, a in enumerate(addrs)}
    spans = []
    for k, o in enumerate(ops):
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*(0x[0-9a-f]+)", o) or re.search(r"\bBRA\S*.*?(0x[0-9a-f]+)", o)
        if m:
            t = int(m.group(1), 16)
            if t in idx and idx[t] <= k:
                spans.append((idx[t], k))
    inner = [sp for sp in spans if not any(o != sp and sp[0] <= o[0] and o[1] <= sp[1] for o in spans)]
    return inner


def regs(s):
    return [int(x) for x in re.findall(r"\bR(\d+)\b", s)]


def analyse(ops):
    last = {}
    dist = collections.Counter()
    nfp = 0
    for k, o in enumerate(ops):
        m = re.match(r"(?:@!?U?P\d+\s+)?(\S+)\s*(.*)", o)
        if not m:
            continue
        opn, args = m.group(1), m.group(2)
        parts = [p.strip() for p in args.split(",")] if args else []
        is_store = opn.startswith(("ST", "RED", "ATOM", "BAR", "BRA", "EXIT", "WARPSYNC", "BSYNC", "BSSY", "SYNCS", "UBLKCP"))
        dst = [] if is_store or not parts else regs(parts[0])
        src = regs(args) if is_store else regs(",".join(parts[1:]))
        wide = opn.startswith(("D", "LDS.64", "LDG.E.64", "LD.E.64", "LDS.128", "LDG.E.128", "IMAD.WIDE", "F2F.F64", "I2F.F64", "MUFU.RCP64H"))
        if re.match(r"D(ADD|MUL|FMA)", opn):
            nfp += 1
            srcs = set()
            for r in src:
                srcs.add(r); srcs.add(r + 1)
            d = [k - last[r] for r in srcs if r in last]
            if d:
                dist[min(min(d), 9)] += 1
        for r in dst:
            n = 4 if "128" in opn else 2 if wide else 1
            for q in range(n):
                last[r + q] = k
    return nfp, dist


if __name__ == "__main__":
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    for name, ops, addrs in kernels(sys.argv[1]):
        if flt and flt not in name:
            continue
        short = name.replace("_Z16klb_chain_kernel", "chain").replace("_Z17klb_hmc_ws_kernel", "ws").replace("5KArgs", "")
        rows = [("whole kernel", ops)] + [("loop %#x..%#x" % (addrs[a], addrs[b]), ops[a:b + 1]) for a, b in loops(ops, addrs)]
        first = True
        for label, body in rows:
            nfp, dist = analyse(body)
            if nfp < 24:
                continue
            if first:
                print(short[:110])
                first = False
            tot = sum(dist.values()) or 1
            print("    %-28s instr %5d fp64 %5d  d1 %4.0f%%  d2 %4.0f%%  d3 %4.0f%%  d>=4 %4.0f%%" % (
                label, len(body), nfp, 100.0 * dist[1] / tot, 100.0 * dist[2] / tot, 100.0 * dist[3] / tot,
                100.0 * sum(v for k, v in dist.items() if k >= 4) / tot))
