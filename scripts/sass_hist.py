#!/usr/bin/env python
"""SASS opcode histogram of a kernel (or of an address range inside it), from `cuobjdump -sass`.

  python scripts/sass_hist.py LIB KERNEL_SUBSTR [--range 0xa9d0 0xe490] [--loops]

Used for profiles/*_sass_hist.txt: proves which instructions the hot loops consist of (UTMALDG / SYNCS for the TMA +
mbarrier pipeline, DADD/DMUL/DFMA for the f64 arithmetic) and how many issue slots go to anything else.
"""
import argparse
import collections
import re
import subprocess


def disasm(lib, pattern):
    names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kernels, cur = {}, None
    for l in names.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and cur:
            kernels[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return {k: v for k, v in kernels.items() if pattern in k}


def opcode(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    return t.split()[0].split(".")[0]


CLASSES = [("fp64", r"^(DADD|DMUL|DFMA|DSETP|MUFU)$"), ("lds/sts", r"^(LDS|STS|LDSM)$"), ("ldg/stg", r"^(LDG|STG|LD|ST|RED|ATOM)$"),
           ("tma/mbarrier", r"^(UTMALDG|UTMASTG|SYNCS|UTMACMDFLUSH|UBLKCP)$"), ("branch/ctl", r"^(BRA|BSSY|BSYNC|CALL|RET|EXIT|WARPSYNC|NANOSLEEP|BAR|YIELD)$"),
           ("int/logic", r"^(IADD3?|IADD|IMAD|LOP3|LEA|SHF|ISETP|PLOP3|SEL|FSEL|IABS|VIMNMX\d?|PRMT|POPC|LOP|UIADD3?|ULOP3|UISETP|UMOV|USEL|ULEA|UIMAD|USHF|UPLOP3|UP2UR|R2UR|S2R|S2UR|CS2R|SHFL|VOTE|VOTEU|R2P|P2R)$"),
           ("mov", r"^(MOV|IMAD\.MOV|UMOV)$")]


def classify(op):
    for name, rx in CLASSES:
        if re.match(rx, op):
            return name
    return "other"


def drop_cold_blocks(sel):
    """remove forward-branched-over regions that only marshal arguments for out-of-line CALLs (the cold division
    paths): they contain CALL but no f64 arithmetic of their own"""
    addr = {a: i for i, (a, t) in enumerate(sel)}
    skip = set()
    for i, (a, t) in enumerate(sel):
        m = re.search(r"BRA(?:\.\S+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tg = int(m.group(1), 16)
            if tg > a and tg in addr:
                body = [x[1] for x in sel[i + 1:addr[tg]]]
                if any("CALL" in x for x in body) and not any(re.search(r"\b(DADD|DMUL|DFMA)\b", x) for x in body):
                    skip.update(range(i + 1, addr[tg]))
    return [x for i, x in enumerate(sel) if i not in skip]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib")
    ap.add_argument("kernel")
    ap.add_argument("--range", nargs=2, default=None)
    ap.add_argument("--loops", action="store_true", help="list backward branches (loop candidates)")
    ap.add_argument("--hot", action="store_true", help="leave out the cold out-of-line-call blocks inside the range")
    a = ap.parse_args()
    for name, ins in disasm(a.lib, a.kernel).items():
        print("== %s: %d instructions" % (name, len(ins)))
        if a.loops:
            for addr, t in ins:
                m = re.search(r"BRA(?:\.\S+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
                if m and int(m.group(1), 16) < addr:
                    print("   loop %#x..%#x  (%d instr)  %s" % (int(m.group(1), 16), addr, (addr - int(m.group(1), 16)) // 16 + 1, t))
        lo, hi = (int(a.range[0], 16), int(a.range[1], 16)) if a.range else (0, 1 << 60)
        sel = [(addr, t) for addr, t in ins if lo <= addr <= hi]
        if a.hot:
            sel = drop_cold_blocks(sel)
        sel = [t for _, t in sel]
        ops = collections.Counter(opcode(t) for t in sel)
        cls = collections.Counter()
        for op, n in ops.items():
            cls[classify(op)] += n
        print("   range %#x..%#x: %d instructions" % (lo, min(hi, ins[-1][0]), len(sel)))
        print("   classes: " + ", ".join("%s %d (%.0f%%)" % (k, v, 100.0 * v / max(len(sel), 1)) for k, v in cls.most_common()))
        print("   opcodes: " + ", ".join("%s %d" % kv for kv in ops.most_common()))


if __name__ == "__main__":
    main()
