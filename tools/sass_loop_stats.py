#!/usr/bin/env python
"""Static evidence for the candidate-loop variants: per generated step kernel (ABL_MODE instance)
the register count, the number of SASS instructions and the extent of every loop (backward
branch), from `cuobjdump -sass` / `-res-usage` of a built model.  No GPU needed.

  python tools/sass_loop_stats.py build/models/<key>/model_kernels.o [kernel-name-substring]
"""
import re
import subprocess
import sys


def main():
    obj = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else "abl_kernel_"
    res = subprocess.run(["cuobjdump", "-res-usage", obj], stdout=subprocess.PIPE, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
    sass = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    fn, ins = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            ins[fn] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and fn:
            ins[fn].append((int(m.group(1), 16), m.group(2).strip()))
    for fn in sorted(ins):
        if want not in fn or "abl_kernel_" not in fn:
            continue
        mode = re.search(r"ILi(\d+)E", fn)
        body = ins[fn]
        exit_at = next((k for k, (_, t) in enumerate(body) if t.startswith("EXIT") and k > len(body) // 3), len(body))
        print("%s  ABL_MODE %s  registers %s  instructions %d" % (fn, mode.group(1) if mode else "-", regs.get(fn, "?"), len(body)))
        for addr, text in body:
            m = re.search(r"BRA(?:\.\w+)* (?:\w+, )?(0x[0-9a-f]+)", text)
            if m and "BRA" in text:
                tgt = int(m.group(1), 16)
                if tgt < addr:
                    n = (addr - tgt) // 16 + 1
                    fp64 = sum(1 for a, t in body if tgt <= a <= addr and re.match(r"(@!?U?P\d+ )?D(ADD|MUL|FMA|SETP)", t))
                    ld = sum(1 for a, t in body if tgt <= a <= addr and re.search(r"\bLD[GS]\b|\bLDG\.", t))
                    print("    loop 0x%04x..0x%04x: %3d instructions (%d FP64, %d loads)" % (tgt, addr, n, fp64, ld))
        print()


if __name__ == "__main__":
    main()
