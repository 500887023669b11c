#!/usr/bin/env python
"""Annotate the SASS of one kernel with its scheduling control fields (stall count, yield, write/read barrier index,
wait mask), decoded from the upper instruction word cuobjdump prints.  Lets the barrier (scoreboard) assignment of the
loads on the rollout kernels' dependent chain be checked without a GPU: which barrier the CDF-row load sets, and which
instruction is the first to wait on it.

usage: sass_ctrl.py <lib.so|cubin> <kernel-name-regex> [--loads]
Control word layout (Volta..Blackwell, bits of the 128-bit instruction): 105-108 stall, 109 yield, 110-112 write
barrier (7 = none), 113-115 read barrier (7 = none), 116-121 wait mask, 122-125 reuse."""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur:
            body[cur].append(line)
    return body


def decode(lines):
    ins = []
    i = 0
    while i < len(lines):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ctrl = hi >> 41
                stall, yld = ctrl & 0xf, (ctrl >> 4) & 1
                wbar, rbar, wait = (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3f
                ins.append((m.group(1), m.group(2).strip(), stall, yld, wbar, rbar, wait))
                i += 2
                continue
        i += 1
    return ins


def main():
    path, pat = sys.argv[1], sys.argv[2]
    only_loads = "--loads" in sys.argv
    for name, lines in kernels(path).items():
        if not re.search(pat, name):
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        print(f"== {dem[:160]}")
        for addr, text, stall, yld, wbar, rbar, wait in decode(lines):
            w = "".join(str(b) for b in range(6) if wait >> b & 1)
            tag = f"st{stall:2d}{'Y' if yld else ' '} W{wbar if wbar != 7 else '-'} R{rbar if rbar != 7 else '-'} wait[{w:6s}]"
            if only_loads and not (wbar != 7 or wait):
                continue
            print(f"{addr} {tag}  {text}")


if __name__ == "__main__":
    main()
