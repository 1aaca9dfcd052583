#!/usr/bin/env python
"""Decodes the scheduling control fields (stall count, yield, scoreboard set / wait) of
sm_100a SASS from `cuobjdump -sass` text, so that fixed-latency waits around DMMA chains can
be read without a GPU.  Usage: cuobjdump -sass x.o | python tools/sass_ctrl.py [first] [last]"""
import re
import sys

lines = sys.stdin.read().split("\n")
lo = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0
hi_ = int(sys.argv[2], 0) if len(sys.argv) > 2 else 1 << 30
ins = re.compile(r"^\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/")
enc2 = re.compile(r"^\s*/\* 0x([0-9a-f]{16}) \*/")
i = 0
while i < len(lines):
    m = ins.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = enc2.match(lines[i + 1])
        if m2:
            addr = int(m.group(1), 16)
            if lo <= addr <= hi_:
                hi = int(m2.group(1), 16)
                stall = (hi >> 41) & 0xf
                yld = (hi >> 45) & 1
                wbar = (hi >> 46) & 7
                rbar = (hi >> 49) & 7
                wait = (hi >> 52) & 0x3f
                print("%05x  st=%2d y=%d w=%s r=%s wait=%s  %s" % (
                    addr, stall, yld, "-" if wbar == 7 else wbar, "-" if rbar == 7 else rbar,
                    "".join(str(b) if wait >> b & 1 else "." for b in range(6)), m.group(2).strip()))
            i += 2
            continue
    i += 1
