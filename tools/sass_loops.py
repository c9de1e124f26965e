#!/usr/bin/env python
"""Backward branches (loops) of a SASS listing made by tools/sass_fn.sh: start, end, length in instructions --
the quick check of a hot loop's size without a GPU.

    tools/sass_fn.sh birp_cpp_b200/csrc/build/birp_kernels.o march_cube_kernelILb1ELi0ELi1ELi4E > /tmp/k.sass
    python tools/sass_loops.py /tmp/k.sass
"""
import re
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` closes the pipe early

for line in open(sys.argv[1]):
    m = re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?)\s*;', line)
    if not m:
        continue
    addr, ins = int(m.group(1), 16), m.group(2)
    b = re.search(r'BRA\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)', ins)
    if b and int(b.group(1), 16) < addr:
        target = int(b.group(1), 16)
        print(f"loop {target:#x}..{addr:#x}: {(addr - target) // 16 + 1} instrs   [{ins}]")
