#!/usr/bin/env python3
"""Developer aid for restating glibc's libm (the reference's un-vendored third-party dependency):
prints the disassembly of one address range of libm.so.6 with every rip-relative .rodata operand
annotated with the double it holds.  usage: libm_annot.py START END [libm.so.6]"""
import re, struct, subprocess, sys
start, end = int(sys.argv[1], 16), int(sys.argv[2], 16)
lib = sys.argv[3] if len(sys.argv) > 3 else "/lib/x86_64-linux-gnu/libm.so.6"
data = open(lib, "rb").read()
out = subprocess.run(["objdump", "-d", "--no-show-raw-insn", "--start-address=%#x" % start,
                      "--stop-address=%#x" % end, lib], capture_output=True, text=True).stdout
for line in out.splitlines():
    m = re.match(r"\s+([0-9a-f]+):\s+(.*)", line)
    if not m:
        continue
    txt = re.sub(r"\s*<[^>]*>", "", m.group(2))
    c = re.search(r"#\s*([0-9a-f]+)", txt)
    note = ""
    if c:
        a = int(c.group(1), 16)
        if 0x8f000 <= a < 0x8f000 + 0x4e2a0:
            q = struct.unpack_from("<Q", data, a)[0]
            d = struct.unpack_from("<d", data, a)[0]
            note = "   ; [%x] = %s (%r) %#018x" % (a, d.hex(), d, q)
        txt = re.sub(r"\s*#.*", "", txt)
    print("%6x: %-58s%s" % (int(m.group(1), 16), txt, note))
