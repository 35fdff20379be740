"""Decode the scheduling control words of a kernel's SASS (sm_100a: the 128-bit encoding cuobjdump prints): per
instruction the stall count, yield flag, write / read scoreboard it sets and the scoreboards it WAITS for.  Shows where
ptxas put the waits for the loads of the traversal loop -- which is not always where the source consumes them.
usage: python scripts/sass_scoreboards.py lib.so 'mangled_kernel_name' [first_addr_hex last_addr_hex]
(no GPU needed)"""
import re
import subprocess
import sys

lib, fun = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi_addr = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
raw = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], stdout=subprocess.PIPE, text=True).stdout.splitlines()
i = 0
while i < len(raw):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", raw[i])
    m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", raw[i + 1]) if m and i + 1 < len(raw) else None
    if m and m2:
        hi = int(m2.group(1), 16)
        stall, yld, wb, rb, wait = (hi >> 41) & 0xf, (hi >> 45) & 1, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f
        a = int(m.group(1), 16)
        if lo <= a <= hi_addr:
            waits = ",".join("SB%d" % b for b in range(6) if wait >> b & 1)
            print("%04x  %-60s stall %2d %s %s %s %s" % (
                a, m.group(2).strip()[:60], stall, "Y" if yld else " ", "sets SB%d" % wb if wb != 7 else "        ",
                "read-SB%d" % rb if rb != 7 else "        ", "WAITS " + waits if waits else ""))
        i += 2
    else:
        i += 1
