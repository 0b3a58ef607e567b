#!/usr/bin/env python
"""Filter `ncu --page source --csv` (stdin -> stdout): keep the first launch of every distinct kernel."""
import sys
seen, keep = set(), False
for line in sys.stdin:
    if line.startswith('"Kernel Name"'):
        name = line.split(',', 1)[1]
        keep = name not in seen
        seen.add(name)
    if keep:
        sys.stdout.write(line)
