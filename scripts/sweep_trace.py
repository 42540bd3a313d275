"""Read the stage clock of one dataflow forward sweep (PLFEM_TRACE_FILE, written by plfem_profile_last) and print, per
elimination-tree level, when its tasks started and finished and how long each stage of a task took (microseconds, medians)."""
import struct
import sys

import numpy as np

raw = open(sys.argv[1], "rb").read()
nt, nl = struct.unpack("qq", raw[:16])
fptr = np.frombuffer(raw, dtype=np.int32, count=nl + 1, offset=16)
st = np.frombuffer(raw, dtype=np.int64, count=nt * 8, offset=16 + 4 * (nl + 1)).reshape(nt, 8).astype(np.float64)
t0 = st[:, 1][st[:, 1] > 0].min()
us = (st - t0) / 1e3
names = ["pdl wait", "dep wait", "gather", "compute+store", "signal", "drain"]
print(f"{nt} tasks, {nl} levels; CTA entries span {(st[:, 0].max() - st[:, 0].min()) / 1e3:.1f} us; sweep part {us[:, 7].max():.1f} us")
print("level  tasks  first-start  last-ready(dep)  last-done |  medians: " + "  ".join(names))
for l in range(nl):
    a, b = fptr[l], fptr[l + 1]
    if a == b:
        continue
    u = us[a:b]
    d = np.diff(u, axis=1)[:, 1:]          # stage lengths from stamp 1 on (stamp 0 is indexed by CTA, not by ticket)
    print(f"{l:5d} {b - a:6d} {u[:, 1].min():12.1f} {u[:, 3].max():16.1f} {u[:, 6].max():10.1f} |  " +
          "  ".join(f"{np.median(d[:, i]):{len(n)}.2f}" for i, n in enumerate(names)))
