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
# stamps (thread 0 of the task): 1 record read, 2 after the PDL wait, 3 dependencies ready, 4 pivot right-hand sides assembled,
# 7 panel multiplied (warp 0), 5 results stored, 6 signalled
order = [1, 2, 3, 4, 7, 5, 6]
names = ["pdl wait", "dep wait", "gather", "multiply", "store", "signal"]
print(f"{nt} tasks, {nl} levels; sweep part {us[:, 6].max():.1f} us")
print("level  tasks  first-start  last-ready(dep)  last-done |  medians: " + "  ".join(names))
for l in range(nl):
    a, b = fptr[l], fptr[l + 1]
    if a == b:
        continue
    u = us[a:b][:, order]
    d = np.diff(u, axis=1)
    print(f"{l:5d} {b - a:6d} {u[:, 0].min():12.1f} {u[:, 2].max():16.1f} {u[:, -1].max():10.1f} |  " +
          "  ".join(f"{np.median(d[:, i]):{len(n)}.2f}" for i, n in enumerate(names)))
