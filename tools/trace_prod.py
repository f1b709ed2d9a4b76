"""Producer-side view of the in-kernel trace (lab build): thread 0 of producer group 0, CTA rank 0, cluster 0."""
import sys
import numpy as np
d = np.loadtxt(sys.argv[1], dtype=np.int64)
t0 = d[d[:, 2] > 0][:, 2].min()
a = d[d[:, 0] == 1][:, 2:] - t0
b = d[d[:, 0] == 2][:, 2:] - t0
print('own iteration: compute+idle before wait | wait for empty slot | stores | fence.proxy.async | arrive | (period)')
for i in range(20, 50):
    if a[i, 0] <= 0 or a[i + 1, 0] <= 0:
        break
    print('  it %3d: %5d | %5d | %5d | %5d | %5d | %6d' % (i, a[i, 0] - b[i - 1, 1], a[i, 1] - a[i, 0], a[i, 2] - a[i, 1],
                                                     b[i, 0] - a[i, 2], b[i, 1] - b[i, 0], a[i + 1, 0] - a[i, 0]))
