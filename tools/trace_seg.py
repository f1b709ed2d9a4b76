import sys
import numpy as np
d = np.loadtxt(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/trace.txt', dtype=np.int64)
ks = int(sys.argv[2]) if len(sys.argv) > 2 else 19
t0 = d[d[:, 2] > 0][:, 2].min()
def role(r):
    x = d[d[:, 0] == r]
    return x[:, 2:] - t0
m, e0, f0, pr = role(0), role(3), role(4), role(1)
print('MMA warp per tile (ns): k-loop span | per k-step: wait full / issue / loop | gap to next tile')
for t in range(1, 6):
    k = m[t * ks:(t + 1) * ks]
    print('  tile %d: k-loop %6d | wait %4d issue %4d other %4d | gap %5d' % (
        t, k[-1, 2] - k[0, 0], np.mean(k[:, 1] - k[:, 0]), np.mean(k[:, 2] - k[:, 1]),
        np.mean(k[1:, 0] - k[:-1, 2]), m[(t + 1) * ks, 0] - k[-1, 2]))
print('epilogue warp 0 of rank 0 (ns): waited for tmem_full | D1 | D2')
for t in range(1, 6):
    print('  tile %d: waited %5d | D1 %5d | D2 %5d' % (t, e0[t, 1] - e0[t, 0], e0[t, 2] - e0[t, 1], f0[t, 0] - e0[t, 2]))
