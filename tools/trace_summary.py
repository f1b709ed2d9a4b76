import sys
import numpy as np
d = np.loadtxt(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/trace.txt', dtype=np.int64)
ks = int(sys.argv[2]) if len(sys.argv) > 2 else 19
t0 = d[d[:, 2] > 0][:, 2].min()
def role(r):
    x = d[d[:, 0] == r]
    return x[:, 1], x[:, 2:] - t0
s, t = role(0)
print('MMA tile boundaries (ns):')
for i in range(ks - 1, 6 * ks, ks):
    print('  tile %d: MMA span %6d  (wait %5d issue %5d other %5d per stage avg) | gap to next tile %6d' % (
        i // ks, t[i, 2] - t[i - ks + 1, 0],
        np.mean(t[i - ks + 1:i + 1, 1] - t[i - ks + 1:i + 1, 0]), np.mean(t[i - ks + 1:i + 1, 2] - t[i - ks + 1:i + 1, 1]),
        np.mean(t[i - ks + 2:i + 1, 0] - t[i - ks + 1:i, 2]), t[i + 1, 0] - t[i, 2]))
s3, t3 = role(3)
s5, t5 = role(5)
print('epi0 (warp 0): after tmem_full: i2_arrived, i1_arrived')
for i in range(2, 6):
    print('  tile %d full %7d  i2arr +%5d  i1arr +%5d ' % (s3[i], t3[i, 1], t5[i, 0] - t3[i, 1], t5[i, 2] - t3[i, 1]))
