"""Time the graph build alone (8 synthetic 20 000-point frames per call, car_auto_T3 graph kwargs): wall clock per call
(it contains the build's one host round trip) and CUDA-event time.  PG_LIB_VARIANT=NAME loads lab/NAME.so."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from pointgnn_b200 import _lib  # noqa: E402
if os.environ.get('PG_LIB_VARIANT'):
    _lib.LIB_PATH = os.path.join(ROOT, 'lab', os.environ['PG_LIB_VARIANT'] + '.so')
from pointgnn_b200.models import graph_gen  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = json.load(open(os.path.join(ROOT, 'tests/golden/config_car_auto_T3_train.json')))
batches = []
for b in range(4):
    pts = np.vstack([synth.lidar_frame(100 + b * frames + i, 20000)[0] for i in range(frames)])
    batches.append(torch.from_numpy(pts).cuda())
fp = torch.arange(frames + 1, dtype=torch.int32, device='cuda') * 20000
for b in batches:
    graph_gen.gen_multi_level_local_graph_v3(b, frame_ptr=fp, **cfg['runtime_graph_gen_kwargs'])
torch.cuda.synchronize()
a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for r in range(reps):
    graph_gen.gen_multi_level_local_graph_v3(batches[r % 4], frame_ptr=fp, **cfg['runtime_graph_gen_kwargs'])
z.record()
z.synchronize()
wall = (time.perf_counter() - t0) / reps * 1e3
print('graph build, %d frames per call: %.3f ms wall per call, %.3f ms between CUDA events' % (frames, wall, a.elapsed_time(z) / reps))
