"""Generate tests/golden/* (run in the BUILD container, where /root/reference exists).

* weights_<cfg>.npz / config_<cfg>.json : the reference's trained checkpoints
  (checkpoints/<cfg>/model-*.{index,data}) decoded with utils/tf_checkpoint.py, and the frozen
  JSON config saved beside them (train.py:591-592).  Data, not code.
* graph_<cfg>.npz : a seeded synthetic frame, the keypoints of the oracle's voxel restatement,
  and the edge lists produced by the REFERENCE's own models/graph_gen.py
  (gen_disjointed_rnn_local_graph_v3, scikit-learn ball tree) on those vertices, in canonical
  (dst, src) order -> pins oracle/graph.py and the CUDA radius kernels to the reference.
* gnn_<cfg>.npz : logits / box encodings / per-layer features of oracle/gnn.py (fp32) on that
  graph with the real weights.  TensorFlow 1.15 cannot run here, so these are regression
  vectors of the restatement, not outputs of the reference ("parity unpinned").
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gnn, graph, reference_graph, synth  # noqa: E402
from pointgnn_b200.utils import tf_checkpoint  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
CONFIGS = {
    'car_auto_T3_train': dict(num_points=3000, frame=7),
    'ped_cyl_auto_T3_trainval': dict(num_points=3000, frame=8),
}


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = reference_graph.load()
    for name, spec in CONFIGS.items():
        ckpt_dir = os.path.join(reference_graph.REFERENCE_ROOT, 'checkpoints', name)
        with open(os.path.join(ckpt_dir, 'config')) as f:
            config = json.load(f)
        with open(os.path.join(GOLDEN, 'config_%s.json' % name), 'w') as f:
            json.dump(config, f, indent=1, sort_keys=True)
        weights = {k: v for k, v in tf_checkpoint.load_checkpoint(ckpt_dir).items()
                   if k.endswith('/weights') or k.endswith('/biases')}
        np.savez(os.path.join(GOLDEN, 'weights_%s.npz' % name), **weights)

        xyz, intensity = synth.lidar_frame(spec['frame'], spec['num_points'])
        kw = config['runtime_graph_gen_kwargs']
        coords, keypoints, edges = graph.gen_multi_level_local_graph_v3(xyz, **kw)
        ref_edges = []
        for lvl, cfg in enumerate(kw['level_configs']):
            e = ref.gen_disjointed_rnn_local_graph_v3(coords[lvl], coords[lvl + 1], **cfg['graph_gen_kwargs'])
            assert np.all(np.diff(e[:, 1]) >= 0), 'reference edges are not grouped by destination'
            e = graph.canonical_edges(e)
            assert np.array_equal(e, edges[lvl]), 'oracle radius graph != reference graph_gen'
            ref_edges.append(e.astype(np.int32))
        np.savez_compressed(
            os.path.join(GOLDEN, 'graph_%s.npz' % name), xyz=xyz, intensity=intensity,
            keypoint_idx=keypoints[0][:, 0].astype(np.int32),
            edges0=ref_edges[0], edges1=ref_edges[1])
        logits, boxes, feats = gnn.predict(weights, config['model_kwargs']['layer_configs'],
                                           config['num_classes'], 7, intensity, coords, keypoints, edges,
                                           return_features=True)
        np.savez_compressed(os.path.join(GOLDEN, 'gnn_%s.npz' % name), logits=logits, boxes=boxes,
                            features_pool=feats[1], features_last=feats[-1])
        print(name, 'K=%d E0=%d E1=%d' % (len(keypoints[0]), len(edges[0]), len(edges[1])),
              'logits', logits.shape, float(np.abs(logits).max()))


if __name__ == '__main__':
    main()
