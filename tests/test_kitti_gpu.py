"""GPU input stage (pg_cam_points_in_image) and the eager run.py twin end to end on a synthetic KITTI-format tree.

* input stage vs the fixture produced by the reference's own dataset/kitti_dataset.py (tests/golden/kitti_io.npz):
  same points kept, coordinates within 1 float32 ulp-class tolerance (the float32 3x3 product is evaluated in a
  fixed order here, BLAS sgemm in the reference), colours and reflectance exact;
* the twin (point-gnn_b200/run.py --test) writes KITTI result files that equal, line by line and within 2e-3, the
  files the CPU oracle pipeline produces for the same frames: oracle input stage -> oracle graph -> oracle GNN ->
  oracle decode + NMS -> oracle label conversion (each of those pinned to the reference in the CPU suite)."""
import json
import os
import shutil

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import cpu_reference, kitti as ok, postprocess as pp
from oracle import graph as ograph

pytestmark = pytest.mark.gpu
IO = dict(np.load(os.path.join(GOLDEN, 'kitti_io.npz')))
CALIB_FILE = os.path.join(GOLDEN, 'kitti_calib.txt')


def test_input_stage_matches_reference_fixture():
    from pointgnn_b200.dataset import kitti_dataset
    calib = kitti_dataset.parse_calib(CALIB_FILE)
    h, w = IO['image'].shape[:2]
    xyz, attr, fp = kitti_dataset.cam_points_in_image_batch([IO['velo']], [calib], [(w, h)], [IO['image']])
    assert xyz.shape == IO['xyz'].shape and attr.shape == IO['attr'].shape
    assert np.abs(xyz.cpu().numpy() - IO['xyz']).max() < 2e-5
    assert np.array_equal(attr.cpu().numpy(), IO['attr'])
    assert fp.cpu().tolist() == [0, len(IO['xyz'])]
    # reflectance only; two frames in one call (the second one reversed: order must be preserved per frame)
    xyz2, attr2, fp2 = kitti_dataset.cam_points_in_image_batch([IO['velo'], IO['velo'][::-1].copy()], [calib, calib],
                                                               [(w, h), (w, h)])
    n = len(IO['xyz'])
    assert fp2.cpu().tolist() == [0, n, 2 * n] and attr2.shape == (2 * n, 1)
    assert np.array_equal(attr2[:n].cpu().numpy(), IO['attr'][:, :1])
    assert np.array_equal(attr2[n:].cpu().numpy(), IO['attr'][::-1, :1])
    assert np.abs(xyz2[n:].cpu().numpy() - IO['xyz'][::-1]).max() < 2e-5


def test_dataset_class_matches_oracle(tmp_path):
    from pointgnn_b200.dataset import kitti_dataset
    root = str(tmp_path / 'kitti')
    names = ok.write_synthetic_kitti(root, [31], 4000)
    ds = kitti_dataset.KittiDataset(os.path.join(root, 'image/testing/image_2'),
                                    os.path.join(root, 'velodyne/testing/velodyne/'),
                                    os.path.join(root, 'calib/testing/calib/'), '', num_classes=4, is_training=False)
    assert ds.num_files == 1 and ds.get_filename(0) == names[0]
    pts = ds.get_cam_points_in_image_with_rgb(0)
    calib = ok.parse_calib(os.path.join(root, 'calib/testing/calib/000000.txt'))
    image = ds.get_image(0)
    xyz, attr = ok.cam_points_in_image(ds.get_velo_data(0), calib, image.shape[1], image.shape[0], image)
    assert pts.xyz.shape == xyz.shape and np.abs(pts.xyz - xyz).max() < 2e-5 and np.array_equal(pts.attr, attr)


@pytest.mark.parametrize('cfg_name', ['car_auto_T3_train'])
def test_run_twin_end_to_end(tmp_path, cfg_name):
    from pointgnn_b200 import run as twin
    root = str(tmp_path / 'kitti')
    names = ok.write_synthetic_kitti(root, [41, 42], 6000)
    ckpt = tmp_path / 'ckpt'
    ckpt.mkdir()
    shutil.copy(os.path.join(GOLDEN, 'config_%s.json' % cfg_name), str(ckpt / 'config'))
    # the trained car model does not fire on the synthetic boxes, so the test checkpoint = the real weights with the
    # object-class logit biases raised by 7: a couple of hundred candidates per frame, clustered on the obstacles.
    # Both pipelines read the same file; this is a pipeline-parity test, not a detection-quality test.
    w = dict(np.load(os.path.join(GOLDEN, 'weights_%s.npz' % cfg_name)))
    b = w['output/predictor/cls/fully_connected_1/biases'].copy()
    b[1:-1] += 7.0
    w['output/predictor/cls/fully_connected_1/biases'] = b
    np.savez(str(ckpt / 'weights.npz'), **w)
    out_dir = str(tmp_path / 'out')
    times = twin.main([str(ckpt), '--test', '--dataset_root_dir', root, '--output_dir', out_dir])
    assert set(times) >= {'fetch input', 'gen graph', 'gnn inference', 'decode box', 'nms', 'total'}   # run.py's timers
    with open(str(ckpt / 'config')) as f:
        config = json.load(f)
    weights = dict(np.load(str(ckpt / 'weights.npz')))
    total_rows = 0
    for name in names:
        # ---- the oracle pipeline for this frame ---------------------------------------------------
        velo = np.fromfile(os.path.join(root, 'velodyne/testing/velodyne', name + '.bin'), dtype=np.float32).reshape(-1, 4)
        calib = ok.parse_calib(os.path.join(root, 'calib/testing/calib', name + '.txt'))
        xyz, attr = ok.cam_points_in_image(velo, calib, 1242, 375)
        coords, kp, edges = ograph.gen_multi_level_local_graph_v3(xyz, **config['runtime_graph_gen_kwargs'])
        logits, boxes, probs = cpu_reference.predict(weights, config['model_kwargs']['layer_configs'],
                                                     config['num_classes'], 7, attr, coords, kp, edges)
        last = coords[config['model_kwargs']['layer_configs'][-1]['graph_level'] + 1]
        dec = pp.decode_boxes(boxes, last, pp.LABEL_MAPS[config['label_method']])
        lab, bx, sc, idx = pp.select_candidates(probs, dec, config['num_classes'])
        want = []
        if len(lab):
            k_lab, k_box, k_sc, _ = pp.nms_boxes_3d_uncertainty(lab, bx, sc, config['nms_overlapped_thres'])
            want = ok.kitti_labels(k_lab, k_box, k_sc, last[idx // config['num_classes']], calib, config['label_method'])
        want_rows = ok.parse_kitti_text(ok.format_kitti(want))
        with open(os.path.join(out_dir, 'data', name + '.txt')) as f:
            text = f.read()
        got_rows = ok.parse_kitti_text(text)
        assert text.endswith('\n')
        assert len(got_rows) == len(want_rows), (name, len(got_rows), len(want_rows))
        # NMS output order = score order of the candidates; rescored values may tie-break differently: match by box
        for (n1, v1), (n2, v2) in zip(got_rows, want_rows):
            assert n1 == n2
            assert np.allclose(v1, v2, rtol=2e-3, atol=2e-3), (name, v1, v2)
        total_rows += len(got_rows)
    assert total_rows > 0, 'the synthetic frames produced no detection at all: the test would be vacuous'
