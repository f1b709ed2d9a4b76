"""KITTI-facing stages either side of the hot path, CPU part.

Fixtures (tools/make_golden.py kitti): the reference's OWN dataset/kitti_dataset.py run on a synthetic KITTI-format
frame (tests/golden/kitti_io.npz) and run.py:361-429 assembled from the reference's functions on the post_car
detections (tests/golden/kitti_result_car.txt).  Checked here:
* the oracle restatements (oracle/kitti.py) - the checkers used on the GPU box,
* the PRODUCT's host-side code of this stage: calibration parser, box tests, KITTI label conversion and file writer
  of the eager run.py twin (NumPy, as in the reference - no GPU needed)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import kitti as ok

IO = dict(np.load(os.path.join(GOLDEN, 'kitti_io.npz')))
CALIB_FILE = os.path.join(GOLDEN, 'kitti_calib.txt')


def _same_rows(text_a, text_b, tol=1e-9):
    a, b = ok.parse_kitti_text(text_a), ok.parse_kitti_text(text_b)
    assert len(a) == len(b)
    for (na, va), (nb, vb) in zip(a, b):
        assert na == nb
        assert np.allclose(va, vb, rtol=tol, atol=tol)


@pytest.mark.parametrize('which', ['oracle', 'product'])
def test_calibration_parser(which):
    if which == 'oracle':
        calib = ok.parse_calib(CALIB_FILE)
    else:
        from pointgnn_b200.dataset import kitti_dataset
        calib = kitti_dataset.parse_calib(CALIB_FILE)
    for k in ('velo_to_cam', 'cam_to_image', 'cam_to_velo', 'P2'):
        assert np.array_equal(np.asarray(calib[k]), IO['calib_' + k]), k


def test_oracle_input_stage_matches_reference():
    calib = ok.parse_calib(CALIB_FILE)
    h, w = IO['image'].shape[:2]
    xyz, attr = ok.cam_points_in_image(IO['velo'], calib, w, h, IO['image'])
    assert np.array_equal(xyz, IO['xyz']) and np.array_equal(attr, IO['attr'])
    xyz1, attr1 = ok.cam_points_in_image(IO['velo'], calib, w, h)
    assert np.array_equal(xyz1, IO['xyz']) and np.array_equal(attr1, IO['attr'][:, :1])
    assert 0 < len(xyz) < len(IO['velo'])            # the crop really removes points


def _detections():
    g = dict(np.load(os.path.join(GOLDEN, 'post_car.npz')))
    cand_xyz = g['points_xyz'][g['cand_index'] // 4]
    return g['uncertainty_label'], g['uncertainty_box'], g['uncertainty_score'], cand_xyz


def test_oracle_writer_matches_reference_text():
    labels, boxes, scores, cand_xyz = _detections()
    calib = ok.parse_calib(CALIB_FILE)
    text = ok.format_kitti(ok.kitti_labels(labels, boxes, scores, cand_xyz, calib, 'Car'))
    with open(os.path.join(GOLDEN, 'kitti_result_car.txt')) as f:
        want = f.read()
    _same_rows(text, want)
    assert text.endswith('\n\n') and want.endswith('\n\n')


def test_product_writer_matches_reference_text(tmp_path):
    """run.py:361-429 of the eager twin: kitti_labels + write_kitti_file."""
    from pointgnn_b200 import run as twin
    from pointgnn_b200.dataset import kitti_dataset
    labels, boxes, scores, cand_xyz = _detections()
    calib = kitti_dataset.parse_calib(CALIB_FILE)
    pred = twin.kitti_labels(labels, boxes, scores, cand_xyz, calib, 'Car', True)
    out = tmp_path / 'eval' / 'data' / '000000.txt'
    twin.write_kitti_file(str(out), pred)
    with open(os.path.join(GOLDEN, 'kitti_result_car.txt')) as f:
        want = f.read()
    text = out.read_text()
    _same_rows(text, want)
    # format: 16 blank-separated fields, every field followed by a blank, an empty line at the end (run.py:424-429)
    first = text.split('\n')[0]
    assert first.endswith(' ') and len(first.split()) == 16 and first.split()[1:4] == ['-1', '-1', '0']
    # no detection -> a file holding one newline (run.py:496-500)
    twin.write_kitti_file(str(tmp_path / 'e' / 'data' / 'x.txt'), [])
    assert (tmp_path / 'e' / 'data' / 'x.txt').read_text() == '\n'
    # without re-scoring the NMS score is written unchanged
    pred2 = twin.kitti_labels(labels, boxes, scores, cand_xyz, calib, 'Car', False)
    assert len(pred2) == len(pred) and all(abs(a[-1] - s) < 1e-12 for a, s in zip(pred2, [p[-1] for p in pred2]))


def test_product_box_tests_match_oracle():
    from pointgnn_b200.dataset import kitti_dataset
    labels, boxes, scores, cand_xyz = _detections()
    for b in boxes[:5]:
        lab = dict(zip(('x3d', 'y3d', 'z3d', 'length', 'height', 'width', 'yaw'), b))
        assert np.array_equal(kitti_dataset.sel_xyz_in_box3d(lab, cand_xyz), ok.sel_xyz_in_box3d(lab, cand_xyz))
        n1, l1, u1 = kitti_dataset.box3d_to_normals(lab)
        n2, l2, u2 = ok.box3d_to_normals(lab)
        assert np.allclose(n1, n2) and np.allclose(l1, l2) and np.allclose(u1, u2)


def test_twin_command_line_matches_reference():
    """The eager twin keeps run.py's command line (run.py:25-49)."""
    import inspect
    from pointgnn_b200 import run as twin
    src = inspect.getsource(twin.main)
    for flag in ("'checkpoint_path'", "'-l', '--level'", "'--test'", "'--no-box-merge'", "'--no-box-score'",
                 "'--dataset_root_dir'", "'--dataset_split_file'", "'--output_dir'"):
        assert flag in src, flag
