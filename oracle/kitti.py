"""CPU oracle for the stages either side of the hot path that touch KITTI files
(SURVEY.md section 8f-2 / 8f-3): the input stage and the KITTI result writer.

TEST INFRASTRUCTURE (see oracle/__init__.py).

* ``reference_dataset_module()`` imports the reference's OWN ``dataset/kitti_dataset.py`` (build container only;
  ``open3d`` is stubbed - it is used for visualisation only, OpenCV is present).  tools/make_golden.py uses it to
  commit tests/golden/kitti_*.npz (calibration matrices, camera points in image, projections, box tests).
* restatements that travel to the GPU box: ``parse_calib`` (kitti_dataset.py:483-522), ``cam_points_in_image``
  (:587-609, 666-689, 998-1006, 1036-1052), ``kitti_labels`` / ``format_kitti`` (run.py:88-100, 361-429),
  checked against those fixtures in tests/test_kitti_cpu.py.
* ``write_synthetic_kitti`` - a tiny KITTI-format directory (velodyne .bin, calib .txt, image .png) made from
  oracle/synth.py frames, for the end-to-end test of the eager run.py twin.
"""
import os
import sys
import types

import numpy as np

from . import postprocess as opost
from . import synth

REFERENCE_ROOT = '/root/reference'

# a real KITTI calibration (object training set, frame 000000), rounded: only its structure matters here
CALIB_TEXT = """P0: 7.070493e+02 0.000000e+00 6.040814e+02 0.000000e+00 0.000000e+00 7.070493e+02 1.805066e+02 0.000000e+00 0.000000e+00 0.000000e+00 1.000000e+00 0.000000e+00
P1: 7.070493e+02 0.000000e+00 6.040814e+02 -3.797842e+02 0.000000e+00 7.070493e+02 1.805066e+02 0.000000e+00 0.000000e+00 0.000000e+00 1.000000e+00 0.000000e+00
P2: 7.070493e+02 0.000000e+00 6.040814e+02 4.575831e+01 0.000000e+00 7.070493e+02 1.805066e+02 -3.454157e-01 0.000000e+00 0.000000e+00 1.000000e+00 4.981016e-03
P3: 7.070493e+02 0.000000e+00 6.040814e+02 -3.341081e+02 0.000000e+00 7.070493e+02 1.805066e+02 2.330660e+00 0.000000e+00 0.000000e+00 1.000000e+00 3.201153e-03
R0_rect: 9.999128e-01 1.009263e-02 -8.511932e-03 -1.012729e-02 9.999406e-01 -4.037671e-03 8.470675e-03 4.123522e-03 9.999556e-01
Tr_velo_to_cam: 6.927964e-03 -9.999722e-01 -2.757829e-03 -2.457729e-02 -1.162982e-03 2.749836e-03 -9.999955e-01 -6.127237e-02 9.999753e-01 6.931141e-03 -1.143899e-03 -3.321029e-01
Tr_imu_to_velo: 9.999976e-01 7.553071e-04 -2.035826e-03 -8.086759e-01 -7.854027e-04 9.998898e-01 -1.482298e-02 3.195559e-01 2.024406e-03 1.482454e-02 9.998881e-01 -7.997231e-01
"""


def reference_dataset_module():
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError('the reference tree is not present (only in the build container)')
    if 'open3d' not in sys.modules:
        sys.modules['open3d'] = types.ModuleType('open3d')
    import importlib.util
    spec = importlib.util.spec_from_file_location('pg_reference_kitti_dataset',
                                                  os.path.join(REFERENCE_ROOT, 'dataset', 'kitti_dataset.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ---------------------------------------------------------------------------------------------
# restatements
# ---------------------------------------------------------------------------------------------
def parse_calib(calib_file):
    """KittiDataset.get_calib, kitti_dataset.py:483-522."""
    with open(calib_file, 'r') as f:
        calib = {}
        for line in f:
            fields = line.split(' ')
            calib[fields[0].rstrip(':')] = np.array(fields[1:], dtype=np.float32)
    calib['P2'] = calib['P2'].reshape(3, 4)
    calib['R0_rect'] = calib['R0_rect'].reshape(3, 3)
    calib['Tr_velo_to_cam'] = calib['Tr_velo_to_cam'].reshape(3, 4)
    calib['velo_to_rect'] = np.vstack([calib['Tr_velo_to_cam'], [0, 0, 0, 1]])
    calib['cam_to_image'] = np.hstack([calib['P2'][:, 0:3], [[0], [0], [0]]])
    calib['rect_to_cam'] = np.hstack([calib['R0_rect'],
                                      np.matmul(np.linalg.inv(calib['P2'][:, 0:3]), calib['P2'][:, [3]])])
    calib['rect_to_cam'] = np.vstack([calib['rect_to_cam'], [0, 0, 0, 1]])
    calib['velo_to_cam'] = np.matmul(calib['rect_to_cam'], calib['velo_to_rect'])
    calib['cam_to_velo'] = np.linalg.inv(calib['velo_to_cam'])
    calib['velo_to_image'] = np.matmul(calib['cam_to_image'], calib['velo_to_cam'])
    return calib


def cam_points_in_image(velo_data, calib, width, height, image=None):
    """get_cam_points_in_image_with_rgb (kitti_dataset.py:666-689) from the raw [M,4] velodyne array.
    -> (xyz [N,3] float32, attr [N,1] or [N,4] float32)."""
    xyz = velo_data[:, :3]
    cam = np.matmul(xyz, np.transpose(calib['velo_to_cam'])[:3, :3].astype(np.float32))
    cam += np.transpose(calib['velo_to_cam'])[[3], :3].astype(np.float32)
    front = cam[:, 2] > 0.1
    cam, attr = cam[front], velo_data[:, [3]][front]
    img = np.matmul(np.hstack([cam, np.ones([cam.shape[0], 1])]), np.transpose(calib['cam_to_image']))
    img = img / img[:, [2]]
    keep = np.logical_and.reduce([img[:, 0] > 0, img[:, 0] < width, img[:, 1] > 0, img[:, 1] < height])
    cam, attr, img = cam[keep], attr[keep], img[keep]
    if image is not None:
        rgb = image[np.int32(img[:, 1]), np.int32(img[:, 0]), ::-1].astype(np.float32) / 255
        attr = np.hstack([attr, rgb])
    return cam, attr


def box3d_to_normals(label):
    """kitti_dataset.py:85-141."""
    yaw = label['yaw']
    r = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    h, w, l = label['height'], label['width'], label['length']
    corners = np.array([[l / 2, 0.0, w / 2], [l / 2, 0.0, -w / 2], [-l / 2, 0.0, -w / 2], [-l / 2, 0.0, w / 2],
                        [l / 2, -h, w / 2], [l / 2, -h, -w / 2], [-l / 2, -h, -w / 2], [-l / 2, -h, w / 2]])
    p = corners.dot(np.transpose(r)) + np.array([label['x3d'], label['y3d'], label['z3d']])
    wx, wy, wz = p[[0], :] - p[[4], :], p[[0], :] - p[[1], :], p[[0], :] - p[[3], :]
    lower = np.concatenate([np.matmul(wx, p[4, :]), np.matmul(wy, p[1, :]), np.matmul(wz, p[3, :])])
    upper = np.concatenate([np.matmul(wx, p[0, :]), np.matmul(wy, p[0, :]), np.matmul(wz, p[0, :])])
    return np.concatenate([wx, wy, wz], axis=0), lower, upper


def sel_xyz_in_box3d(label, xyz):
    """kitti_dataset.py:143-162."""
    normals, lower, upper = box3d_to_normals(label)
    proj = np.matmul(xyz, np.transpose(normals))
    return np.logical_and.reduce([(proj[:, i] > lower[i]) & (proj[:, i] < upper[i]) for i in range(3)])


def occlusion(label, xyz):
    """run.py:88-100."""
    if xyz.shape[0] == 0:
        return 0
    normals, lower, upper = box3d_to_normals(label)
    proj = np.matmul(xyz, np.transpose(normals))
    rates = [(np.max(proj[:, i]) - np.min(proj[:, i])) / (upper[i] - lower[i]) for i in range(3)]
    return rates[0] * rates[1] * rates[2]


CLASS_NAMES = {   # run.py:371-383
    'yaw': ['Background', 'Car', 'Car', 'Pedestrian', 'Pedestrian', 'Cyclist', 'Cyclist', 'DontCare'],
    'Car': ['Background', 'Car', 'Car', 'DontCare'],
    'Pedestrian_and_Cyclist': ['Background', 'Pedestrian', 'Pedestrian', 'Cyclist', 'Cyclist', 'DontCare'],
}


def kitti_labels(class_labels, boxes, scores, candidate_xyz, calib, label_method, use_box_score=True):
    """run.py:361-408."""
    corners = opost.boxes_3d_to_corners(boxes)
    out = []
    for i in range(len(corners)):
        img = np.matmul(np.hstack([corners[i], np.ones([8, 1])]), np.transpose(calib['cam_to_image']))
        xy = (img / img[:, [2]])[:, :2]
        xmin, ymin = np.amin(xy, axis=0)
        xmax, ymax = np.amax(xy, axis=0)
        cx0, cy0, cx1, cy1 = max(xmin, 0.0), max(ymin, 0.0), min(xmax, 1242.0), min(ymax, 375.0)
        if 1.0 - (cy1 - cy0) * (cx1 - cx0) / ((ymax - ymin) * (xmax - xmin)) > 0.4:
            continue
        x3d, y3d, z3d, l, h, w, yaw = boxes[i]
        score = scores[i]
        if use_box_score:
            lab = {'x3d': x3d, 'y3d': y3d, 'z3d': z3d, 'yaw': yaw, 'height': h, 'width': w, 'length': l}
            inside = sel_xyz_in_box3d(lab, candidate_xyz)
            score = (1 + occlusion(lab, candidate_xyz[inside])) * score
        out.append((CLASS_NAMES[label_method][class_labels[i]], -1, -1, 0, cx0, cy0, cx1, cy1, h, w, l, x3d, y3d, z3d,
                    yaw, score))
    return out


def format_kitti(pred_labels):
    """run.py:421-429: the text of one result file."""
    return ''.join(''.join(str(field) + ' ' for field in lab) + '\n' for lab in pred_labels) + '\n'


def parse_kitti_text(text):
    """-> list of (class name, [15 floats]) of a result file."""
    rows = []
    for line in text.split('\n'):
        fields = line.split()
        if fields:
            rows.append((fields[0], [float(v) for v in fields[1:]]))
    return rows


# ---------------------------------------------------------------------------------------------
# synthetic KITTI directory
# ---------------------------------------------------------------------------------------------
def write_synthetic_kitti(root, frame_seeds, num_points=6000, test_split=True):
    """A KITTI-object-style tree under ``root`` (image/, velodyne/, calib/ for the `testing` split as run.py --test
    expects them) whose clouds are oracle/synth.py frames mapped back into velodyne coordinates, so that the
    camera-frame points the pipeline sees are KITTI-shaped.  Returns the frame names."""
    import cv2
    split = 'testing' if test_split else 'training'
    dirs = {k: os.path.join(root, k, split, v) for k, v in (('image', 'image_2'), ('velodyne', 'velodyne'), ('calib', 'calib'))}
    for d in dirs.values():
        os.makedirs(d, exist_ok=True)
    names = []
    tmp_calib = os.path.join(root, '_calib_tmp.txt')
    with open(tmp_calib, 'w') as f:
        f.write(CALIB_TEXT)
    calib = parse_calib(tmp_calib)
    os.remove(tmp_calib)
    for i, seed in enumerate(frame_seeds):
        name = '%06d' % i
        names.append(name)
        xyz, intensity = synth.lidar_frame(seed, num_points)
        velo = np.matmul(np.hstack([xyz.astype(np.float64), np.ones([len(xyz), 1])]), np.transpose(calib['cam_to_velo']))[:, :3]
        # a few points behind the sensor / outside the image so that the crop has something to remove
        rng = np.random.default_rng(seed)
        extra = np.c_[rng.uniform(-30, 5, 400), rng.uniform(-40, 40, 400), rng.uniform(-2, 1, 400)]
        data = np.vstack([np.hstack([velo, intensity]), np.hstack([extra, rng.uniform(0, 1, (400, 1))])]).astype(np.float32)
        data = data[rng.permutation(len(data))]
        data.tofile(os.path.join(dirs['velodyne'], name + '.bin'))
        with open(os.path.join(dirs['calib'], name + '.txt'), 'w') as f:
            f.write(CALIB_TEXT)
        img = rng.integers(0, 255, (375, 1242, 3), dtype=np.uint8)
        cv2.imwrite(os.path.join(dirs['image'], name + '.png'), img)
    return names
