"""CPU oracle for the step right after the hot path: box decoding + NMS / merge / rescore
(reference run.py:265-325, models/box_encoding.py:265-299,
models/nms.py:9-27, 64-88, 133-170, 256-270).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Two things live here:

* ``reference_modules()`` - imports the reference's OWN ``models/box_encoding.py`` and
  ``models/nms.py`` from /root/reference (build container only).  ``nms.py`` needs
  ``shapely.geometry.Polygon`` (absent here, no network); the only members it uses are
  ``Polygon(points).area`` and ``p1.intersection(p2).area`` on the 4-corner footprints of boxes, so a
  stand-in ``Polygon`` for CONVEX polygons (shoelace area, Sutherland-Hodgman clipping, float64) is
  installed as ``shapely.geometry`` before the import.  tools/make_golden.py runs the reference code
  that way and commits the results as tests/golden/post_*.npz.
* a NumPy restatement of the same pipeline (``decode_boxes``, ``overlapped_boxes_3d_fast_poly``,
  ``nms_boxes_3d_uncertainty`` ...) that travels to the GPU box, checked against those fixtures in
  tests/test_postprocess_cpu.py and used as the checker of the CUDA kernels.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'

# reference models/box_encoding.py:211-229 (l, h, w medians used by the class-aware codecs)
MEDIAN_OBJECT_SIZE = {
    'Cyclist': (1.76, 1.75, 0.6),
    'Van': (4.98, 2.13, 1.88),
    'Tram': (14.66, 3.61, 2.6),
    'Car': (3.88, 1.5, 1.63),
    'Misc': (2.52, 1.65, 1.51),
    'Pedestrian': (0.88, 1.77, 0.65),
    'Truck': (10.81, 3.34, 2.63),
    'Person_sitting': (0.75, 1.26, 0.59),
}


# ---------------------------------------------------------------------------------------------
# convex polygon geometry (stand-in for shapely on box footprints)
# ---------------------------------------------------------------------------------------------
def polygon_area(pts):
    """Shoelace formula, absolute value (shapely's .area is unsigned)."""
    pts = np.asarray(pts, dtype=np.float64)
    if len(pts) < 3:
        return 0.0
    x, y = pts[:, 0], pts[:, 1]
    return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))


def _signed_area(pts):
    x, y = pts[:, 0], pts[:, 1]
    return 0.5 * float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))


def clip_convex(subject, clip):
    """Sutherland-Hodgman: the part of convex polygon ``subject`` inside convex polygon ``clip``."""
    subject = np.asarray(subject, dtype=np.float64)
    clip = np.asarray(clip, dtype=np.float64)
    if _signed_area(clip) < 0:
        clip = clip[::-1]
    out = [tuple(p) for p in subject]
    n = len(clip)
    for i in range(n):
        if not out:
            break
        a, b = clip[i], clip[(i + 1) % n]
        ex, ey = b[0] - a[0], b[1] - a[1]
        inp, out = out, []
        for j in range(len(inp)):
            p, q = inp[j], inp[(j + 1) % len(inp)]
            sp = ex * (p[1] - a[1]) - ey * (p[0] - a[0])
            sq = ex * (q[1] - a[1]) - ey * (q[0] - a[0])
            if sp >= 0:
                out.append(p)
            if (sp >= 0) != (sq >= 0):
                t = sp / (sp - sq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return np.array(out, dtype=np.float64).reshape(-1, 2)


class Polygon(object):
    """The subset of shapely.geometry.Polygon the reference's nms.py touches, for convex polygons."""

    def __init__(self, points):
        self.points = np.asarray(points, dtype=np.float64).reshape(-1, 2)

    @property
    def area(self):
        return polygon_area(self.points)

    def intersection(self, other):
        if len(self.points) < 3 or len(other.points) < 3 or self.area == 0.0 or other.area == 0.0:
            return Polygon(np.zeros((0, 2)))
        return Polygon(clip_convex(self.points, other.points))


def reference_modules():
    """-> (box_encoding, nms): the reference's unmodified modules, with the Polygon stand-in as shapely."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError('the reference tree is not present (only in the build container)')
    if 'shapely' not in sys.modules:
        shapely = types.ModuleType('shapely')
        geometry = types.ModuleType('shapely.geometry')
        geometry.Polygon = Polygon
        shapely.geometry = geometry
        sys.modules['shapely'] = shapely
        sys.modules['shapely.geometry'] = geometry
    import importlib.util
    mods = []
    for name in ('box_encoding', 'nms'):
        spec = importlib.util.spec_from_file_location('pg_reference_%s' % name,
                                                      os.path.join(REFERENCE_ROOT, 'models', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


# ---------------------------------------------------------------------------------------------
# restatement: box decoding (box_encoding.py:265-299) and the run.py:265-286 candidate filter
# ---------------------------------------------------------------------------------------------
LABEL_MAPS = {   # run.py:243-250
    'yaw': {'Background': 0, 'Car': 1, 'Pedestrian': 3, 'Cyclist': 5, 'DontCare': 7},
    'Car': {'Background': 0, 'Car': 1, 'DontCare': 3},
    'Pedestrian_and_Cyclist': {'Background': 0, 'Pedestrian': 1, 'Cyclist': 3, 'DontCare': 5},
}


def class_size_table(label_map, num_classes):
    """Per class label: (l, h, w, yaw offset) of classaware_all_class_box_decoding; NaN rows = not decoded."""
    table = np.full((num_classes, 4), np.nan, dtype=np.float64)
    for name, label in label_map.items():
        if name in ('Background', 'DontCare'):
            continue
        l, h, w = MEDIAN_OBJECT_SIZE[name]
        table[label] = (l, h, w, 0.0)                 # "horizontal" class
        table[label + 1] = (l, h, w, 0.5 * np.pi)     # "vertical" class
    return table


def decode_boxes(box_encodings, points_xyz, label_map):
    """classaware_all_class_box_decoding for every (vertex, class): [K, C, 7] float32 encodings at the K
    last-level vertices -> [K, C, 7] float64-valued decoded boxes (x, y, z, l, h, w, yaw), computed as the
    reference does (float32 encodings, float64 Python scalars -> NumPy keeps float32 arrays)."""
    k, c, _ = box_encodings.shape
    enc = box_encodings.astype(np.float32)
    out = np.copy(enc)
    table = class_size_table(label_map, c)
    for cls in range(c):
        if np.isnan(table[cls, 0]):
            continue
        l, h, w, yaw0 = table[cls]
        out[:, cls, 0] = enc[:, cls, 0] * l
        out[:, cls, 1] = enc[:, cls, 1] * h
        out[:, cls, 2] = enc[:, cls, 2] * w
        out[:, cls, 3] = np.exp(enc[:, cls, 3]) * l
        out[:, cls, 4] = np.exp(enc[:, cls, 4]) * h
        out[:, cls, 5] = np.exp(enc[:, cls, 5]) * w
        out[:, cls, 6] = enc[:, cls, 6] * (np.pi * 0.25) + (0.5 * np.pi if yaw0 else 0.0)
    out[:, :, :3] += points_xyz.astype(np.float32)[:, None, :]
    return out


def select_candidates(probs, decoded, num_classes):
    """run.py:281-296: class c of vertex v is a candidate iff 0 < c < C-1 and prob > 1/C; "vertical" labels
    are folded onto their class (2 -> 1, 4 -> 3, 6 -> 5).  -> (labels, boxes [B,7], scores, flat indices)."""
    k = probs.shape[0]
    labels = np.tile(np.arange(num_classes)[None, :], (k, 1)).reshape(-1)
    p = probs.reshape(-1)
    mask = (labels > 0) & (labels < num_classes - 1) & (p > 1.0 / num_classes)
    idx = np.nonzero(mask)[0]
    lab = labels[idx].copy()
    for a, b in ((2, 1), (4, 3), (6, 5)):
        lab[lab == a] = b
    return lab, decoded.reshape(-1, 7)[idx].copy(), p[idx].copy(), idx


# ---------------------------------------------------------------------------------------------
# restatement: nms.py
# ---------------------------------------------------------------------------------------------
def boxes_3d_to_corners(boxes_3d):
    """nms.py:9-27 -> [B, 8, 3] float64.  Like the reference, trigonometry and the half extents are evaluated
    in the dtype of ``boxes_3d`` (float32 at run.py's call sites) and only then promoted to float64."""
    boxes_3d = np.asarray(boxes_3d).reshape(-1, 7)
    out = np.zeros((len(boxes_3d), 8, 3))
    for i, (x, y, z, l, h, w, yaw) in enumerate(boxes_3d):
        r = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        c = np.array([[l / 2, 0.0, w / 2], [l / 2, 0.0, -w / 2], [-l / 2, 0.0, -w / 2], [-l / 2, 0.0, w / 2],
                      [l / 2, -h, w / 2], [l / 2, -h, -w / 2], [-l / 2, -h, -w / 2], [-l / 2, -h, w / 2]])
        out[i] = c.dot(r.T) + np.array([x, y, z])
    return out


def overlapped_boxes_3d_fast_poly(single_box, box_list):
    """nms.py:64-88: 3-D IoU of one box against a list, footprints as exact convex polygons."""
    box_list = np.asarray(box_list, dtype=np.float64).reshape(-1, 8, 3)
    overlap = np.zeros(len(box_list))
    if len(box_list) == 0:
        return overlap
    mx0, mn0 = single_box.max(axis=0), single_box.min(axis=0)
    mx, mn = box_list.max(axis=1), box_list.min(axis=1)
    apart = np.any((mx0 < mn) | (mn0 > mx), axis=1)
    p1 = Polygon(single_box[:4, [0, 2]])
    area1 = p1.area
    for i in range(len(box_list)):
        if apart[i]:
            continue
        p2 = Polygon(box_list[i][:4, [0, 2]])
        shared = p1.intersection(p2).area
        area2 = p2.area
        shared_y = min(mx[i, 1], mx0[1]) - max(mn[i, 1], mn0[1])
        inter = shared_y * shared
        union = (mx[i, 1] - mn[i, 1]) * area2 + (mx0[1] - mn0[1]) * area1
        overlap[i] = np.float32(inter) / (union - inter)
    return overlap


def nms_boxes_3d_uncertainty(class_labels, boxes, scores, overlapped_thres=0.5, merge=True, rescore=True):
    """nms.py:133-170 + 256-270 (top_k = -1): sort by score, greedy suppression within a class; the kept box
    becomes the coordinate-wise MEDIAN of itself and the boxes it suppresses (merge) and its score grows by
    sum(score_j * IoU(median box, box_j)) (rescore).  merge / rescore False give nms.py:172-240's variants.
    -> (labels, boxes, scores, order indices into the input)."""
    order = np.argsort(-scores)                      # nms.py:93 (default quicksort; ties are not a contract)
    classes = np.asarray(class_labels)[order].copy()
    scores = np.asarray(scores)[order].copy()        # dtypes kept: float32 at run.py's call sites
    boxes = np.asarray(boxes)[order].copy()
    corners = boxes_3d_to_corners(boxes)
    keep = np.ones(len(scores), dtype=bool)
    for i in range(len(scores) - 1):
        if not keep[i]:
            continue
        valid = np.nonzero(keep[i + 1:])[0] + i + 1
        ov = overlapped_boxes_3d_fast_poly(corners[i], corners[valid])
        rem = valid[(ov > overlapped_thres) & (classes[valid] == classes[i])]
        if merge:
            boxes[i] = np.median(np.concatenate([boxes[rem], boxes[[i]]], axis=0), axis=0)
        if rescore:
            mean_corners = boxes_3d_to_corners(boxes[[i]])[0]
            scores[i] += np.sum(scores[rem] * overlapped_boxes_3d_fast_poly(mean_corners, corners[rem]))
        keep[rem] = False
    sel = np.nonzero(keep)[0]
    return classes[sel], boxes[sel], scores[sel], order[sel]


def postprocess_frame(probs, box_encodings, points_xyz, label_method, nms_overlapped_thres, merge=True, rescore=True):
    """run.py:265-325 for one frame -> (class labels [D], boxes [D,7], scores [D])."""
    num_classes = probs.shape[1]
    decoded = decode_boxes(box_encodings, points_xyz, LABEL_MAPS[label_method])
    labels, boxes, scores, _ = select_candidates(probs, decoded, num_classes)
    if len(labels) == 0:
        return labels, boxes, scores
    lab, bx, sc, _ = nms_boxes_3d_uncertainty(labels, boxes, scores, nms_overlapped_thres, merge, rescore)
    return lab, bx, sc


# ---------------------------------------------------------------------------------------------
# seeded synthetic network outputs (clusters of proposals around a few objects)
# ---------------------------------------------------------------------------------------------
def synthetic_outputs(seed, num_objects=12, per_object=25, num_classes=4, spread=0.6):
    """-> (points_xyz [K,3] f32, box_encodings [K,C,7] f32, probs [K,C] f32): K = num_objects * per_object
    vertices scattered around object centres, each proposing a noisy box per class."""
    rng = np.random.default_rng(seed)
    centers = np.c_[rng.uniform(-20, 20, num_objects), rng.uniform(0.5, 2, num_objects), rng.uniform(5, 60, num_objects)]
    k = num_objects * per_object
    pts = (np.repeat(centers, per_object, axis=0) + rng.normal(0, spread, (k, 3))).astype(np.float32)
    enc = rng.normal(0, 0.15, (k, num_classes, 7)).astype(np.float32)
    probs = rng.dirichlet(np.ones(num_classes) * 0.6, k).astype(np.float32)
    return pts, enc, probs


def reference_postprocess_frame(be, nms, probs, box_encodings, points_xyz, label_method, thres, variant='uncertainty'):
    """run.py:265-325 executed with the reference's OWN modules (be, nms = reference_modules())."""
    num_classes = probs.shape[1]
    label_map = LABEL_MAPS[label_method]
    box_probs = probs
    box_labels = np.tile(np.expand_dims(np.arange(num_classes), axis=0), (box_probs.shape[0], 1)).reshape((-1))
    box_probs = box_probs.reshape((-1))
    pred_boxes = box_encodings.reshape((-1, 1, 7))
    xyz = np.tile(np.expand_dims(points_xyz, axis=1), (1, num_classes, 1)).reshape((-1, 3))
    decoded = be.classaware_all_class_box_decoding(np.expand_dims(box_labels, axis=1), xyz, pred_boxes, label_map)
    mask = (box_labels > 0) * (box_labels < num_classes - 1) * (box_probs > 1. / num_classes)
    idx = np.nonzero(mask)[0]
    lab = box_labels[idx]
    sc = box_probs[idx]
    dec = decoded[idx, 0]
    lab[lab == 2] = 1
    lab[lab == 4] = 3
    lab[lab == 6] = 5
    fn = {'uncertainty': nms.nms_boxes_3d_uncertainty, 'merge_only': nms.nms_boxes_3d_merge_only,
          'score_only': nms.nms_boxes_3d_score_only, 'plain': nms.nms_boxes_3d}[variant]
    out = fn(lab, dec, sc, overlapped_fn=nms.overlapped_boxes_3d_fast_poly, overlapped_thres=thres, appr_factor=100.0,
             top_k=-1, attributes=np.arange(len(idx)))
    return dict(label=out[0], box=out[1], score=out[2], nms_index=out[3], cand_index=idx, decoded=decoded[:, 0])
