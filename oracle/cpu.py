"""numpy front-end of the C oracle (`dgod_oracle.c`).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Every function takes/returns numpy arrays (fp32 / int64) and mirrors one reference operation;
the docstrings cite the reference lines the C code restates.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "libdgod_oracle.so"


def build(force: bool = False) -> Path:
    src = _DIR / "dgod_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "-B", "libdgod_oracle.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.o_nms.restype = C.c_int
        _lib.o_batched_nms.restype = C.c_int
        _lib.o_rpn_filter_image.restype = C.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------- boxes
def box_iou(b1, b2):
    """TV ops/boxes.py:344-370."""
    b1, b2 = _f32(b1).reshape(-1, 4), _f32(b2).reshape(-1, 4)
    out = np.empty((len(b1), len(b2)), np.float32)
    lib().o_box_iou(_p(b1), C.c_int(len(b1)), _p(b2), C.c_int(len(b2)), _p(out))
    return out


def nms(boxes, scores, thr):
    """torchvision::nms CPU kernel (TV ops/boxes.py:20-48)."""
    boxes, scores = _f32(boxes).reshape(-1, 4), _f32(scores)
    keep = np.empty(len(boxes), np.int64)
    n = lib().o_nms(_p(boxes), _p(scores), C.c_int(len(boxes)), C.c_double(thr), _p(keep))
    return keep[:n].copy()


def batched_nms(boxes, scores, groups, thr, mode=-1):
    """TV ops/boxes.py:51-120; mode -1 = torchvision's CPU rule, 0 vanilla, 1 coordinate trick."""
    boxes, scores, groups = _f32(boxes).reshape(-1, 4), _f32(scores), _i64(groups)
    keep = np.empty(len(boxes), np.int64)
    n = lib().o_batched_nms(_p(boxes), _p(scores), _p(groups), C.c_int(len(boxes)), C.c_double(thr),
                            C.c_int(mode), _p(keep))
    return keep[:n].copy()


def clip_boxes(boxes, img_h, img_w):
    """TV ops/boxes.py:149-182."""
    b = _f32(boxes).reshape(-1, 4).copy()
    lib().o_clip_boxes(_p(b), C.c_int(len(b)), C.c_float(img_h), C.c_float(img_w))
    return b


# ---------------------------------------------------------------------------------- matcher
def matcher(q, high, low, allow_low_quality):
    """TV models/detection/_utils.py:357-416."""
    q = _f32(q)
    m, n = q.shape
    out = np.empty(n, np.int64)
    lib().o_matcher(_p(q), C.c_int(m), C.c_int(n), C.c_double(high), C.c_double(low),
                    C.c_int(int(allow_low_quality)), _p(out))
    return out


def rpn_assign(gt, anchors, high=0.7, low=0.3):
    """TV models/detection/rpn.py:193-229 for one image -> (matched_idx, labels_f32, matched_boxes)."""
    gt, anchors = _f32(gt).reshape(-1, 4), _f32(anchors).reshape(-1, 4)
    n = len(anchors)
    idx = np.empty(n, np.int64)
    labels = np.empty(n, np.float32)
    mb = np.empty((n, 4), np.float32)
    lib().o_rpn_assign(_p(gt), C.c_int(len(gt)), _p(anchors), C.c_int(n), C.c_double(high),
                       C.c_double(low), _p(idx), _p(labels), _p(mb))
    return idx, labels, mb


def roi_assign(gt, gt_labels, props, high=0.5, low=0.5):
    """TV models/detection/roi_heads.py:580-613 for one image -> (clamped_idx, labels)."""
    gt, props, gt_labels = _f32(gt).reshape(-1, 4), _f32(props).reshape(-1, 4), _i64(gt_labels)
    n = len(props)
    idx = np.empty(n, np.int64)
    labels = np.empty(n, np.int64)
    lib().o_roi_assign(_p(gt), _p(gt_labels), C.c_int(len(gt)), _p(props), C.c_int(n),
                       C.c_double(high), C.c_double(low), _p(idx), _p(labels))
    return idx, labels


# ---------------------------------------------------------------------------------- FCOS
def fcos_assign(anchors, n_first, n_last, gt, gt_labels, radius=1.5):
    """fcos.py:510-548 + fcos.py:136-147 for one image -> (matched_idx, cls_targets, box_targets)."""
    anchors, gt, gt_labels = _f32(anchors).reshape(-1, 4), _f32(gt).reshape(-1, 4), _i64(gt_labels)
    n = len(anchors)
    idx = np.empty(n, np.int64)
    cls = np.empty(n, np.int64)
    bt = np.empty((n, 4), np.float32)
    lib().o_fcos_assign(_p(anchors), C.c_int(n), C.c_int(n_first), C.c_int(n_last),
                        C.c_double(radius), _p(gt), _p(gt_labels), C.c_int(len(gt)), _p(idx),
                        _p(cls), _p(bt))
    return idx, cls, bt


def fcos_loss(cls_logits, bbox_regression, bbox_ctrness, anchors, cls_targets, box_targets, alpha=0.25):
    """fcos.py:149-202 -> float32 [4]: classification, bbox_regression, bbox_ctrness, #foreground."""
    cl, rg, ct = _f32(cls_logits), _f32(bbox_regression), _f32(bbox_ctrness).reshape(-1)
    B, N, Cn = cl.shape
    an, tg, bt = _f32(anchors).reshape(-1, 4), _i64(cls_targets).reshape(-1), _f32(box_targets).reshape(-1, 4)
    out = np.empty(4, np.float32)
    lib().o_fcos_loss(_p(cl), _p(rg), _p(ct), _p(an), _p(tg), _p(bt), C.c_int(B), C.c_int(N), C.c_int(Cn),
                      C.c_float(alpha), _p(out))
    return out


def fcos_candidates(cls_logits, bbox_regression, bbox_ctrness, anchors, num_anchors_per_level, img_h, img_w,
                    score_thresh=0.2, topk=1000):
    """fcos.py:576-597 for one image: head outputs [N,C] / [N,4] / [N] over all levels -> (boxes [L*topk,4], scores,
    labels int64, counts [L]); level l's survivors start at l*topk."""
    cl, rg, ct = _f32(cls_logits), _f32(bbox_regression).reshape(-1, 4), _f32(bbox_ctrness).reshape(-1)
    an = _f32(anchors).reshape(-1, 4)
    off = np.zeros(len(num_anchors_per_level) + 1, np.int32)
    off[1:] = np.cumsum(num_anchors_per_level)
    L = len(num_anchors_per_level)
    boxes = np.empty((L * topk, 4), np.float32)
    scores = np.empty(L * topk, np.float32)
    labels = np.empty(L * topk, np.int64)
    counts = np.empty(L, np.int32)
    lib().o_fcos_candidates(_p(cl), _p(rg), _p(ct), _p(an), C.c_int(cl.shape[1]), _p(off), C.c_int(L), C.c_float(img_h),
                            C.c_float(img_w), C.c_float(score_thresh), C.c_int(topk), _p(boxes), _p(scores), _p(labels),
                            _p(counts))
    return boxes, scores, labels, counts


def balanced_sample(labels, keys, num_pos, batch_size):
    """BalancedPositiveNegativeSampler with explicit keys (TV models/detection/_utils.py:11-71): per image the
    min(#pos, num_pos) positives and min(#neg, batch - #picked_pos) negatives with the smallest keys (ties: lower index),
    returned as ascending index arrays."""
    out = []
    for lab, key in zip(np.asarray(labels), np.asarray(keys, dtype=np.float32)):
        pos, neg = np.nonzero(lab >= 1)[0], np.nonzero(lab == 0)[0]
        p = pos[np.lexsort((pos, key[pos]))][:num_pos]
        n = neg[np.lexsort((neg, key[neg]))][:max(batch_size - len(p), 0)]
        out.append((np.sort(p), np.sort(n)))
    return out


def image_batch(images, mean, std, min_size, max_size, size_divisible=32):
    """GeneralizedRCNNTransform.forward for a list of [C,H,W] images -> ([B,C,Hp,Wp] float32, [(h,w)])."""
    import math
    imgs = [_f32(i) for i in images]
    sizes = []
    for im in imgs:
        h, w = im.shape[1:]
        sf = min(float(min_size) / min(h, w), float(max_size) / max(h, w))
        sizes.append((int(math.floor(h * sf)), int(math.floor(w * sf))))
    ph = int(math.ceil(max(s[0] for s in sizes) / float(size_divisible)) * size_divisible)
    pw = int(math.ceil(max(s[1] for s in sizes) / float(size_divisible)) * size_divisible)
    out = np.empty((len(imgs), imgs[0].shape[0], ph, pw), np.float32)
    m, s_ = _f32(np.asarray(mean)), _f32(np.asarray(std))
    for i, (im, (oh, ow)) in enumerate(zip(imgs, sizes)):
        o = np.empty(out.shape[1:], np.float32)
        lib().o_image_resize_pad(_p(im), C.c_int(im.shape[0]), C.c_int(im.shape[1]), C.c_int(im.shape[2]), C.c_int(oh),
                                 C.c_int(ow), _p(m), _p(s_), _p(o), C.c_int(ph), C.c_int(pw))
        out[i] = o
    return out, sizes


# ---------------------------------------------------------------------------------- RoIAlign
def roi_align_fwd(x, rois, scale, ph, pw, sr, aligned=False):
    """torchvision::roi_align CPU forward (TV ops/roi_align.py:204-260)."""
    x, rois = _f32(x), _f32(rois).reshape(-1, 5)
    B, Cc, H, W = x.shape
    out = np.empty((len(rois), Cc, ph, pw), np.float32)
    lib().o_roi_align_fwd(_p(x), C.c_int(B), C.c_int(Cc), C.c_int(H), C.c_int(W), _p(rois),
                          C.c_int(len(rois)), C.c_float(scale), C.c_int(ph), C.c_int(pw),
                          C.c_int(sr), C.c_int(int(aligned)), _p(out))
    return out


def roi_align_bwd(grad_out, shape, rois, scale, sr, aligned=False):
    """torchvision::_roi_align_backward CPU kernel."""
    g, rois = _f32(grad_out), _f32(rois).reshape(-1, 5)
    B, Cc, H, W = shape
    ph, pw = g.shape[2], g.shape[3]
    gi = np.zeros(shape, np.float32)
    lib().o_roi_align_bwd(_p(g), C.c_int(B), C.c_int(Cc), C.c_int(H), C.c_int(W), _p(rois),
                          C.c_int(len(rois)), C.c_float(scale), C.c_int(ph), C.c_int(pw),
                          C.c_int(sr), C.c_int(int(aligned)), _p(gi))
    return gi


def level_map(boxes, k_min, k_max, s0=224.0, lvl0=4.0, eps=1e-6):
    """TV ops/poolers.py:73-84."""
    boxes = _f32(boxes).reshape(-1, 4)
    out = np.empty(len(boxes), np.int64)
    lib().o_level_map(_p(boxes), C.c_int(len(boxes)), C.c_int(k_min), C.c_int(k_max),
                      C.c_float(s0), C.c_float(lvl0), C.c_float(eps), _p(out))
    return out


def _level_args(feats, scales):
    n = len(feats)
    H = (C.c_int * n)(*[f.shape[2] for f in feats])
    W = (C.c_int * n)(*[f.shape[3] for f in feats])
    sc = (C.c_float * n)(*scales)
    return n, H, W, sc


def msroi_align_fwd(feats, rois, scales, ph, pw, sr, k_min, k_max, s0=224.0, lvl0=4.0, eps=1e-6):
    """TV ops/poolers.py:147-227 over `feats` (list of NCHW arrays), rois [K,5]."""
    feats = [_f32(f) for f in feats]
    rois = _f32(rois).reshape(-1, 5)
    n, H, W, sc = _level_args(feats, scales)
    B, Cc = feats[0].shape[:2]
    ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in feats])
    out = np.empty((len(rois), Cc, ph, pw), np.float32)
    lib().o_msroi_align_fwd(ptrs, C.c_int(n), H, W, sc, C.c_int(B), C.c_int(Cc), _p(rois),
                            C.c_int(len(rois)), C.c_int(ph), C.c_int(pw), C.c_int(sr),
                            C.c_int(k_min), C.c_int(k_max), C.c_float(s0), C.c_float(lvl0),
                            C.c_float(eps), _p(out))
    return out


def msroi_align_bwd(grad_out, shapes, rois, scales, sr, k_min, k_max, s0=224.0, lvl0=4.0, eps=1e-6):
    """Backward of msroi_align_fwd: returns one gradient array per level."""
    g = _f32(grad_out)
    rois = _f32(rois).reshape(-1, 5)
    grads = [np.zeros(s, np.float32) for s in shapes]
    n, H, W, sc = _level_args(grads, scales)
    B, Cc = shapes[0][:2]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in grads])
    lib().o_msroi_align_bwd(_p(g), C.c_int(n), H, W, sc, C.c_int(B), C.c_int(Cc), _p(rois),
                            C.c_int(len(rois)), C.c_int(g.shape[2]), C.c_int(g.shape[3]),
                            C.c_int(sr), C.c_int(k_min), C.c_int(k_max), C.c_float(s0),
                            C.c_float(lvl0), C.c_float(eps), ptrs)
    return grads


# ---------------------------------------------------------------------------------- RPN
def grid_anchors(cell, H, W, stride_h, stride_w):
    """TV models/detection/anchor_utils.py:84-113 for one level."""
    cell = _f32(cell).reshape(-1, 4)
    out = np.empty((H * W * len(cell), 4), np.float32)
    lib().o_grid_anchors(_p(cell), C.c_int(len(cell)), C.c_int(H), C.c_int(W), C.c_int(stride_h),
                         C.c_int(stride_w), _p(out))
    return out


def box_decode(rel, boxes, weights=(1.0, 1.0, 1.0, 1.0), clip=float(np.log(1000.0 / 16))):
    """TV models/detection/_utils.py:186-224: rel [n, n_cls*4], boxes [n,4] -> [n, n_cls*4]."""
    boxes = _f32(boxes).reshape(-1, 4)
    rel = _f32(rel).reshape(len(boxes), -1)
    n_cls = rel.shape[1] // 4
    out = np.empty_like(rel)
    lib().o_box_decode(_p(rel), _p(boxes), C.c_int(len(boxes)), C.c_int(n_cls),
                       *[C.c_float(w) for w in weights], C.c_float(clip), _p(out))
    return out


def rpn_filter_image(proposals, objectness, n_per_level, pre_top_n, post_top_n, nms_thresh,
                     min_size, score_thresh, img_h, img_w):
    """TV models/detection/rpn.py:242-297 for one image -> (boxes [k,4], scores [k])."""
    proposals, objectness = _f32(proposals).reshape(-1, 4), _f32(objectness).reshape(-1)
    npl = (C.c_int * len(n_per_level))(*n_per_level)
    ob = np.zeros((post_top_n, 4), np.float32)
    os_ = np.zeros(post_top_n, np.float32)
    n = lib().o_rpn_filter_image(_p(proposals), _p(objectness), npl, C.c_int(len(n_per_level)),
                                 C.c_int(pre_top_n), C.c_int(post_top_n), C.c_double(nms_thresh),
                                 C.c_float(min_size), C.c_float(score_thresh), C.c_float(img_h),
                                 C.c_float(img_w), _p(ob), _p(os_))
    return ob[:n].copy(), os_[:n].copy()


def detect_candidates(logits, reg, proposals, img_h, img_w, weights=(10.0, 10.0, 5.0, 5.0),
                      clip=float(np.log(1000.0 / 16)), score_thresh=0.05, min_size=1e-2):
    """TV models/detection/roi_heads.py:692-724 for the rows of one image."""
    logits, reg, proposals = _f32(logits), _f32(reg), _f32(proposals).reshape(-1, 4)
    n, n_cls = logits.shape
    cb = np.empty((n, n_cls - 1, 4), np.float32)
    cs = np.empty((n, n_cls - 1), np.float32)
    cl = np.empty((n, n_cls - 1), np.int64)
    cv = np.empty((n, n_cls - 1), np.uint8)
    lib().o_detect_candidates(_p(logits), _p(reg), _p(proposals), C.c_int(n), C.c_int(n_cls),
                              *[C.c_float(w) for w in weights], C.c_float(clip), C.c_float(img_h),
                              C.c_float(img_w), C.c_float(score_thresh), C.c_float(min_size),
                              _p(cb), _p(cs), _p(cl), _p(cv))
    return cb, cs, cl, cv


def grl_scale(g, alpha=0.1):
    """DGcommon.py:40-42."""
    g = _f32(g)
    out = np.empty_like(g)
    lib().o_grl_scale(_p(g), _p(out), C.c_int64(g.size), C.c_float(alpha))
    return out
