"""CPU oracle for the iDispNet hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference's instance
disparity path so the CUDA product path can be checked against it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``disprcnn_b200/``
imports this module; the product path fails loudly without its CUDA library.

Parity pinning: the reference ships no tests and no golden vectors
(SURVEY.md section 4), so this restatement is pinned by executing the
reference's own modules, imported unmodified from ``/root/reference`` in the
authoring container, on seeded inputs.  ``tests/golden/make_golden.py`` is the
committed generator; ``tests/test_oracle.py`` checks this file against those
fixtures.  The ROIAlign restatement additionally has a C twin
(``roi_align_oracle.c``) and is pinned against the reference's CPU kernel
compiled from its own source (``oracle/build_ref.py`` -> ``oracle/_ref``).

Reference locations restated here (paths relative to the reference root):
  cost volume ............ disprcnn/modeling/psmnet/stackhourglass.py:115-128
  convbn_3d .............. disprcnn/modeling/psmnet/submodule.py:19-22
  dres0/dres1 ............ stackhourglass.py:63-70, applied :130-131
  hourglass .............. stackhourglass.py:7-51, wired :133-140
  classif1..3 ............ stackhourglass.py:78-88, applied :142-144
  trilinear + softmax .... stackhourglass.py:169-172
  disparityregression .... submodule.py:51-57
  feature_extraction ..... submodule.py:60-139 (convbn :13-16, BasicBlock :25-48)
  PSMNet.forward ......... stackhourglass.py:106-174 (image crops -> extractor -> the rows above)
  ROIAlign (CPU kernel) .. disprcnn/csrc/cpu/ROIAlign_cpu.cpp:18-219
  crop + normalise ....... disprcnn/modeling/detector/disprcnn3d.py:44-50
  ROI box alignment ...... disprcnn3d.py:126-146, utils/stereo_utils.py:219-229
  ROI disparity hand-off . disprcnn3d.py:161-190 (roi_disp_postprocess), structures/disparity.py:39-78 (resize / crop),
                           modeling/pointnet_module/point_rcnn/lib/net/point_rcnn.py:113-136 (depth maps)

The arithmetic that the reference delegates to PyTorch (Conv3d,
ConvTranspose3d, BatchNorm3d in eval mode, F.interpolate, F.softmax) is
delegated to the same ``torch.nn.functional`` CPU ops here; the reference pins
``pytorch=1.2.0`` (environment.yaml:9), this image has 2.11.0.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm3d default, submodule.py:22


# ----------------------------------------------------------------------------
# cost volume (stackhourglass.py:115-128)
# ----------------------------------------------------------------------------
def cost_volume(left_fea, right_fea, mindisp, maxdisp):
    """[B,C,H,W] x2 -> [B,2C,D,H,W], D=(maxdisp-mindisp)//4.

    Follows the reference's slice-copy loop literally (including Python floor
    division of a negative ``mindisp``) rather than the closed form, so the
    closed form used by the kernel is checked against the loop.
    """
    bsz, C, Hp, Wp = left_fea.shape
    D = (maxdisp - mindisp) // 4
    cost = torch.zeros(bsz, C * 2, D, Hp, Wp, dtype=left_fea.dtype)
    for i in range(mindisp // 4, maxdisp // 4):
        k = i - mindisp // 4
        if i < 0:
            cost[:, :C, k, :, :i] = left_fea[:, :, :, :i]
            cost[:, C:, k, :, :i] = right_fea[:, :, :, -i:]
        elif i > 0:
            cost[:, :C, k, :, i:] = left_fea[:, :, :, i:]
            cost[:, C:, k, :, i:] = right_fea[:, :, :, :-i]
        else:
            cost[:, :C, k, :, :] = left_fea
            cost[:, C:, k, :, :] = right_fea
    return cost.contiguous()


# ----------------------------------------------------------------------------
# 3-D stack (stackhourglass.py:130-144)
# ----------------------------------------------------------------------------
def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'],
                        sd[prefix + '.weight'], sd[prefix + '.bias'], False, 0.0, BN_EPS)


def _convbn(x, sd, prefix, stride=1):
    """convbn_3d (submodule.py:19-22): Conv3d(k3,pad1,bias=False) + BatchNorm3d (eval)."""
    y = F.conv3d(x, sd[prefix + '.0.weight'], None, stride, 1)
    return _bn(y, sd, prefix + '.1')


def _deconvbn(x, sd, prefix):
    """ConvTranspose3d(k3,s2,p1,op1,bias=False) + BatchNorm3d (stackhourglass.py:22-30)."""
    y = F.conv_transpose3d(x, sd[prefix + '.0.weight'], None, 2, 1, 1)
    return _bn(y, sd, prefix + '.1')


def _hourglass(x, presqu, postsqu, sd, p):
    """hourglass.forward (stackhourglass.py:32-51)."""
    out = F.relu(_convbn(x, sd, p + '.conv1.0', 2))
    pre = _convbn(out, sd, p + '.conv2')
    pre = F.relu(pre + postsqu) if postsqu is not None else F.relu(pre)
    out = F.relu(_convbn(pre, sd, p + '.conv3.0', 2))
    out = F.relu(_convbn(out, sd, p + '.conv4.0'))
    if presqu is not None:
        post = F.relu(_deconvbn(out, sd, p + '.conv5') + presqu)
    else:
        post = F.relu(_deconvbn(out, sd, p + '.conv5') + pre)
    out = _deconvbn(post, sd, p + '.conv6')
    return out, pre, post


def stack3d(cost, sd, return_intermediates=False):
    """[B,2C,D,H,W] -> cost3 logits [B,1,D,H,W]  (stackhourglass.py:130-144)."""
    cost0 = F.relu(_convbn(cost, sd, 'dres0.0'))
    cost0 = F.relu(_convbn(cost0, sd, 'dres0.2'))
    t = F.relu(_convbn(cost0, sd, 'dres1.0'))
    cost0 = _convbn(t, sd, 'dres1.2') + cost0

    out1, pre1, post1 = _hourglass(cost0, None, None, sd, 'dres2')
    out1 = out1 + cost0
    out2, pre2, post2 = _hourglass(out1, pre1, post1, sd, 'dres3')
    out2 = out2 + cost0
    out3, pre3, post3 = _hourglass(out2, pre1, post2, sd, 'dres4')  # presqu is pre1 (:139)
    out3 = out3 + cost0

    def classif(x, p):
        y = F.relu(_convbn(x, sd, p + '.0'))
        return F.conv3d(y, sd[p + '.2.weight'], None, 1, 1)

    cost1 = classif(out1, 'classif1')
    cost2 = classif(out2, 'classif2') + cost1
    cost3 = classif(out3, 'classif3') + cost2
    if return_intermediates:
        return cost3, dict(cost0=cost0, out1=out1, out2=out2, out3=out3, pre1=pre1, post1=post1,
                           pre2=pre2, post2=post2, cost1=cost1, cost2=cost2)
    return cost3


# ----------------------------------------------------------------------------
# trilinear upsample + softmax + regression (stackhourglass.py:169-174, submodule.py:51-57)
# ----------------------------------------------------------------------------
def disparityregression(x, maxdisp, mindisp=0):
    assert x.shape[1] == int(maxdisp - mindisp)
    disp = torch.arange(mindisp, maxdisp, dtype=x.dtype).reshape(1, -1, 1, 1)
    return torch.sum(x * disp, 1)


def upsample_softargmin(cost3, mindisp, maxdisp, H, W):
    """[B,1,D,Hf,Wf] -> [B,H,W]."""
    c = F.interpolate(cost3, [maxdisp - mindisp, H, W], mode='trilinear', align_corners=True)
    c = torch.squeeze(c, 1)
    p = F.softmax(c, dim=1)
    return disparityregression(p, maxdisp, mindisp)


def idispnet_from_features(left_fea, right_fea, sd, mindisp, maxdisp, H=None, W=None):
    """PSMNet.forward from the feature tensors on (stackhourglass.py:115-174, eval branch)."""
    _, _, Hf, Wf = left_fea.shape
    H = 4 * Hf if H is None else H
    W = 4 * Wf if W is None else W
    cost = cost_volume(left_fea, right_fea, mindisp, maxdisp)
    cost3 = stack3d(cost, sd)
    return upsample_softargmin(cost3, mindisp, maxdisp, H, W)


# ----------------------------------------------------------------------------
# 2-D feature extractor (submodule.py:60-139) and the whole PSMNet.forward (stackhourglass.py:106-174)
# ----------------------------------------------------------------------------
def _bn2d(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], sd[prefix + '.weight'], sd[prefix + '.bias'],
                        False, 0.0, BN_EPS)


def _convbn2d(x, sd, prefix, stride, pad, dilation):
    """submodule.py:13-16: Conv2d(k, stride, padding = dilation if dilation > 1 else pad, dilation, bias=False) + BatchNorm2d."""
    w = sd[prefix + '.0.weight']
    return _bn2d(F.conv2d(x, w, None, stride, dilation if dilation > 1 else pad, dilation), sd, prefix + '.1')


def _basic_block(x, sd, p, stride, pad, dilation):
    """submodule.py:25-48: conv1 (convbn + ReLU), conv2 (convbn), optional 1x1 downsample of the input, add (NO ReLU after)."""
    out = F.relu(_convbn2d(x, sd, p + '.conv1.0', stride, pad, dilation))
    out = _convbn2d(out, sd, p + '.conv2', 1, pad, dilation)
    if (p + '.downsample.0.weight') in sd:
        x = _bn2d(F.conv2d(x, sd[p + '.downsample.0.weight'], None, stride), sd, p + '.downsample.1')
    return out + x


def feature_extraction(x, sd, prefix='feature_extraction.', return_intermediates=False):
    """submodule.py:112-139.  x [B,3,H,W] -> [B,32,H/4,W/4]."""
    pre = prefix
    o = F.relu(_convbn2d(x, sd, pre + 'firstconv.0', 2, 1, 1))
    o = F.relu(_convbn2d(o, sd, pre + 'firstconv.2', 1, 1, 1))
    o = F.relu(_convbn2d(o, sd, pre + 'firstconv.4', 1, 1, 1))
    first = o
    for b in range(3):
        o = _basic_block(o, sd, f'{pre}layer1.{b}', 1, 1, 1)
    for b in range(16):
        o = _basic_block(o, sd, f'{pre}layer2.{b}', 2 if b == 0 else 1, 1, 1)
    raw = o
    for b in range(3):
        o = _basic_block(o, sd, f'{pre}layer3.{b}', 1, 1, 1)
    for b in range(3):
        o = _basic_block(o, sd, f'{pre}layer4.{b}', 1, 1, 2)
    skip = o
    size = skip.shape[-2:]
    branches = {}
    for k, name in ((56, 'branch1'), (32, 'branch2'), (16, 'branch3'), (8, 'branch4')):
        bo = F.relu(_convbn2d(F.avg_pool2d(skip, (k, k), (k, k)), sd, f'{pre}{name}.1', 1, 0, 1))
        branches[name] = F.interpolate(bo, size, mode='bilinear', align_corners=True)
    cat = torch.cat((raw, skip, branches['branch4'], branches['branch3'], branches['branch2'], branches['branch1']), 1)
    o = F.relu(_convbn2d(cat, sd, pre + 'lastconv.0', 1, 1, 1))
    out = F.conv2d(o, sd[pre + 'lastconv.2.weight'])
    if return_intermediates:
        return out, dict(first=first, raw=raw, skip=skip, cat=cat)
    return out


def psmnet_forward(left, right, sd, mindisp, maxdisp):
    """stackhourglass.py:106-174 in eval mode: image crops [B,3,H,W] x2 -> disparity [B,H,W] (pred3)."""
    H, W = left.shape[-2:]
    return idispnet_from_features(feature_extraction(left, sd), feature_extraction(right, sd), sd, mindisp, maxdisp, H, W)


# ----------------------------------------------------------------------------
# ROIAlign forward (csrc/cpu/ROIAlign_cpu.cpp:18-219), float32 arithmetic
# ----------------------------------------------------------------------------
def roi_align_forward(inp, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio):
    """numpy float32 restatement; inp [N,C,H,W] f32, rois [R,5] f32 -> [R,C,ph,pw] f32.

    Every intermediate is kept in float32 in the same operation order as the
    reference's ``T = float`` instantiation so the sampling indices and
    weights are bit-identical.
    """
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    f = np.float32
    N, C, H, W = inp.shape
    R = rois.shape[0]
    out = np.zeros((R, C, pooled_h, pooled_w), dtype=np.float32)
    scale = f(spatial_scale)
    for n in range(R):
        b = int(rois[n, 0])
        rsw = f(rois[n, 1] * scale)
        rsh = f(rois[n, 2] * scale)
        rew = f(rois[n, 3] * scale)
        reh = f(rois[n, 4] * scale)
        roi_w = max(f(rew - rsw), f(1.0))
        roi_h = max(f(reh - rsh), f(1.0))
        bin_h = f(roi_h / f(pooled_h))
        bin_w = f(roi_w / f(pooled_w))
        gh = sampling_ratio if sampling_ratio > 0 else int(math.ceil(f(roi_h / f(pooled_h))))
        gw = sampling_ratio if sampling_ratio > 0 else int(math.ceil(f(roi_w / f(pooled_w))))
        count = f(gh * gw)
        # sample coordinates (ROIAlign_cpu.cpp:36-43), vectorised over (ph, iy) / (pw, ix)
        ph = np.arange(pooled_h, dtype=np.float32)[:, None]
        iy = (np.arange(gh, dtype=np.float32)[None, :] + f(0.5)).astype(np.float32)
        yy = (rsh + (ph * bin_h).astype(np.float32)).astype(np.float32) + \
             ((iy * bin_h).astype(np.float32) / f(gh)).astype(np.float32)
        yy = yy.astype(np.float32)  # [ph, gh]
        pw = np.arange(pooled_w, dtype=np.float32)[:, None]
        ix = (np.arange(gw, dtype=np.float32)[None, :] + f(0.5)).astype(np.float32)
        xx = (rsw + (pw * bin_w).astype(np.float32)).astype(np.float32) + \
             ((ix * bin_w).astype(np.float32) / f(gw)).astype(np.float32)
        xx = xx.astype(np.float32)  # [pw, gw]

        def axis_terms(v, size):
            oob = (v < f(-1.0)) | (v > f(size))
            v = np.where(v <= 0, f(0), v).astype(np.float32)
            lo = v.astype(np.int32)
            clamp = lo >= size - 1
            lo = np.where(clamp, size - 1, lo)
            hi = np.where(clamp, size - 1, lo + 1)
            v = np.where(clamp, lo.astype(np.float32), v).astype(np.float32)
            l = (v - lo.astype(np.float32)).astype(np.float32)
            h = (f(1.0) - l).astype(np.float32)
            return oob, lo, hi, l, h

        oy, ylo, yhi, ly, hy = axis_terms(yy, H)
        ox, xlo, xhi, lx, hx = axis_terms(xx, W)
        img = inp[b]  # [C,H,W]
        # [ph,gh,pw,gw] broadcast
        Y = lambda a: a[:, :, None, None]
        X = lambda a: a[None, None, :, :]
        w1 = (Y(hy) * X(hx)).astype(np.float32)
        w2 = (Y(hy) * X(lx)).astype(np.float32)
        w3 = (Y(ly) * X(hx)).astype(np.float32)
        w4 = (Y(ly) * X(lx)).astype(np.float32)
        oob = Y(oy) | X(ox)
        v1 = img[:, Y(ylo), X(xlo)]
        v2 = img[:, Y(ylo), X(xhi)]
        v3 = img[:, Y(yhi), X(xlo)]
        v4 = img[:, Y(yhi), X(xhi)]
        # ROIAlign_cpu.cpp:196-199: ((w1*v1 + w2*v2) + w3*v3) + w4*v4 in float
        val = ((w1 * v1).astype(np.float32) + (w2 * v2).astype(np.float32)).astype(np.float32)
        val = (val + (w3 * v3).astype(np.float32)).astype(np.float32)
        val = (val + (w4 * v4).astype(np.float32)).astype(np.float32)
        val = np.where(oob[None], f(0), val).astype(np.float32)  # [C,ph,gh,pw,gw]
        # accumulate in (iy, ix) order like the reference's scalar loop
        acc = np.zeros((C, pooled_h, pooled_w), dtype=np.float32)
        for a in range(gh):
            for c in range(gw):
                acc = (acc + val[:, :, a, :, c]).astype(np.float32)
        out[n] = (acc / count).astype(np.float32)
    return out


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def crop_and_transform_roi_img(im, rois, resolution=224):
    """disprcnn3d.py:44-50: ROIAlign((res,res),1.0,0) then (x-mean)/std per channel."""
    out = roi_align_forward(im, rois, 1.0, resolution, resolution, 0)
    mean = np.asarray(IMAGENET_MEAN, dtype=np.float32)[None, :, None, None]
    std = np.asarray(IMAGENET_STD, dtype=np.float32)[None, :, None, None]
    return ((out - mean).astype(np.float32) / std).astype(np.float32)


def align_stereo_boxes(left_boxes, right_boxes, width, height):
    """disprcnn3d.py:126-146 (+ stereo_utils.py:219-229): per-ROI left/right crop boxes.

    left_boxes/right_boxes: lists of lists ``[[(x1,y1,x2,y2), ...] per image]``.
    width/height: the (unpadded) BoxList size ``left_result[i].width/.height`` -- one int for all images or a
    per-image sequence (KITTI frames differ in size; the ImageList batch is padded, the BoxLists are not).
    Returns (rois_left [R,5], rois_right [R,5]) as python lists.
    """
    rl, rr = [], []
    widths = width if hasattr(width, '__len__') else [width] * len(left_boxes)
    heights = height if hasattr(height, '__len__') else [height] * len(left_boxes)
    for i, (lbs, rbs) in enumerate(zip(left_boxes, right_boxes)):
        width, height = widths[i], heights[i]
        for lb, rb in zip(lbs, rbs):
            x1, y1, x2, y2 = math.floor(lb[0]), math.floor(lb[1]), math.ceil(lb[2]), math.ceil(lb[3])
            x1p, x2p = math.floor(rb[0]), math.ceil(rb[2])
            x1 = max(0, x1)
            x1p = max(0, x1p)
            y1 = max(0, y1)
            y2 = min(y2, height - 1)
            x2 = min(x2, width - 1)
            x2p = min(x2p, width - 1)
            mw = max(x2 - x1, x2p - x1p)
            mw = min(mw, min(width - x1, width - x1p))
            rl.append([i, x1, y1, x1 + mw, y2])
            rr.append([i, x1p, y1, x1p + mw, y2])
    return rl, rr


# ----------------------------------------------------------------------------
# hand-off of the per-ROI disparity maps (disprcnn3d.py:161-190, point_rcnn.py:113-136, structures/disparity.py:39-78)
# ----------------------------------------------------------------------------
def _resize_crop_shift(disp_roi, lb, rb):
    """DisparityMap(out[i]).resize((max(x2-x1, x2p-x1p), y2-y1)).crop((0, 0, x2-x1, y2-y1)).data
    (disprcnn3d.py:171-175 / point_rcnn.py:124-129; resize / crop: structures/disparity.py:39-78); the shift by x1 - x1p is
    written differently at the two call sites (one float add at disprcnn3d.py:178, two at point_rcnn.py:130) and stays there."""
    x1, y1, x2, y2 = math.floor(lb[0]), math.floor(lb[1]), math.ceil(lb[2]), math.ceil(lb[3])
    x1p, x2p = math.floor(rb[0]), math.ceil(rb[2])
    dst_w, dst_h = max(x2 - x1, x2p - x1p), y2 - y1
    src_w = disp_roi.shape[1]
    t = F.interpolate(disp_roi[None, None], (dst_h, dst_w), mode='bilinear', align_corners=True)[0, 0]
    t = t / src_w * dst_w                       # disparity.py:60
    t = t[:y2 - y1, :x2 - x1]                   # crop (the resized map is at least that large)
    return t, (x1, y1, x2, y2), x1p


def roi_disp_postprocess(roi_disp, left_boxes, right_boxes, masks, height, width):
    """disprcnn3d.py:161-190.  roi_disp [R,S,S]; left_boxes / right_boxes: lists (per image) of box lists; masks [R,H,W] (0/1) or
    None.  Returns [N,H,W]."""
    outs, r = [], 0
    for lbs, rbs in zip(left_boxes, right_boxes):
        per_img = []
        for lb, rb in zip(lbs, rbs):
            d, (x1, y1, x2, y2), x1p = _resize_crop_shift(roi_disp[r], lb, rb)
            m = torch.zeros((height, width))
            m[y1:y1 + d.shape[0], x1:x1 + d.shape[1]] = d + (x1 - x1p)          # disprcnn3d.py:178
            m = m.clamp(min=0)
            if masks is not None:
                m = m * masks[r].float()
            per_img.append(m)
            r += 1
        outs.append(torch.stack(per_img).max(dim=0)[0] if per_img else torch.zeros((height, width)))
    return torch.stack(outs) if outs else torch.zeros((0, height, width))


def roi_depth_maps(roi_disp, left_boxes, right_boxes, fu_baseline, height, width):
    """point_rcnn.py:124-134: per-ROI image-sized depth maps, fu*baseline / (disp + 1e-6) inside the box."""
    outs = []
    for r, (lb, rb) in enumerate(zip(left_boxes, right_boxes)):
        d, (x1, y1, x2, y2), x1p = _resize_crop_shift(roi_disp[r], lb, rb)
        d = d + x1 - x1p                                                         # point_rcnn.py:130: (disp + x1) - x1p
        m = torch.zeros((height, width))
        m[y1:y2, x1:x2] = float(fu_baseline[r]) / (d + 1e-6)
        outs.append(m)
    return torch.stack(outs) if outs else torch.zeros((0, height, width))
