"""CPU restatement (torch fp32, functional) of the two networks on the hot path and of the
small tensor ops around them.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every function cites the reference lines it restates.  ``params`` are dicts
``{state_dict_key: torch.float32 tensor}`` with the reference's own key names.
All ``grid_sample``-type sampling uses the pinned torch-1.1 semantics
(``align_corners=True``) written out explicitly as pixel-space gathers.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .synth import LFN_BACKWARD, LFN_KLAST, LFN_LEVELS


def to_torch(d):
    return {k: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v)
            for k, v in d.items()}


# ----------------------------------------------------------------------------------------
# correlation  (lite_flow_net/correlation.py:11-106 kernels, :281-341 wrapper)
# ----------------------------------------------------------------------------------------
def correlation(first, second, stride):
    """49-channel cost volume.  ``out[b, (dy+3)*7+(dx+3), y, x] =
    mean_c first[b,c,y*s,x*s] * second[b,c,y*s+dy*s,x*s+dx*s]`` with zeros outside the image;
    output size ceil(H/s) x ceil(W/s) (correlation.py:294).

    Follows the kernel's index math: padded coords ``x1=(bx+3)*s`` (:50-51), displacement
    ``s2o=(top_channel%7-3)*s``, ``s2p=(top_channel/7-3)*s`` (:74-75), division by the channel
    count (:101-103)."""
    B, C, H, W = first.shape
    s = int(stride)
    Ho, Wo = int(math.ceil(H / s)), int(math.ceil(W / s))
    pad = 3 * s
    # rbot1 = zero-padded second (correlation.py:283-284, rearrange kernel :11-36)
    sp = F.pad(second, (pad, pad, pad, pad))
    # the padded buffers are (H+6s) x (W+6s); output pixel (by,bx) reads first at (by*s, bx*s).
    # For odd sizes with s=2 the last output row/col reads first at index (Ho-1)*s <= H-1: ok.
    f = first[:, :, 0:(Ho - 1) * s + 1:s, 0:(Wo - 1) * s + 1:s]
    out = first.new_zeros(B, 49, Ho, Wo)
    for ch in range(49):
        dx = (ch % 7 - 3) * s
        dy = (ch // 7 - 3) * s
        y0, x0 = pad + dy, pad + dx
        g = sp[:, :, y0:y0 + (Ho - 1) * s + 1:s, x0:x0 + (Wo - 1) * s + 1:s]
        out[:, ch] = (f * g).sum(1) / float(C)
    return out


# ----------------------------------------------------------------------------------------
# Backward warp  (lite_flow_net.py:10-28)
# ----------------------------------------------------------------------------------------
def bilinear_sample_zeros(inp, px, py):
    """Bilinear gather of ``inp`` [B,C,H,W] at pixel coords (px,py) [B,h,w], zeros outside
    (= ``grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True)``)."""
    B, C, H, W = inp.shape
    x0 = torch.floor(px)
    y0 = torch.floor(py)
    wx1 = px - x0
    wy1 = py - y0
    wx0 = 1.0 - wx1
    wy0 = 1.0 - wy1
    out = inp.new_zeros(B, C, px.shape[1], px.shape[2])
    flat = inp.reshape(B, C, H * W)
    for (xx, wx) in ((x0, wx0), (x0 + 1, wx1)):
        for (yy, wy) in ((y0, wy0), (y0 + 1, wy1)):
            valid = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
            xi = xx.clamp(0, W - 1).long()
            yi = yy.clamp(0, H - 1).long()
            idx = (yi * W + xi).reshape(B, 1, -1).expand(B, C, -1)
            v = torch.gather(flat, 2, idx).reshape(B, C, px.shape[1], px.shape[2])
            out = out + v * (wx * wy * valid.to(inp.dtype)).unsqueeze(1)
    return out


def backward_warp(inp, flow):
    """``Backward`` (lite_flow_net.py:10-28): the grid is ``linspace(-1,1,W)[x] + flow_x /
    ((W-1)/2)`` which under align_corners=True un-normalises to pixel ``x + flow_x``."""
    B, _, H, W = flow.shape
    xs = torch.arange(W, dtype=flow.dtype).view(1, 1, W).expand(B, H, W)
    ys = torch.arange(H, dtype=flow.dtype).view(1, H, 1).expand(B, H, W)
    return bilinear_sample_zeros(inp, xs + flow[:, 0], ys + flow[:, 1])


# ----------------------------------------------------------------------------------------
# LiteFlowNet  (lite_flow_net.py:31-325)
# ----------------------------------------------------------------------------------------
def _conv(p, name, x, stride=1, pad=None, leaky=True):
    w = p[name + ".weight"]
    b = p.get(name + ".bias")
    if pad is None:
        pad = (w.shape[2] // 2, w.shape[3] // 2)
    y = F.conv2d(x, w, b, stride=stride, padding=pad)
    return F.leaky_relu(y, 0.1) if leaky else y


def lfn_features(p, x):
    """``Features.forward`` (lite_flow_net.py:77-86)."""
    f = "moduleFeatures."
    one = _conv(p, f + "moduleOne.0", x)
    two = _conv(p, f + "moduleTwo.0", one, stride=2)
    two = _conv(p, f + "moduleTwo.2", two)
    two = _conv(p, f + "moduleTwo.4", two)
    thr = _conv(p, f + "moduleThr.0", two, stride=2)
    thr = _conv(p, f + "moduleThr.2", thr)
    fou = _conv(p, f + "moduleFou.0", thr, stride=2)
    fou = _conv(p, f + "moduleFou.2", fou)
    fiv = _conv(p, f + "moduleFiv.0", fou, stride=2)
    six = _conv(p, f + "moduleSix.0", fiv, stride=2)
    return [one, two, thr, fou, fiv, six]


def lfn_matching(p, k, lv, feat1, feat2, flow):
    """``Matching.forward`` (lite_flow_net.py:132-152)."""
    m = "moduleMatching.%d." % k
    if lv == 2:
        feat1 = _conv(p, m + "moduleFeat.0", feat1)
        feat2 = _conv(p, m + "moduleFeat.0", feat2)
    if flow is not None:
        flow = F.conv_transpose2d(flow, p[m + "moduleUpflow.weight"], None, stride=2, padding=1,
                                  groups=2)
        feat2 = backward_warp(feat2, flow * LFN_BACKWARD[lv])
    if lv >= 4:
        corr = F.leaky_relu(correlation(feat1, feat2, 1), 0.1)
    else:
        corr = F.leaky_relu(correlation(feat1, feat2, 2), 0.1)
        corr = F.conv_transpose2d(corr, p[m + "moduleUpcorr.weight"], None, stride=2, padding=1,
                                  groups=49)
    x = _conv(p, m + "moduleMain.0", corr)
    x = _conv(p, m + "moduleMain.2", x)
    x = _conv(p, m + "moduleMain.4", x)
    x = _conv(p, m + "moduleMain.6", x, leaky=False)
    return x if flow is None else flow + x


def lfn_subpixel(p, k, lv, feat1, feat2, flow):
    """``Subpixel.forward`` (lite_flow_net.py:182-190)."""
    m = "moduleSubpixel.%d." % k
    if lv == 2:
        feat1 = _conv(p, m + "moduleFeat.0", feat1)
        feat2 = _conv(p, m + "moduleFeat.0", feat2)
    feat2 = backward_warp(feat2, flow * LFN_BACKWARD[lv])
    x = torch.cat([feat1, feat2, flow], 1)
    x = _conv(p, m + "moduleMain.0", x)
    x = _conv(p, m + "moduleMain.2", x)
    x = _conv(p, m + "moduleMain.4", x)
    x = _conv(p, m + "moduleMain.6", x, leaky=False)
    return flow + x


def lfn_regularization(p, k, lv, img1, img2, feat1, flow):
    """``Regularization.forward`` (lite_flow_net.py:243-264)."""
    m = "moduleRegularization.%d." % k
    diff = img1 - backward_warp(img2, flow * LFN_BACKWARD[lv])
    diff = (diff.pow(2.0).sum(1, True) + 1e-6).sqrt()
    B = flow.shape[0]
    fmean = flow.reshape(B, 2, -1).mean(2, True).reshape(B, 2, 1, 1)
    if lv < 5:
        feat1 = _conv(p, m + "moduleFeat.0", feat1)
    x = torch.cat([diff, flow - fmean, feat1], 1)
    for i in (0, 2, 4, 6, 8, 10):
        x = _conv(p, m + "moduleMain.%d" % i, x)
    kd = LFN_KLAST[lv]
    if lv >= 5:
        d = _conv(p, m + "moduleDist.0", x, leaky=False)
    else:
        d = _conv(p, m + "moduleDist.0", x, pad=(kd // 2, 0), leaky=False)
        d = _conv(p, m + "moduleDist.1", d, pad=(0, kd // 2), leaky=False)
    d = d.pow(2.0).neg()
    d = (d - d.max(1, True)[0]).exp()
    div = d.sum(1, True).reciprocal()
    ux = F.unfold(flow[:, 0:1], kd, stride=1, padding=(kd - 1) // 2).view_as(d)
    uy = F.unfold(flow[:, 1:2], kd, stride=1, padding=(kd - 1) // 2).view_as(d)
    sx = _conv(p, m + "moduleScaleX", d * ux, leaky=False) * div
    sy = _conv(p, m + "moduleScaleY", d * uy, leaky=False) * div
    return torch.cat([sx, sy], 1)


def liteflownet_forward(p, img1, img2, return_intermediates=False):
    """``LiteFlowNet.forward`` (lite_flow_net.py:285-325).  Returns {1..5: flow} already
    multiplied by ``20*0.5**i`` (:322-324)."""
    f1 = lfn_features(p, img1)
    f2 = lfn_features(p, img2)
    im1, im2 = [img1], [img2]
    for lv in range(1, 6):
        size = (f1[lv].shape[2], f1[lv].shape[3])
        im1.append(F.interpolate(im1[-1], size=size, mode="bilinear", align_corners=False))
        im2.append(F.interpolate(im2[-1], size=size, mode="bilinear", align_corners=False))
    flow = None
    flows = {}
    inter = {}
    for cnt, idx in enumerate([-1, -2, -3, -4, -5]):
        k = LFN_LEVELS.index(6 - cnt)
        lv = 6 - cnt
        flow = lfn_matching(p, k, lv, f1[idx], f2[idx], flow)
        inter[("matching", lv)] = flow
        flow = lfn_subpixel(p, k, lv, f1[idx], f2[idx], flow)
        inter[("subpixel", lv)] = flow
        flow = lfn_regularization(p, k, lv, im1[idx], im2[idx], f1[idx], flow)
        inter[("regularization", lv)] = flow
        flows[5 - cnt] = flow
    flows = {i: v * (20.0 * (0.5 ** i)) for i, v in flows.items()}
    if return_intermediates:
        return flows, inter, f1, f2
    return flows


def get_target_size(h, w):
    """``DeepFlow.get_target_size`` (deep_flow.py:89-105), restated operation for operation.
    The arguments are shadowed by the 1x2 candidate arrays before ``h / w`` is evaluated, so the
    "ratio" matrix is ``|h_i * (1/w_j) - h_j / w_j|`` whose diagonal is zero *up to one rounding*:
    the result is the floor multiple of 32 when ``h0 * (1/w0) == h0 / w0`` in float64 (376x1241,
    370x1226 -> 352x1216) and the ceil multiples otherwise (192x640 -> 224x672).  Golden-pinned."""
    hh = 32 * np.array([[math.floor(h / 32), math.floor(h / 32) + 1]])
    ww = 32 * np.array([[math.floor(w / 32), math.floor(w / 32) + 1]])
    ratio = np.abs(np.matmul(np.transpose(hh), 1 / ww) - hh / ww)
    index = int(np.argmin(ratio))
    return int(hh[0, index // 2]), int(ww[0, index % 2])


def resize_dense_flow(flow, H, W):
    """``DeepFlow.resize_dense_flow`` (deep_flow.py:107-129)."""
    rh = float(H / flow.shape[2])
    rw = float(W / flow.shape[3])
    flow = F.interpolate(flow, (H, W), mode="bilinear", align_corners=True)
    return torch.stack([flow[:, 0] * rw, flow[:, 1] * rh], 1)


def fb_consistency(flow_fwd, flow_bwd):
    """``FlowToPix`` + ``forward_backward_consistency`` (layers.py:213-229, deep_flow.py:171-196):
    ``diff = || flow_fwd - bilinear(-flow_bwd at (x,y)+flow_fwd) ||_2`` -> [N,H,W,1]."""
    B, _, H, W = flow_fwd.shape
    xs = torch.arange(W, dtype=flow_fwd.dtype).view(1, 1, W).expand(B, H, W)
    ys = torch.arange(H, dtype=flow_fwd.dtype).view(1, H, 1).expand(B, H, W)
    # the reference normalises then grid_sample un-normalises; keep that arithmetic
    gx = ((xs + flow_fwd[:, 0]) / (W - 1) - 0.5) * 2
    gy = ((ys + flow_fwd[:, 1]) / (H - 1) - 0.5) * 2
    px = ((gx + 1) / 2) * (W - 1)
    py = ((gy + 1) / 2) * (H - 1)
    warp = bilinear_sample_zeros(-flow_bwd, px, py)
    d = flow_fwd - warp
    return d.norm(dim=1, keepdim=True).permute(0, 2, 3, 1)


def liteflow_inference_flow(p, img_ref, img_cur):
    """``LiteFlow.inference_flow(forward_backward=True)`` (lite_flow.py:89-148) on one pair of
    [1,3,H,W] float images in [0,1]: returns forward, backward [1,2,H,W] and flow_diff [1,H,W,1]."""
    a = torch.cat([img_ref, img_cur], 0)
    b = torch.cat([img_cur, img_ref], 0)
    _, _, h, w = a.shape
    th, tw = get_target_size(h, w)
    ra = F.interpolate(a, (th, tw), mode="bilinear", align_corners=True)
    rb = F.interpolate(b, (th, tw), mode="bilinear", align_corners=True)
    out = liteflownet_forward(p, ra, rb)
    flow = resize_dense_flow(out[1], h, w)
    fwd, bwd = flow[0:1], flow[1:2]
    return {"forward": fwd, "backward": bwd, "flow_diff": fb_consistency(fwd, bwd)}


# ----------------------------------------------------------------------------------------
# monodepth2  (resnet_encoder.py:87-98, depth_decoder.py:50-65, layers.py:16-25,106-136,347-350,
#              monodepth2.py:91-139)
# ----------------------------------------------------------------------------------------
def _bn(p, name, x):
    return F.batch_norm(x, p[name + ".running_mean"], p[name + ".running_var"],
                        p[name + ".weight"], p[name + ".bias"], False, 0.0, 1e-5)


def resnet18_encoder(p, x):
    """``ResnetEncoder.forward`` (resnet_encoder.py:87-98) with torchvision BasicBlocks."""
    x = (x - 0.45) / 0.225
    x = F.conv2d(x, p["encoder.conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(p, "encoder.bn1", x))
    feats = [x]
    x = F.max_pool2d(x, 3, 2, 1)
    for li in (1, 2, 3, 4):
        for b in (0, 1):
            pre = "encoder.layer%d.%d." % (li, b)
            stride = 2 if (li > 1 and b == 0) else 1
            idt = x
            y = F.conv2d(x, p[pre + "conv1.weight"], None, stride=stride, padding=1)
            y = F.relu(_bn(p, pre + "bn1", y))
            y = F.conv2d(y, p[pre + "conv2.weight"], None, stride=1, padding=1)
            y = _bn(p, pre + "bn2", y)
            if (pre + "downsample.0.weight") in p:
                idt = F.conv2d(x, p[pre + "downsample.0.weight"], None, stride=stride)
                idt = _bn(p, pre + "downsample.1", idt)
            x = F.relu(y + idt)
        feats.append(x)
    return feats


def _conv3x3_refl(p, name, x):
    """``Conv3x3`` (layers.py:121-136): ReflectionPad2d(1) + 3x3 conv."""
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), p[name + ".weight"], p[name + ".bias"])


def depth_decoder(p, feats, scales=(0, 1, 2, 3)):
    """``DepthDecoder.forward`` (depth_decoder.py:50-65)."""
    out = {}
    x = feats[-1]
    idx = 0
    for i in range(4, -1, -1):
        x = F.elu(_conv3x3_refl(p, "decoder.%d.conv.conv" % idx, x))
        idx += 1
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if i > 0:
            x = torch.cat([x, feats[i - 1]], 1)
        x = F.elu(_conv3x3_refl(p, "decoder.%d.conv.conv" % idx, x))
        idx += 1
        if i in scales:
            out[("disp", i)] = torch.sigmoid(_conv3x3_refl(p, "decoder.%d.conv" % (10 + i), x))
    return out


def monodepth2_inference_depth(enc, dec, img, min_depth=0.1, max_depth=100.0, baseline=5.4):
    """``Monodepth2DepthNet.inference_depth`` (monodepth2.py:91-139), kitti constants (:74-77)."""
    feats = resnet18_encoder(enc, img)
    disp = depth_decoder(dec, feats, scales=(0,))[("disp", 0)]
    disp = F.interpolate(disp, img.shape[2:], mode="bilinear", align_corners=False)
    min_disp, max_disp = 1 / max_depth, 1 / min_depth
    scaled = min_disp + (max_disp - min_disp) * disp           # layers.py:16-25
    return (1.0 / scaled) * baseline
