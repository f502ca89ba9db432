"""Range-view post-processing after the lidar decode, on the device.

Drop-ins, with the reference's names and argument meaning, for the host-side NumPy / cv2 / numba code the reference
runs per sample after `decode_first_stage`:

  depth_normalization, inverse_depth_normalization          ldm/data/utils.py:537-580
  postprocess_range_depth_int, postprocess_range_depth      ldm/data/utils.py:471-534
  LidarConverter.range2pcd / .undo_default_transforms       ldm/data/lidar_converter.py:122-176, 436-485
  points_in_bbox_corners                                    ldm/data/box_np_ops.py:453-471
  composite_range_samples / postprocess_lidar_samples       scripts/inference_test_bench.py:567-629 (+ ddpm.py:1503-1543)

Differences from the reference, all of them about WHERE the data lives: inputs and outputs are CUDA tensors (the reference
takes tensors, calls .cpu().numpy() and returns ndarrays); batched calls return padded point buffers plus counts instead
of Python lists, so nothing synchronises with the host unless the caller asks for a ragged list.  CUDA only — there is
no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib as L
from .ops import _cuda, _timed

MAP_NONE, MAP_DEPTH_NORM, MAP_DEPTH_UNNORM, MAP_INT_UNNORM = 0, 1, 2, 3


def _f32c(t):
    if not t.is_cuda:
        raise RuntimeError("mobi_b200.lidar runs on CUDA tensors only (no CPU fallback)")
    return t.detach().to(torch.float32).contiguous()


def _per_sample(v, batch, device):
    """min_d / max_d as the reference passes them (0-dim tensors, floats) or per-sample [B] tensors -> f32 [batch]."""
    if not torch.is_tensor(v):
        v = torch.tensor(float(v))
    if not v.is_cuda:                                   # the reference's range assertion, when it costs no sync
        assert bool((v >= -1).all() and (v <= 1).all()), "min_d and max_d must be in the range -1 to 1 and min_d < max_d"
    v = v.detach().to(device=device, dtype=torch.float32).reshape(-1)
    if v.numel() == 1 and batch != 1:
        v = v.expand(batch)
    assert v.numel() == batch, "min_d / max_d: one value, or one per sample"
    return v.contiguous()


def _range_map(x, mode, min_d=None, max_d=None, alpha=0.75, clamp=False, out=None):
    assert 0 < alpha <= 1, "alpha must be in the range 0 to 1"
    x = _f32c(x)
    per_sample = min_d is not None and torch.is_tensor(min_d) and min_d.numel() > 1
    batch = x.shape[0] if per_sample else 1
    if out is None:
        out = torch.empty_like(x)
    a = L.RangeMapArgs()
    a.in_, a.out = x.data_ptr(), out.data_ptr()
    if min_d is not None:
        mn, mx = _per_sample(min_d, batch, x.device), _per_sample(max_d, batch, x.device)
        a.min_d, a.max_d = mn.data_ptr(), mx.data_ptr()
    a.n = x.numel() // batch
    a.in_stride = a.out_stride = a.n
    a.batch, a.mode, a.clamp_input, a.alpha = batch, mode, 1 if clamp else 0, alpha
    if x.numel():
        with _timed("range_map", nbytes=x.numel() * 8):
            L.check(L.load().mobi_range_map(C.byref(a), L.stream()), "range_map")
    return out


def depth_normalization(depth, min_d, max_d, alpha=0.75):
    """utils.py:537-557.  min_d / max_d: scalars (the reference's per-sample call) or [B] tensors for depth [B, ...]."""
    return _range_map(depth, MAP_DEPTH_NORM, min_d, max_d, alpha)


def inverse_depth_normalization(normalized_depth, min_d, max_d, alpha=0.75):
    """utils.py:560-580."""
    return _range_map(normalized_depth, MAP_DEPTH_UNNORM, min_d, max_d, alpha)


def intensity_unnormalization(x):
    """torch.clamp(-0.5 * torch.log(1 - (x + 1) / 2) - 1, -1, 1)  (ddpm.py:1541)."""
    return _range_map(x, MAP_INT_UNNORM)


def _undo(crops, origs, crop_left, width_crop, *, zero_context=False, maps=(MAP_NONE, MAP_NONE), min_d=None, max_d=None,
          alpha=0.75, clamp=False):
    """crops: 1-2 f32 views [B, h, w] that share one batch stride (e.g. channels of the decoded [B, C, h, w] image);
    origs: matching [B, H, W] tensors."""
    B, H, W = origs[0].shape
    ch, cw = crops[0].shape[-2:]
    a = L.RangeUndoArgs()
    outs = []
    keep = []
    for c, (crop, orig) in enumerate(zip(crops, origs)):
        assert crop.is_cuda and crop.dtype == torch.float32 and crop.shape[0] == B and tuple(crop.shape[-2:]) == (ch, cw)
        assert crop.stride(-1) == 1 and crop.stride(-2) == cw and crop.stride(0) == crops[0].stride(0)
        orig = _f32c(orig)
        out = torch.empty((B, H, W), device=orig.device, dtype=torch.float32)
        a.crop[c], a.orig[c], a.out[c], a.map[c] = crop.data_ptr(), orig.data_ptr(), out.data_ptr(), maps[c]
        outs.append(out)
        keep.append(orig)
    cl = crop_left.detach().to(device=outs[0].device, dtype=torch.int64).reshape(-1).contiguous()
    wc = width_crop.detach().to(device=outs[0].device, dtype=torch.int64).reshape(-1).contiguous()
    assert cl.numel() == B and wc.numel() == B
    a.crop_left, a.width_crop = cl.data_ptr(), wc.data_ptr()
    if min_d is not None:
        mn, mx = _per_sample(min_d, B, outs[0].device), _per_sample(max_d, B, outs[0].device)
        a.min_d, a.max_d = mn.data_ptr(), mx.data_ptr()
    a.crop_batch_stride = crops[0].stride(0)
    a.batch, a.channels, a.crop_h, a.crop_w, a.H, a.W = B, len(crops), ch, cw, H, W
    a.zero_context, a.clamp_input, a.alpha = 1 if zero_context else 0, 1 if clamp else 0, alpha
    with _timed("range_undo", nbytes=len(crops) * B * (ch * cw + 2 * H * W) * 4):
        L.check(L.load().mobi_range_undo_transforms(C.byref(a), L.stream()), "range_undo_transforms")
    return outs


def _crop_view(t):
    """[B, 1, h, w] / [B, h, w] f32 CUDA tensor -> [B, h, w] view with dense rows (copy only when it has to be)."""
    if not t.is_cuda:
        raise RuntimeError("mobi_b200.lidar runs on CUDA tensors only (no CPU fallback)")
    if t.dim() == 4:
        assert t.shape[1] == 1
        t = t[:, 0]
    t = t.detach()
    if t.dtype != torch.float32 or t.stride(-1) != 1 or t.stride(-2) != t.shape[-1]:
        t = t.to(torch.float32).contiguous()
    return t


def postprocess_range_depth_int(*, range_depth, range_depth_orig, range_int, range_int_orig, crop_left, width_crop,
                                zero_context=False):
    """utils.py:471-505: [B, 1, h, w] crops -> un-cropped [B, H, W] depth and intensity (CUDA tensors)."""
    d, i = _crop_view(range_depth), _crop_view(range_int)
    if d.stride(0) != i.stride(0):
        d, i = d.contiguous(), i.contiguous()
    return tuple(_undo([d, i], [range_depth_orig, range_int_orig], crop_left, width_crop, zero_context=zero_context))


def postprocess_range_depth(*, range_depth, range_depth_orig, crop_left, width_crop, zero_context=False):
    """utils.py:507-534."""
    return _undo([_crop_view(range_depth)], [range_depth_orig], crop_left, width_crop, zero_context=zero_context)[0]


def points_in_bbox_corners(points, rbbox_corners):
    """box_np_ops.py:453-471: points [N, 3+], corners [M, 8, 3] -> bool [N, M]."""
    points, rbbox_corners = _f32c(points), _f32c(rbbox_corners)
    n, m = points.shape[0], rbbox_corners.shape[0]
    assert points.dim() == 2 and points.shape[1] >= 3 and tuple(rbbox_corners.shape[1:]) == (8, 3)
    out = torch.empty((n, m), device=points.device, dtype=torch.uint8)
    if n:
        with _timed("points_in_boxes", nbytes=n * (12 + m)):
            L.check(L.load().mobi_points_in_boxes(points.data_ptr(), points.shape[1], rbbox_corners.data_ptr(), out.data_ptr(),
                                                  n, m, L.stream()), "points_in_boxes")
    return out.bool()


class LidarConverter:
    """ldm/data/lidar_converter.py:22-38, the device-side half: range view -> point cloud and crop -> sweep."""

    def __init__(self, H=32, W=1096, depth_interval=(1.4, 54), log_scale=False, depth_scale=5.8):
        if log_scale:
            raise NotImplementedError("mobi_b200.lidar.LidarConverter: log_scale is never enabled by the reference configs")
        self.base_size = (H, W)
        self.depth_interval = depth_interval
        self.log_scale, self.depth_scale = log_scale, depth_scale

    def range2pcd_batch(self, range_depth, range_pitch, range_yaw, label=None, point_stride=None, want_index=False):
        """[B, H, W] sweeps -> (points [B, H*W, stride] padded, index [B, H*W] or None, count int32 [B]); no host sync.
        points[b, :count[b]] are the kept points in pixel order; columns: x, y, z, (label), (beam index)."""
        d, p, y = _f32c(range_depth), _f32c(range_pitch), _f32c(range_yaw)
        B, H, W = d.shape
        assert (H, W) == tuple(self.base_size), "range2pcd expects sweeps at the converter's base size"
        stride = point_stride or (5 if label is not None else 3)
        a = L.Range2PcdArgs()
        points = torch.empty((B, H * W, stride), device=d.device, dtype=torch.float32)
        count = torch.empty((B,), device=d.device, dtype=torch.int32)
        index = torch.empty((B, H * W), device=d.device, dtype=torch.int32) if want_index else None
        lab = None if label is None else _f32c(label)
        a.depth, a.pitch, a.yaw, a.label = d.data_ptr(), p.data_ptr(), y.data_ptr(), L.ptr(lab)
        a.points, a.index, a.count = points.data_ptr(), L.ptr(index), count.data_ptr()
        a.batch, a.H, a.W, a.point_stride = B, H, W, stride
        a.depth_min, a.depth_max = self.depth_interval
        with _timed("range_cloud", nbytes=B * H * W * (12 + 4 * stride)):
            L.check(L.load().mobi_range2pcd(C.byref(a), L.stream()), "range2pcd")
        return points, index, count

    def range2pcd(self, range_depth, range_pitch, range_yaw, label=None):
        """lidar_converter.py:122-176 for one [H, W] sweep: (pcd [N, 3], label [N] or None, beam_index [N]).  Reads N
        back (one host sync), as the ragged return type demands.  An integer `label` comes back as int64 pixel labels
        only for the `arange` use (inference_test_bench.py:587); other labels are carried as f32."""
        is_index = label is not None and not label.is_floating_point()
        pts, idx, cnt = self.range2pcd_batch(range_depth[None], range_pitch[None], range_yaw[None],
                                             None if (label is None or is_index) else label[None], point_stride=5,
                                             want_index=is_index)
        n = int(cnt.item())
        pts = pts[0, :n]
        if label is None:
            lab = None
        elif is_index:
            lab = label.reshape(-1)[idx[0, :n].long()]
        else:
            lab = pts[:, 3]
        return pts[:, :3], lab, pts[:, 4].long()

    def undo_default_transforms(self, crop_left, width_crop, range_depth_crop, range_depth, range_int_crop=None,
                                range_int=None, mask=None):
        """lidar_converter.py:436-485 for one sample ([h, w] crop, [H, W] sweep).  `mask` is not supported (no caller in the
        reference passes it)."""
        if mask is not None:
            raise NotImplementedError("undo_default_transforms(mask=...) has no caller in the reference")
        assert range_int is None or range_int_crop is not None, "If range_int is not None, range_int_crop must be provided"
        dev = range_depth.device
        cl = torch.tensor([int(crop_left)], device=dev)
        wc = torch.tensor([int(width_crop)], device=dev)
        crops, origs = [_crop_view(range_depth_crop[None])], [range_depth[None]]
        if range_int is not None:
            crops.append(_crop_view(range_int_crop[None]))
            origs.append(range_int[None])
        outs = _undo(crops, origs, cl, wc)
        return outs[0][0], (outs[1][0] if range_int is not None else None)


def composite_range_samples(range_sample_depth, range_sample_int, *, range_depth_orig, range_int_orig, range_pitch, range_yaw,
                            range_instance_mask_orig, bbox_3d, depth_interval=(1.4, 54)):
    """inference_test_bench.py:580-629 for the whole batch in ONE launch.  Returns a dict of CUDA tensors:
      pred_instance_mask u8 [B, H, W]; range_pred f32 [B, 4, H, W] = (depth, intensity, pitch, yaw) after the paste;
      pred_points f32 [B, H*W, 5] (padded) = (x, y, z, intensity, beam index) in pixel order; n_points int32 [B]."""
    sd, si = _f32c(range_sample_depth), _f32c(range_sample_int)
    B, H, W = sd.shape
    d0, i0, p, y = _f32c(range_depth_orig), _f32c(range_int_orig), _f32c(range_pitch), _f32c(range_yaw)
    gt, box = _f32c(range_instance_mask_orig), _f32c(bbox_3d)
    _cuda(sd, si, d0, i0, p, y, gt, box)
    assert tuple(box.shape) == (B, 8, 3), "one box [8, 3] per sample"
    for t in (si, d0, i0, p, y, gt):
        assert t.numel() == B * H * W
    dev = sd.device
    range_pred = torch.empty((B, 4, H, W), device=dev, dtype=torch.float32)
    pred_mask = torch.empty((B, H, W), device=dev, dtype=torch.uint8)
    points = torch.empty((B, H * W, 5), device=dev, dtype=torch.float32)
    count = torch.empty((B,), device=dev, dtype=torch.int32)
    a = L.RangeCompositeArgs()
    a.sample_depth, a.sample_int, a.depth_orig, a.int_orig = sd.data_ptr(), si.data_ptr(), d0.data_ptr(), i0.data_ptr()
    a.gt_mask, a.bbox, a.pitch, a.yaw = gt.data_ptr(), box.data_ptr(), p.data_ptr(), y.data_ptr()
    a.range_pred, a.pred_mask, a.points, a.count = range_pred.data_ptr(), pred_mask.data_ptr(), points.data_ptr(), count.data_ptr()
    a.batch, a.H, a.W = B, H, W
    a.depth_min, a.depth_max = depth_interval
    with _timed("range_cloud", nbytes=B * H * W * (7 * 4 + 4 * 4 + 1 + 5 * 4)):
        L.check(L.load().mobi_range_composite(C.byref(a), L.stream()), "range_composite")
    return dict(pred_instance_mask=pred_mask, range_pred=range_pred, pred_points=points, n_points=count)


def postprocess_lidar_samples(lidar_sample, lidar_batch, bbox_3d, *, range_object_norm=True, range_object_norm_scale=0.75,
                              unnormalize_intensity=False):
    """Decoded lidar image -> edited sweep and point cloud, two launches, nothing leaves HBM.

    lidar_sample: f32 [B, C >= 2, h, w] straight from `decode_first_stage(..., module_name="lidar_stage_model")`;
    lidar_batch: the `batch["lidar"]` dict of the reference dataset (nuscenes.py:470-489) on the device.
    Follows LatentDiffusion.log_images (ddpm.py:1503-1543) then inference_test_bench.py:567-629: clamp to [-1, 1]; depth
    channel through inverse_depth_normalization when `range_object_norm`; the intensity the reference LOGS is left
    normalised (ddpm.py:1541 rebinds `sample_int`, so log["range_sample_int"] keeps the network's scale) unless
    `unnormalize_intensity` asks for the un-normalised one.
    Returns composite_range_samples(...) plus range_sample_depth / range_sample_int [B, H, W]."""
    if not lidar_sample.is_cuda:
        raise RuntimeError("mobi_b200.lidar runs on CUDA tensors only (no CPU fallback)")
    x = lidar_sample.detach()
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.to(torch.float32).contiguous()
    maps = (MAP_DEPTH_UNNORM if range_object_norm else MAP_NONE, MAP_INT_UNNORM if unnormalize_intensity else MAP_NONE)
    sd, si = _undo([x[:, 0], x[:, 1]], [lidar_batch["range_depth_orig"], lidar_batch["range_int_orig"]],
                   lidar_batch["range_shift_left"], lidar_batch["width_crop"], maps=maps,
                   min_d=lidar_batch["min_depth_obj"] if range_object_norm else None,
                   max_d=lidar_batch["max_depth_obj"] if range_object_norm else None,
                   alpha=range_object_norm_scale, clamp=True)
    out = composite_range_samples(sd, si, range_depth_orig=lidar_batch["range_depth_orig"],
                                  range_int_orig=lidar_batch["range_int_orig"], range_pitch=lidar_batch["range_pitch"],
                                  range_yaw=lidar_batch["range_yaw"],
                                  range_instance_mask_orig=lidar_batch["range_instance_mask_orig"], bbox_3d=bbox_3d)
    out["range_sample_depth"], out["range_sample_int"] = sd, si
    return out
