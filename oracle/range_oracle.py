"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product
never does).  NumPy restatement of the range-view post-processing that follows the lidar decode:

  depth_normalization / inverse_depth_normalization      ldm/data/utils.py:537-580
  intensity un-normalisation                              ldm/models/diffusion/ddpm.py:1540-1543
  LidarConverter.resize (avg pool | cv2 INTER_NEAREST)    ldm/data/lidar_converter.py:8-19, 230-287
  LidarConverter.undo_default_transforms                  ldm/data/lidar_converter.py:436-485
  postprocess_range_depth_int                             ldm/data/utils.py:471-505
  LidarConverter.range2pcd                                ldm/data/lidar_converter.py:122-176
  corner_to_surfaces_3d / surface_equ_3d / point-in-box   ldm/data/box_np_ops.py:406-427, 712-771, 453-471
  the save_samples sequence (instance mask, paste, cloud) scripts/inference_test_bench.py:567-629

Pinned against the unmodified reference by tests/golden/range_post.npz (oracle/make_golden_range.py, which imports
ldm.data.* with the real cv2 / numba / torch CPU kernels).  All arithmetic is float32 in the reference's operation order.
"""
import numpy as np

F32 = np.float32
DEPTH_INTERVAL = (1.4, 54)
IGNORE = -1000


# ------------------------------------------------------------------------------------------------ normalisations
def depth_normalization(depth, min_d, max_d, alpha=0.75):
    """utils.py:537-557 (three linear pieces; values outside [-1, 1] are left untouched here, the reference leaves
    them uninitialised)."""
    d = np.asarray(depth, F32)
    min_d, max_d, a = F32(min_d), F32(max_d), F32(alpha)
    out = d.copy()
    mid = (d >= min_d) & (d <= max_d)
    out[mid] = -a + F32(2 * alpha) * (d[mid] - min_d) / (max_d - min_d)
    low = (d >= -1) & (d < min_d)
    out[low] = F32(-1) + F32(-(alpha - 1)) * (d[low] + F32(1)) / (min_d + F32(1))
    high = (d > max_d) & (d <= 1)
    out[high] = a + F32(1 - alpha) * (d[high] - max_d) / (F32(1) - max_d)
    return out


def inverse_depth_normalization(normalized_depth, min_d, max_d, alpha=0.75):
    """utils.py:560-580."""
    x = np.asarray(normalized_depth, F32)
    min_d, max_d, a = F32(min_d), F32(max_d), F32(alpha)
    out = x.copy()
    mid = (x >= -a) & (x <= a)
    out[mid] = min_d + (x[mid] + a) * (max_d - min_d) / F32(2 * alpha)
    low = (x >= -1) & (x < -a)
    out[low] = F32(-1) + -(x[low] + F32(1)) * (min_d + F32(1)) / F32(alpha - 1)
    high = (x > a) & (x <= 1)
    out[high] = max_d + (x[high] - a) * (F32(1) - max_d) / F32(1 - alpha)
    return out


def intensity_unnormalization(x):
    """ddpm.py:1541: clamp(-0.5 * log(1 - (x + 1) / 2) - 1, -1, 1)."""
    x = np.asarray(x, F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        y = F32(-0.5) * np.log(F32(1) - (x + F32(1)) / F32(2)) - F32(1)
    return np.clip(y, F32(-1), F32(1)).astype(F32)


# ------------------------------------------------------------------------------------------------------ resizing
def avg_pool_resize(x, new_H, new_W):
    """pool_resize(mode="avg_pool") (lidar_converter.py:8-13): F.avg_pool2d with kernel (H // new_H, W // new_W); the
    CPU kernel sums the window row by row in fp32 and divides once."""
    x = np.asarray(x, F32)
    H, W = x.shape
    kh, kw = H // new_H, W // new_W
    oh, ow = (H - kh) // kh + 1, (W - kw) // kw + 1
    acc = np.zeros((oh, ow), F32)
    for i in range(kh):
        for j in range(kw):
            acc = acc + x[i:i + oh * kh:kh, j:j + ow * kw:kw]
    return acc / F32(kh * kw)


def nearest_resize(x, new_H, new_W):
    """cv2.resize(..., interpolation=cv2.INTER_NEAREST): src index = min(floor(dst * (1 / (dst_size / src_size))), src - 1)
    in double precision."""
    x = np.asarray(x)
    H, W = x.shape
    ifx, ify = 1.0 / (float(new_W) / W), 1.0 / (float(new_H) / H)
    sx = np.minimum(np.floor(np.arange(new_W) * ifx).astype(np.int64), W - 1)
    sy = np.minimum(np.floor(np.arange(new_H) * ify).astype(np.int64), H - 1)
    return x[sy][:, sx]


def resize(x, new_H, new_W):
    """LidarConverter.resize for one float array (lidar_converter.py:259-266)."""
    if x.shape == (new_H, new_W):
        return np.asarray(x, F32).copy()
    if x.shape[0] % new_H == 0 and x.shape[1] % new_W == 0:
        return avg_pool_resize(x, new_H, new_W)
    return nearest_resize(np.asarray(x, F32), new_H, new_W)


def undo_default_transforms(crop_left, width_crop, range_depth_crop, range_depth, range_int_crop=None, range_int=None):
    """lidar_converter.py:436-485 without the optional `mask` (no caller passes it): shrink the square crop back to
    [H, width_crop] and paste it at column `crop_left % W`, wrapping around the 360 degree seam."""
    H, W = range_depth.shape
    crop_left = int(crop_left) % W
    width_crop = int(width_crop)
    outs = []
    for crop, full in ((range_depth_crop, range_depth), (range_int_crop, range_int)):
        if full is None:
            outs.append(None)
            continue
        small = resize(crop, H, width_crop)
        aux = np.asarray(full, F32).copy()
        right = min(crop_left + small.shape[1], W)
        aux[:, crop_left:right] = small[:, :right - crop_left]
        aux[:, :width_crop - (right - crop_left)] = small[:, right - crop_left:]
        outs.append(np.where(aux == IGNORE, full, aux).astype(F32))
    return outs[0], outs[1]


def postprocess_range_depth_int(range_depth, range_depth_orig, range_int, range_int_orig, crop_left, width_crop,
                                zero_context=False):
    """utils.py:471-505: batched undo_default_transforms; inputs [B, 1, h, w] crops and [B, H, W] originals."""
    if zero_context:
        range_depth_orig = range_depth_orig * 0 - 1
    d_all, i_all = [], []
    for b in range(len(range_depth)):
        d, i = undo_default_transforms(crop_left[b], width_crop[b], range_depth[b, 0], range_depth_orig[b],
                                       range_int[b, 0], range_int_orig[b])
        d_all.append(d)
        i_all.append(i)
    return np.stack(d_all), np.stack(i_all)


# ---------------------------------------------------------------------------------------------------- point cloud
def range2pcd(range_depth, range_pitch, range_yaw, label=None, depth_interval=DEPTH_INTERVAL):
    """lidar_converter.py:122-176 for an image already at the base size (log_scale False, the default)."""
    d = (np.asarray(range_depth, F32) + F32(1)) / F32(2)
    depth = (d * F32(depth_interval[1])).flatten()
    yaw = np.asarray(range_yaw, F32).flatten()
    pitch = np.asarray(range_pitch, F32).flatten()
    pcd = np.zeros((len(yaw), 3), F32)
    pcd[:, 0] = np.cos(yaw) * np.cos(pitch) * depth
    pcd[:, 1] = -np.sin(yaw) * np.cos(pitch) * depth
    pcd[:, 2] = np.sin(pitch) * depth
    mask = np.logical_and(depth > depth_interval[0], depth < depth_interval[1])
    H, W = range_pitch.shape
    beam = np.tile(np.arange(H - 1, -1, -1).reshape(H, 1), (1, W)).flatten()[mask]
    return pcd[mask], (None if label is None else np.asarray(label).flatten()[mask]), beam


SURFACE_CORNERS = ((0, 1, 2, 3), (7, 6, 5, 4), (0, 3, 7, 4), (1, 5, 6, 2), (0, 4, 5, 1), (3, 2, 6, 7))


def box_planes(corners):
    """corner_to_surfaces_3d + surface_equ_3d (box_np_ops.py:406-427, 712-732): [8, 3] corners -> normals [6, 3], d [6]
    of a x + b y + c z + d = 0 with the normals pointing inwards."""
    c = np.asarray(corners)
    surf = np.stack([c[list(idx)] for idx in SURFACE_CORNERS])                     # [6, 4, 3]
    vec = surf[:, :2] - surf[:, 1:3]
    n = np.cross(vec[:, 0], vec[:, 1])
    d = -np.einsum("ij,ij->i", n, surf[:, 0])
    return n, d


def points_in_bbox_corners(points, rbbox_corners):
    """box_np_ops.py:453-471 / 736-771: [N, 3+] points, [M, 8, 3] corners -> bool [N, M] (inside <=> every plane
    value < 0)."""
    pts = np.asarray(points)[:, :3]
    ret = np.ones((pts.shape[0], len(rbbox_corners)), bool)
    for j, corners in enumerate(rbbox_corners):
        n, d = box_planes(corners)
        for k in range(6):
            sign = pts[:, 0] * n[k, 0] + pts[:, 1] * n[k, 1] + pts[:, 2] * n[k, 2] + d[k]
            ret[:, j] &= ~(sign >= 0)
    return ret


def box_corners(center, dims, yaw):
    """center_to_corner_box3d(origin=(0.5, 0.5, 0.5), axis=2) for one box (box_np_ops.py:48-78, 178-237): the corner
    order the surface table above expects."""
    norm = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 1], [0, 1, 0], [1, 0, 0], [1, 0, 1], [1, 1, 1], [1, 1, 0]], F32) - F32(0.5)
    c = norm * np.asarray(dims, F32)[None]
    s, co = F32(np.sin(yaw)), F32(np.cos(yaw))
    rot_t = np.array([[co, -s, 0], [s, co, 0], [0, 0, 1]], F32)
    return (c @ rot_t + np.asarray(center, F32)[None]).astype(F32)


# -------------------------------------------------------------------------------- the save_samples sequence per sample
def composite_sample(range_sample_depth, range_sample_int, depth_orig, int_orig, pitch, yaw, bbox_3d, gt_instance_mask):
    """inference_test_bench.py:580-629 for one sample: instance mask of the generated object, paste into the original
    sweep, edited point cloud [N, 5] = (x, y, z, intensity, beam index)."""
    shape = gt_instance_mask.shape
    label = np.arange(0, int(np.prod(shape))).reshape(shape)
    points, points_label, _ = range2pcd(range_sample_depth, pitch, yaw, label)
    inside = points_in_bbox_corners(points, bbox_3d)
    pred_mask = np.zeros(int(np.prod(shape)))
    pred_mask[points_label[inside[:, 0]]] = 1
    pred_mask = pred_mask.reshape(shape)
    instance_mask = np.logical_or(pred_mask, gt_instance_mask)
    depth_final = np.where(instance_mask, range_sample_depth, depth_orig)
    int_final = np.where(instance_mask, range_sample_int, int_orig)
    range_pred = np.stack([depth_final, int_final, pitch, yaw])
    xyz, pts_int, beam = range2pcd(depth_final, pitch, yaw, int_final)
    pred_points = np.concatenate([xyz, pts_int[:, None], beam[:, None]], axis=1)
    return dict(pred_instance_mask=pred_mask, range_pred=range_pred, pred_points=pred_points)


# ------------------------------------------------------------------------------------------------ synthetic sweeps
def default_pitch_yaw(H=32, W=1096):
    """The empty-pixel tables of pcd2range (lidar_converter.py:86-94)."""
    beam = np.array([0.0232 * x for x in range(-23, 9)])
    scan_x = (np.arange(W, dtype=F32) / F32(W))[None].repeat(H, 0)
    yaw = (np.pi * (scan_x * 2 - 1)).astype(F32)
    pitch = np.zeros((H, W), F32)
    rows = np.linspace(0, 31, H).round().astype(int) if H != 32 else np.arange(32)
    for i, r in enumerate(rows):
        pitch[i, :] = beam[31 - r]
    return pitch, yaw


def synth_range_inputs(seed, B=2, H=32, W=1096, crop=128, width_crops=(64, 128), obj_range_m=10.0):
    """Deterministic synthetic sweeps (uniform RNG and arithmetic only, so every platform regenerates the same bits):
    a box ~obj_range_m ahead of the sensor whose direction sits in the middle of the crop window, a decoded crop whose
    middle columns carry object-normalised depths inside [min_d, max_d], holes (-1) in the original sweep."""
    rng = np.random.default_rng(seed)
    pitch0, yaw0 = default_pitch_yaw(H, W)
    out = dict(range_depth=[], range_int=[], range_depth_orig=[], range_int_orig=[], range_pitch=[], range_yaw=[],
               range_instance_mask_orig=[], crop_left=[], width_crop=[], min_depth_obj=[], max_depth_obj=[], centers=[])
    for b in range(B):
        wc = int(width_crops[b % len(width_crops)])
        col = int(rng.integers(0, W))                                       # object direction (column of the sweep)
        if b == 1:
            col = W - 7                                                     # this window wraps the 360 degree seam
        crop_left = (col - wc // 2) % W + W                                 # tiled (x3) coordinates, may wrap the seam
        d_orig = rng.uniform(-0.95, 0.9, (H, W)).astype(F32)
        d_orig[rng.uniform(size=(H, W)) < 0.15] = -1
        i_orig = rng.uniform(0, 0.6, (H, W)).astype(F32)
        pitch = (pitch0 + rng.uniform(-0.002, 0.002, (H, W)).astype(F32)).astype(F32)
        yaw = (yaw0 + rng.uniform(-0.001, 0.001, (H, W)).astype(F32)).astype(F32)
        gt = np.zeros((H, W), F32)
        gt[H // 2:H // 2 + 3, [(col + k) % W for k in range(-2, 3)]] = 1
        d_obj = F32(2 * obj_range_m / DEPTH_INTERVAL[1] - 1)
        min_d, max_d = F32(d_obj - 0.03), F32(d_obj + 0.03)
        dc = rng.uniform(-1, 1, (crop, crop)).astype(F32)
        lo, hi = int(crop * 0.3), int(crop * 0.7)
        dc[:, lo:hi] = rng.uniform(-0.75, 0.75, (crop, hi - lo)).astype(F32)
        ic = rng.uniform(-1, 1, (crop, crop)).astype(F32)
        out["range_depth"].append(dc[None])
        out["range_int"].append(ic[None])
        out["range_depth_orig"].append(d_orig)
        out["range_int_orig"].append(i_orig)
        out["range_pitch"].append(pitch)
        out["range_yaw"].append(yaw)
        out["range_instance_mask_orig"].append(gt)
        out["crop_left"].append(crop_left)
        out["width_crop"].append(wc)
        out["min_depth_obj"].append(min_d)
        out["max_depth_obj"].append(max_d)
        out["centers"].append((float(yaw0[0, col]), obj_range_m))
    res = {k: np.stack(v) if k not in ("centers",) else v for k, v in out.items()}
    res["crop_left"] = res["crop_left"].astype(np.int64)
    res["width_crop"] = res["width_crop"].astype(np.int64)
    return res


def synth_boxes(inputs, dims=(4.6, 2.2, 1.8)):
    """One box per sample, long axis along the viewing ray, centred obj_range_m ahead and 0.4 m below the sensor."""
    boxes = []
    for yaw_c, r in inputs["centers"]:
        center = (r * np.cos(yaw_c), -r * np.sin(yaw_c), -0.4)
        boxes.append(box_corners(center, dims, -yaw_c))
    return np.stack(boxes).astype(F32)


def run_pipeline(inp, bbox_3d, alpha=0.75, int_unnorm=False):
    """ddpm.py:1503-1543 (clamp, in-place inverse depth normalisation of the logged sample; the logged intensity is
    NOT un-normalised — `sample_int` is rebound, ddpm.py:1541) followed by inference_test_bench.py:567-629."""
    B = len(inp["range_depth"])
    depth = np.clip(inp["range_depth"], -1, 1).astype(F32)
    inten = np.clip(inp["range_int"], -1, 1).astype(F32)
    for b in range(B):
        depth[b] = inverse_depth_normalization(depth[b], inp["min_depth_obj"][b], inp["max_depth_obj"][b], alpha)
    if int_unnorm:
        inten = intensity_unnormalization(inten)
    sd, si = postprocess_range_depth_int(depth, inp["range_depth_orig"], inten, inp["range_int_orig"],
                                         inp["crop_left"], inp["width_crop"])
    outs = [composite_sample(sd[b], si[b], inp["range_depth_orig"][b], inp["range_int_orig"][b], inp["range_pitch"][b],
                             inp["range_yaw"][b], bbox_3d[[b]], inp["range_instance_mask_orig"][b]) for b in range(B)]
    return dict(range_sample_depth=sd, range_sample_int=si,
                pred_instance_mask=np.stack([o["pred_instance_mask"] for o in outs]),
                range_pred=np.stack([o["range_pred"] for o in outs]),
                pred_points=[o["pred_points"] for o in outs])
