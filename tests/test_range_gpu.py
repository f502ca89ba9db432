"""Range-view post-processing on the device (mobi_b200.lidar -> C ABI -> range_view.cu) against the NumPy oracle and the
fixtures the unmodified reference produced (tests/golden/range_post.npz).  SURVEY.md §8(f) row 3.

Bars: every integer / index / mask result and every value that involves no sin / cos (the avg-pool or nearest shrink,
the wrap-around paste, the piecewise depth maps) is BIT-EXACT; point coordinates are within 2e-6 relative of the range
(sinf / cosf are <= 2 ulp on the device, NumPy's SIMD ones likewise); the intensity un-normalisation (logf) within 2e-6.
"""
import os

import numpy as np
import pytest
import torch

from oracle import range_oracle as ro

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "range_post.npz")
DEV = "cuda:0"


def cu(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    return t if dtype is None else t.to(dtype)


def lidar_batch(inp):
    return dict(range_depth_orig=cu(inp["range_depth_orig"]), range_int_orig=cu(inp["range_int_orig"]),
                range_pitch=cu(inp["range_pitch"]), range_yaw=cu(inp["range_yaw"]),
                range_instance_mask_orig=cu(inp["range_instance_mask_orig"]), range_shift_left=cu(inp["crop_left"]),
                width_crop=cu(inp["width_crop"]), min_depth_obj=cu(inp["min_depth_obj"]), max_depth_obj=cu(inp["max_depth_obj"]))


def check_against(out, want, inp, coord_tol=2e-6 * 54):
    B = len(inp["range_depth"])
    assert np.array_equal(out["range_sample_depth"].cpu().numpy(), want["range_sample_depth"])
    assert np.array_equal(out["range_sample_int"].cpu().numpy(), want["range_sample_int"])
    assert np.array_equal(out["pred_instance_mask"].cpu().numpy(), want["pred_instance_mask"].astype(np.uint8))
    assert np.array_equal(out["range_pred"].cpu().numpy(), want["range_pred"].astype(np.float32))
    n = out["n_points"].cpu().numpy()
    pts = out["pred_points"].cpu().numpy()
    for b in range(B):
        w = want["pred_points"][b]
        assert n[b] == len(w), (b, n[b], len(w))
        got = pts[b, :n[b]]
        assert np.abs(got[:, :3] - w[:, :3]).max() <= coord_tol
        assert np.array_equal(got[:, 3], w[:, 3].astype(np.float32))            # intensity carried through untouched
        assert np.array_equal(got[:, 4], w[:, 4].astype(np.float32))            # beam index


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN, allow_pickle=False)


@pytest.fixture(scope="module")
def case(gold):
    inp = ro.synth_range_inputs(int(gold["seed"]), B=int(gold["B"]), H=int(gold["H"]), W=int(gold["W"]),
                                crop=int(gold["crop"]), width_crops=tuple(int(w) for w in gold["width_crops"]))
    return inp, gold["bbox_3d"]


def test_scalar_maps(gold):
    from mobi_b200 import lidar
    grid = np.linspace(-1, 1, 4001, dtype=np.float64).astype(np.float32)
    g = cu(grid)
    mn, mx = float(gold["grid_min"]), float(gold["grid_max"])
    fwd = lidar.depth_normalization(g, torch.tensor(mn), torch.tensor(mx)).cpu().numpy()
    inv = lidar.inverse_depth_normalization(g, torch.tensor(mn), torch.tensor(mx)).cpu().numpy()
    assert np.array_equal(fwd, ro.depth_normalization(grid, mn, mx))
    assert np.array_equal(inv, ro.inverse_depth_normalization(grid, mn, mx))
    assert np.abs(fwd - gold["grid_fwd"]).max() < 2e-5 and np.abs(inv - gold["grid_inv"]).max() < 2e-6
    un = lidar.intensity_unnormalization(g).cpu().numpy()
    want = ro.intensity_unnormalization(grid)
    assert np.abs(un - want).max() <= 2e-6 and un[-1] == 1.0 and un[0] == -1.0
    # a second alpha and per-sample bounds; round trip inverse(forward(x)) == x within fp32 rounding
    x = np.random.default_rng(0).uniform(-1, 1, (3, 5, 7)).astype(np.float32)
    mns, mxs = np.array([-0.9, -0.2, 0.3], np.float32), np.array([-0.5, 0.4, 0.95], np.float32)
    f = lidar.depth_normalization(cu(x), cu(mns), cu(mxs), alpha=0.6)
    for b in range(3):
        assert np.array_equal(f[b].cpu().numpy(), ro.depth_normalization(x[b], mns[b], mxs[b], 0.6))
    back = lidar.inverse_depth_normalization(f, cu(mns), cu(mxs), alpha=0.6).cpu().numpy()
    assert np.abs(back - x).max() < 5e-6


def test_staged_pipeline_matches_reference_fixture(gold, case):
    """The reference's own call sequence, function by function, through the drop-in names."""
    from mobi_b200 import lidar
    inp, bbox = case
    B, H, W = int(gold["B"]), int(gold["H"]), int(gold["W"])
    lb = lidar_batch(inp)
    sample = torch.clamp(cu(np.concatenate([inp["range_depth"], inp["range_int"]], 1)), -1., 1.)
    sample_depth, sample_int = sample[:, [0]], sample[:, [1]]
    for i in range(B):                                                            # ddpm.py:1533-1536
        sample_depth[i] = lidar.inverse_depth_normalization(sample_depth[i], lb["min_depth_obj"][i], lb["max_depth_obj"][i],
                                                            alpha=0.75)
    assert np.array_equal(sample_depth[0].cpu().numpy(), gold["unnorm_depth_0"])
    sd, si = lidar.postprocess_range_depth_int(range_depth=sample_depth, range_depth_orig=lb["range_depth_orig"],
                                               range_int=sample_int, range_int_orig=lb["range_int_orig"],
                                               crop_left=lb["range_shift_left"], width_crop=lb["width_crop"])
    masks = np.unpackbits(gold["pred_instance_mask"])[:B * H * W].reshape(B, H, W)
    conv = lidar.LidarConverter()
    for b in range(B):
        win = (int(inp["crop_left"][b]) % W + np.arange(int(inp["width_crop"][b]))) % W
        assert np.array_equal(sd[b].cpu().numpy()[:, win], gold["window_depth_%d" % b])
        assert np.array_equal(si[b].cpu().numpy()[:, win], gold["window_int_%d" % b])
        rest = np.setdiff1d(np.arange(W), win)
        assert np.array_equal(sd[b].cpu().numpy()[:, rest], inp["range_depth_orig"][b][:, rest])
        # inference_test_bench.py:586-598
        label = torch.arange(H * W, device=DEV).reshape(H, W)
        points, points_label, _ = conv.range2pcd(sd[b], lb["range_pitch"][b], lb["range_yaw"][b], label)
        inside = lidar.points_in_bbox_corners(points, cu(bbox[[b]]))
        pred = torch.zeros(H * W, device=DEV)
        pred[points_label[inside[:, 0]]] = 1
        assert np.array_equal(pred.reshape(H, W).cpu().numpy().astype(np.uint8), masks[b])
        instance = torch.logical_or(pred.reshape(H, W) > 0, lb["range_instance_mask_orig"][b] > 0)
        depth_final = torch.where(instance, sd[b], lb["range_depth_orig"][b])
        int_final = torch.where(instance, si[b], lb["range_int_orig"][b])
        xyz, pint, beam = conv.range2pcd(depth_final, lb["range_pitch"][b], lb["range_yaw"][b], int_final)
        pp = torch.cat([xyz, pint[:, None], beam[:, None].float()], 1).cpu().numpy()
        assert len(pp) == int(gold["n_points"][b])
        assert np.abs(pp[::7] - gold["points_every7_%d" % b]).max() < 2e-4
        assert np.abs(pp.astype(np.float64).sum(0) - gold["points_colsum_%d" % b]).max() < 5e-2


def test_fused_pipeline_matches_oracle_and_fixture(gold, case):
    from mobi_b200 import lidar
    inp, bbox = case
    B, H, W = int(gold["B"]), int(gold["H"]), int(gold["W"])
    dec = cu(np.concatenate([inp["range_depth"] * 1.3, inp["range_int"] * 1.3, inp["range_int"]], 1))   # un-clamped, 3 channels
    inp2 = dict(inp, range_depth=np.clip(inp["range_depth"] * 1.3, -1, 1), range_int=np.clip(inp["range_int"] * 1.3, -1, 1))
    out = lidar.postprocess_lidar_samples(dec, lidar_batch(inp), cu(bbox))
    check_against(out, ro.run_pipeline(inp2, bbox), inp2)
    # and on the fixture's own (already in-range) inputs
    dec = cu(np.concatenate([inp["range_depth"], inp["range_int"]], 1))
    out = lidar.postprocess_lidar_samples(dec, lidar_batch(inp), cu(bbox))
    check_against(out, ro.run_pipeline(inp, bbox), inp)
    masks = np.unpackbits(gold["pred_instance_mask"])[:B * H * W].reshape(B, H, W)
    assert np.array_equal(out["pred_instance_mask"].cpu().numpy(), masks)
    assert out["n_points"].cpu().tolist() == gold["n_points"].tolist()
    pts = out["pred_points"].cpu().numpy()
    for b in range(B):
        assert np.abs(pts[b, :int(gold["n_points"][b])][::7] - gold["points_every7_%d" % b]).max() < 2e-4


def test_unnormalised_intensity_option(case):
    from mobi_b200 import lidar
    inp, bbox = case
    dec = cu(np.concatenate([inp["range_depth"], inp["range_int"] * 0.98], 1))
    out = lidar.postprocess_lidar_samples(dec, lidar_batch(inp), cu(bbox), unnormalize_intensity=True)
    inp2 = dict(inp, range_int=inp["range_int"] * np.float32(0.98))
    want = ro.run_pipeline(inp2, bbox, int_unnorm=True)
    assert np.array_equal(out["range_sample_depth"].cpu().numpy(), want["range_sample_depth"])
    assert np.abs(out["range_sample_int"].cpu().numpy() - want["range_sample_int"]).max() < 5e-6   # logf then a 16-256 mean
    assert np.array_equal(out["pred_instance_mask"].cpu().numpy(), want["pred_instance_mask"].astype(np.uint8))


@pytest.mark.parametrize("crop,wcs", [(512, (64, 128, 256, 512)), (256, (256, 64, 100, 37))])
def test_full_size_batch(crop, wcs):
    """BASELINE-size inputs: 8 sweeps of 32 x 1096 with 512 x 512 (and 256 x 256) decoded crops, every width class."""
    from mobi_b200 import lidar
    inp = ro.synth_range_inputs(7, B=8, crop=crop, width_crops=wcs)
    bbox = ro.synth_boxes(inp)
    dec = cu(np.concatenate([inp["range_depth"], inp["range_int"]], 1))
    out = lidar.postprocess_lidar_samples(dec, lidar_batch(inp), cu(bbox))
    check_against(out, ro.run_pipeline(inp, bbox), inp)
    assert int(out["pred_instance_mask"].sum()) > 8 * 50


def test_properties_at_scale():
    """Size-independent properties on 32 sweeps: counts equal the number of in-range pixels, the cloud is in pixel order
    (beam index never increases), clouding the pasted sweep again reproduces the cloud, pasted pixels come from the
    sample and all others from the original."""
    from mobi_b200 import lidar
    inp = ro.synth_range_inputs(11, B=32, crop=128, width_crops=(64, 128))
    bbox = ro.synth_boxes(inp)
    lb = lidar_batch(inp)
    out = lidar.postprocess_lidar_samples(cu(np.concatenate([inp["range_depth"], inp["range_int"]], 1)), lb, cu(bbox))
    rp, n = out["range_pred"], out["n_points"]
    metres = (rp[:, 0] + 1) / 2 * 54
    assert torch.equal(((metres > 1.4) & (metres < 54)).flatten(1).sum(1).int(), n)
    conv = lidar.LidarConverter()
    pts2, idx2, n2 = conv.range2pcd_batch(rp[:, 0], rp[:, 2], rp[:, 3], label=rp[:, 1], want_index=True)
    assert torch.equal(n2, n)
    for b in range(32):
        k = int(n[b])
        assert torch.equal(pts2[b, :k], out["pred_points"][b, :k])
        assert bool((idx2[b, 1:k] > idx2[b, :k - 1]).all())
        beam = out["pred_points"][b, :k, 4]
        assert bool((beam[1:] <= beam[:-1]).all()) and beam.max() <= 31 and beam.min() >= 0
    paste = (out["pred_instance_mask"] > 0) | (lb["range_instance_mask_orig"] > 0)
    assert torch.equal(rp[:, 0], torch.where(paste, out["range_sample_depth"], lb["range_depth_orig"]))
    assert torch.equal(rp[:, 1], torch.where(paste, out["range_sample_int"], lb["range_int_orig"]))


def test_edge_cases():
    from mobi_b200 import lidar
    H, W = 5, 37                                                           # H * W not a multiple of 32, tiny clusters
    conv = lidar.LidarConverter(H=H, W=W)
    rng = np.random.default_rng(3)
    pitch = cu(rng.uniform(-0.5, 0.2, (2, H, W)).astype(np.float32))
    yaw = cu(rng.uniform(-3.1, 3.1, (2, H, W)).astype(np.float32))
    empty = torch.full((2, H, W), -1.0, device=DEV)                        # every pixel is a hole: zero points
    pts, _, cnt = conv.range2pcd_batch(empty, pitch, yaw)
    assert cnt.tolist() == [0, 0]
    full = torch.zeros((2, H, W), device=DEV)                               # 27 m everywhere: every pixel kept
    pts, idx, cnt = conv.range2pcd_batch(full, pitch, yaw, want_index=True)
    assert cnt.tolist() == [H * W, H * W] and torch.equal(idx[0], torch.arange(H * W, device=DEV, dtype=torch.int32))
    want, _, beam = ro.range2pcd(full[0].cpu().numpy(), pitch[0].cpu().numpy(), yaw[0].cpu().numpy())
    assert np.abs(pts[0].cpu().numpy() - want).max() < 2e-6 * 54
    # the interval is open: exactly 1.4 m (float32) and exactly 54 m are dropped
    edge = torch.tensor([[np.float32(1.4) / 54 * 2 - 1, 1.0, 0.5]], device=DEV).reshape(1, 1, 3)
    conv3 = lidar.LidarConverter(H=1, W=3)
    got = conv3.range2pcd(edge[0], torch.zeros(1, 3, device=DEV), torch.zeros(1, 3, device=DEV))
    m = ((edge[0].cpu().numpy() + np.float32(1)) / np.float32(2) * np.float32(54)).flatten()
    assert len(got[0]) == int(((m > np.float32(1.4)) & (m < 54)).sum())
    # single-sample undo with and without intensity, zero_context, crop already at the target size
    d_crop = cu(rng.uniform(-1, 1, (H, 8)).astype(np.float32))
    d_full = cu(rng.uniform(-1, 1, (H, W)).astype(np.float32))
    d, i = conv.undo_default_transforms(W + 33, 8, d_crop, d_full)          # wraps: columns 33..36 and 0..3
    assert i is None
    wd, _ = ro.undo_default_transforms(W + 33, 8, d_crop.cpu().numpy(), d_full.cpu().numpy())
    assert np.array_equal(d.cpu().numpy(), wd)
    z = lidar.postprocess_range_depth(range_depth=d_crop[None, None], range_depth_orig=d_full[None],
                                      crop_left=torch.tensor([2]), width_crop=torch.tensor([8]), zero_context=True)
    wz = ro.postprocess_range_depth_int(d_crop[None, None].cpu().numpy(), d_full[None].cpu().numpy(),
                                        d_crop[None, None].cpu().numpy(), d_full[None].cpu().numpy(), [2], [8], True)[0]
    assert np.array_equal(z.cpu().numpy(), wz)
    with pytest.raises(RuntimeError):
        lidar.inverse_depth_normalization(torch.zeros(4), 0.0, 0.5)            # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        lidar.LidarConverter(log_scale=True)


def test_points_in_boxes_matches_oracle():
    from mobi_b200 import lidar
    rng = np.random.default_rng(5)
    boxes = np.stack([ro.box_corners((3, 1, 0), (4, 2, 1.5), 0.3), ro.box_corners((-2, 5, 0.5), (1, 1, 1), -1.1),
                      ro.box_corners((0, 0, 0), (2, 2, 2), 0.0)]).astype(np.float32)
    pts = rng.uniform(-6, 6, (20000, 4)).astype(np.float32)
    pts[:5, :3] = [[0, 0, 0], [0.99, 0.99, 0.99], [1.0, 0, 0], [1.01, 0, 0], [0, -1.5, 0]]
    got = lidar.points_in_bbox_corners(cu(pts), cu(boxes)).cpu().numpy()
    want = ro.points_in_bbox_corners(pts, boxes)
    assert np.array_equal(got, want) and want.sum() > 200
    assert got[:5, 2].tolist() == [True, True, False, False, False]
    assert lidar.points_in_bbox_corners(cu(pts[:0]), cu(boxes)).shape == (0, 3)
