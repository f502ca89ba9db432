"""Pins oracle/range_oracle.py (NumPy restatement of the range-view post-processing) to
tests/golden/range_post.npz, which oracle/make_golden_range.py produced with the UNMODIFIED reference (real cv2, numba
and torch CPU kernels).  Inputs are regenerated from the seed; only outputs and the box corners are stored."""
import os

import numpy as np
import pytest

from oracle import range_oracle as ro

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "range_post.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN, allow_pickle=False)


def load_case(g):
    inp = ro.synth_range_inputs(int(g["seed"]), B=int(g["B"]), H=int(g["H"]), W=int(g["W"]), crop=int(g["crop"]),
                                width_crops=tuple(int(w) for w in g["width_crops"]))
    return inp, g["bbox_3d"]


def test_scalar_maps_bit_exact(gold):
    grid = np.linspace(-1, 1, 4001, dtype=np.float64).astype(np.float32)
    # torch.linspace(-1, 1, 4001) in fp32 is not bit-identical to the fp64 linspace; both sides see the same grid only
    # within 1 ulp, so compare with a tolerance of a few ulps scaled by the steepest slope (12.5 on the low piece)
    fwd = ro.depth_normalization(grid, gold["grid_min"], gold["grid_max"])
    inv = ro.inverse_depth_normalization(grid, gold["grid_min"], gold["grid_max"])
    assert np.abs(fwd - gold["grid_fwd"]).max() < 2e-5
    assert np.abs(inv - gold["grid_inv"]).max() < 2e-6
    un = ro.intensity_unnormalization(grid)
    assert np.abs(un - gold["grid_int"]).max() < 2e-4            # the log is steep near x -> 1
    assert un[-1] == 1.0 and un[0] == -1.0


def test_inverse_normalisation_of_the_decoded_crop_is_bit_exact(gold):
    inp, _ = load_case(gold)
    d = np.clip(inp["range_depth"][0], -1, 1)
    mine = ro.inverse_depth_normalization(d, inp["min_depth_obj"][0], inp["max_depth_obj"][0])
    assert np.array_equal(mine, gold["unnorm_depth_0"])


def test_pipeline_matches_the_reference(gold):
    inp, bbox = load_case(gold)
    out = ro.run_pipeline(inp, bbox)
    B, H, W = int(gold["B"]), int(gold["H"]), int(gold["W"])
    masks = np.unpackbits(gold["pred_instance_mask"])[:B * H * W].reshape(B, H, W)
    for b in range(B):
        win = (int(inp["crop_left"][b]) % W + np.arange(int(inp["width_crop"][b]))) % W
        # avg-pool (divisible) and cv2-nearest (width 100) shrink + wrap-around paste: bit-exact
        assert np.array_equal(out["range_sample_depth"][b][:, win], gold["window_depth_%d" % b]), b
        assert np.array_equal(out["range_sample_int"][b][:, win], gold["window_int_%d" % b]), b
        rest = np.setdiff1d(np.arange(W), win)
        assert np.array_equal(out["range_sample_depth"][b][:, rest], inp["range_depth_orig"][b][:, rest])
        assert np.array_equal(out["pred_instance_mask"][b].astype(np.uint8), masks[b]), b
        pp = out["pred_points"][b]
        assert len(pp) == int(gold["n_points"][b])
        assert np.abs(pp[::7] - gold["points_every7_%d" % b]).max() < 1e-5
        assert np.abs(pp.astype(np.float64).sum(0) - gold["points_colsum_%d" % b]).max() < 1e-2
        assert masks[b].sum() > 100                                   # the synthetic object really lands in its box


def test_wraparound_and_nearest_cases_are_present(gold):
    inp, _ = load_case(gold)
    W = int(gold["W"])
    assert any(int(c) % W + int(w) > W for c, w in zip(inp["crop_left"], inp["width_crop"]))
    assert any(int(gold["crop"]) % int(w) != 0 for w in inp["width_crop"])


def test_resize_edge_cases():
    x = np.arange(32 * 64, dtype=np.float32).reshape(32, 64)
    assert np.array_equal(ro.resize(x, 32, 64), x)                       # same size: copy
    assert ro.resize(x, 16, 16).shape == (16, 16)                         # avg pool 2 x 4
    assert ro.resize(x, 16, 16)[0, 0] == np.float32((0 + 1 + 2 + 3 + 64 + 65 + 66 + 67) / 8)
    n = ro.resize(x, 32, 48)                                              # 64 % 48 != 0: nearest
    assert n.shape == (32, 48) and n[0, 47] == x[0, min(int(np.floor(47 * (64 / 48))), 63)]


def test_points_in_box_axis_aligned():
    corners = ro.box_corners((0, 0, 0), (2, 2, 2), 0.0)[None]
    pts = np.array([[0, 0, 0], [0.99, 0.99, 0.99], [1.0, 0, 0], [1.01, 0, 0], [0, -1.5, 0]], np.float32)
    assert ro.points_in_bbox_corners(pts, corners)[:, 0].tolist() == [True, True, False, False, False]


def test_device_path_refuses_cpu_tensors():
    """The product path has no CPU fallback: every entry of mobi_b200.lidar raises on host tensors before any launch."""
    import torch
    from mobi_b200 import lidar
    x = torch.zeros(2, 1, 8, 8)
    with pytest.raises(RuntimeError):
        lidar.inverse_depth_normalization(x, 0.0, 0.5)
    with pytest.raises(RuntimeError):
        lidar.postprocess_range_depth(range_depth=x, range_depth_orig=torch.zeros(2, 4, 16), crop_left=torch.tensor([0, 0]),
                                      width_crop=torch.tensor([8, 8]))
    with pytest.raises(RuntimeError):
        lidar.points_in_bbox_corners(torch.zeros(4, 3), torch.zeros(1, 8, 3))
    with pytest.raises(RuntimeError):
        lidar.postprocess_lidar_samples(torch.zeros(1, 3, 8, 8), {}, torch.zeros(1, 8, 3))
    with pytest.raises(RuntimeError):
        lidar.LidarConverter(H=4, W=16).range2pcd(torch.zeros(4, 16), torch.zeros(4, 16), torch.zeros(4, 16))
