"""TEST INFRASTRUCTURE ONLY — generates tests/golden/range_post.npz by running the UNMODIFIED reference's range-view
post-processing (ldm.data.utils / ldm.data.lidar_converter / ldm.data.box_np_ops with the real cv2, numba and torch
CPU kernels) on the deterministic synthetic sweeps of oracle/range_oracle.py.

Run here (the build container) with:  python -m oracle.make_golden_range
Only OUTPUTS (and the box corners, whose sin/cos are platform-rounded) are stored; tests regenerate the inputs from
the seed.  The sequence is ddpm.py:1503-1543 (clamp + in-place inverse depth normalisation of the logged sample) followed
by scripts/inference_test_bench.py:567-629 (save_samples).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import range_oracle, ref_shims  # noqa: E402

SEED = 20240917
CASES = dict(B=3, H=32, W=1096, crop=128, width_crops=(64, 128, 100))      # 100: 128 % 100 != 0 -> the cv2 nearest path


def reference_pipeline(inp, bbox_3d, alpha=0.75):
    from ldm.data.box_np_ops import points_in_bbox_corners
    from ldm.data.lidar_converter import LidarConverter
    from ldm.data.utils import inverse_depth_normalization, postprocess_range_depth_int

    B = len(inp["range_depth"])
    lidar_sample = torch.cat([torch.from_numpy(inp["range_depth"]), torch.from_numpy(inp["range_int"])], 1)
    lidar_sample = torch.clamp(lidar_sample, -1., 1.)
    sample_depth, sample_int = lidar_sample[:, [0]], lidar_sample[:, [1]]
    for i in range(B):
        sample_depth[i] = inverse_depth_normalization(sample_depth[i], torch.tensor(inp["min_depth_obj"][i]),
                                                      torch.tensor(inp["max_depth_obj"][i]), alpha=alpha)
    sd, si = postprocess_range_depth_int(
        range_depth=sample_depth, range_depth_orig=torch.from_numpy(inp["range_depth_orig"]),
        range_int=sample_int, range_int_orig=torch.from_numpy(inp["range_int_orig"]),
        crop_left=torch.from_numpy(inp["crop_left"]), width_crop=torch.from_numpy(inp["width_crop"]))
    out = dict(unnorm_depth=sample_depth.numpy(), range_sample_depth=sd, range_sample_int=si, pred_instance_mask=[],
               range_pred=[], pred_points=[], n_points=[])
    for i in range(B):
        conv = LidarConverter()
        pitch, yaw = inp["range_pitch"][i], inp["range_yaw"][i]
        gt = inp["range_instance_mask_orig"][i]
        pred_mask = np.zeros(np.prod(gt.shape))
        label = np.arange(0, np.prod(gt.shape)).reshape(gt.shape)
        points, points_label, _ = conv.range2pcd(sd[i], pitch, yaw, label)
        object_points = points_in_bbox_corners(points, bbox_3d[[i]])
        pred_mask[points_label[object_points[:, 0]]] = 1
        pred_mask = pred_mask.reshape(gt.shape)
        instance_mask = np.logical_or(pred_mask, gt)
        depth_final = np.where(instance_mask, sd[i], inp["range_depth_orig"][i])
        int_final = np.where(instance_mask, si[i], inp["range_int_orig"][i])
        xyz, pts_int, beam = conv.range2pcd(depth_final, pitch, yaw, int_final)
        out["pred_instance_mask"].append(pred_mask.astype(np.uint8))
        out["range_pred"].append(np.stack([depth_final, int_final, pitch, yaw]).astype(np.float32))
        pp = np.concatenate([xyz, pts_int[:, None], beam[:, None]], axis=1)
        out["pred_points"].append(pp.astype(np.float32))
        out["n_points"].append(len(pp))
    return out


def main():
    ref_shims.install()
    from ldm.data.box_np_ops import center_to_corner_box3d
    from ldm.data.utils import depth_normalization, inverse_depth_normalization

    inp = range_oracle.synth_range_inputs(SEED, **CASES)
    bbox = range_oracle.synth_boxes(inp)
    # the oracle's corner convention against the reference's center_to_corner_box3d (lidar: axis 2)
    for (yaw_c, r), mine in zip(inp["centers"], bbox):
        center = np.array([[r * np.cos(yaw_c), -r * np.sin(yaw_c), -0.4]], np.float32)
        ref_c = center_to_corner_box3d(center, np.array([[4.6, 2.2, 1.8]], np.float32), np.array([-yaw_c], np.float32),
                                       origin=(0.5, 0.5, 0.5), axis=2)[0]
        assert np.abs(ref_c - mine).max() < 1e-4, np.abs(ref_c - mine).max()
    ref = reference_pipeline(inp, bbox)
    print("points per sample:", ref["n_points"], " generated-object pixels:",
          [int(m.sum()) for m in ref["pred_instance_mask"]])

    # scalar maps over a grid that crosses every piece boundary (utils.py:537-580)
    grid = torch.linspace(-1, 1, 4001)
    mn, mx = torch.tensor(-0.66), torch.tensor(-0.6)
    fwd = depth_normalization(grid.clone(), mn, mx, alpha=0.75).numpy()
    inv = inverse_depth_normalization(grid.clone(), mn, mx, alpha=0.75).numpy()
    int_un = torch.clamp(-0.5 * torch.log(1 - (grid + 1) / 2) - 1, -1, 1).numpy()                 # ddpm.py:1541

    # keep the fixture small: only the pasted window of the un-cropped images (everything else must equal the original
    # sweep, which tests regenerate), every 7th point of the ordered clouds plus float64 column sums, one un-normalised crop
    W = CASES["W"]
    win = [(int(inp["crop_left"][b]) % W + np.arange(int(inp["width_crop"][b]))) % W for b in range(CASES["B"])]
    extra = {}
    for b in range(CASES["B"]):
        extra["window_depth_%d" % b] = ref["range_sample_depth"][b][:, win[b]]
        extra["window_int_%d" % b] = ref["range_sample_int"][b][:, win[b]]
        extra["points_every7_%d" % b] = ref["pred_points"][b][::7]
        extra["points_colsum_%d" % b] = ref["pred_points"][b].astype(np.float64).sum(0)
    np.savez_compressed(
        os.path.join(GOLDEN, "range_post.npz"), seed=SEED, B=CASES["B"], H=CASES["H"], W=W, crop=CASES["crop"],
        width_crops=np.array(CASES["width_crops"]), bbox_3d=bbox, unnorm_depth_0=ref["unnorm_depth"][0],
        pred_instance_mask=np.packbits(np.stack(ref["pred_instance_mask"]).astype(bool)),
        n_points=np.array(ref["n_points"]), grid_min=mn.numpy(), grid_max=mx.numpy(), grid_fwd=fwd, grid_inv=inv,
        grid_int=int_un, **extra)
    print("wrote range_post.npz:", os.path.getsize(os.path.join(GOLDEN, "range_post.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
