"""The per-batch body of the reference's canonical caller, scripts/inference_test_bench.py:403-464 + 567-629, on the drop-in
model: get_input (4 VAE encodes + latent assembly) -> conditioning tokens -> 50-step sampling with classifier-free
guidance -> decode_sample -> both first-stage decodes -> range-view post-processing (un-crop, instance mask, paste, edited
point cloud).  Every step runs on the device; nothing here is model code, it only sequences the drop-in classes the way
the reference script does, so that `bench.py` and `tools/inference_test_bench.py` time / run the same thing a user runs.

Out of scope: the nuScenes dataset / dataloader, the CLIP vision tower (an injected module: the
conditioning stage is entered at its pooler_output), the per-sample cv2 compositing and file writing on the host.
"""
import torch

from . import lidar


def make_sampler(model, plms=False, **kw):
    """inference_test_bench.py:346-349."""
    from .ddim import DDIMSampler
    from .plms import PLMSSampler
    return PLMSSampler(model, **kw) if plms else DDIMSampler(model, **kw)


def unconditional_conditioning(model, rows):
    """inference_test_bench.py:423-428: [learnable_vector, bbox_uncond_vector] repeated for every UNet row."""
    uc = [model.learnable_vector.detach().float().repeat(rows, 1, 1)]
    keys = model.cond_stage_key if isinstance(model.cond_stage_key, (list, tuple)) else [model.cond_stage_key]
    if "ref_bbox" in keys:
        uc.append(model.bbox_uncond_vector.detach().float().repeat(rows, 1, 1))
    return torch.cat(uc, dim=1).contiguous()


@torch.no_grad()
def inpaint_batch(model, sampler, batch, *, ddim_steps=50, scale=5.0, ddim_eta=0.0, start_code=None, noise=None,
                  postprocess=True, range_object_norm=True, range_object_norm_scale=0.75):
    """One batch of the reference dataset's layout (ldm/data/nuscenes.py:452-489), already on the device:
      batch["image"] = {GT, inpaint_image, inpaint_mask, cond: {ref_image, ref_bbox}}
      batch["lidar"] = {range_data, range_data_inpaint, range_mask, cond: {ref_image, ref_bbox}, range_depth_orig,
                        range_int_orig, range_pitch, range_yaw, range_instance_mask_orig, range_shift_left, width_crop,
                        min_depth_obj, max_depth_obj}
      batch["bbox_3d"] = [N, 8, 3] box corners in the lidar frame.
    Returns samples (interleaved latents), image_sample [N, 3, H, W] and range_sample [N, 2, H, W] clamped to [-1, 1]
    (LatentDiffusion.log_data, ddpm.py:1491-1538), and, with `postprocess`, the outputs of
    lidar.postprocess_lidar_samples (edited sweeps, instance masks, compacted point clouds)."""
    data = model.get_input(batch, model.first_stage_key, noise=noise)
    c = model.get_learned_conditioning(data["cond"])
    rows = data["z"].shape[0]
    uc = unconditional_conditioning(model, rows) if scale != 1.0 else None
    shape = [model.channels, model.image_size, model.image_size]
    z = data["z"]
    rest = dict(inpaint_image=z[:, 4:8].contiguous(), inpaint_mask=z[:, 8:9].contiguous())
    from .plms import PLMSSampler
    kw = rest if isinstance(sampler, PLMSSampler) else dict(test_model_kwargs=rest)     # plms.py:218 vs ddim.py:170-172
    samples, _ = sampler.sample(S=ddim_steps, conditioning=c, batch_size=rows, shape=shape, verbose=False,
                                unconditional_guidance_scale=scale, unconditional_conditioning=uc, eta=ddim_eta,
                                x_T=start_code, **kw)
    h_camera, h_lidar = model.decode_sample(samples, data.get("z_lidar"))
    out = {"samples": samples, "cond": c, "z": z}
    img = model.decode_first_stage(h_camera.contiguous())
    out["image_decoded"] = img
    out["image_sample"] = lidar._range_map(img, lidar.MAP_NONE, clamp=True)                # torch.clamp(x, -1, 1)
    rng = model.decode_first_stage(h_lidar.contiguous(), module_name="lidar_stage_model")
    out["range_sample"] = rng
    if postprocess:
        out.update(lidar.postprocess_lidar_samples(rng, batch["lidar"], batch["bbox_3d"],
                                                   range_object_norm=range_object_norm,
                                                   range_object_norm_scale=range_object_norm_scale))
    return out


def batch_to_device(batch, device, non_blocking=True):
    """inference_test_bench.py:move_to_device: nested dict of tensors -> device (pinned host memory makes it asynchronous)."""
    if isinstance(batch, dict):
        return {k: batch_to_device(v, device, non_blocking) for k, v in batch.items()}
    if torch.is_tensor(batch):
        return batch.to(device, non_blocking=non_blocking)
    return batch


def batch_bytes(batch):
    if isinstance(batch, dict):
        return sum(batch_bytes(v) for v in batch.values())
    return batch.numel() * batch.element_size() if torch.is_tensor(batch) else 0


def batch_slice(batch, a, b):
    if isinstance(batch, dict):
        return {k: batch_slice(v, a, b) for k, v in batch.items()}
    return batch[a:b] if torch.is_tensor(batch) else batch
