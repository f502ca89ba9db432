"""CPU checks of the training-step infrastructure: the training oracle (oracle/train_oracle.py) against golden
gradients produced by the UNMODIFIED reference (LatentDiffusion.p_losses + loss.backward() + torch.optim.AdamW,
tests/golden/train_tiny.npz), and the data-parallel host logic (flat parameter store + ONE all-reduce over the flat
gradient buffer) with world_size 2 on gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sampler_oracle as so
from oracle import train_oracle as to
from oracle import unet_oracle as uo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4  # fp32 vs fp32, reduction order only


def _golden():
    return np.load(os.path.join(GOLDEN, "train_tiny.npz"))


def _oracle_run():
    g = _golden()
    cfg = uo.tiny_unet_config()
    sd = uo.synth_state_dict(uo.state_dict_shapes(cfg), seed=0)
    loss, grads = to.loss_and_grads(sd, cfg, so.register_schedule(), torch.from_numpy(g["x_start"]),
                                    torch.from_numpy(g["t"]), torch.from_numpy(g["noise"]), torch.from_numpy(g["cond"]))
    return g, sd, loss, grads


def test_train_oracle_matches_reference_loss_and_grads():
    g, sd, loss, grads = _oracle_run()
    assert abs(loss.item() - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    names = [str(n) for n in g["names"]]
    assert sorted(grads) == names                       # the trainable set of DiffusionWrapper.__init__
    assert all(to.is_trainable(n) for n in names) and len(names) == 189
    l2 = np.array([grads[n].norm().item() for n in names])
    assert np.all(np.abs(l2 - g["grad_l2"]) <= TOL * np.maximum(g["grad_l2"], 1e-12)), np.abs(l2 - g["grad_l2"]).max()
    full = [k[2:] for k in g.files if k.startswith("g:")]
    assert len(full) == 81
    for n in full:
        ref = torch.from_numpy(g["g:" + n])
        err = ((grads[n] - ref).abs().max() / ref.abs().max()).item()
        assert err <= TOL, (n, err)


def test_adamw_restatement_matches_torch_optim():
    g, sd, loss, grads = _oracle_run()
    for k in g.files:
        if not k.startswith("p1:"):
            continue
        n = k[3:]
        p1, _, _ = to.adamw_step(sd[n], grads[n], torch.zeros_like(sd[n]), torch.zeros_like(sd[n]), step=1)
        ref = torch.from_numpy(g[k])
        # the first AdamW step moves every weight by ~lr: compare the UPDATE, not the weights
        upd, upd_ref = p1 - sd[n], ref - sd[n]
        assert ((upd - upd_ref).abs().max() / upd_ref.abs().max()).item() <= 2e-3, n


def test_flat_params_views_and_selection():
    from mobi_b200.training import FlatParams, is_trainable
    torch.manual_seed(0)

    class Blk(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.attn1 = torch.nn.Linear(8, 8)
            self.cross_modal_attn_camera = torch.nn.ModuleDict(dict(
                to_k=torch.nn.Linear(8, 8, bias=False), to_v=torch.nn.Linear(8, 8, bias=False)))
            self.cond_adapter_norm = torch.nn.LayerNorm(6)

    m = Blk()
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    fp = FlatParams(m, is_trainable, torch.device("cpu"))
    assert [n for n in fp.names] == [n for n, _ in m.named_parameters() if is_trainable(n)]
    for n, p in m.named_parameters():
        assert torch.equal(p.detach(), before[n])
        assert p.requires_grad == is_trainable(n)
        if is_trainable(n):
            assert p.data_ptr() == fp.view(fp.params, n).data_ptr() and fp.offsets[n] % 4 == 0
    fp.params.add_(1.0)                                   # the module sees optimizer updates of the flat buffer
    assert torch.allclose(m.cond_adapter_norm.weight.detach(), before["cond_adapter_norm.weight"] + 1.0)
    kv = fp.pair_view(fp.params, "cross_modal_attn_camera.to_k.weight", "cross_modal_attn_camera.to_v.weight")
    assert kv.shape == (16, 8) and torch.equal(kv[8:], m.cross_modal_attn_camera.to_v.weight.detach())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    from mobi_b200.training import allreduce_mean_
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        # each rank computes the oracle gradients of ITS joint sample; the flat all-reduce must give the gradients of the
        # two-sample batch (mean loss over the batch = mean of the per-rank mean losses for equal shard sizes)
        g = _golden()
        cfg = uo.tiny_unet_config()
        sd = uo.synth_state_dict(uo.state_dict_shapes(cfg), seed=0)
        sl = slice(2 * rank, 2 * rank + 2)
        _, grads = to.loss_and_grads(sd, cfg, so.register_schedule(), torch.from_numpy(g["x_start"])[sl],
                                     torch.from_numpy(g["t"])[sl], torch.from_numpy(g["noise"])[sl],
                                     torch.from_numpy(g["cond"])[sl])
        names = sorted(grads)
        flat = torch.cat([grads[n].reshape(-1) for n in names])
        allreduce_mean_(flat)
        if rank == 0:
            ret["flat"] = flat.clone()
            ret["names"] = names
            ret["sizes"] = [grads[n].numel() for n in names]
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_gradient_allreduce_matches_full_batch():
    g = _golden()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    flat, names, sizes = ret["flat"], ret["names"], ret["sizes"]
    off = 0
    checked = 0
    for n, sz in zip(names, sizes):
        if ("g:" + n) in g.files:
            ref = torch.from_numpy(g["g:" + n]).reshape(-1)
            got = flat[off:off + sz]
            assert ((got - ref).abs().max() / ref.abs().max()).item() <= TOL, n
            checked += 1
        off += sz
    assert checked == 81


def test_wgrad_split_k_policy():
    """Host logic of the split-K weight gradient: slices are a power of two that divides the token count, at least 512 rows
    each, and only as many as it takes for tiles x slices to fill the SMs."""
    from mobi_b200.train_ops import wgrad_splits
    assert wgrad_splits(16384, 320, 320) == 32            # 6 output tiles on 148 SMs
    assert wgrad_splits(8192, 320, 320) == 16             # slices never shorter than 512 rows
    assert wgrad_splits(4096, 640, 640) == 8
    assert wgrad_splits(1024, 1280, 1280) == 2
    assert wgrad_splits(512, 1280, 1280) == 1
    assert wgrad_splits(16384, 2560, 1280) == 1           # 160 tiles already fill the machine
    for M in (520, 1000, 12288):
        s = wgrad_splits(M, 320, 320)
        assert M % s == 0 and (s == 1 or M // s >= 512)
