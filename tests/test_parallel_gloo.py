"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding and the data-parallel
gradient exchange.  The per-rank gradients come from the oracle; the check is SURVEY 8e's equivalence:
SUM all-reduce with the TV term scaled by 1/world == single-process gradient at the global batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import stylenet_oracle as O
from fast_neural_style_transfer_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        p = O.make_net_params(seed=0, random_affine=True)
        vp = O.make_vgg_params(seed=1)
        content = O.make_image(2 * world, 16, 16, seed=5, normalized=True)
        targets = O.style_targets(vp, O.make_image(1, 16, 16, seed=6, normalized=True))
        drop = O.make_dropout_scales(2 * world, seed=7)
        lo, hi = parallel.shard_bounds(content.shape[0], rank, world)
        assert torch.equal(parallel.shard_batch(content, rank, world), content[lo:hi])
        _, grads = O.loss_and_grads(p, vp, content[lo:hi], targets, [d[lo:hi] for d in drop],
                                    tv_weight=10.0 * parallel.tv_weight_scale(world))
        holder = torch.nn.ParameterDict({k.replace(".", "/"): torch.nn.Parameter(v.clone()) for k, v in p.items()})
        for k, v in grads.items():
            holder[k.replace(".", "/")].grad = v.clone()
        parallel.GradientAllReduce(holder, world).all_reduce()
        # zero-copy path: gradients that are views of one flat buffer are reduced in place
        holder2 = torch.nn.ParameterDict({k.replace(".", "/"): torch.nn.Parameter(v.clone()) for k, v in p.items()})
        order = [n.replace("/", ".") for n, _ in holder2.named_parameters()]       # the module's own parameter order
        flat = torch.cat([grads[k].reshape(-1) for k in order])
        for k, piece in zip(order, torch.split(flat, [p[k].numel() for k in order])):
            holder2[k.replace(".", "/")].grad = piece.view_as(p[k])
        ar = parallel.GradientAllReduce(holder2, world)
        assert ar._flat_view() is not None and ar._flat_view().data_ptr() == flat.data_ptr()
        ar.all_reduce()
        for k in p:
            assert torch.allclose(holder2[k.replace(".", "/")].grad, holder[k.replace(".", "/")].grad, rtol=1e-6, atol=0)
        # bucket-wise exchange (the staged backward calls the hook as each slice of the flat buffer becomes final, back to front;
        # asynchronous all-reduces, waited for at the last bucket): same result, and the all_reduce() after backward() is a no-op
        holder3 = torch.nn.ParameterDict({k.replace(".", "/"): torch.nn.Parameter(v.clone()) for k, v in p.items()})
        flat3 = torch.cat([grads[k].reshape(-1) for k in order])
        for k, piece in zip(order, torch.split(flat3, [p[k].numel() for k in order])):
            holder3[k.replace(".", "/")].grad = piece.view_as(p[k])
        ar3 = parallel.GradientAllReduce(holder3, world)
        cuts = [flat3.numel(), flat3.numel() * 2 // 5, 1000, 0]
        for i in range(3):
            ar3._bucket_ready(flat3, cuts[i + 1], cuts[i], i == 2)
        assert ar3._reduced is flat3 and not ar3._pending
        before = flat3.clone()
        assert ar3.all_reduce().data_ptr() == flat3.data_ptr() and torch.equal(flat3, before)      # already reduced: untouched
        assert torch.allclose(flat3, flat, rtol=1e-6, atol=0)
        ar3.all_reduce()                                                                            # a second call reduces again
        assert torch.allclose(flat3, world * before, rtol=1e-6, atol=0)
        assert parallel.all_finite(torch.tensor(1.0), world)
        assert not parallel.all_finite(torch.tensor(float("nan") if rank == 1 else 1.0), world)
        if rank == 0:
            torch.save({k: holder[k.replace(".", "/")].grad for k in p}, os.path.join(out_dir, "dp.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for total in (1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_data_parallel_gradients_match_global_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    dp = torch.load(os.path.join(tmp_path, "dp.pt"))
    p = O.make_net_params(seed=0, random_affine=True)
    vp = O.make_vgg_params(seed=1)
    content = O.make_image(2 * world, 16, 16, seed=5, normalized=True)
    targets = O.style_targets(vp, O.make_image(1, 16, 16, seed=6, normalized=True))
    _, ref = O.loss_and_grads(p, vp, content, targets, O.make_dropout_scales(2 * world, seed=7))
    gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in ref.values())))
    for k in ref:
        err = float((dp[k].double() - ref[k].double()).norm()) / max(float(ref[k].double().norm()), 1e-4 * gn)
        assert err < 2e-4, (k, err)
