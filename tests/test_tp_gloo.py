"""world_size-2 check of the N>1 host path on CPU (gloo): weight sharding + the all-reduce that
follows each row-parallel projection reproduce the unsharded layer (oracle math)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import lia_b200  # noqa: F401
    from lia_b200 import tp, weights
    r, w_ = tp.init_from_env("gloo")
    assert (r, w_) == (rank, world) and tp.world_size() == world
    h, f, H = 128, 512, 2
    w = weights.random_layer(h, f, seed=11, bias_std=0.05, ln_std=0.1)
    lay = weights.LayerLayout(h, f, world)
    v = lay.views(weights.pack_layer(w, lay, rank))
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 4, h, generator=g).to(torch.bfloat16)
    # row-parallel fc1 -> fc2 with the product's sharding, reduced with the product's all_reduce
    ln = torch.nn.functional.layer_norm(x.float(), (h,), w["ln2_w"].float(), w["ln2_b"].float(), 1e-5)
    part = torch.relu(ln @ v["fc1_w"].float().t() + v["fc1_b"].float()) @ v["fc2_w"].float().t() + v["fc2_b"].float()
    tp.all_reduce(part)
    full = torch.relu(ln @ w["fc1_w"].float().t() + w["fc1_b"].float()) @ w["fc2_w"].float().t() + w["fc2_b"].float()
    ok = torch.allclose(part, full, atol=2e-2, rtol=2e-2)
    t = tp.max_over_ranks(float(rank), "cpu")
    ret[rank] = bool(ok) and t == float(world - 1)
    tp.barrier()
    dist.destroy_process_group()


def test_tp_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
