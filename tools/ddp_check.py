"""2+ GPU check of the data-parallel step (torchrun): every rank evaluates its row shard and
all-reduces the flat gradient buffer (NCCL AVG); rank 0 compares with the gradient of the whole
global minibatch evaluated on one GPU.  Prints the max relative difference."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200.architectures import Config, build_constant_input_dim_deep_gp  # noqa: E402
from gpflux_b200.distributed import FlatGradients, shard_rows  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(0)
N, D, M = 2048, 4, 64
X = rng.standard_normal((N, D))
Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
eps = torch.from_numpy(rng.standard_normal((N, D))).cuda()
model = build_constant_input_dim_deep_gp(X, 2, Config(M, 1e-2, 1e-2), z_init=X[:M]).cuda()
Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
a, b = shard_rows(N, rank, world)
flat = FlatGradients(model.parameters())
flat.zero_()
model.loss(Xd[a:b], Yd[a:b], eps=[eps[a:b], None]).backward()
flat.allreduce_mean()
got = flat.flat.clone()
if rank == 0:
    flat.zero_()
    model.loss(Xd, Yd, eps=[eps, None]).backward()
    ref = flat.pack()
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"world={world}: max |sharded - full| / max|full| = {err:.3e}")
    assert err < 1e-11, err
dist.barrier()
dist.destroy_process_group()
