"""2+ GPU check of the data-parallel step (torchrun).

1. Every rank evaluates its row shard; the per-layer gradient buckets are all-reduced through the
   library's own NCCL communicator (gpfx_allreduce_grads), overlapped with the backward; rank 0
   compares with the gradient of the whole global minibatch evaluated on one GPU.
2. The same step captured in ONE CUDA graph (all-reduces included) and replayed on fresh minibatches:
   the gradients of replays 2 and 3 must equal the eager gradients of the same minibatches (the
   round-1 FlatGradients re-pointed .grad after the first pack, so replays reduced stale buffers).
Prints the max relative differences; exits non-zero on failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200.architectures import Config, build_constant_input_dim_deep_gp  # noqa: E402
from gpflux_b200.distributed import FlatGradients, GpfxComm, shard_rows  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = GpfxComm()
rng = np.random.default_rng(0)
N, D, M = 2048, 4, 64
NB = 4  # minibatches
X = rng.standard_normal((NB * N, D))
Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((NB * N, 1))
eps_all = torch.from_numpy(rng.standard_normal((NB * N, D))).cuda()
model = build_constant_input_dim_deep_gp(X, 3, Config(M, 1e-2, 1e-2), z_init=X[:M]).cuda()
for layer in model.f_layers:
    layer.q_mu.assign(0.2 * rng.standard_normal(tuple(layer.q_mu.shape)))
Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
a, b = shard_rows(N, rank, world)
n_loc = b - a
assert n_loc * world == N

# everything runs on ONE non-default stream from the first step on: autograd's AccumulateGrad nodes remember
# the stream of their first use, and a graph capture may not depend on the legacy default stream
work_stream = torch.cuda.Stream()
work_stream.wait_stream(torch.cuda.current_stream())
torch.cuda.set_stream(work_stream)

flat = FlatGradients.for_model(model, comm=comm, overlap=True)
sX = torch.empty((n_loc, D), dtype=torch.float64, device="cuda")
sY = torch.empty((n_loc, 1), dtype=torch.float64, device="cuda")
sE = torch.empty((n_loc, D), dtype=torch.float64, device="cuda")


def load(i):
    sX.copy_(Xd[i * N + a : i * N + b])
    sY.copy_(Yd[i * N + a : i * N + b])
    sE.copy_(eps_all[i * N + a : i * N + b])


def step():
    flat.zero_()
    model.loss(sX, sY, eps=[sE, sE, None]).backward()
    flat.allreduce_mean()


# ---- 1. sharded + bucketed all-reduce == full batch
load(0)
step()
torch.cuda.synchronize()
got = flat.flat.clone()
ok = True
if rank == 0:
    single = FlatGradients.for_model(model, comm=None, overlap=False)  # re-points .grad at its own buffer
    single.zero_()
    with flat.paused():  # rank-local evaluation: no collective
        model.loss(Xd[:N], Yd[:N], eps=[eps_all[:N], eps_all[:N], None]).backward()
    ref = single.flat.clone()
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"world={world}: buckets {flat.bucket_numels()}: max |sharded - full| / max|full| = {err:.3e}")
    ok = ok and err < 1e-11
flat._attach()  # .grad views back on the bucketed buffer

# ---- 2. graph replay == eager on later steps
eager = []
for i in range(NB):
    load(i)
    step()
    torch.cuda.synchronize()
    eager.append(flat.flat.clone())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=work_stream):
    step()
worst = 0.0
for i in range(NB):
    load(i)
    g.replay()
    torch.cuda.synchronize()
    e = float((flat.flat - eager[i]).abs().max() / eager[i].abs().max())
    worst = max(worst, e)
    # a stale-buffer bug would reproduce minibatch 0's gradient on every replay
    if i > 0:
        assert float((flat.flat - eager[0]).abs().max()) > 0.0
t = torch.tensor([worst], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world={world}: CUDA-graph replay vs eager over {NB} minibatches: max rel diff = {float(t):.3e}")
    ok = ok and float(t) < 1e-12
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.broadcast(flag, src=0)
passed = float(flag) == 1.0
torch.cuda.synchronize()
del g  # a live graph holds the communicator's captured collectives
torch.cuda.set_stream(torch.cuda.default_stream())
torch.cuda.synchronize()
print(f"rank {rank}: tearing down", flush=True)
comm.destroy()
print(f"rank {rank}: communicator destroyed", flush=True)
dist.destroy_process_group()
print(f"rank {rank}: done", flush=True)
sys.exit(0 if passed else 1)
