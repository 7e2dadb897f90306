"""CPU, world_size 2 over gloo: the batch-sharded loss path gives the single-process result.
The per-shard operator is stood in by the CPU oracle (the CUDA operators need a GPU); what is under test is the
sharding + reduction plumbing of pointcloudcounterfactual_b200.sharding."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_chamfer(recon: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    import oracle

    d1, _, d2, _ = oracle.nn_distance(recon.numpy(), ref.numpy())
    return torch.from_numpy(d1.mean(1) + d2.mean(1))


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pointcloudcounterfactual_b200 import sharding, synthetic

    r, w, _ = sharding.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    recon, ref = synthetic.s1_near(5, 128)  # 5 clouds over 2 ranks: uneven shards (3 + 2)
    local, mean = sharding.ShardedLoss(_oracle_chamfer, rank, world)(recon, ref)
    grad = torch.full((4,), float(rank + 1))
    sharding.all_reduce_mean_(grad)
    # the asynchronous form bench.py uses: two collectives in flight, consumed in order
    h1 = sharding.global_mean_loss_async(local)
    h2 = sharding.global_mean_loss_async(local * 2)
    amean, amean2 = h1.wait(), h2.wait()
    # the per-interval form: local accumulation over several "steps", ONE collective when the value is read
    acc = sharding.LossAccumulator(torch.device("cpu"))
    for step in range(3):
        acc.add(local * (step + 1))
    acc_mean = acc.reduce()
    acc.add(local)
    acc_mean2 = acc.reduce()  # the reset worked: only the last add counts
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), local=local.numpy(), mean=mean.numpy(), grad=grad.numpy(),
             amean=amean.numpy(), amean2=amean2.numpy(), acc_mean=acc_mean.numpy(), acc_mean2=acc_mean2.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loss_matches_single_process(tmp_path):
    from pointcloudcounterfactual_b200 import sharding, synthetic

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    recon, ref = synthetic.s1_near(5, 128)
    full = _oracle_chamfer(recon, ref).numpy()
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    assert np.array_equal(np.concatenate([p["local"] for p in parts]), full)
    for p in parts:
        assert abs(float(p["mean"]) - full.mean()) < 1e-7
        assert np.allclose(p["grad"], 1.5)
        assert abs(float(p["amean"]) - full.mean()) < 1e-7 and abs(float(p["amean2"]) - 2 * full.mean()) < 1e-6
        assert abs(float(p["acc_mean"]) - 2 * full.mean()) < 1e-6  # (1 + 2 + 3) / 3 times the mean over all ranks' clouds
        assert abs(float(p["acc_mean2"]) - full.mean()) < 1e-7
    lo, hi = sharding.shard_bounds(5, 2, 0)
    assert (lo, hi) == (0, 3)
