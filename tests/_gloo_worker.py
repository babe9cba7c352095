"""Worker of tests/test_multirank_cpu.py: one of WORLD_SIZE CPU processes under torchrun (gloo).

Restates, with gloo in place of NCCL and the oracle in place of the kernels, the protocol the library runs per compute
call when particles are sharded over ranks (csrc/ta_b200.cu finish_timeseries): every rank sums the per-particle rows of
its own contiguous particle range, appends its particle count, ONE all-reduce(sum) of T + 1 doubles, divide by the total
count.  Writes the result of each rank to <out>/rank<r>.npz.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (test infrastructure)


def main():
    out, T, N, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vel = np.random.default_rng(seed).standard_normal((T, N, 3)).astype(np.float32).astype(np.float64)   # same on every rank
    a0, a1 = rank * N // world, (rank + 1) * N // world                                                  # SURVEY 8e partition
    bp, _ = oracle.vacf_fft(vel[:, a0:a1])
    payload = torch.zeros(T + 1, dtype=torch.float64)
    payload[:T] = torch.from_numpy(bp.sum(axis=1))
    payload[T] = a1 - a0
    dist.all_reduce(payload, op=dist.ReduceOp.SUM)
    ts = (payload[:T] / payload[T]).numpy()
    # max-over-ranks of a per-rank time, as bench.py does for every timed step
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    np.savez(os.path.join(out, f"rank{rank}.npz"), ts=ts, count=float(payload[T]), tmax=float(t[0]), a0=a0, a1=a1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
