"""Worker for tests/test_multi_rank_cpu.py: world_size-2 gloo run of the sharding logic on CPU.
Each rank takes its contiguous shard of a 10-instance batch (qb200_shard_range), generates exactly
its slice of the global synthetic stream (counter RNG with first_index), evaluates it with the CPU
oracle (test infrastructure; there is no GPU here), and the shards are all-gathered and compared
with the unsharded evaluation.  No collective is on the data path: the gather is only the check."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from qublas_b200 import ops  # noqa: E402
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    total, M = 10, 8
    a = QType(7, 8)
    plan = ops.make_gemul_plan(a, a, QType(17, 4, True, RND_POS_INF, SAT_TCPL), M, M, M, mul=QType(14, 8, True, RND_CONV, SAT_TCPL),
                               add=[QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(3)])
    first, count = ops.shard_range(total, world, rank)
    words = M * M
    A = oracle.synth_fill(count * words, 4, a.raw_min, a.raw_max, 1, 0, first * words)
    B = oracle.synth_fill(count * words, 4, a.raw_min, a.raw_max, 1, 1, first * words)
    mine = oracle.gemul_batched(plan, A, B, count)
    # gather (variable shard sizes -> pad to the largest)
    cap = (total + world - 1) // world * words
    buf = torch.zeros(cap, dtype=torch.int32); buf[:count * words] = torch.from_numpy(mine)
    outs = [torch.zeros(cap, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(outs, buf)
    meta = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(meta, torch.tensor([first, count]))
    if rank == 0:
        Af = oracle.synth_fill(total * words, 4, a.raw_min, a.raw_max, 1, 0, 0)
        Bf = oracle.synth_fill(total * words, 4, a.raw_min, a.raw_max, 1, 1, 0)
        full = oracle.gemul_batched(plan, Af, Bf, total)
        got = np.concatenate([outs[r][:int(meta[r][1]) * words].numpy() for r in range(world)])
        assert [int(m[0]) for m in meta] == [0, 5] and [int(m[1]) for m in meta] == [5, 5]
        assert np.array_equal(got, full)
        print("SHARD_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
