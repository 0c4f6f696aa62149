"""BASELINE configs[3]: 65536 x 65536 fp32 dense multiply, 2048-block, sharded over the GPUs of one box
(C-stationary grid, NCCL all-gather of the panels, tcgen05 kind::tf32 3xTF32 kernel on every rank).
   torchrun --nproc-per-node 8 tools/bench_cfg4_fp32.py [N] [BLK] [steps] [gemm_algo=3]
Prints one JSON line on rank 0 (TFLOP/s, max over ranks) and checks one output block against a float64 numpy product of the
(fp32-rounded, for gemm_algo 3) panels.  gemm_algo 0 runs the same sharded multiply on the exact fp64 DMMA kernel (the
north star's N = 65536 fp64 point), 4 on the tcgen05 Ozaki-II path."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import matrel_b200 as mb
from matrel_b200.distributed import GridGroups, GridPlan, ShardedMatrix, sharded_multiply

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
algo = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ALGO_NAME = {0: "fp64 DMMA kernel (mma.sync.m8n8k4.f64)", 3: "tcgen05 kind::tf32, 3xTF32 split, fp32 TMEM accumulation",
             4: "tcgen05 kind::i8, Ozaki-II (16 moduli) fp64 emulation"}[algo]
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
device = torch.device("cuda", local_rank)
torch.cuda.set_device(device)
dist.init_process_group("nccl", device_id=device)
rank, world = dist.get_rank(), dist.get_world_size()
plan = GridPlan(world, n, n, blk)
groups = GridGroups(plan, rank)
stream = torch.cuda.Stream(device=device)
with torch.cuda.stream(stream):
    s = mb.MatfastSession(device=local_rank, stream=stream.cuda_stream, gemm_algo=algo)
    A = ShardedMatrix.rand(s, plan, rank, 42, device)
    B = ShardedMatrix.rand(s, plan, rank, 43, device)
    s.sync()
    for _ in range(1 if n >= 65536 and algo != 3 else 2):
        out = sharded_multiply(s, groups, A, B, plan, plan)
        del out
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        out = sharded_multiply(s, groups, A, B, plan, plan)
        if _ < steps - 1:
            del out
    e1.record(stream)
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    err = None
    if rank == 0:
        dC, (gA, gB, dA, dB) = out
        i, j = plan.owned(0)[0]
        rnd = (lambda x: x.astype(np.float32).astype(np.float64)) if algo == 3 else (lambda x: x)
        Arow = np.concatenate([rnd(dA.get_block(i, k).to_numpy()[:256]) for k in range(plan.nbc)], axis=1)
        Bcol = np.concatenate([rnd(dB.get_block(k, j).to_numpy()[:, :256]) for k in range(plan.nbr)], axis=0)
        want = Arow @ Bcol
        got = dC.get_block(i, j).to_numpy()[:256, :256]
        err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
        print(json.dumps({"config": f"{n}x{n} {'fp32' if algo == 3 else 'fp64'} dense multiply, {blk}-block, {plan.pr}x{plan.pc} grid over {world}xB200",
                          "algo": ALGO_NAME, "ms_per_step": ms,
                          ("TFLOPs_fp32_equiv" if algo == 3 else "TFLOPs"): 2.0 * n ** 3 / (ms * 1e-3) / 1e12, "n_gpus": world,
                          ("max_rel_err_vs_fp64_of_fp32_inputs_sampled_block" if algo == 3 else "max_rel_err_vs_numpy_fp64_sampled_block"): err}), flush=True)
    del out
    s.stop()
dist.barrier()
dist.destroy_process_group()
