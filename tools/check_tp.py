"""NCCL check: column-parallel QLinear (sharded canonical int4 weights + all-gather) == single-GPU QLinear."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optimum-quanto_b200"))
import quanto_b200 as q
from quanto_b200.parallel import ColumnParallelQLinear

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
ok = True
for (M, N, K) in ((8, 2048, 1024), (300, 4096, 2048)):
    lin = torch.nn.Linear(K, N, bias=True).to(torch.bfloat16)
    ql = q.QLinear.from_module(lin, weights=q.qint4)
    ql.freeze()
    x = torch.randn(M, K).to(torch.bfloat16).to(dev)
    tp = ColumnParallelQLinear(ql.weight, ql.bias.detach(), rank, world).to(dev)
    y_tp = tp(x)
    y_full = ql.to(dev)(x)
    torch.cuda.synchronize()
    err = (y_tp.float() - y_full.float()).abs().max().item()
    ref = y_full.float().abs().max().item()
    same = torch.equal(y_tp, y_full)
    if rank == 0:
        print(f"M={M} N={N} K={K} world={world}: max|diff|={err:.3e} (max|y|={ref:.3f}) bit-identical={same}")
    ok = ok and err <= 2e-2 * ref
    # all-gather fused into the GEMM epilogue (peer stores): twice, the buffer is reused
    tpf = ColumnParallelQLinear(ql.weight, ql.bias.detach(), rank, world, fused=True).to(dev)
    for rep in range(2):
        y_f = tpf(x).clone()
        torch.cuda.synchronize()
        errf = (y_f.float() - y_full.float()).abs().max().item()
        samef = torch.equal(y_f, y_tp)
        print(f"   rank {rank} fused rep {rep}: max|diff|={errf:.3e} identical-to-NCCL-path={samef}", flush=True)
        ok = ok and errf <= 2e-2 * ref
# timing: bench-shaped problem, NCCL gather vs fused gather (device time, max over ranks)
from quanto_b200.parallel import FusedGather, gather_columns, shard_weight
M, N, K = 4096, 14336, 4096
n_local = N // world
g = torch.Generator().manual_seed(1)
lin = torch.nn.Linear(K, n_local, bias=False).to(torch.bfloat16)
ql = q.QLinear.from_module(lin, weights=q.qint4)
ql.freeze()
ql = ql.to(dev)
x = torch.randn(M, K).to(torch.bfloat16).to(dev)
fg = FusedGather(n_local)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


t_local = timeit(lambda: ql(x))
t_nccl = timeit(lambda: gather_columns(ql(x)))
t_fused = timeit(lambda: fg.forward(x, ql.weight, None))
if rank == 0:
    fl = 2 * M * N * K
    print(f"world={world} M={M} N={N} K={K}: local GEMM {t_local:.3f} ms | GEMM + NCCL all-gather {t_nccl:.3f} ms "
          f"({fl / t_nccl / 1e9:.0f} TF/s) | fused peer-store gather {t_fused:.3f} ms ({fl / t_fused / 1e9:.0f} TF/s)")
dist.barrier()
if rank == 0:
    print("TP CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
