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
dist.barrier()
if rank == 0:
    print("TP CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
