"""2+ ranks: the peer-memory exchange and the query-sharded control of ShardedLongTermReadout against the NCCL
all-gather exchange of the N-sharded bank (same results)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tests import synth
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
from vos_e_sam_b200.sharded import ShardedLongTermReadout
g = torch.Generator().manual_seed(5)
n, h, w = 20000, 20, 30
k, s, _ = synth.keys(g, n)
v = torch.randn(1, 128, n, generator=g)
outs = {}
for mode, shard in (('nccl', 'n'), ('peer', 'n'), ('nccl', 'queries')):
    eng = ShardedLongTermReadout(dict(top_k=30, vosmem_exchange=mode, vosmem_shard=shard), rank, world, dev)
    eng.load_long_term(k, s, v)
    gq = torch.Generator().manual_seed(9)
    res = []
    for _ in range(5):                      # several frames: both slots of the double buffer get reused
        qk, qe = synth.query(gq, h, w)
        res.append(eng.match(qk.to(dev), qe.to(dev)).clone())
    torch.cuda.synchronize()
    outs[mode if shard == 'n' else 'queries'] = res
err = max(float((a - b).abs().max()) for a, b in zip(outs['nccl'], outs['peer']))
print(f'rank {rank}: max |nccl - peer| over 5 frames = {err:.3e}', flush=True)
assert err < 1e-5
# query sharding runs other split counts: the same candidates except where the k / k+1 gap is a near-tie (either pick
# is a correct top-k member), fp32 sums in another order
agree = min(float(((a - b).abs().amax(0) < 1e-3).float().mean()) for a, b in zip(outs['nccl'], outs['queries']))
print(f'rank {rank}: n-sharded vs query-sharded: {100 * agree:.1f} % of the query columns agree to 1e-3', flush=True)
assert agree > 0.9
dist.barrier()
dist.destroy_process_group()
