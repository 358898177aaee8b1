"""2+ ranks (torchrun): the N-sharded long-term readout -- peer-memory exchange ('peer') and the all-to-all exchange
('nccl'), gathered and as query slices -- and the query-sharded control against the UNSHARDED readout of the same bank
computed on every rank's own GPU (vosmem_match).  Several frames, so both slots of the double buffers are reused."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tests import synth
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
import vos_e_sam_b200 as vos
from vos_e_sam_b200 import ops
from vos_e_sam_b200.sharded import ShardedLongTermReadout
g = torch.Generator().manual_seed(5)
n, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (20000, 20, 30)
frames = 6
k, s, _ = synth.keys(g, n)
v = torch.randn(1, 128, n, generator=g)
gq = torch.Generator().manual_seed(9)
queries = [tuple(t.to(dev) for t in synth.query(gq, h, w)) for _ in range(frames)]

# unsharded: the whole bank on this GPU
store = vos.KeyValueMemoryStore(count_usage=False)
store.add(k.to(dev), [v.to(dev)], s.to(dev), None, None)
want = [ops.match(qk.flatten(2)[0], qe.flatten(2)[0], [store.key_segment(0, n)], [store.value_segment(0, 0, with_usage=False)],
                  128, 30).clone() for qk, qe in queries]

worst = {}
for mode, shard in (('nccl', 'n'), ('peer', 'n'), ('nccl', 'queries')):
    eng = ShardedLongTermReadout(dict(top_k=30, vosmem_exchange=mode, vosmem_shard=shard), rank, world, dev)
    eng.load_long_term(k, s, v)
    q_lo, q_hi = eng.query_range(h * w)
    full = [eng.match(qk, qe).clone() for qk, qe in queries]
    part = [eng.match(qk, qe, gather=False).clone() for qk, qe in queries]
    torch.cuda.synchronize()
    eng.check_status()
    name = mode if shard == 'n' else 'queries'
    # same kernels, same scores: only near-tie picks (k / k+1 gap ~ 0) and the fp32 summation order may differ
    agree = min(float(((a - b).abs().amax(0) < 1e-3).float().mean()) for a, b in zip(full, want))
    err_part = max(float((p - f[:, q_lo:q_hi]).abs().max()) for p, f in zip(part, full))
    worst[name] = agree
    print(f'rank {rank}: {name:8s} vs unsharded: {100 * agree:.2f} % of the query columns agree to 1e-3; '
          f'slice vs gathered max diff {err_part:.1e}', flush=True)
    assert agree > 0.98 and err_part < 1e-6
    if name == 'nccl':
        ref = full
    elif name == 'peer':
        err = max(float((a - b).abs().max()) for a, b in zip(ref, full))
        print(f'rank {rank}: max |nccl - peer| over {frames} frames = {err:.3e}', flush=True)
        assert err < 1e-5
dist.barrier()
dist.destroy_process_group()
