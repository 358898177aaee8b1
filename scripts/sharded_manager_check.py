"""2+ ranks (torchrun): MemoryManager(config['vosmem_shard'] = 'n') -- one tracker spread over the ranks -- replays the
lifecycle recorded from the reference InferenceCore (tests/golden/lifecycle_*.npz: working memory growth, consolidation
into long-term prototypes, eviction, several object groups) and must give what the single-GPU manager gives: the
readout of every frame (also against the reference's own recorded readout), the bank sizes after every event, and the
usage counters that drive consolidation / eviction."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
import vos_e_sam_b200 as vos
from oracle import readout_oracle as orc
from tests.replay import load, lifecycle_config, replay_lifecycle

for name in ('lifecycle_evict.npz', 'lifecycle_groups.npz'):
    z = load(name)
    for dtype, tol in (('fp32', 2e-3), ('bf16', 1e-2)):
        cfg = dict(lifecycle_config(z), vosmem_value_dtype=dtype)
        single = vos.MemoryManager(cfg)
        sharded = vos.MemoryManager(dict(cfg, vosmem_shard='n'))
        want = replay_lifecycle(z, single, device=dev)
        got = replay_lifecycle(z, sharded, device=dev)          # (asserts the bank sizes after every event)
        torch.cuda.synchronize()
        sharded._sharded.sync_usage()
        sharded._sharded.engine.check_status()
        worst_ref = max(orc.rel_err(g.cpu(), ref) for _, g, ref in got)
        worst_single = max(orc.rel_err(g.cpu(), w.cpu()) for (_, g, _), (_, w, _) in zip(got, want))
        use_err = 0.0
        for a, b in ((single.work_mem, sharded.work_mem), (single.long_mem, sharded.long_mem)):
            if a.engaged() and a.count_usage:
                use_err = max(use_err, orc.rel_err(b.use_count.cpu(), a.use_count.cpu()))
                assert torch.equal(a.life_count, b.life_count)
        print(f'rank {rank}: {name} {dtype}: {len(got)} frames, readout vs reference {worst_ref:.2e}, vs single-GPU manager '
              f'{worst_single:.2e}, usage vs single-GPU {use_err:.2e}', flush=True)
        assert worst_ref < tol and worst_single < tol and use_err < 1e-3
dist.barrier()
dist.destroy_process_group()
